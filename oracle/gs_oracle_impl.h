/*
 * CPU ORACLE (test infrastructure, NOT the product) -- restatement of the
 * reference differentiable 3DGS rasterizer
 *   RAST = /root/reference/submodules/depth-diff-gaussian-rasterization-min
 * written from the maths in SURVEY.md Appendix A.  Every function cites the
 * reference lines whose behaviour it reproduces.  No reference source is
 * copied: data layout, control flow and the sort are our own.
 *
 * This file is included twice by gs_oracle.c:
 *   REAL=float   SUFFIX=f32   "faithful" mode: same arithmetic type as the
 *                             reference kernels (IEEE fp32, expf, ndc2Pix in
 *                             double), build with -ffp-contract=off
 *   REAL=double  SUFFIX=f64   "truth" mode used for tolerance budgeting and
 *                             finite-difference gradient checks
 *
 * Parity status: the reference has no tests / golden vectors for this path
 * (SURVEY.md section 4); the oracle is pinned against outputs of the
 * reference CUDA sources themselves, built by oracle/Makefile into
 * oracle/_ref/ and run on the B200 box (tests/golden/).
 */

#ifndef REAL
#error "include from gs_oracle.c"
#endif

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct FN(GsoState) {
    int P, W, H, gx, gy;
    long long R;           /* num_rendered */
    /* geometry state (rasterizer_impl.h:33-47) */
    REAL* depths;          /* P */
    unsigned char* clamped;/* 3P */
    int* radii;            /* P */
    REAL* means2D;         /* 2P */
    REAL* cov3D;           /* 6P */
    REAL* conic_opacity;   /* 4P */
    REAL* rgb;             /* 3P */
    uint32_t* tiles_touched; /* P */
    uint32_t* point_offsets; /* P (inclusive scan) */
    /* binning state */
    uint64_t* keys;        /* R sorted */
    uint32_t* point_list;  /* R sorted */
    /* image state */
    uint32_t* ranges;      /* 2*G */
    REAL* final_T;         /* N */
    uint32_t* n_contrib;   /* N */
} FN(GsoState);

typedef struct FN(GsoParams) {
    int P, D, M, W, H;
    REAL tanfovx, tanfovy, scale_modifier;
    const REAL* bg;             /* 3 */
    const REAL* means3D;        /* 3P */
    const REAL* shs;            /* P*M*3 or NULL */
    const REAL* colors_precomp; /* 3P or NULL */
    const REAL* opacities;      /* P */
    const REAL* scales;         /* 3P or NULL */
    const REAL* rotations;      /* 4P or NULL */
    const REAL* cov3D_precomp;  /* 6P or NULL */
    const REAL* viewmatrix;     /* 16, m[r+4c] = V[r][c] */
    const REAL* projmatrix;     /* 16 */
    const REAL* campos;         /* 3 */
} FN(GsoParams);

#define RC(x) ((REAL)(x))

#if IS_F32
#define R_EXP expf
#define R_SQRT sqrtf
#define R_CEIL ceilf
#else
#define R_EXP exp
#define R_SQRT sqrt
#define R_CEIL ceil
#endif

/* SH constants: auxiliary.h:22-39 */
static const REAL FN(SH_C0) = RC(0.28209479177387814);
static const REAL FN(SH_C1) = RC(0.4886025119029199);
static const REAL FN(SH_C2)[5] = {RC(1.0925484305920792), RC(-1.0925484305920792), RC(0.31539156525252005),
                                  RC(-1.0925484305920792), RC(0.5462742152960396)};
static const REAL FN(SH_C3)[7] = {RC(-0.5900435899266435), RC(2.890611442640554), RC(-0.4570457994644658),
                                  RC(0.3731763325901154), RC(-0.4570457994644658), RC(1.445305721320277),
                                  RC(-0.5900435899266435)};

/* auxiliary.h:41-44 -- evaluated in double, rounded to REAL */
static inline REAL FN(ndc2pix)(REAL v, int S) { return (REAL)((((double)v + 1.0) * S - 1.0) * 0.5); }

/* auxiliary.h:58-77: V.[p;1]; m[r + 4c] = V[r][c]; summation order left to right */
static inline void FN(xf4x3)(const REAL* p, const REAL* m, REAL* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void FN(xf4x4)(const REAL* p, const REAL* m, REAL* o) {
    FN(xf4x3)(p, m, o);
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* auxiliary.h:46-56. Float quotient truncated toward zero, clamped to [0, grid]. */
static inline void FN(get_rect)(REAL px, REAL py, int radius, int gx, int gy, int* rmin, int* rmax) {
    int a;
    a = (int)((px - radius) / TILE);            rmin[0] = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)((py - radius) / TILE);            rmin[1] = a < 0 ? 0 : (a > gy ? gy : a);
    a = (int)((px + radius + TILE - 1) / TILE); rmax[0] = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)((py + radius + TILE - 1) / TILE); rmax[1] = a < 0 ? 0 : (a > gy ? gy : a);
}

/* Standard rotation matrix of the UN-normalised quaternion (r,x,y,z); forward.cu:127-139.
 * R is the ordinary (row,col) matrix; the reference's glm matrix is its transpose. */
static inline void FN(quat_to_R)(const REAL* q, REAL R[3][3]) {
    REAL r = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = RC(1) - RC(2) * (y * y + z * z); R[0][1] = RC(2) * (x * y - r * z);         R[0][2] = RC(2) * (x * z + r * y);
    R[1][0] = RC(2) * (x * y + r * z);         R[1][1] = RC(1) - RC(2) * (x * x + z * z); R[1][2] = RC(2) * (y * z - r * x);
    R[2][0] = RC(2) * (x * z - r * y);         R[2][1] = RC(2) * (y * z + r * x);         R[2][2] = RC(1) - RC(2) * (x * x + y * y);
}

/* forward.cu:118-152: Sigma = R S S R^T, upper triangle. M[i][j] = s_i R[j][i]; Sigma[a][b] = sum_i M[i][a] M[i][b]. */
static inline void FN(cov3d)(const REAL* scale, REAL mod, const REAL* q, REAL* c6) {
    REAL R[3][3], M[3][3];
    FN(quat_to_R)(q, R);
    for (int i = 0; i < 3; i++) {
        REAL s = mod * scale[i];
        for (int j = 0; j < 3; j++) M[i][j] = s * R[j][i];
    }
#define SIG(a, b) (M[0][a] * M[0][b] + M[1][a] * M[1][b] + M[2][a] * M[2][b])
    c6[0] = SIG(0, 0); c6[1] = SIG(0, 1); c6[2] = SIG(0, 2);
    c6[3] = SIG(1, 1); c6[4] = SIG(1, 2); c6[5] = SIG(2, 2);
#undef SIG
}

/* Shared by forward (forward.cu:74-113) and backward (backward.cu:144-207):
 * A = J.V3 (rows 0,1 non-zero), cov2D = A Sigma A^T (+0.3 on the diagonal). */
typedef struct FN(Cov2DCtx) {
    REAL t[3];        /* clamped view-space mean */
    REAL xmul, ymul;  /* 0 when clamped (backward.cu:175-176) */
    REAL A[2][3];
    REAL a, b, c;     /* cov00+0.3, cov01, cov11+0.3 */
} FN(Cov2DCtx);

static inline void FN(cov2d)(const REAL* mean, REAL fx, REAL fy, REAL tanfovx, REAL tanfovy, const REAL* c6,
                             const REAL* vm, FN(Cov2DCtx)* o) {
    REAL t[3];
    FN(xf4x3)(mean, vm, t);
    const REAL limx = RC(1.3) * tanfovx, limy = RC(1.3) * tanfovy;
    const REAL txtz = t[0] / t[2], tytz = t[1] / t[2];
    REAL cx = txtz < -limx ? -limx : txtz; cx = cx > limx ? limx : cx;   /* min(limx, max(-limx, txtz)) */
    REAL cy = tytz < -limy ? -limy : tytz; cy = cy > limy ? limy : cy;
    t[0] = cx * t[2];
    t[1] = cy * t[2];
    o->xmul = (txtz < -limx || txtz > limx) ? RC(0) : RC(1);
    o->ymul = (tytz < -limy || tytz > limy) ? RC(0) : RC(1);
    o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
    const REAL J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    const REAL J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* V3[r][c] = vm[r + 4c];  A[i][j] = sum_k V3[k][j] * J[i][k]  (k = 0,1,2 in that order) */
    for (int j = 0; j < 3; j++) {
        const REAL v0 = vm[0 + 4 * j], v1 = vm[1 + 4 * j], v2 = vm[2 + 4 * j];
        o->A[0][j] = v0 * J00 + v1 * RC(0) + v2 * J02;
        o->A[1][j] = v0 * RC(0) + v1 * J11 + v2 * J12;
    }
    const REAL S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    REAL B[2][3];
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 3; k++) B[i][k] = o->A[i][0] * S[0][k] + o->A[i][1] * S[1][k] + o->A[i][2] * S[2][k];
    o->a = (B[0][0] * o->A[0][0] + B[0][1] * o->A[0][1] + B[0][2] * o->A[0][2]) + RC(0.3);
    o->b = (B[0][0] * o->A[1][0] + B[0][1] * o->A[1][1] + B[0][2] * o->A[1][2]);
    o->c = (B[1][0] * o->A[1][0] + B[1][1] * o->A[1][1] + B[1][2] * o->A[1][2]) + RC(0.3);
}

/* SH basis polynomials for unit direction (x,y,z): forward.cu:30-59 == utils/sh.py:74-100 */
static inline void FN(sh_basis)(int deg, REAL x, REAL y, REAL z, REAL* b) {
    b[0] = FN(SH_C0);
    if (deg > 0) {
        b[1] = -FN(SH_C1) * y; b[2] = FN(SH_C1) * z; b[3] = -FN(SH_C1) * x;
        if (deg > 1) {
            REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = FN(SH_C2)[0] * xy;
            b[5] = FN(SH_C2)[1] * yz;
            b[6] = FN(SH_C2)[2] * (RC(2) * zz - xx - yy);
            b[7] = FN(SH_C2)[3] * xz;
            b[8] = FN(SH_C2)[4] * (xx - yy);
            if (deg > 2) {
                b[9] = FN(SH_C3)[0] * y * (RC(3) * xx - yy);
                b[10] = FN(SH_C3)[1] * xy * z;
                b[11] = FN(SH_C3)[2] * y * (RC(4) * zz - xx - yy);
                b[12] = FN(SH_C3)[3] * z * (RC(2) * zz - RC(3) * xx - RC(3) * yy);
                b[13] = FN(SH_C3)[4] * x * (RC(4) * zz - xx - yy);
                b[14] = FN(SH_C3)[5] * z * (xx - yy);
                b[15] = FN(SH_C3)[6] * x * (xx - RC(3) * yy);
            }
        }
    }
}

static inline int FN(n_active)(int deg) { return (deg + 1) * (deg + 1); }

/* ---------------------------------------------------------------- forward */

/* forward.cu:155-256 (preprocessCUDA) + auxiliary.h:139-164 (in_frustum) */
static void FN(preprocess_one)(const FN(GsoParams)* p, FN(GsoState)* s, int i, REAL fx, REAL fy) {
    s->radii[i] = 0;
    s->tiles_touched[i] = 0;
    const REAL* m = p->means3D + 3 * i;
    REAL pv[3], ph[4];
    FN(xf4x3)(m, p->viewmatrix, pv);
    if (pv[2] <= RC(0.2)) return;                       /* only the near plane culls */
    FN(xf4x4)(m, p->projmatrix, ph);
    const REAL pw = RC(1) / (ph[3] + RC(0.0000001));
    const REAL ppx = ph[0] * pw, ppy = ph[1] * pw;

    const REAL* c6;
    if (p->cov3D_precomp) c6 = p->cov3D_precomp + 6 * i;
    else {
        FN(cov3d)(p->scales + 3 * i, p->scale_modifier, p->rotations + 4 * i, s->cov3D + 6 * i);
        c6 = s->cov3D + 6 * i;
    }
    FN(Cov2DCtx) cc;
    FN(cov2d)(m, fx, fy, p->tanfovx, p->tanfovy, c6, p->viewmatrix, &cc);
    const REAL det = cc.a * cc.c - cc.b * cc.b;
    if (det == RC(0)) return;
    const REAL det_inv = RC(1) / det;
    const REAL conx = cc.c * det_inv, cony = -cc.b * det_inv, conz = cc.a * det_inv;
    const REAL mid = RC(0.5) * (cc.a + cc.c);
    REAL disc = mid * mid - det; if (!(disc > RC(0.1))) disc = RC(0.1);   /* max(0.1f, .) */
    const REAL l1 = mid + R_SQRT(disc), l2 = mid - R_SQRT(disc);
    const REAL radf = R_CEIL(RC(3) * R_SQRT(l1 > l2 ? l1 : l2));
    const REAL px = FN(ndc2pix)(ppx, p->W), py = FN(ndc2pix)(ppy, p->H);
    int rmin[2], rmax[2];
    FN(get_rect)(px, py, (int)radf, s->gx, s->gy, rmin, rmax);
    if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) return;

    if (!p->colors_precomp) {                            /* forward.cu:20-71 */
        REAL d[3] = {m[0] - p->campos[0], m[1] - p->campos[1], m[2] - p->campos[2]};
        const REAL len = R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] = d[0] / len; d[1] = d[1] / len; d[2] = d[2] / len;
        REAL b[16];
        FN(sh_basis)(p->D, d[0], d[1], d[2], b);
        const REAL* sh = p->shs + (size_t)i * p->M * 3;
        const int na = FN(n_active)(p->D);
        for (int ch = 0; ch < 3; ch++) {
            REAL acc = b[0] * sh[ch];
            for (int k = 1; k < na; k++) acc = acc + b[k] * sh[3 * k + ch];
            acc += RC(0.5);
            s->clamped[3 * i + ch] = (acc < RC(0));
            s->rgb[3 * i + ch] = acc < RC(0) ? RC(0) : acc;
        }
    }
    s->depths[i] = pv[2];
    s->radii[i] = (int)radf;
    s->means2D[2 * i] = px; s->means2D[2 * i + 1] = py;
    s->conic_opacity[4 * i] = conx; s->conic_opacity[4 * i + 1] = cony;
    s->conic_opacity[4 * i + 2] = conz; s->conic_opacity[4 * i + 3] = p->opacities[i];
    s->tiles_touched[i] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
}

/* Stable LSD radix sort of (key,val) pairs, 16-bit digits over `bits` low bits. Our own
 * implementation of the semantics of cub::DeviceRadixSort::SortPairs(begin=0,end=bits)
 * (rasterizer_impl.cu:304-309): stable, ascending, on the selected bit range. */
static void gso_sort_pairs(uint64_t* k, uint32_t* v, size_t n, int bits);

/* forward.cu:261-391 (renderCUDA), one tile */
static void FN(blend_tile)(const FN(GsoParams)* p, FN(GsoState)* s, int tx, int ty, REAL* out_color, REAL* out_depth) {
    const int W = p->W, H = p->H;
    const uint32_t r0 = s->ranges[2 * (ty * s->gx + tx)], r1 = s->ranges[2 * (ty * s->gx + tx) + 1];
    const REAL* feat = p->colors_precomp ? p->colors_precomp : s->rgb;
    for (int ly = 0; ly < TILE; ly++) for (int lx = 0; lx < TILE; lx++) {
        const int x = tx * TILE + lx, y = ty * TILE + ly;
        if (x >= W || y >= H) continue;
        const size_t pix = (size_t)W * y + x;
        REAL T = RC(1), C[3] = {0, 0, 0}, Dd = 0, acc = RC(0.000001);
        uint32_t contributor = 0, last = 0;
        for (uint32_t j = r0; j < r1; j++) {
            contributor++;
            const uint32_t g = s->point_list[j];
            const REAL dx = s->means2D[2 * g] - (REAL)x, dy = s->means2D[2 * g + 1] - (REAL)y;
            const REAL* co = s->conic_opacity + 4 * g;
            const REAL power = RC(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > RC(0)) continue;
            REAL alpha = co[3] * R_EXP(power); if (alpha > RC(0.99)) alpha = RC(0.99);
            if (alpha < RC(1) / RC(255)) continue;
            const REAL test_T = T * (RC(1) - alpha);
            if (test_T < RC(0.0001)) break;              /* pixel done; this entry not accumulated */
            for (int ch = 0; ch < 3; ch++) C[ch] += feat[3 * g + ch] * alpha * T;
            Dd += s->depths[g] * alpha * T;
            acc += alpha * T;
            T = test_T;
            last = contributor;
        }
        s->final_T[pix] = T;
        s->n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pix] = C[ch] + T * p->bg[ch];
        out_depth[pix] = (acc > RC(0.5)) ? Dd / acc : RC(0);
    }
}

static void FN(gso_free_state)(FN(GsoState)* s) {
    if (!s) return;
    free(s->depths); free(s->clamped); free(s->radii); free(s->means2D); free(s->cov3D);
    free(s->conic_opacity); free(s->rgb); free(s->tiles_touched); free(s->point_offsets);
    free(s->keys); free(s->point_list); free(s->ranges); free(s->final_T); free(s->n_contrib);
    free(s);
}

/* rasterizer_impl.cu:198-339 (Rasterizer::forward) + rasterize_points.cu:57-117 (zero-init outputs, P==0) */
FN(GsoState)* FN(gso_forward)(const FN(GsoParams)* p, REAL* out_color, REAL* out_depth, int* radii_out,
                              long long* num_rendered) {
    const int P = p->P, W = p->W, H = p->H;
    const size_t N = (size_t)W * H;
    FN(GsoState)* s = (FN(GsoState)*)calloc(1, sizeof(*s));
    s->P = P; s->W = W; s->H = H;
    s->gx = (W + TILE - 1) / TILE; s->gy = (H + TILE - 1) / TILE;
    const size_t G = (size_t)s->gx * s->gy;
    memset(out_color, 0, 3 * N * sizeof(REAL));
    memset(out_depth, 0, N * sizeof(REAL));
    size_t Pa = P > 0 ? P : 1;
    s->depths = (REAL*)calloc(Pa, sizeof(REAL));
    s->clamped = (unsigned char*)calloc(3 * Pa, 1);
    s->radii = (int*)calloc(Pa, sizeof(int));
    s->means2D = (REAL*)calloc(2 * Pa, sizeof(REAL));
    s->cov3D = (REAL*)calloc(6 * Pa, sizeof(REAL));
    s->conic_opacity = (REAL*)calloc(4 * Pa, sizeof(REAL));
    s->rgb = (REAL*)calloc(3 * Pa, sizeof(REAL));
    s->tiles_touched = (uint32_t*)calloc(Pa, sizeof(uint32_t));
    s->point_offsets = (uint32_t*)calloc(Pa, sizeof(uint32_t));
    s->ranges = (uint32_t*)calloc(2 * G, sizeof(uint32_t));
    s->final_T = (REAL*)calloc(N, sizeof(REAL));
    s->n_contrib = (uint32_t*)calloc(N, sizeof(uint32_t));
    *num_rendered = 0;
    if (P == 0) return s;                               /* zero images, not background */

    const REAL fy = H / (RC(2) * p->tanfovy), fx = W / (RC(2) * p->tanfovx);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) FN(preprocess_one)(p, s, i, fx, fy);

    /* inclusive scan (rasterizer_impl.cu:278) */
    uint32_t run = 0;
    for (int i = 0; i < P; i++) { run += s->tiles_touched[i]; s->point_offsets[i] = run; }
    const size_t R = run;
    s->R = (long long)R; *num_rendered = (long long)R;
    s->keys = (uint64_t*)malloc((R ? R : 1) * sizeof(uint64_t));
    s->point_list = (uint32_t*)malloc((R ? R : 1) * sizeof(uint32_t));

    /* duplicateWithKeys: rasterizer_impl.cu:70-111 */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < P; i++) {
        if (s->radii[i] <= 0) continue;
        size_t off = i == 0 ? 0 : s->point_offsets[i - 1];
        int rmin[2], rmax[2];
        FN(get_rect)(s->means2D[2 * i], s->means2D[2 * i + 1], s->radii[i], s->gx, s->gy, rmin, rmax);
        const float df = (float)s->depths[i];            /* key packs the fp32 bit pattern */
        uint32_t dbits; memcpy(&dbits, &df, 4);
        for (int y = rmin[1]; y < rmax[1]; y++)
            for (int x = rmin[0]; x < rmax[0]; x++) {
                s->keys[off] = ((uint64_t)(uint32_t)(y * s->gx + x) << 32) | dbits;
                s->point_list[off] = (uint32_t)i;
                off++;
            }
    }
    /* getHigherMsb (rasterizer_impl.cu:35-50) only selects enough bits to cover the tile id:
     * sorting on all 64 bits gives the same order. */
    gso_sort_pairs(s->keys, s->point_list, R, 64);

    /* identifyTileRanges: rasterizer_impl.cu:116-138 */
    for (size_t j = 0; j < R; j++) {
        const uint32_t cur = (uint32_t)(s->keys[j] >> 32);
        if (j == 0) s->ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(s->keys[j - 1] >> 32);
            if (cur != prev) { s->ranges[2 * prev + 1] = (uint32_t)j; s->ranges[2 * cur] = (uint32_t)j; }
        }
        if (j == R - 1) s->ranges[2 * cur + 1] = (uint32_t)R;
    }

#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < (int)G; t++) FN(blend_tile)(p, s, t % s->gx, t / s->gx, out_color, out_depth);

    if (radii_out) memcpy(radii_out, s->radii, (size_t)P * sizeof(int));
    return s;
}

/* --------------------------------------------------------------- backward */

static inline void FN(atomic_add)(REAL* a, REAL v) {
#pragma omp atomic
    *a += v;
}

/* backward.cu:399-586 (renderCUDA backward), one tile. dL_dout_depth is ignored (depth
 * gradient commented out in the reference: backward.cu:443-469,539-554). */
static void FN(blend_tile_bwd)(const FN(GsoParams)* p, const FN(GsoState)* s, int tx, int ty, const REAL* dL_dpix,
                               REAL* dL_dmean2D /*3P*/, REAL* dL_dconic /*4P*/, REAL* dL_dopacity, REAL* dL_dcolors) {
    const int W = p->W, H = p->H;
    const uint32_t r0 = s->ranges[2 * (ty * s->gx + tx)], r1 = s->ranges[2 * (ty * s->gx + tx) + 1];
    const REAL* feat = p->colors_precomp ? p->colors_precomp : s->rgb;
    const REAL ddelx_dx = (REAL)(0.5 * W), ddely_dy = (REAL)(0.5 * H);
    for (int ly = 0; ly < TILE; ly++) for (int lx = 0; lx < TILE; lx++) {
        const int x = tx * TILE + lx, y = ty * TILE + ly;
        if (x >= W || y >= H) continue;
        const size_t pix = (size_t)W * y + x;
        const REAL T_final = s->final_T[pix];
        REAL T = T_final;
        const uint32_t last_contributor = s->n_contrib[pix];
        REAL accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
        REAL dL_dpixel[3];
        for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpix[(size_t)ch * H * W + pix];
        uint32_t contributor = r1 - r0;
        for (uint32_t jj = r1; jj > r0; jj--) {
            contributor--;
            if (contributor >= last_contributor) continue;
            const uint32_t g = s->point_list[jj - 1];
            const REAL dx = s->means2D[2 * g] - (REAL)x, dy = s->means2D[2 * g + 1] - (REAL)y;
            const REAL* co = s->conic_opacity + 4 * g;
            const REAL power = RC(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > RC(0)) continue;
            const REAL G = R_EXP(power);
            REAL alpha = co[3] * G; if (alpha > RC(0.99)) alpha = RC(0.99);
            if (alpha < RC(1) / RC(255)) continue;
            T = T / (RC(1) - alpha);
            const REAL dchannel_dcolor = alpha * T;
            REAL dL_dalpha = 0;
            for (int ch = 0; ch < 3; ch++) {
                const REAL c = feat[3 * g + ch];
                accum_rec[ch] = last_alpha * last_color[ch] + (RC(1) - last_alpha) * accum_rec[ch];
                last_color[ch] = c;
                dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                FN(atomic_add)(&dL_dcolors[3 * g + ch], dchannel_dcolor * dL_dpixel[ch]);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            REAL bg_dot = 0;
            for (int ch = 0; ch < 3; ch++) bg_dot += p->bg[ch] * dL_dpixel[ch];
            dL_dalpha += (-T_final / (RC(1) - alpha)) * bg_dot;
            const REAL dL_dG = co[3] * dL_dalpha;       /* clamp at 0.99 passes gradient */
            const REAL gdx = G * dx, gdy = G * dy;
            const REAL dG_ddelx = -gdx * co[0] - gdy * co[1];
            const REAL dG_ddely = -gdy * co[2] - gdx * co[1];
            FN(atomic_add)(&dL_dmean2D[3 * g + 0], dL_dG * dG_ddelx * ddelx_dx);
            FN(atomic_add)(&dL_dmean2D[3 * g + 1], dL_dG * dG_ddely * ddely_dy);
            FN(atomic_add)(&dL_dconic[4 * g + 0], RC(-0.5) * gdx * dx * dL_dG);
            FN(atomic_add)(&dL_dconic[4 * g + 1], RC(-0.5) * gdx * dy * dL_dG);
            FN(atomic_add)(&dL_dconic[4 * g + 3], RC(-0.5) * gdy * dy * dL_dG);
            FN(atomic_add)(&dL_dopacity[g], G * dL_dalpha);
        }
    }
}

/* backward.cu:144-274 (computeCov2DCUDA), one Gaussian. Assigns dL_dmean3D. */
static void FN(cov2d_bwd_one)(const FN(GsoParams)* p, const FN(GsoState)* s, int i, REAL fx, REAL fy,
                              const REAL* dL_dconic, REAL* dL_dmean3D, REAL* dL_dcov) {
    if (!(s->radii[i] > 0)) return;
    const REAL* c6 = p->cov3D_precomp ? p->cov3D_precomp + 6 * i : s->cov3D + 6 * i;
    FN(Cov2DCtx) cc;
    FN(cov2d)(p->means3D + 3 * i, fx, fy, p->tanfovx, p->tanfovy, c6, p->viewmatrix, &cc);
    const REAL dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
    const REAL a = cc.a, b = cc.b, c = cc.c;
    const REAL denom = a * c - b * b;
    REAL dL_da = 0, dL_db = 0, dL_dc = 0;
    const REAL denom2inv = RC(1) / ((denom * denom) + RC(0.0000001));
    REAL (*T)[3] = cc.A;                                /* T[i][j] == (J.V3)[i][j] */
    REAL* o = dL_dcov + 6 * i;
    if (denom2inv != RC(0)) {
        dL_da = denom2inv * (-c * c * dcx + RC(2) * b * c * dcy + (denom - a * c) * dcz);
        dL_dc = denom2inv * (-a * a * dcz + RC(2) * a * b * dcy + (denom - a * c) * dcx);
        dL_db = denom2inv * RC(2) * (b * c * dcx - (denom + RC(2) * b * b) * dcy + a * b * dcz);
        o[0] = (T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc);
        o[3] = (T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc);
        o[5] = (T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc);
        o[1] = RC(2) * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + RC(2) * T[1][0] * T[1][1] * dL_dc;
        o[2] = RC(2) * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + RC(2) * T[1][0] * T[1][2] * dL_dc;
        o[4] = RC(2) * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + RC(2) * T[1][1] * T[1][2] * dL_dc;
    } else {
        for (int k = 0; k < 6; k++) o[k] = 0;
    }
    const REAL S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    REAL dT[2][3];
    for (int k = 0; k < 3; k++) {
        const REAL u0 = T[0][0] * S[k][0] + T[0][1] * S[k][1] + T[0][2] * S[k][2];
        const REAL u1 = T[1][0] * S[k][0] + T[1][1] * S[k][1] + T[1][2] * S[k][2];
        dT[0][k] = RC(2) * u0 * dL_da + u1 * dL_db;
        dT[1][k] = RC(2) * u1 * dL_dc + u0 * dL_db;
    }
    const REAL* vm = p->viewmatrix;
#define V3(r, c_) vm[(r) + 4 * (c_)]
    const REAL dJ00 = V3(0, 0) * dT[0][0] + V3(0, 1) * dT[0][1] + V3(0, 2) * dT[0][2];
    const REAL dJ02 = V3(2, 0) * dT[0][0] + V3(2, 1) * dT[0][1] + V3(2, 2) * dT[0][2];
    const REAL dJ11 = V3(1, 0) * dT[1][0] + V3(1, 1) * dT[1][1] + V3(1, 2) * dT[1][2];
    const REAL dJ12 = V3(2, 0) * dT[1][0] + V3(2, 1) * dT[1][1] + V3(2, 2) * dT[1][2];
    const REAL tz = RC(1) / cc.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const REAL dtx = cc.xmul * -fx * tz2 * dJ02;
    const REAL dty = cc.ymul * -fy * tz2 * dJ12;
    const REAL dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (RC(2) * fx * cc.t[0]) * tz3 * dJ02 + (RC(2) * fy * cc.t[1]) * tz3 * dJ12;
    /* V3^T . (dtx,dty,dtz): auxiliary.h:90-97 */
    dL_dmean3D[3 * i + 0] = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;
    dL_dmean3D[3 * i + 1] = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
    dL_dmean3D[3 * i + 2] = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;
#undef V3
}

/* backward.cu:346-396 (preprocessCUDA bwd) + :20-139 (SH bwd) + :278-341 (cov3D bwd), one Gaussian */
static void FN(preprocess_bwd_one)(const FN(GsoParams)* p, const FN(GsoState)* s, int i, const REAL* dL_dmean2D,
                                   REAL* dL_dmeans, const REAL* dL_dcolor, const REAL* dL_dcov3D, REAL* dL_dsh,
                                   REAL* dL_dscale, REAL* dL_drot) {
    if (!(s->radii[i] > 0)) return;
    const REAL* m = p->means3D + 3 * i;
    const REAL* proj = p->projmatrix;
    REAL mh[4];
    FN(xf4x4)(m, proj, mh);
    const REAL m_w = RC(1) / (mh[3] + RC(0.0000001));
    const REAL mul1 = (proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12]) * m_w * m_w;
    const REAL mul2 = (proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13]) * m_w * m_w;
    const REAL gx = dL_dmean2D[3 * i], gy = dL_dmean2D[3 * i + 1];
    dL_dmeans[3 * i + 0] += (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
    dL_dmeans[3 * i + 1] += (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
    dL_dmeans[3 * i + 2] += (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;

    if (p->shs) {
        const REAL d0[3] = {m[0] - p->campos[0], m[1] - p->campos[1], m[2] - p->campos[2]};
        const REAL len = R_SQRT(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
        const REAL x = d0[0] / len, y = d0[1] / len, z = d0[2] / len;
        const REAL* sh = p->shs + (size_t)i * p->M * 3;
        REAL* dsh = dL_dsh + (size_t)i * p->M * 3;
        REAL dRGB[3];
        for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * i + ch] * (s->clamped[3 * i + ch] ? RC(0) : RC(1));
        REAL b[16];
        FN(sh_basis)(p->D, x, y, z, b);
        const int na = FN(n_active)(p->D);
        for (int k = 0; k < na; k++) for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] = b[k] * dRGB[ch];
        /* d(rgb)/d(dir): backward.cu:58-60,78-80,99-122 */
        REAL dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
        const int D = p->D;
#define SHK(k) (sh[3 * (k) + ch])
        for (int ch = 0; ch < 3; ch++) {
            REAL ax = 0, ay = 0, az = 0;
            if (D > 0) {
                ax = -FN(SH_C1) * SHK(3); ay = -FN(SH_C1) * SHK(1); az = FN(SH_C1) * SHK(2);
                if (D > 1) {
                    const REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    ax += FN(SH_C2)[0] * y * SHK(4) + FN(SH_C2)[2] * RC(2) * -x * SHK(6) + FN(SH_C2)[3] * z * SHK(7) + FN(SH_C2)[4] * RC(2) * x * SHK(8);
                    ay += FN(SH_C2)[0] * x * SHK(4) + FN(SH_C2)[1] * z * SHK(5) + FN(SH_C2)[2] * RC(2) * -y * SHK(6) + FN(SH_C2)[4] * RC(2) * -y * SHK(8);
                    az += FN(SH_C2)[1] * y * SHK(5) + FN(SH_C2)[2] * RC(2) * RC(2) * z * SHK(6) + FN(SH_C2)[3] * x * SHK(7);
                    if (D > 2) {
                        ax += (FN(SH_C3)[0] * SHK(9) * RC(3) * RC(2) * xy + FN(SH_C3)[1] * SHK(10) * yz + FN(SH_C3)[2] * SHK(11) * RC(-2) * xy +
                               FN(SH_C3)[3] * SHK(12) * RC(-3) * RC(2) * xz + FN(SH_C3)[4] * SHK(13) * (RC(-3) * xx + RC(4) * zz - yy) +
                               FN(SH_C3)[5] * SHK(14) * RC(2) * xz + FN(SH_C3)[6] * SHK(15) * RC(3) * (xx - yy));
                        ay += (FN(SH_C3)[0] * SHK(9) * RC(3) * (xx - yy) + FN(SH_C3)[1] * SHK(10) * xz + FN(SH_C3)[2] * SHK(11) * (RC(-3) * yy + RC(4) * zz - xx) +
                               FN(SH_C3)[3] * SHK(12) * RC(-3) * RC(2) * yz + FN(SH_C3)[4] * SHK(13) * RC(-2) * xy +
                               FN(SH_C3)[5] * SHK(14) * RC(-2) * yz + FN(SH_C3)[6] * SHK(15) * RC(-3) * RC(2) * xy);
                        az += (FN(SH_C3)[1] * SHK(10) * xy + FN(SH_C3)[2] * SHK(11) * RC(4) * RC(2) * yz + FN(SH_C3)[3] * SHK(12) * RC(3) * (RC(2) * zz - xx - yy) +
                               FN(SH_C3)[4] * SHK(13) * RC(4) * RC(2) * xz + FN(SH_C3)[5] * SHK(14) * (xx - yy));
                    }
                }
            }
            dx_[ch] = ax; dy_[ch] = ay; dz_[ch] = az;
        }
#undef SHK
        const REAL ddir[3] = {dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2],
                              dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2],
                              dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2]};
        /* dnormvdv: auxiliary.h:107-117 */
        const REAL sum2 = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
        const REAL inv32 = RC(1) / R_SQRT(sum2 * sum2 * sum2);
        dL_dmeans[3 * i + 0] += ((+sum2 - d0[0] * d0[0]) * ddir[0] - d0[1] * d0[0] * ddir[1] - d0[2] * d0[0] * ddir[2]) * inv32;
        dL_dmeans[3 * i + 1] += (-d0[0] * d0[1] * ddir[0] + (sum2 - d0[1] * d0[1]) * ddir[1] - d0[2] * d0[1] * ddir[2]) * inv32;
        dL_dmeans[3 * i + 2] += (-d0[0] * d0[2] * ddir[0] - d0[1] * d0[2] * ddir[1] + (sum2 - d0[2] * d0[2]) * ddir[2]) * inv32;
    }

    if (p->scales) {
        const REAL* q = p->rotations + 4 * i;
        const REAL r = q[0], x = q[1], y = q[2], z = q[3];
        REAL R[3][3], M[3][3], sv[3];
        FN(quat_to_R)(q, R);
        for (int a = 0; a < 3; a++) {
            sv[a] = p->scale_modifier * p->scales[3 * i + a];
            for (int bb = 0; bb < 3; bb++) M[a][bb] = sv[a] * R[bb][a];
        }
        const REAL* g = dL_dcov3D + 6 * i;
        const REAL dS[3][3] = {{g[0], RC(0.5) * g[1], RC(0.5) * g[2]},
                               {RC(0.5) * g[1], g[3], RC(0.5) * g[4]},
                               {RC(0.5) * g[2], RC(0.5) * g[4], g[5]}};
        /* dL_dM = 2 M dSigma */
        REAL dM[3][3];
        for (int a = 0; a < 3; a++)
            for (int bb = 0; bb < 3; bb++)
                dM[a][bb] = RC(2) * (M[a][0] * dS[0][bb] + M[a][1] * dS[1][bb] + M[a][2] * dS[2][bb]);
        for (int a = 0; a < 3; a++)
            dL_dscale[3 * i + a] = R[0][a] * dM[a][0] + R[1][a] * dM[a][1] + R[2][a] * dM[a][2];
        REAL Q[3][3];                                    /* Q[a][b] = s_a dM[a][b] = dL/dR[b][a] */
        for (int a = 0; a < 3; a++) for (int bb = 0; bb < 3; bb++) Q[a][bb] = dM[a][bb] * sv[a];
        REAL* dq = dL_drot + 4 * i;                       /* no normalisation Jacobian: backward.cu:340 */
        dq[0] = RC(2) * z * (Q[0][1] - Q[1][0]) + RC(2) * y * (Q[2][0] - Q[0][2]) + RC(2) * x * (Q[1][2] - Q[2][1]);
        dq[1] = RC(2) * y * (Q[1][0] + Q[0][1]) + RC(2) * z * (Q[2][0] + Q[0][2]) + RC(2) * r * (Q[1][2] - Q[2][1]) - RC(4) * x * (Q[2][2] + Q[1][1]);
        dq[2] = RC(2) * x * (Q[1][0] + Q[0][1]) + RC(2) * r * (Q[2][0] - Q[0][2]) + RC(2) * z * (Q[1][2] + Q[2][1]) - RC(4) * y * (Q[2][2] + Q[0][0]);
        dq[3] = RC(2) * r * (Q[0][1] - Q[1][0]) + RC(2) * x * (Q[2][0] + Q[0][2]) + RC(2) * y * (Q[1][2] + Q[2][1]) - RC(4) * z * (Q[1][1] + Q[0][0]);
    }
}

/* rasterizer_impl.cu:343-444 (Rasterizer::backward) + rasterize_points.cu:144-199 (zero-filled outputs).
 * All outputs are fully overwritten (zero-initialised here). dL_dconic is [P,2,2] scratch. */
void FN(gso_backward)(const FN(GsoParams)* p, const FN(GsoState)* s, const REAL* dL_dpix,
                      REAL* dL_dmeans2D /*3P*/, REAL* dL_dcolors /*3P*/, REAL* dL_dopacity /*P*/,
                      REAL* dL_dmeans3D /*3P*/, REAL* dL_dcov3D /*6P*/, REAL* dL_dsh /*P*M*3*/,
                      REAL* dL_dscales /*3P*/, REAL* dL_drots /*4P*/, REAL* dL_dconic /*4P*/) {
    const int P = p->P;
    const size_t Ps = (size_t)P;
    memset(dL_dmeans2D, 0, 3 * Ps * sizeof(REAL)); memset(dL_dcolors, 0, 3 * Ps * sizeof(REAL));
    memset(dL_dopacity, 0, Ps * sizeof(REAL));     memset(dL_dmeans3D, 0, 3 * Ps * sizeof(REAL));
    memset(dL_dcov3D, 0, 6 * Ps * sizeof(REAL));   memset(dL_dsh, 0, Ps * p->M * 3 * sizeof(REAL));
    memset(dL_dscales, 0, 3 * Ps * sizeof(REAL));  memset(dL_drots, 0, 4 * Ps * sizeof(REAL));
    memset(dL_dconic, 0, 4 * Ps * sizeof(REAL));
    if (P == 0) return;
    const REAL fy = p->H / (RC(2) * p->tanfovy), fx = p->W / (RC(2) * p->tanfovx);
    const int G = s->gx * s->gy;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < G; t++)
        FN(blend_tile_bwd)(p, s, t % s->gx, t / s->gx, dL_dpix, dL_dmeans2D, dL_dconic, dL_dopacity, dL_dcolors);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        FN(cov2d_bwd_one)(p, s, i, fx, fy, dL_dconic, dL_dmeans3D, dL_dcov3D);
        FN(preprocess_bwd_one)(p, s, i, dL_dmeans2D, dL_dmeans3D, dL_dcolors, dL_dcov3D, dL_dsh, dL_dscales, dL_drots);
    }
}

/* rasterizer_impl.cu:54-66,141-153 (markVisible) */
void FN(gso_mark_visible)(int P, const REAL* means3D, const REAL* viewmatrix, unsigned char* present) {
    for (int i = 0; i < P; i++) {
        REAL pv[3];
        FN(xf4x3)(means3D + 3 * i, viewmatrix, pv);
        present[i] = pv[2] > RC(0.2);
    }
}

/* accessors for tests */
long long FN(gso_num_rendered)(const FN(GsoState)* s) { return s->R; }
const uint32_t* FN(gso_point_list)(const FN(GsoState)* s) { return s->point_list; }
const uint64_t* FN(gso_keys)(const FN(GsoState)* s) { return s->keys; }
const uint32_t* FN(gso_ranges)(const FN(GsoState)* s) { return s->ranges; }
const uint32_t* FN(gso_n_contrib)(const FN(GsoState)* s) { return s->n_contrib; }
const uint32_t* FN(gso_tiles_touched)(const FN(GsoState)* s) { return s->tiles_touched; }
const REAL* FN(gso_final_T)(const FN(GsoState)* s) { return s->final_T; }
const REAL* FN(gso_means2D)(const FN(GsoState)* s) { return s->means2D; }
const REAL* FN(gso_conic_opacity)(const FN(GsoState)* s) { return s->conic_opacity; }
const REAL* FN(gso_rgb)(const FN(GsoState)* s) { return s->rgb; }
const REAL* FN(gso_depths)(const FN(GsoState)* s) { return s->depths; }
const REAL* FN(gso_cov3D)(const FN(GsoState)* s) { return s->cov3D; }
const unsigned char* FN(gso_clamped)(const FN(GsoState)* s) { return s->clamped; }
void FN(gso_free)(FN(GsoState)* s) { FN(gso_free_state)(s); }

#undef R_EXP
#undef R_SQRT
#undef R_CEIL
#undef RC
#undef FN
#undef CAT
#undef CAT_
