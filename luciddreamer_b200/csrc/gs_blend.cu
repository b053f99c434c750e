// Tile compositing, forward and backward (sm_100a), round-2 kernels.
//
// k_blend_fwd replaces RAST/cuda_rasterizer/forward.cu:261-391 (renderCUDA): front-to-back alpha compositing of
//   colour + depth per 16x16 tile, same formulas per (pixel, splat).
// k_blend_bwd replaces backward.cu:399-586 (renderCUDA backward), same formulas per (pixel, splat).
//
// Both kernels are instruction-issue bound (ncu round 1: 83 % / 73 % issue utilisation, < 2 % DRAM), so the
// design minimises issued instructions per (pixel, splat) evaluation:
//  * one CTA of 64 threads per tile; a warp owns a 16x8 pixel block and every thread FOUR pixels of one column
//    (x, y), (x, y+2), (x, y+4), (x, y+6), held as two float2 PAIRS: all per-pixel arithmetic is issued as
//    sm_100 packed FP32 (fma.rn.f32x2 / mul.f32x2 / add.f32x2 -> FFMA2 / FMUL2 / FADD2, one issue slot for
//    two pixels; per-splat scalars ride along as broadcast operands);
//  * the per-splat record (xy, conic, opacity, rgb, depth) is fetched ONCE per (tile, splat) with 128-bit loads
//    into shared memory; the staging thread also tests its splat against the tile's two 16x8 blocks (exact
//    ellipse-vs-box test) and the warps ballot the results into a bit mask per block, walked with one FLO per hit;
//  * per-splat FAST-PATH eligibility is decided once at staging (warp-uniform): opacity <= 0.99 (the
//    min(0.99, alpha) clamp of forward.cu:343 can never bind) and a well-conditioned positive-definite conic
//    (power <= 0 holds with a 10x margin over fp32 rounding, so forward.cu:336's `power > 0` test can never
//    fire).  Eligible splats skip both per-pixel tests; everything else takes the generic path, which is the
//    reference's test sequence verbatim;
//  * exp(): ex2.approx of the log2-scaled power (conic pre-multiplied at staging).  alpha is compared against
//    1/255; inside a band of 1e-7 around that threshold a thread re-evaluates its four alphas with expf and
//    the reference's expression (generic path), so the skip decision is the reference's.  Forward and
//    backward run the SAME instruction sequence up to that decision, hence skip exactly the same
//    (pixel, splat) pairs;
//  * backward: the reference issues 9 global float atomicAdds per contributing (pixel, splat); here the 9 partials
//    of the four pixels are summed in registers, reduced across the warp with a value-halving shuffle butterfly
//    (12 shuffles for 9 values), written to a per-warp shared-memory slot with a plain store (each warp visits a
//    staged splat at most once), and leave the CTA as 128-bit vector reductions: one per (tile, splat).
//    1/(1-alpha) is a single MUFU.RCP (the gradient tolerance is 1e-3 relative).
#include <cstdlib>

#include "gs_common.cuh"

namespace {

constexpr int kBatch = 128;           // splats staged per round
constexpr int kWords = kBatch / 32;

// Two geometries, chosen per launch by the number of tiles (gs_launch_blend_*):
//   NH = 2  a warp owns a 16x8 block, a thread FOUR pixels (two float2 pairs); 2 warps (64 threads) per tile
//   NH = 1  a warp owns a 16x4 strip, a thread TWO pixels (one pair); 4 warps (128 threads) per tile -- twice the warps
//           for images with few tiles (512x512: 1024 tiles would leave the SMs at 14 resident warps), finer culling
template <int NH> struct Geo {
    static constexpr int RW = 4 * NH;             // pixel rows per warp
    static constexpr int NW = 16 / RW;            // warps per tile
    static constexpr int NT = 32 * NW;            // threads per CTA
    static constexpr int NQ = 2 * NH;             // pixels per thread
};

struct __align__(16) SRec {           // shared-memory copy of a splat record
    float4 a;                         // x, y, A2 = -0.5 log2e conic_a, B2 = -log2e conic_b
    float4 b;                         // C2 = -0.5 log2e conic_c, opacity, r, g
    float2 c;                         // b, depth
    uint32_t id;
    uint32_t generic;                 // != 0: not eligible for the fast path
};

constexpr float kHalfLog2e = -0.5f * 1.4426950408889634f;     // staged conic scale (see stage_batch)
constexpr float kUnscale = -2.0f * 0.6931471805599453f;       // back to the conic for the flush of the backward pass
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kBand = 1e-7f;
constexpr float kNegInf = -__builtin_huge_valf();

// ---- packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2); scalar broadcasts fold into the operand (Rx.F32)
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 bc(const float s) { return make_float2(s, s); }
__device__ __forceinline__ float ex2_approx(const float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(const float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// NW-bit mask: which of the tile's 16 x RW pixel strips the splat can touch
template <int NH>
__device__ __forceinline__ uint32_t strip_mask(const float4 a, const float4 b, const float thr, int tx0, int ty0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < Geo<NH>::NW; w++) {
        const float x0 = (float)tx0, y0 = (float)(ty0 + Geo<NH>::RW * w);
        if (gs_box_hit(a.x, a.y, a.z, a.w, b.x, thr, x0, y0, x0 + 15.f, y0 + (float)(Geo<NH>::RW - 1))) m |= 1u << w;
    }
    return m;
}

// Stages up to kBatch splats into shared memory and publishes the per-strip hit masks.
// slot j of the batch holds list position pos(j).  Mask words are bit-REVERSED (slot 32k+i -> bit 31-i) so the
// traversal finds the next slot with one FLO (count-leading-zeros).  All threads must call.
template <int NH, typename PosFn>
__device__ __forceinline__ void stage_batch(SRec* sRec, uint32_t (*sMask)[kWords], const uint32_t* __restrict__ list,
                                            const float4* __restrict__ rec, uint32_t beg, int cnt, int tx0, int ty0,
                                            PosFn pos_of) {
    constexpr int NW = Geo<NH>::NW, NT = Geo<NH>::NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int h = 0; h < kBatch / NT; h++) {
        const int j = tid + h * NT;
        uint32_t m = 0;
        if (j < cnt) {
            const uint32_t id = list[beg + pos_of(j)];
            const float4* r = rec + (size_t)GS_REC_V4 * id;
            const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
            // the conic is staged pre-multiplied for the exponent in base 2:
            //   log2(e) * power = A2 dx dx + C2 dy dy + B2 dx dy,  A2 = -0.5 log2e A, C2 = -0.5 log2e C, B2 = -log2e B
            SRec s; s.a = make_float4(a.x, a.y, a.z * kHalfLog2e, a.w * (2.f * kHalfLog2e));
            s.b = make_float4(b.x * kHalfLog2e, b.y, b.z, b.w); s.c = make_float2(c.x, c.y); s.id = id;
            // fast-path eligibility (NaNs answer "generic"): alpha = opacity * G can exceed neither 0.99 nor, with
            // det >= 1e-5 trace^2 (smallest eigenvalue >= 1e-5 trace: ~100x the fp32 rounding of the three products),
            // can the computed power be positive
            const float A = a.z, B = a.w, Cc = b.x, tr = A + Cc;
            const bool fast = (b.y <= 0.99f) && (A > 0.f) && (Cc > 0.f) && (A * Cc - B * B > 1e-5f * tr * tr) &&
                              (fabsf(a.x) < 1e7f) && (fabsf(a.y) < 1e7f) && (tr < 1e7f);
            s.generic = fast ? 0u : 1u;
            sRec[j] = s;
            m = strip_mask<NH>(a, b, c.w, tx0, ty0);
        }
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t bits = __brev(__ballot_sync(0xffffffffu, (m >> w) & 1u));
            if (lane == 0) sMask[w][wid + NW * h] = bits;     // word index = j / 32
        }
    }
}

#define GS_Q(v, q) (((q) & 1) ? (v)[(q) >> 1].y : (v)[(q) >> 1].x)

// The approximate evaluation shared by forward and backward: p = log2(e) * power, g = 2^p, a = opacity * g,
// d = a + nam (nam = -1/255 for a live pixel: sign-exact alpha >= 1/255 test; -inf for a dead one).  Returns true
// when one of the alphas sits inside the re-evaluation band around 1/255.
template <int NH>
__device__ __forceinline__ bool eval_alpha(const SRec& r, const float dx, const float2* dy, const float2* nam, float2* p,
                                           float2* g, float2* a, float2* d) {
    const float hA = (r.a.z * dx) * dx, hB = r.a.w * dx;
#pragma unroll
    for (int h = 0; h < NH; h++) {
        const float2 u = fma2(bc(r.b.x), dy[h], bc(hB));          // C2 dy + B2 dx
        p[h] = fma2(u, dy[h], bc(hA));                            // (C2 dy + B2 dx) dy + A2 dx dx
        g[h] = make_float2(ex2_approx(p[h].x), ex2_approx(p[h].y));
        a[h] = mul2(bc(r.b.y), g[h]);
        d[h] = add2(a[h], nam[h]);
    }
    float m = fminf(fabsf(d[0].x), fabsf(d[0].y));
    if (NH == 2) m = fminf(m, fminf(fabsf(d[NH - 1].x), fabsf(d[NH - 1].y)));
    return m < kBand;
}

// power and alpha (before the 0.99 clamp) of the thread's pixels with the reference's arithmetic (forward.cu:330-343),
// from the unscaled record
template <int NH>
__device__ __forceinline__ void exact_eval(const SRec& r, const float dx, const float2* dy,
                                           const float4* __restrict__ rec, float* power, float* alpha) {
    const float4 ra = __ldg(rec + (size_t)GS_REC_V4 * r.id), rb = __ldg(rec + (size_t)GS_REC_V4 * r.id + 1);
#pragma unroll
    for (int q = 0; q < 2 * NH; q++) {
        const float dyq = GS_Q(dy, q);
        power[q] = -0.5f * (ra.z * dx * dx + rb.x * dyq * dyq) - ra.w * dx * dyq;
        alpha[q] = r.b.y * expf(power[q]);
    }
}

// the thread's pixels as packed pairs: pair h holds image rows pyb + 4h and pyb + 4h + 2
template <int NH> struct FwdPix {
    float2 T[NH];                     // transmittance in front of the next splat; frozen once the pixel is dead
    float2 nam[NH];                   // -1/255 while the pixel is live; -inf once it terminated (T < 1e-4) or if it
                                      // lies outside the image: a + nam < 0 then fails the alpha test for every splat
    float2 C0[NH], C1[NH], C2[NH], D[NH], acc[NH];
    uint32_t last[2 * NH];
};

// forward.cu:330-369, generic path: the reference's full test sequence per pixel (power > 0, min(0.99, alpha),
// alpha < 1/255, T < 1e-4).  Taken by splats that are not fast-path eligible, by threads with an alpha in the band and
// whenever a pixel of the warp is about to terminate.
template <int NH>
__device__ __forceinline__ void fwd_generic(FwdPix<NH>& P, const SRec& r, const float dx, const float2* dy, const float2* p,
                                            const float2* a, const bool band, const uint32_t pos1,
                                            const float4* __restrict__ rec) {
    // in the band around 1/255, and for every splat that is not fast-path eligible (ill-conditioned conics: the
    // factored exponent rounds differently from the reference's expression, and the difference is amplified by the
    // cancellation inside `power`), the decision operands are re-evaluated in the reference's own operation order
    float alpha[2 * NH], power[2 * NH];
#pragma unroll
    for (int q = 0; q < 2 * NH; q++) { alpha[q] = GS_Q(a, q); power[q] = GS_Q(p, q); }
    if (band || r.generic) exact_eval<NH>(r, dx, dy, rec, power, alpha);
#pragma unroll
    for (int q = 0; q < 2 * NH; q++) {
        float& T = GS_Q(P.T, q);
        const bool live = GS_Q(P.nam, q) > -1.f;
        const float al = fminf(0.99f, alpha[q]);
        const bool ok = live && !(power[q] > 0.0f) && !(al < kAlphaMin);
        const float test_T = T * (1.f - al);
        const bool low = test_T < 0.0001f;                // forward.cu:348-352
        if (ok && !low) {
            const float w = al * T;                       // one weight for colour, depth and coverage
            GS_Q(P.C0, q) += r.b.z * w;
            GS_Q(P.C1, q) += r.b.w * w;
            GS_Q(P.C2, q) += r.c.x * w;
            GS_Q(P.D, q) += r.c.y * w;
            GS_Q(P.acc, q) += w;
            P.last[q] = pos1;
            T = test_T;
        }
        if (ok && low) GS_Q(P.nam, q) = kNegInf;          // done = true: T stays, nothing behind contributes
    }
}

template <int NH, int OCC>
__global__ void __launch_bounds__(Geo<NH>::NT, OCC)
k_blend_fwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const GsDevStatus* __restrict__ status, long long capacity,
            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
            float* __restrict__ out_depth) {
    constexpr int NW = Geo<NH>::NW, NQ = Geo<NH>::NQ;
    if ((long long)status->num_pairs > capacity) return;
    __shared__ SRec sRec[kBatch];
    __shared__ uint32_t sMask[NW][kWords];               // [pixel strip][32-splat word], bit-reversed
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float bg0 = __ldg(v.bg), bg1 = __ldg(v.bg + 1), bg2 = __ldg(v.bg + 2);
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + (lane & 15);
    const uint32_t pyb = ty0 + Geo<NH>::RW * wid + (lane >> 4);    // rows pyb + 2q
    const float pixx = (float)px;
    float2 npy[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) npy[h] = make_float2(-(float)(pyb + 4 * h), -(float)(pyb + 4 * h + 2));
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    const int n = (int)(end - beg);

    FwdPix<NH> P;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const bool in = px < (uint32_t)v.W && (pyb + 2 * q) < (uint32_t)v.H;
        GS_Q(P.T, q) = 1.0f;
        GS_Q(P.nam, q) = in ? -kAlphaMin : kNegInf;
        GS_Q(P.C0, q) = 0.f; GS_Q(P.C1, q) = 0.f; GS_Q(P.C2, q) = 0.f; GS_Q(P.D, q) = 0.f;
        GS_Q(P.acc, q) = 0.000001f; P.last[q] = 0u;
    }
    auto all_dead = [&]() {
        float m = fmaxf(P.nam[0].x, P.nam[0].y);
        if (NH == 2) m = fmaxf(m, fmaxf(P.nam[NH - 1].x, P.nam[NH - 1].y));
        return m < -1.f;
    };

    for (int base = 0; base < n; base += kBatch) {
        if (__syncthreads_and(all_dead())) break;
        const int cnt = min(kBatch, n - base);
        stage_batch<NH>(sRec, sMask, list, rec, beg, cnt, tx0, ty0, [&](int j) { return base + j; });
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            if (bits == 0) continue;
            if (__all_sync(0xffffffffu, all_dead())) break;
            while (bits) {
                const int lz = __clz(bits);
                bits &= ~(0x80000000u >> lz);
                const int j = 32 * k + lz;
                const SRec& r = sRec[j];
                const uint32_t pos1 = (uint32_t)(base + j + 1);
                const float dx = r.a.x - pixx;
                float2 dy[NH], p[NH], g[NH], a[NH], d[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) dy[h] = add2(bc(r.a.y), npy[h]);
                const bool band = eval_alpha<NH>(r, dx, dy, P.nam, p, g, a, d);
                // fast path: alpha = a (no clamp can bind), contributes <=> a >= 1/255 on a live pixel <=> d >= 0
                // (forward.cu:336-346).  A pixel that does not contribute runs with alpha = 0: test_T = T exactly.
                float2 ae[NH], tT[NH];
                bool okq[NQ];
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    okq[q] = GS_Q(d, q) >= 0.f;
                    GS_Q(ae, q) = okq[q] ? GS_Q(a, q) : 0.f;
                }
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const float2 om = fma2(ae[h], bc(-1.f), bc(1.f));     // 1 - alpha
                    tT[h] = mul2(P.T[h], om);                             // test_T = T * (1 - alpha)
                }
                float tmin = fminf(tT[0].x, tT[0].y);
                if (NH == 2) tmin = fminf(tmin, fminf(tT[NH - 1].x, tT[NH - 1].y));
                // a pixel about to terminate (forward.cu:348-352; dead pixels keep T >= 1e-4 and never trigger), an
                // alpha in the band or an ineligible splat: this thread takes the reference's sequence for its pixels
                if (band || tmin < 0.0001f || r.generic) {
                    fwd_generic<NH>(P, r, dx, dy, p, a, band, pos1, rec);
                    continue;
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) P.last[q] = okq[q] ? pos1 : P.last[q];
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const float2 w = mul2(ae[h], P.T[h]);                 // alpha * T
                    P.T[h] = tT[h];
                    P.C0[h] = fma2(bc(r.b.z), w, P.C0[h]);
                    P.C1[h] = fma2(bc(r.b.w), w, P.C1[h]);
                    P.C2[h] = fma2(bc(r.c.x), w, P.C2[h]);
                    P.D[h] = fma2(bc(r.c.y), w, P.D[h]);
                    P.acc[h] = add2(P.acc[h], w);
                }
            }
        }
    }
    const size_t HW = (size_t)v.H * v.W;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const uint32_t py = pyb + 2 * q;
        if (px < (uint32_t)v.W && py < (uint32_t)v.H) {
            const uint32_t pix_id = (uint32_t)v.W * py + px;
            const float T = GS_Q(P.T, q);
            final_T[pix_id] = T;
            n_contrib[pix_id] = P.last[q];
            out_color[pix_id] = GS_Q(P.C0, q) + T * bg0;
            out_color[HW + pix_id] = GS_Q(P.C1, q) + T * bg1;
            out_color[2 * HW + pix_id] = GS_Q(P.C2, q) + T * bg2;
            out_depth[pix_id] = (GS_Q(P.acc, q) > 0.5f) ? GS_Q(P.D, q) / GS_Q(P.acc, q) : 0.f;
        }
    }
}

// Reduce 9 per-lane values across the warp with value halving: after step s every lane keeps only the half of
// the values selected by one bit of its lane id, so the 5 steps cost 5+3+2+1+1 = 12 shuffles (a plain
// butterfly needs 45).  On return even lanes whose (bit3,bit2,bit1) pattern is 000,001,010,100,101 hold in
// v[0] the full warp sum of value slot 5*bit4 + 3*bit3 + 2*bit2 + bit1; other lanes hold zeros/partials.
template <int N, int OFF>
__device__ __forceinline__ void halving_step(float* v, const int lane) {
    constexpr int Hh = (N + 1) / 2;
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int k = 0; k < Hh; k++) {
        const float lo = v[k];
        const float hi = (k + Hh < N) ? v[k + Hh] : 0.f;
        const float send = up ? lo : hi;
        const float recv = __shfl_xor_sync(0xffffffffu, send, OFF);
        v[k] = (up ? hi : lo) + recv;
    }
}
__device__ __forceinline__ void warp_reduce9(float* v, const int lane) {
    halving_step<9, 16>(v, lane);
    halving_step<5, 8>(v, lane);
    halving_step<3, 4>(v, lane);
    halving_step<2, 2>(v, lane);
    halving_step<1, 1>(v, lane);
}

template <int NH> struct BwdPix {
    float2 T[NH];
    float2 tb[NH];                    // -T_final * (bg . dL_dpixel)
    float2 AR[NH];                    // sum_ch accum_rec_ch * dL_dpixel_ch, already advanced past the last contributing splat
    float2 g0[NH], g1[NH], g2[NH];    // dL_dpixel
    int last[2 * NH];                 // last contributor (1-based list position)
};

// What both backward paths sum per (thread, splat), for w = opacity * G * dL_dalpha of the contributing pixels:
//   vv[0..2] = sum alpha T dL_dpixel_ch          (dL_dcolor)
//   vv[3..8] = sum w, dx w, sum w dy, dx dx w, dx sum w dy, sum w dy dy
// backward.cu:563-583 needs G dL_dalpha times {1, dx, dy, dx dx, dx dy, dy dy} and the splat's constants (A, B, C,
// -0.5 W, -0.5 H, -0.5, 1/opacity), which are applied once per (tile, splat) at the flush.
//
// backward.cu:520-528 keeps accum_rec per channel and forms sum_ch (c_ch - accum_rec_ch) * dL_dpixel_ch.  dL_dpixel is
// constant per pixel, so only the scalar AR = sum_ch accum_rec_ch * dL_dpixel_ch is carried: it obeys the same
// recurrence (AR' = a cg + (1 - a) AR = AR + a (cg - AR), cg = c . dL_dpixel) and is advanced eagerly.

// generic path of the backward: the reference's test sequence per pixel, exact alpha inside the band
template <int NH>
__device__ __forceinline__ bool bwd_generic(BwdPix<NH>& Q, const SRec& r, const float dx, const float2* dy, const float2* p,
                                            const float2* g, const float2* a, const bool band, const int pos,
                                            const float4* __restrict__ rec, float* vv) {
    (void)g;
    float alpha[2 * NH], power[2 * NH];          // alpha = opacity * G, unclamped
#pragma unroll
    for (int q = 0; q < 2 * NH; q++) { alpha[q] = GS_Q(a, q); power[q] = GS_Q(p, q); }
    if (band || r.generic) exact_eval<NH>(r, dx, dy, rec, power, alpha);
    bool any = false;
    float s0 = 0.f, sy = 0.f, syy = 0.f;
#pragma unroll
    for (int q = 0; q < 2 * NH; q++) {
        const float al = fminf(0.99f, alpha[q]);
        const bool ok = (pos < Q.last[q]) && !(power[q] > 0.0f) && !(al < kAlphaMin);
        any = any || ok;
        const float a_ = ok ? al : 0.f;                     // alpha = 0 makes every update below a no-op
        const float inv = rcp_approx(1.f - a_);
        float& T = GS_Q(Q.T, q);
        T = T * inv;
        const float g0 = GS_Q(Q.g0, q), g1 = GS_Q(Q.g1, q), g2 = GS_Q(Q.g2, q);
        const float cg = fmaf(r.c.x, g2, fmaf(r.b.w, g1, r.b.z * g0));
        float& AR = GS_Q(Q.AR, q);
        const float dcol = cg - AR;
        AR = fmaf(a_, dcol, AR);
        const float dch = a_ * T;
        vv[0] = fmaf(dch, g0, vv[0]); vv[1] = fmaf(dch, g1, vv[1]); vv[2] = fmaf(dch, g2, vv[2]);
        const float dL_dalpha = fmaf(dcol, T, GS_Q(Q.tb, q) * inv);
        const float oG = ok ? alpha[q] : 0.f;               // opacity * G; zero keeps an inf of a skipped splat out of the sums
        const float wq = oG * dL_dalpha;
        const float dyq = GS_Q(dy, q);
        s0 += wq;
        const float t = wq * dyq;
        sy += t;
        syy = fmaf(t, dyq, syy);
    }
    vv[3] = s0;  vv[4] = dx * s0;  vv[5] = sy;
    vv[6] = dx * vv[4];  vv[7] = dx * sy;  vv[8] = syy;
    return any;
}

template <int NH, int OCC>
__global__ void __launch_bounds__(Geo<NH>::NT, OCC)
k_blend_bwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const float* __restrict__ final_Ts,
            const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float4* __restrict__ acc,
            const GsDevStatus* __restrict__ status, long long capacity) {
    constexpr int NW = Geo<NH>::NW, NT = Geo<NH>::NT, NQ = Geo<NH>::NQ;
    // a forward whose pair count exceeded the binning capacity rendered nothing (device-side guard): there is no list
    // to walk, the accumulators stay zero (static-capacity mode reports the overflow at the next host call)
    if ((long long)status->num_pairs > capacity) return;
    __shared__ SRec sRec[kBatch];
    __shared__ float sAcc[NW][kBatch * 9];               // per warp: plain stores, no atomics
    __shared__ uint32_t sMask[NW][kWords];
    __shared__ uint32_t sDone[NW][kWords];               // which slots of sAcc a warp has written this round
    __shared__ int sMax[NW];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + (lane & 15);
    const uint32_t pyb = ty0 + Geo<NH>::RW * wid + (lane >> 4);
    const float pixx = (float)px;
    float2 npy[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) npy[h] = make_float2(-(float)(pyb + 4 * h), -(float)(pyb + 4 * h + 2));
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    if (beg == end) return;

    const size_t HW = (size_t)v.H * v.W;
    const float bgc0 = __ldg(v.bg), bgc1 = __ldg(v.bg + 1), bgc2 = __ldg(v.bg + 2);
    BwdPix<NH> Q;
    int wmax = 0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const uint32_t py = pyb + 2 * q;
        const bool in = px < (uint32_t)v.W && py < (uint32_t)v.H;
        const uint32_t pix_id = (uint32_t)v.W * py + px;
        const float T_final = in ? final_Ts[pix_id] : 0.f;
        GS_Q(Q.T, q) = T_final;
        Q.last[q] = in ? (int)n_contrib[pix_id] : 0;
        const float g0 = in ? dL_dpix[pix_id] : 0.f, g1 = in ? dL_dpix[HW + pix_id] : 0.f, g2 = in ? dL_dpix[2 * HW + pix_id] : 0.f;
        GS_Q(Q.g0, q) = g0; GS_Q(Q.g1, q) = g1; GS_Q(Q.g2, q) = g2;
        float bd = 0.f;
        bd += bgc0 * g0; bd += bgc1 * g1; bd += bgc2 * g2;
        GS_Q(Q.tb, q) = -T_final * bd;
        GS_Q(Q.AR, q) = 0.f;
        wmax = max(wmax, Q.last[q]);
    }
    const float ddelx_dx = 0.5 * v.W, ddely_dy = 0.5 * v.H;

    // max of n_contrib over the warp's strip / over the tile: nothing behind it contributes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) sMax[wid] = wmax;
    __syncthreads();
    int maxc = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) maxc = max(maxc, sMax[w]);

    // owner lanes / slots of the halving reduction
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    const int slot = 5 * b4 + 3 * b3 + 2 * b2 + b1;
    const bool owner = ((lane & 1) == 0) && !(b3 && b2) && !(b3 == 0 && b2 && b1) && slot < 9;
    float* const myAcc = sAcc[wid] + (owner ? slot : 0);

    for (int hi = maxc; hi > 0; hi -= kBatch) {
        __syncthreads();
        const int cnt = min(kBatch, hi);
        // slot j holds list position hi-1-j: reverse traversal = increasing j
        stage_batch<NH>(sRec, sMask, list, rec, beg, cnt, tx0, ty0, [&](int j) { return hi - 1 - j; });
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            uint32_t done = 0;
            while (bits) {
                const int lz = __clz(bits);
                const uint32_t bit = 0x80000000u >> lz;
                bits &= ~bit;
                const int j = 32 * k + lz;
                const int pos = hi - 1 - j;      // 0-based list position
                if (pos >= wmax) continue;       // behind every pixel of this strip
                const SRec& r = sRec[j];
                const float dx = r.a.x - pixx;
                float2 dy[NH], p[NH], g[NH], a[NH], d[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) dy[h] = add2(bc(r.a.y), npy[h]);
                float2 nam[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) nam[h] = bc(-kAlphaMin);
                const bool band = eval_alpha<NH>(r, dx, dy, nam, p, g, a, d);
                float vv[9];
                bool any;
                if (band || r.generic) {
#pragma unroll
                    for (int q = 0; q < 9; q++) vv[q] = 0.f;
                    any = bwd_generic<NH>(Q, r, dx, dy, p, g, a, band, pos, rec, vv);
                } else {
                    // fast path: alpha = a (no clamp), contributes <=> pos < last && d >= 0; a skipped pixel runs with
                    // alpha = 0, which makes every update a no-op (T / 1 = T, AR + 0, zero weights)
                    float2 ae[NH];
#pragma unroll
                    for (int q = 0; q < NQ; q++) GS_Q(ae, q) = (pos < Q.last[q] && GS_Q(d, q) >= 0.f) ? GS_Q(a, q) : 0.f;
                    uint32_t anyb = __float_as_uint(ae[0].x) | __float_as_uint(ae[0].y);
                    if (NH == 2) anyb |= __float_as_uint(ae[NH - 1].x) | __float_as_uint(ae[NH - 1].y);
                    any = anyb != 0u;
                    float2 c0 = bc(0.f), c1 = bc(0.f), c2 = bc(0.f), s0 = bc(0.f), sy = bc(0.f), syy = bc(0.f);
#pragma unroll
                    for (int h = 0; h < NH; h++) {
                        const float2 om = fma2(ae[h], bc(-1.f), bc(1.f));                     // 1 - alpha
                        const float2 inv = make_float2(rcp_approx(om.x), rcp_approx(om.y));
                        Q.T[h] = mul2(Q.T[h], inv);                                           // T of the splats in front
                        const float2 cg = fma2(bc(r.c.x), Q.g2[h], fma2(bc(r.b.w), Q.g1[h], mul2(bc(r.b.z), Q.g0[h])));
                        const float2 dcol = fma2(Q.AR[h], bc(-1.f), cg);                      // cg - AR
                        Q.AR[h] = fma2(ae[h], dcol, Q.AR[h]);
                        const float2 dch = mul2(ae[h], Q.T[h]);                               // dchannel_dcolor
                        c0 = fma2(dch, Q.g0[h], c0); c1 = fma2(dch, Q.g1[h], c1); c2 = fma2(dch, Q.g2[h], c2);
                        const float2 dLda = fma2(dcol, Q.T[h], mul2(Q.tb[h], inv));           // dL_dalpha
                        const float2 wq = mul2(ae[h], dLda);                                  // opacity G dL_dalpha
                        s0 = add2(s0, wq);
                        const float2 t = mul2(wq, dy[h]);
                        sy = add2(sy, t);
                        syy = fma2(t, dy[h], syy);
                    }
                    vv[0] = c0.x + c0.y; vv[1] = c1.x + c1.y; vv[2] = c2.x + c2.y;
                    const float S0 = s0.x + s0.y, SY = sy.x + sy.y;
                    vv[3] = S0; vv[4] = dx * S0; vv[5] = SY;
                    vv[6] = dx * vv[4]; vv[7] = dx * SY; vv[8] = syy.x + syy.y;
                }
                if (!__any_sync(0xffffffffu, any)) continue;
                warp_reduce9(vv, lane);
                if (owner) myAcc[j * 9] = vv[0];
                done |= bit;
            }
            if (lane == 0) sDone[wid][k] = done;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < kBatch / NT; h++) {
            const int j = tid + h * NT;
            if (j < cnt) {
                const uint32_t bit = 0x80000000u >> (j & 31);
                bool dn[NW];
                bool anyd = false;
#pragma unroll
                for (int w = 0; w < NW; w++) { dn[w] = (sDone[w][j >> 5] & bit) != 0u; anyd = anyd || dn[w]; }
                if (anyd) {
                    float r[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) {
                        float x = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; w++) x += dn[w] ? sAcc[w][j * 9 + k] : 0.f;
                        r[k] = x;
                    }
                    float4* dst = acc + (size_t)3 * sRec[j].id;
                    // r[3..8] = sums of w, w dx, w dy, w dx dx, w dx dy, w dy dy (w = opacity G dL_dalpha) over the tile
                    const SRec& sr = sRec[j];
                    const float A = sr.a.z * kUnscale, B = sr.a.w * (0.5f * kUnscale), C = sr.b.x * kUnscale;
                    const float kx = r[4], ky = r[5], hf = -0.5f;
                    atomicAdd(dst, make_float4(-(A * kx + B * ky) * ddelx_dx, -(C * ky + B * kx) * ddely_dy, hf * r[6], hf * r[7]));
                    atomicAdd(dst + 1, make_float4(hf * r[8], r[3] * rcp_approx(sr.b.y), r[0], r[1]));
                    atomicAdd(reinterpret_cast<float*>(dst + 2), r[2]);
                }
            }
        }
    }
}

}  // namespace

int g_gs_blend_variant = 0;           // GS_BLEND_VARIANT=2 / 3 force the 16x8 / 16x4 warp geometry (A/B measurements)
int g_gs_num_sms = 148;

// 16x4 strips (4 warps per tile) when 16x8 blocks would leave the SMs with fewer than ~20 resident warps
static bool use_strips(const GsView& v) {
    if (g_gs_blend_variant == 2) return false;
    if (g_gs_blend_variant == 3) return true;
    return (long long)v.gx * v.gy * 2 < (long long)g_gs_num_sms * 20;
}

void gs_launch_blend_fwd(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const GsDevStatus* status, long long capacity, float* final_T, uint32_t* n_contrib,
                         float* out_color, float* out_depth, cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    // CTAs per SM the register allocation aims at: 14 (72 registers, 28 warps per SM) measured 3 % faster than 12 (80
    // registers) on config 3; GS_BLEND_OCC_FWD=12 keeps the other build for A/B runs
    static const int occ = getenv("GS_BLEND_OCC_FWD") ? atoi(getenv("GS_BLEND_OCC_FWD")) : 14;
    if (use_strips(v))
        k_blend_fwd<1, 6><<<grid, Geo<1>::NT, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color, out_depth);
    else if (occ == 12)
        k_blend_fwd<2, 12><<<grid, Geo<2>::NT, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color, out_depth);
    else
        k_blend_fwd<2, 14><<<grid, Geo<2>::NT, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color, out_depth);
}
void gs_launch_blend_bwd(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float4* acc,
                         const GsDevStatus* status, long long capacity, cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    // 13 CTAs per SM (72 registers, no spills; shared memory then limits at 13): 3 % faster than 10 (87 registers)
    static const int occ = getenv("GS_BLEND_OCC_BWD") ? atoi(getenv("GS_BLEND_OCC_BWD")) : 13;
    if (use_strips(v))
        k_blend_bwd<1, 7><<<grid, Geo<1>::NT, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc, status, capacity);
    else if (occ == 10)
        k_blend_bwd<2, 10><<<grid, Geo<2>::NT, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc, status, capacity);
    else
        k_blend_bwd<2, 13><<<grid, Geo<2>::NT, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc, status, capacity);
}
