// ROUND-1 blend kernels, kept only as the A/B baseline of gs_blend.cu (GS_BLEND_VARIANT=1); not the product path.
// Tile compositing, forward and backward (sm_100a).
//
// k_blend_fwd_r1 replaces RAST/cuda_rasterizer/forward.cu:261-391 (renderCUDA): front-to-back alpha compositing of
//   colour + depth per 16x16 tile, same formulas per (pixel, splat).
// k_blend_bwd_r1 replaces backward.cu:399-586 (renderCUDA backward), same formulas per (pixel, splat).
//
// Both kernels are FP32-issue bound (ncu: ~90 % issue-slot utilisation, <2 % DRAM), so the design minimises
// instructions per (pixel, splat) evaluation and skips evaluations that cannot contribute:
//  * one CTA of 64 threads per tile; a warp owns a 16x8 pixel block and every thread FOUR pixels of one column
//    (x, y), (x, y+2), (x, y+4), (x, y+6): the shared-memory record fetch, the loop bookkeeping, the dx terms of
//    the quadratic form and -- in the backward -- the cross-lane reduction are paid once per four pixels;
//  * the per-splat record (xy, conic, opacity, rgb, depth) is fetched ONCE per (tile, splat) with 128-bit loads
//    into shared memory; the reference gathers colour and depth from global memory per contributing
//    (pixel, splat) (forward.cu:359,364);
//  * while a batch of 128 splats is staged, the staging thread tests its splats against the two 16x8 blocks of
//    the tile (exact ellipse-vs-box test, gs_box_hit) and the warps ballot the results into a 128-bit mask per
//    block.  A warp whose mask is sparse walks only the set bits; a dense mask falls back to the plain loop.
//    Skipped splats cannot reach alpha >= 1/255 anywhere in the block (every pixel would skip them in the
//    reference too, forward.cu:336-346), so no output changes;
//  * exp(): ex2.approx of power*log2(e) (2 instructions instead of the 10 of expf).  Its ~5e-7 relative error is
//    far inside the 1e-4 colour tolerance, but alpha is compared against 1/255; inside a band of 1e-7 around
//    that threshold alpha is re-evaluated with expf so the skip decision is the reference's;
//  * backward: the reference issues 9 global float atomicAdds per contributing (pixel, splat); here the 9 partials
//    of the four pixels are summed in registers, reduced across the warp with a value-halving shuffle butterfly
//    (12 shuffles for 9 values), across the 2 warps in shared memory, and leave the CTA as 128-bit vector
//    reductions: one per (tile, splat).  Warps start the reverse traversal at max(n_contrib) over their pixels.
#include <cstdlib>

#include "gs_common.cuh"

namespace {

constexpr int kThreads = 64;          // 2 warps; warp w -> pixel rows [8w, 8w+8) of the tile, 16 wide
constexpr int kPix = 4;               // pixels per thread: (x, y0 + 2q), q = 0..3
constexpr int kBatch = 128;           // splats staged per round (two per thread)
constexpr int kWords = kBatch / 32;

struct __align__(16) SRec {           // shared-memory copy of a splat record
    float4 a;                         // x, y, conic_a, conic_b
    float4 b;                         // conic_c, opacity, r, g
    float2 c;                         // b, depth
    uint32_t id;
    uint32_t pad;
};

constexpr float kHalfLog2e = -0.5f * 1.4426950408889634f;     // staged conic scale (see stage_batch)
constexpr float kUnscale = -2.0f * 0.6931471805599453f;       // back to the conic for the flush of the backward pass

// 2-bit mask: which of the tile's two 16x8 pixel blocks the splat can touch
__device__ __forceinline__ uint32_t block_mask(const float4 a, const float4 b, const float thr, int tx0, int ty0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 2; w++) {
        const float x0 = (float)tx0, y0 = (float)(ty0 + 8 * w);
        if (gs_box_hit(a.x, a.y, a.z, a.w, b.x, thr, x0, y0, x0 + 15.f, y0 + 7.f)) m |= 1u << w;
    }
    return m;
}

// Stages up to two splats per thread into shared memory and publishes the per-block hit masks.
// slot j of the batch holds list position pos(j); returns nothing, fills sRec / sMask.  All threads must call.
template <typename PosFn>
__device__ __forceinline__ void stage_batch(SRec* sRec, uint32_t (*sMask)[kWords], const uint32_t* __restrict__ list,
                                            const float4* __restrict__ rec, uint32_t beg, int cnt, int tx0, int ty0,
                                            PosFn pos_of) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int h = 0; h < kBatch / kThreads; h++) {
        const int j = tid + h * kThreads;
        uint32_t m = 0;
        if (j < cnt) {
            const uint32_t id = list[beg + pos_of(j)];
            const float4* r = rec + (size_t)GS_REC_V4 * id;
            const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
            // the conic is staged pre-multiplied for the exponent in base 2:
            //   log2(e) * power = A2 dx dx + C2 dy dy + B2 dx dy,  A2 = -0.5 log2e A, C2 = -0.5 log2e C, B2 = -log2e B
            SRec s; s.a = make_float4(a.x, a.y, a.z * kHalfLog2e, a.w * (2.f * kHalfLog2e));
            s.b = make_float4(b.x * kHalfLog2e, b.y, b.z, b.w); s.c = make_float2(c.x, c.y); s.id = id; s.pad = 0;
            sRec[j] = s;
            m = block_mask(a, b, c.w, tx0, ty0);
        }
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const uint32_t bits = __ballot_sync(0xffffffffu, (m >> w) & 1u);
            if (lane == 0) sMask[w][wid + 2 * h] = bits;      // word index = j / 32
        }
    }
}

struct FwdPix {
    float T;                          // > 0: live transmittance; <= 0: pixel terminated (|T| is the value to report)
    float C0, C1, C2, D, acc;         //      or outside the image
    uint32_t last;
};

// forward.cu:330-369 for the thread's four pixels, written branch-free (predicated updates) so that the four
// independent dependency chains interleave; the arithmetic of every taken update is the reference's.
__device__ __forceinline__ void fwd_eval4(FwdPix* P, const SRec& r, const float dx, const float* dy, const uint32_t pos1,
                                          const float4* __restrict__ rec) {
    float power[kPix], alpha[kPix];       // power = log2(e) * the reference's power (same sign)
    bool band = false;
    const float hA = r.a.z * dx * dx, hB = r.a.w * dx;
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        power[q] = fmaf(hB, dy[q], fmaf(r.b.x * dy[q], dy[q], hA));
        float g;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(power[q]));
        alpha[q] = r.b.y * g;
        band = band || (fabsf(alpha[q] - 1.0f / 255.0f) < 1e-7f);
    }
    if (band) {                       // rare: decide the 1/255 test with the reference's arithmetic (forward.cu:336-343)
        const float4 a = __ldg(rec + (size_t)GS_REC_V4 * r.id), b = __ldg(rec + (size_t)GS_REC_V4 * r.id + 1);
#pragma unroll
        for (int q = 0; q < kPix; q++)
            alpha[q] = r.b.y * expf(-0.5f * (a.z * dx * dx + b.x * dy[q] * dy[q]) - a.w * dx * dy[q]);
    }
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        FwdPix& p = P[q];
        const float al = fminf(0.99f, alpha[q]);
        const bool ok = !(power[q] > 0.0f) && !(al < 1.0f / 255.0f);
        const float test_T = p.T * (1.f - al);
        const bool term = ok && (test_T < 0.0001f);       // forward.cu:348-352 (also true for T <= 0)
        const bool upd = ok && !term;
        if (upd) {
            const float w = al * p.T;                     // one weight for colour, depth and coverage
            p.C0 += r.b.z * w;
            p.C1 += r.b.w * w;
            p.C2 += r.c.x * w;
            p.D += r.c.y * w;
            p.acc += w;
            p.last = pos1;
        }
        p.T = upd ? test_T : (term ? -fabsf(p.T) : p.T);
    }
}

__global__ void __launch_bounds__(kThreads, 16)
k_blend_fwd_r1(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const GsDevStatus* __restrict__ status, long long capacity,
            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
            float* __restrict__ out_depth) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ SRec sRec[kBatch];
    __shared__ uint32_t sMask[2][kWords];                // [pixel block][32-splat word]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float bg0 = __ldg(v.bg), bg1 = __ldg(v.bg + 1), bg2 = __ldg(v.bg + 2);
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + (lane & 15);
    const uint32_t pyb = ty0 + 8 * wid + (lane >> 4);    // rows pyb + 2q
    const float pixx = (float)px;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    const int n = (int)(end - beg);

    FwdPix P[kPix];
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        const bool in = px < (uint32_t)v.W && (pyb + 2 * q) < (uint32_t)v.H;
        P[q].T = in ? 1.0f : -1.0f;
        P[q].C0 = P[q].C1 = P[q].C2 = P[q].D = 0.f; P[q].acc = 0.000001f; P[q].last = 0u;
    }

    for (int base = 0; base < n; base += kBatch) {
        bool alldone = true;
#pragma unroll
        for (int q = 0; q < kPix; q++) alldone = alldone && (P[q].T <= 0.f);
        if (__syncthreads_and(alldone)) break;
        const int cnt = min(kBatch, n - base);
        stage_batch(sRec, sMask, list, rec, beg, cnt, tx0, ty0, [&](int j) { return base + j; });
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            if (bits == 0) continue;
            alldone = true;
#pragma unroll
            for (int q = 0; q < kPix; q++) alldone = alldone && (P[q].T <= 0.f);
            if (__all_sync(0xffffffffu, alldone)) break;
            const int wcnt = min(32, cnt - 32 * k);
            // dense mask: visit every splat of the word (the per-pixel tests skip the misses anyway)
            if (__popc(bits) * 4 >= wcnt * 3) bits = wcnt >= 32 ? 0xffffffffu : ((1u << wcnt) - 1u);
            while (bits) {
                const int j = 32 * k + __ffs(bits) - 1;
                bits &= bits - 1;
                const SRec& r = sRec[j];
                const float dx = r.a.x - pixx;
                float dy[kPix];
#pragma unroll
                for (int q = 0; q < kPix; q++) dy[q] = r.a.y - (float)(pyb + 2 * q);
                fwd_eval4(P, r, dx, dy, (uint32_t)(base + j + 1), rec);
            }
        }
    }
    const size_t HW = (size_t)v.H * v.W;
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        const uint32_t py = pyb + 2 * q;
        if (px < (uint32_t)v.W && py < (uint32_t)v.H) {
            const uint32_t pix_id = (uint32_t)v.W * py + px;
            const float T = fabsf(P[q].T);
            final_T[pix_id] = T;
            n_contrib[pix_id] = P[q].last;
            out_color[pix_id] = P[q].C0 + T * bg0;
            out_color[HW + pix_id] = P[q].C1 + T * bg1;
            out_color[2 * HW + pix_id] = P[q].C2 + T * bg2;
            out_depth[pix_id] = (P[q].acc > 0.5f) ? P[q].D / P[q].acc : 0.f;
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

struct BwdPix {
    float T;
    float tb;                         // -T_final * (bg . dL_dpixel)
    float AR;                         // sum_ch accum_rec_ch * dL_dpixel_ch, already advanced past the last contributing splat
    float g0, g1, g2;                 // dL_dpixel
    int last_contributor;
};

// backward.cu:487-584 for the thread's four pixels, branch-free: every pixel evaluates the full expression and a
// predicate zeroes what a skipped (pixel, splat) would add.  Returns whether any of the four contributed.
// 1/(1-alpha) is formed once (reciprocal) for both quotients: the gradient tolerance is 1e-3 relative, the
// difference to two IEEE divisions is ~1e-7.
__device__ __forceinline__ bool bwd_eval4(BwdPix* Q, const SRec& r, const float dx, const float* dy, const int pos,
                                          float* vv) {
    float power[kPix], G[kPix], alpha[kPix];
    float s0 = 0.f, sy = 0.f, syy = 0.f;
    // no exact-exp band here: a borderline alpha ~ 1/255 decided differently from the forward changes one pixel's
    // reconstructed transmittance by 0.4 %, far below the gradient tolerance, and saves 3 instructions per pixel
    const float hA = r.a.z * dx * dx, hB = r.a.w * dx;
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        power[q] = fmaf(hB, dy[q], fmaf(r.b.x * dy[q], dy[q], hA));       // log2(e) * power (staged conic)
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(G[q]) : "f"(power[q]));
        alpha[q] = r.b.y * G[q];
    }
    bool any = false;
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        BwdPix& p = Q[q];
        const float al = fminf(0.99f, alpha[q]);
        const bool ok = (pos < p.last_contributor) && !(power[q] > 0.0f) && !(al < 1.0f / 255.0f);
        any = any || ok;
        const float a_ = ok ? al : 0.f;                     // alpha = 0 makes every update below a no-op
        const float inv = __frcp_rn(1.f - a_);
        const float T = p.T * inv;
        const float dchannel_dcolor = a_ * T;
        // backward.cu:520-528 keeps accum_rec per channel (updated lazily with last_alpha * last_color + (1 - last_alpha)
        // * accum_rec at the NEXT contributing splat) and forms sum_ch (c_ch - accum_rec_ch) * dL_dpixel_ch.  dL_dpixel is
        // constant per pixel, so only the scalar AR = sum_ch accum_rec_ch * dL_dpixel_ch is carried: it obeys the same
        // recurrence (AR' = a cg + (1 - a) AR = AR + a (cg - AR), cg = c . dL_dpixel) and is advanced eagerly.
        const float cg = r.b.z * p.g0 + r.b.w * p.g1 + r.c.x * p.g2;
        const float dcol = cg - p.AR;
        p.AR = fmaf(a_, dcol, p.AR);
        p.T = T;
        vv[0] += dchannel_dcolor * p.g0; vv[1] += dchannel_dcolor * p.g1; vv[2] += dchannel_dcolor * p.g2;
        const float dL_dalpha = fmaf(dcol, T, p.tb * inv);  // finite also for a skipped splat (a_ = 0, inv = 1)
        const float Gq = ok ? G[q] : 0.f;                   // zero weight; keeps an inf of a skipped splat out of the sums
        // backward.cu:563-583 needs, per (pixel, splat), k = o G dL_dalpha times {A dx + B dy, C dy + B dx, dx dx, dx dy,
        // dy dy} and G dL_dalpha itself.  Only the raw moments of g = G dL_dalpha are summed here -- dx is the same for
        // the thread's four pixels, so three sums (g, g dy, g dy dy) per pixel and three products per thread suffice;
        // the splat's constants (o, A, B, C, -0.5 W, -0.5 H, -0.5) are applied once per (tile, splat) at the flush.
        const float gda = Gq * dL_dalpha;
        s0 += gda;
        const float t = gda * dy[q];
        sy += t;
        syy = fmaf(t, dy[q], syy);
    }
    vv[3] = s0;  vv[4] = dx * s0;  vv[5] = sy;
    vv[6] = dx * vv[4];  vv[7] = dx * sy;  vv[8] = syy;
    return any;
}

__global__ void __launch_bounds__(kThreads, 10)
k_blend_bwd_r1(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const float* __restrict__ final_Ts,
            const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float4* __restrict__ acc) {
    __shared__ SRec sRec[kBatch];
    __shared__ float sAcc[kBatch * 9];
    __shared__ uint32_t sMask[2][kWords];
    __shared__ int sMax[kThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + (lane & 15);
    const uint32_t pyb = ty0 + 8 * wid + (lane >> 4);
    const float pixx = (float)px;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    if (beg == end) return;

    const size_t HW = (size_t)v.H * v.W;
    const float bgc0 = __ldg(v.bg), bgc1 = __ldg(v.bg + 1), bgc2 = __ldg(v.bg + 2);
    BwdPix Q[kPix];
    int wmax = 0;
#pragma unroll
    for (int q = 0; q < kPix; q++) {
        const uint32_t py = pyb + 2 * q;
        const bool in = px < (uint32_t)v.W && py < (uint32_t)v.H;
        const uint32_t pix_id = (uint32_t)v.W * py + px;
        BwdPix& p = Q[q];
        const float T_final = in ? final_Ts[pix_id] : 0.f;
        p.T = T_final;
        p.last_contributor = in ? (int)n_contrib[pix_id] : 0;
        p.g0 = in ? dL_dpix[pix_id] : 0.f; p.g1 = in ? dL_dpix[HW + pix_id] : 0.f; p.g2 = in ? dL_dpix[2 * HW + pix_id] : 0.f;
        float bd = 0.f;
        bd += bgc0 * p.g0; bd += bgc1 * p.g1; bd += bgc2 * p.g2;
        p.tb = -T_final * bd;
        p.AR = 0.f;
        wmax = max(wmax, p.last_contributor);
    }
    const float ddelx_dx = 0.5 * v.W, ddely_dy = 0.5 * v.H;

    // max of n_contrib over the warp's block / over the tile: nothing behind it contributes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) sMax[wid] = wmax;
    __syncthreads();
    int maxc = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) maxc = max(maxc, sMax[w]);

    // owner lanes / slots of the halving reduction
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    const int slot = 5 * b4 + 3 * b3 + 2 * b2 + b1;
    const bool owner = ((lane & 1) == 0) && !(b3 && b2) && !(b3 == 0 && b2 && b1) && slot < 9;

    for (int hi = maxc; hi > 0; hi -= kBatch) {
        __syncthreads();
        const int cnt = min(kBatch, hi);
        // slot j holds list position hi-1-j: reverse traversal = increasing j
        stage_batch(sRec, sMask, list, rec, beg, cnt, tx0, ty0, [&](int j) { return hi - 1 - j; });
#pragma unroll
        for (int h = 0; h < kBatch / kThreads; h++)
#pragma unroll
            for (int k = 0; k < 9; k++) sAcc[(tid + h * kThreads) * 9 + k] = 0.f;
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            while (bits) {
                const int j = 32 * k + __ffs(bits) - 1;
                bits &= bits - 1;
                const int p = hi - 1 - j;        // 0-based list position
                if (p >= wmax) continue;         // behind every pixel of this block
                const SRec& r = sRec[j];
                const float dx = r.a.x - pixx;
                float vv[9];
#pragma unroll
                for (int q = 0; q < 9; q++) vv[q] = 0.f;
                float dy[kPix];
#pragma unroll
                for (int q = 0; q < kPix; q++) dy[q] = r.a.y - (float)(pyb + 2 * q);
                const bool any = bwd_eval4(Q, r, dx, dy, p, vv);
                if (!__any_sync(0xffffffffu, any)) continue;
                warp_reduce9(vv, lane);
                if (owner) atomicAdd(&sAcc[j * 9 + slot], vv[0]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < kBatch / kThreads; h++) {
            const int j = tid + h * kThreads;
            if (j < cnt) {
                float r[9];
                bool any = false;
#pragma unroll
                for (int k = 0; k < 9; k++) { r[k] = sAcc[j * 9 + k]; any = any || (r[k] != 0.f); }
                if (any) {
                    float4* dst = acc + (size_t)3 * sRec[j].id;
                    // r[3..8] = sums of g, g dx, g dy, g dx dx, g dx dy, g dy dy (g = G dL_dalpha) over the tile
                    const SRec& sr = sRec[j];
                    const float A = sr.a.z * kUnscale, B = sr.a.w * (0.5f * kUnscale), C = sr.b.x * kUnscale, o = sr.b.y;
                    const float kx = o * r[4], ky = o * r[5], h = -0.5f * o;
                    atomicAdd(dst, make_float4(-(A * kx + B * ky) * ddelx_dx, -(C * ky + B * kx) * ddely_dy, h * r[6], h * r[7]));
                    atomicAdd(dst + 1, make_float4(h * r[8], r[3], r[0], r[1]));
                    atomicAdd(reinterpret_cast<float*>(dst + 2), r[2]);
                }
            }
        }
    }
}

}  // namespace

void gs_launch_blend_fwd_r1(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const GsDevStatus* status, long long capacity, float* final_T, uint32_t* n_contrib,
                         float* out_color, float* out_depth, cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    k_blend_fwd_r1<<<grid, kThreads, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color,
                                          out_depth);
}
void gs_launch_blend_bwd_r1(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float4* acc,
                         cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    k_blend_bwd_r1<<<grid, kThreads, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc);
}
