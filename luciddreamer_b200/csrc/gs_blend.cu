// Tile compositing, forward and backward (sm_100a).
//
// k_blend_fwd replaces RAST/cuda_rasterizer/forward.cu:261-391 (renderCUDA): front-to-back alpha compositing of
//   colour + depth per 16x16 tile, same arithmetic per (pixel, splat).
// k_blend_bwd replaces backward.cu:399-586 (renderCUDA backward), same arithmetic per (pixel, splat).
//
// Both kernels are FP32-issue bound (ncu: >90 % issue-slot utilisation, <2 % DRAM), so the design minimises
// instructions per (pixel, splat) evaluation and skips evaluations that cannot contribute:
//  * one CTA of 128 threads per tile; a warp owns an 8x8 pixel block and every thread TWO pixels (x, y) and
//    (x, y+4): the shared-memory record fetch, the loop bookkeeping, the dx terms of the quadratic form and -- in
//    the backward -- the cross-lane reduction are paid once per pixel pair;
//  * the per-splat record (xy, conic, opacity, rgb, depth) is fetched ONCE per (tile, splat) with 128-bit loads
//    into shared memory; the reference gathers colour and depth from global memory per contributing
//    (pixel, splat) (forward.cu:359,364);
//  * while a batch of 128 splats is staged, the staging thread tests its splat against the four 8x8 blocks of
//    the tile (exact ellipse-vs-box test, gs_box_hit) and the warps ballot the results into a 128-bit mask per
//    block.  A warp whose mask is sparse walks only the set bits; a dense mask falls back to the plain loop.
//    Skipped splats cannot reach alpha >= 1/255 anywhere in the block (every pixel would skip them in the
//    reference too, forward.cu:336-346), so no output changes;
//  * backward: the reference issues 9 global float atomicAdds per contributing (pixel, splat); here the 9 partials
//    of the two pixels are summed in registers, reduced across the warp with a value-halving shuffle butterfly
//    (12 shuffles for 9 values), across the 4 warps in shared memory, and leave the CTA as 128-bit vector
//    reductions: one per (tile, splat).  Warps start the reverse traversal at max(n_contrib) over their pixels.
#include "gs_common.cuh"

namespace {

constexpr int kThreads = 128;         // 4 warps; warp w -> 8x8 pixel block (w & 1, w >> 1)
constexpr int kBatch = 128;           // splats staged per round (one per thread)
constexpr int kWords = kBatch / 32;

struct __align__(16) SRec {           // shared-memory copy of a splat record
    float4 a;                         // x, y, conic_a, conic_b
    float4 b;                         // conic_c, opacity, r, g
    float2 c;                         // b, depth
    uint32_t id;
    uint32_t pad;
};

// 4-bit mask: which of the tile's four 8x8 pixel blocks the splat can touch
__device__ __forceinline__ uint32_t block_mask(const float4 a, const float4 b, const float thr, int tx0, int ty0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const float x0 = (float)(tx0 + 8 * (w & 1)), y0 = (float)(ty0 + 8 * (w >> 1));
        if (gs_box_hit(a.x, a.y, a.z, a.w, b.x, thr, x0, y0, x0 + 7.f, y0 + 7.f)) m |= 1u << w;
    }
    return m;
}

struct FwdPix {
    float T, C0, C1, C2, D, acc;
    uint32_t last;
    bool done;
};

// forward.cu:330-369 for one pixel.  qx = conic_a*dx*dx term etc. are formed exactly like the reference expression
// power = -0.5f * (a*dx*dx + c*dy*dy) - b*dx*dy.
__device__ __forceinline__ void fwd_eval(FwdPix& p, const SRec& r, const float dx, const float dy, const uint32_t pos1) {
    const float power = -0.5f * (r.a.z * dx * dx + r.b.x * dy * dy) - r.a.w * dx * dy;
    if (power > 0.0f) return;
    const float alpha = fminf(0.99f, r.b.y * expf(power));
    if (alpha < 1.0f / 255.0f) return;
    const float test_T = p.T * (1.f - alpha);
    if (test_T < 0.0001f) { p.done = true; return; }
    p.C0 += r.b.z * alpha * p.T;
    p.C1 += r.b.w * alpha * p.T;
    p.C2 += r.c.x * alpha * p.T;
    p.D += r.c.y * alpha * p.T;
    p.acc += alpha * p.T;
    p.T = test_T;
    p.last = pos1;
}

__global__ void __launch_bounds__(kThreads)
k_blend_fwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const GsDevStatus* __restrict__ status, long long capacity,
            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
            float* __restrict__ out_depth) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ SRec sRec[kBatch];
    __shared__ uint32_t sMask[4][kWords];                // [pixel block][32-splat word]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float bg0 = __ldg(v.bg), bg1 = __ldg(v.bg + 1), bg2 = __ldg(v.bg + 2);
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + 8 * (wid & 1) + (lane & 7);
    const uint32_t py0 = ty0 + 8 * (wid >> 1) + (lane >> 3), py1 = py0 + 4;
    const bool in0 = px < (uint32_t)v.W && py0 < (uint32_t)v.H;
    const bool in1 = px < (uint32_t)v.W && py1 < (uint32_t)v.H;
    const float pixx = (float)px, pixy0 = (float)py0, pixy1 = (float)py1;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    const int n = (int)(end - beg);

    FwdPix P0 = {1.0f, 0.f, 0.f, 0.f, 0.f, 0.000001f, 0u, !in0};
    FwdPix P1 = {1.0f, 0.f, 0.f, 0.f, 0.f, 0.000001f, 0u, !in1};

    for (int base = 0; base < n; base += kBatch) {
        if (__syncthreads_and(P0.done && P1.done)) break;
        uint32_t m = 0;
        if (base + tid < n) {
            const uint32_t id = list[beg + base + tid];
            const float4* r = rec + (size_t)GS_REC_V4 * id;
            const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
            SRec s; s.a = a; s.b = b; s.c = make_float2(c.x, c.y); s.id = id; s.pad = 0;
            sRec[tid] = s;
            m = block_mask(a, b, c.w, tx0, ty0);
        }
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t bits = __ballot_sync(0xffffffffu, (m >> w) & 1u);
            if (lane == 0) sMask[w][wid] = bits;
        }
        __syncthreads();
        const int cnt = min(kBatch, n - base);
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            if (bits == 0) continue;
            if (__all_sync(0xffffffffu, P0.done && P1.done)) break;
            const int wcnt = min(32, cnt - 32 * k);
            if (__popc(bits) * 4 >= wcnt * 3) {
                // dense: plain loop over the word, no bit scanning
                for (int jj = 0; jj < wcnt; jj++) {
                    const int j = 32 * k + jj;
                    const SRec& r = sRec[j];
                    const float dx = r.a.x - pixx;
                    if (!P0.done) fwd_eval(P0, r, dx, r.a.y - pixy0, (uint32_t)(base + j + 1));
                    if (!P1.done) fwd_eval(P1, r, dx, r.a.y - pixy1, (uint32_t)(base + j + 1));
                }
            } else {
                while (bits) {
                    const int j = 32 * k + __ffs(bits) - 1;
                    bits &= bits - 1;
                    const SRec& r = sRec[j];
                    const float dx = r.a.x - pixx;
                    if (!P0.done) fwd_eval(P0, r, dx, r.a.y - pixy0, (uint32_t)(base + j + 1));
                    if (!P1.done) fwd_eval(P1, r, dx, r.a.y - pixy1, (uint32_t)(base + j + 1));
                }
            }
        }
    }
    const size_t HW = (size_t)v.H * v.W;
    if (in0) {
        const uint32_t pix_id = (uint32_t)v.W * py0 + px;
        final_T[pix_id] = P0.T;
        n_contrib[pix_id] = P0.last;
        out_color[pix_id] = P0.C0 + P0.T * bg0;
        out_color[HW + pix_id] = P0.C1 + P0.T * bg1;
        out_color[2 * HW + pix_id] = P0.C2 + P0.T * bg2;
        out_depth[pix_id] = (P0.acc > 0.5f) ? P0.D / P0.acc : 0.f;
    }
    if (in1) {
        const uint32_t pix_id = (uint32_t)v.W * py1 + px;
        final_T[pix_id] = P1.T;
        n_contrib[pix_id] = P1.last;
        out_color[pix_id] = P1.C0 + P1.T * bg0;
        out_color[HW + pix_id] = P1.C1 + P1.T * bg1;
        out_color[2 * HW + pix_id] = P1.C2 + P1.T * bg2;
        out_depth[pix_id] = (P1.acc > 0.5f) ? P1.D / P1.acc : 0.f;
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
    float T, T_final, last_alpha;
    float ar0, ar1, ar2;              // accum_rec
    float lc0, lc1, lc2;              // last_color
    float g0, g1, g2, bg_dot;         // dL_dpixel, bg . dL_dpixel
    int last_contributor;
};

// backward.cu:487-584 for one pixel: returns false if this (pixel, splat) does not contribute, else adds its 9
// partial derivatives to vv[]
__device__ __forceinline__ bool bwd_eval(BwdPix& p, const SRec& r, const float dx, const float dy, const int pos,
                                         const float ddelx_dx, const float ddely_dy, float* vv) {
    if (!(pos < p.last_contributor)) return false;
    const float power = -0.5f * (r.a.z * dx * dx + r.b.x * dy * dy) - r.a.w * dx * dy;
    if (power > 0.0f) return false;
    const float G = expf(power);
    const float alpha = fminf(0.99f, r.b.y * G);
    if (alpha < 1.0f / 255.0f) return false;
    p.T = p.T / (1.f - alpha);
    const float dchannel_dcolor = alpha * p.T;
    float dL_dalpha = 0.0f;
    const float c0 = r.b.z, c1 = r.b.w, c2 = r.c.x;
    p.ar0 = p.last_alpha * p.lc0 + (1.f - p.last_alpha) * p.ar0; p.lc0 = c0;
    dL_dalpha += (c0 - p.ar0) * p.g0;
    p.ar1 = p.last_alpha * p.lc1 + (1.f - p.last_alpha) * p.ar1; p.lc1 = c1;
    dL_dalpha += (c1 - p.ar1) * p.g1;
    p.ar2 = p.last_alpha * p.lc2 + (1.f - p.last_alpha) * p.ar2; p.lc2 = c2;
    dL_dalpha += (c2 - p.ar2) * p.g2;
    vv[0] += dchannel_dcolor * p.g0; vv[1] += dchannel_dcolor * p.g1; vv[2] += dchannel_dcolor * p.g2;
    dL_dalpha *= p.T;
    p.last_alpha = alpha;
    dL_dalpha += (-p.T_final / (1.f - alpha)) * p.bg_dot;
    const float dL_dG = r.b.y * dL_dalpha;
    const float gdx = G * dx, gdy = G * dy;
    const float dG_ddelx = -gdx * r.a.z - gdy * r.a.w;
    const float dG_ddely = -gdy * r.b.x - gdx * r.a.w;
    vv[3] += dL_dG * dG_ddelx * ddelx_dx;
    vv[4] += dL_dG * dG_ddely * ddely_dy;
    vv[5] += -0.5f * gdx * dx * dL_dG;
    vv[6] += -0.5f * gdx * dy * dL_dG;
    vv[7] += -0.5f * gdy * dy * dL_dG;
    vv[8] += G * dL_dalpha;
    return true;
}

__global__ void __launch_bounds__(kThreads)
k_blend_bwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const float* __restrict__ final_Ts,
            const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float4* __restrict__ acc) {
    __shared__ SRec sRec[kBatch];
    __shared__ float sAcc[kBatch * 9];
    __shared__ uint32_t sMask[4][kWords];
    __shared__ int sMax[kThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const int tx0 = blockIdx.x * GS_TILE, ty0 = blockIdx.y * GS_TILE;
    const uint32_t px = tx0 + 8 * (wid & 1) + (lane & 7);
    const uint32_t py0 = ty0 + 8 * (wid >> 1) + (lane >> 3), py1 = py0 + 4;
    const bool in0 = px < (uint32_t)v.W && py0 < (uint32_t)v.H;
    const bool in1 = px < (uint32_t)v.W && py1 < (uint32_t)v.H;
    const float pixx = (float)px, pixy0 = (float)py0, pixy1 = (float)py1;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    if (beg == end) return;

    const size_t HW = (size_t)v.H * v.W;
    const float bgc0 = __ldg(v.bg), bgc1 = __ldg(v.bg + 1), bgc2 = __ldg(v.bg + 2);
    BwdPix Q[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const bool in = q ? in1 : in0;
        const uint32_t pix_id = (uint32_t)v.W * (q ? py1 : py0) + px;
        BwdPix& p = Q[q];
        p.T_final = in ? final_Ts[pix_id] : 0.f;
        p.T = p.T_final;
        p.last_contributor = in ? (int)n_contrib[pix_id] : 0;
        p.g0 = in ? dL_dpix[pix_id] : 0.f; p.g1 = in ? dL_dpix[HW + pix_id] : 0.f; p.g2 = in ? dL_dpix[2 * HW + pix_id] : 0.f;
        float bd = 0.f;
        bd += bgc0 * p.g0; bd += bgc1 * p.g1; bd += bgc2 * p.g2;
        p.bg_dot = bd;
        p.last_alpha = 0.f; p.ar0 = p.ar1 = p.ar2 = 0.f; p.lc0 = p.lc1 = p.lc2 = 0.f;
    }
    const float ddelx_dx = 0.5 * v.W, ddely_dy = 0.5 * v.H;

    // max of n_contrib over the warp's block / over the tile: nothing behind it contributes
    int wmax = max(Q[0].last_contributor, Q[1].last_contributor);
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
        const int pos = hi - 1 - tid;            // position in the tile list this thread stages (j = tid)
        uint32_t m = 0;
        if (pos >= 0) {
            const uint32_t id = list[beg + pos];
            const float4* r = rec + (size_t)GS_REC_V4 * id;
            const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
            SRec s; s.a = a; s.b = b; s.c = make_float2(c.x, c.y); s.id = id; s.pad = 0;
            sRec[tid] = s;
            m = block_mask(a, b, c.w, tx0, ty0);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) sAcc[tid * 9 + k] = 0.f;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t bits = __ballot_sync(0xffffffffu, (m >> w) & 1u);
            if (lane == 0) sMask[w][wid] = bits;
        }
        __syncthreads();
        const int cnt = min(kBatch, hi);
#pragma unroll 1
        for (int k = 0; k < kWords; k++) {
            uint32_t bits = sMask[wid][k];
            while (bits) {
                const int j = 32 * k + __ffs(bits) - 1;
                bits &= bits - 1;
                const int p = hi - 1 - j;        // 0-based list position; reverse traversal = increasing j
                if (p >= wmax) continue;         // behind every pixel of this block
                const SRec& r = sRec[j];
                const float dx = r.a.x - pixx;
                float vv[9];
#pragma unroll
                for (int q = 0; q < 9; q++) vv[q] = 0.f;
                const bool c0 = bwd_eval(Q[0], r, dx, r.a.y - pixy0, p, ddelx_dx, ddely_dy, vv);
                const bool c1 = bwd_eval(Q[1], r, dx, r.a.y - pixy1, p, ddelx_dx, ddely_dy, vv);
                if (!__any_sync(0xffffffffu, c0 || c1)) continue;
                warp_reduce9(vv, lane);
                if (owner) atomicAdd(&sAcc[j * 9 + slot], vv[0]);
            }
        }
        __syncthreads();
        if (tid < cnt) {
            float r[9];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 9; k++) { r[k] = sAcc[tid * 9 + k]; any = any || (r[k] != 0.f); }
            if (any) {
                float4* dst = acc + (size_t)3 * sRec[tid].id;
                atomicAdd(dst, make_float4(r[3], r[4], r[5], r[6]));
                atomicAdd(dst + 1, make_float4(r[7], r[8], r[0], r[1]));
                atomicAdd(reinterpret_cast<float*>(dst + 2), r[2]);
            }
        }
    }
}

}  // namespace

void gs_launch_blend_fwd(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const GsDevStatus* status, long long capacity, float* final_T, uint32_t* n_contrib,
                         float* out_color, float* out_depth, cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    k_blend_fwd<<<grid, kThreads, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color,
                                          out_depth);
}
void gs_launch_blend_bwd(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float4* acc,
                         cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    k_blend_bwd<<<grid, kThreads, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc);
}
