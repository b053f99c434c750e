// Tile compositing, forward and backward (sm_100a).
//
// k_blend_fwd replaces RAST/cuda_rasterizer/forward.cu:261-391 (renderCUDA): one CTA per 16x16 tile, one thread
//   per pixel, front-to-back alpha compositing of colour + depth.  Differences in data movement only: the
//   per-splat record (xy, conic, opacity, rgb, depth = 3 x float4) is fetched ONCE per (tile, splat) with
//   128-bit loads into shared memory; the reference gathers colour and depth from global memory per
//   contributing (pixel, splat) (forward.cu:359,364).
// k_blend_bwd replaces backward.cu:399-586 (renderCUDA backward).  The reference issues 9 global float
//   atomicAdds per contributing (pixel, splat); here the 9 partials are reduced across the warp with a
//   value-halving shuffle butterfly (12 shuffles for 9 values), across the 8 warps in shared memory, and
//   leave the CTA as 3 x 128-bit vector reductions per (tile, splat): 256x fewer L2 atomics.
//   Tiles start the reverse traversal at max(n_contrib) over their pixels instead of at the end of the list.
#include "gs_common.cuh"

namespace {

constexpr int kBlock = GS_TILE_PIX;   // 256 threads, thread t -> pixel (t & 15, t >> 4); warp w -> rows 2w, 2w+1

__global__ void __launch_bounds__(kBlock)
k_blend_fwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const GsDevStatus* __restrict__ status, long long capacity,
            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
            float* __restrict__ out_depth) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ float4 sA[kBlock], sB[kBlock], sC[kBlock];
    const int tid = threadIdx.x;
    const float bg0 = __ldg(v.bg), bg1 = __ldg(v.bg + 1), bg2 = __ldg(v.bg + 2);
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const uint32_t px = blockIdx.x * GS_TILE + (tid & 15), py = blockIdx.y * GS_TILE + (tid >> 4);
    const bool inside = px < (uint32_t)v.W && py < (uint32_t)v.H;
    const uint32_t pix_id = (uint32_t)v.W * py + px;
    const float pixx = (float)px, pixy = (float)py;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    int toDo = (int)(end - beg);
    const int rounds = (toDo + kBlock - 1) / kBlock;
    bool done = !inside;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, acc = 0.000001f;
    uint32_t contributor = 0, last_contributor = 0;

    for (int i = 0; i < rounds; i++, toDo -= kBlock) {
        if (__syncthreads_count(done) == kBlock) break;
        const uint32_t progress = (uint32_t)i * kBlock + tid;
        if (beg + progress < end) {
            const uint32_t id = list[beg + progress];
            const float4* r = rec + (size_t)3 * id;
            sA[tid] = __ldg(r); sB[tid] = __ldg(r + 1); sC[tid] = __ldg(r + 2);
        }
        __syncthreads();
        const int lim = min(kBlock, toDo);
        for (int j = 0; !done && j < lim; j++) {
            contributor++;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = a.x - pixx, dy = a.y - pixy;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, b.y * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 c = sC[j];
            C0 += b.z * alpha * T;
            C1 += b.w * alpha * T;
            C2 += c.x * alpha * T;
            D += c.y * alpha * T;
            acc += alpha * T;
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        const size_t HW = (size_t)v.H * v.W;
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        out_color[pix_id] = C0 + T * bg0;
        out_color[HW + pix_id] = C1 + T * bg1;
        out_color[2 * HW + pix_id] = C2 + T * bg2;
        out_depth[pix_id] = (acc > 0.5f) ? D / acc : 0.f;
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

__global__ void __launch_bounds__(kBlock)
k_blend_bwd(const GsView v, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ list,
            const float4* __restrict__ rec, const float* __restrict__ final_Ts,
            const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float4* __restrict__ acc) {
    __shared__ float4 sA[kBlock], sB[kBlock];
    __shared__ float sCol2[kBlock];
    __shared__ uint32_t sId[kBlock];
    __shared__ float sAcc[kBlock * 9];
    __shared__ int sMax[kBlock / 32];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.y * v.gx + blockIdx.x;
    const uint32_t px = blockIdx.x * GS_TILE + (tid & 15), py = blockIdx.y * GS_TILE + (tid >> 4);
    const bool inside = px < (uint32_t)v.W && py < (uint32_t)v.H;
    const uint32_t pix_id = (uint32_t)v.W * py + px;
    const float pixx = (float)px, pixy = (float)py;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    if (beg == end) return;

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    const size_t HW = (size_t)v.H * v.W;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (inside) { g0 = dL_dpix[pix_id]; g1 = dL_dpix[HW + pix_id]; g2 = dL_dpix[2 * HW + pix_id]; }
    float bg_dot = 0.f;
    bg_dot += __ldg(v.bg) * g0; bg_dot += __ldg(v.bg + 1) * g1; bg_dot += __ldg(v.bg + 2) * g2;
    const float ddelx_dx = 0.5 * v.W, ddely_dy = 0.5 * v.H;

    // tile-wide max of n_contrib: nothing behind it contributes to any pixel
    int m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sMax[tid >> 5] = m;
    __syncthreads();
    int maxc = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) maxc = max(maxc, sMax[w]);

    // owner lanes / slots of the halving reduction
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    const int slot = 5 * b4 + 3 * b3 + 2 * b2 + b1;
    const bool owner = ((lane & 1) == 0) && !(b3 && b2) && !(b3 == 0 && b2 && b1) && slot < 9;

    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;      // accum_rec
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;      // last_color
    float last_alpha = 0.f;

    for (int hi = maxc; hi > 0; hi -= kBlock) {
        __syncthreads();
        const int pos = hi - 1 - tid;            // position in the tile list this thread stages
        if (pos >= 0) {
            const uint32_t id = list[beg + pos];
            const float4* r = rec + (size_t)3 * id;
            sA[tid] = __ldg(r); sB[tid] = __ldg(r + 1);
            sCol2[tid] = __ldg(reinterpret_cast<const float*>(r + 2));
            sId[tid] = id;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) sAcc[tid * 9 + k] = 0.f;
        __syncthreads();
        const int cnt = min(kBlock, hi);
        for (int j = 0; j < cnt; j++) {
            const int p = hi - 1 - j;
            bool contrib = inside && p < last_contributor;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = a.x - pixx, dy = a.y - pixy;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            contrib = contrib && !(power > 0.0f);
            const float G = expf(power);
            const float alpha = fminf(0.99f, b.y * G);
            contrib = contrib && !(alpha < 1.0f / 255.0f);
            if (!__any_sync(0xffffffffu, contrib)) continue;
            float vv[9];
#pragma unroll
            for (int k = 0; k < 9; k++) vv[k] = 0.f;
            if (contrib) {
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                const float c0 = b.z, c1 = b.w, c2 = sCol2[j];
                ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0; lc0 = c0;
                dL_dalpha += (c0 - ar0) * g0;
                ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1; lc1 = c1;
                dL_dalpha += (c1 - ar1) * g1;
                ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2; lc2 = c2;
                dL_dalpha += (c2 - ar2) * g2;
                vv[0] = dchannel_dcolor * g0; vv[1] = dchannel_dcolor * g1; vv[2] = dchannel_dcolor * g2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = b.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * a.z - gdy * a.w;
                const float dG_ddely = -gdy * b.x - gdx * a.w;
                vv[3] = dL_dG * dG_ddelx * ddelx_dx;
                vv[4] = dL_dG * dG_ddely * ddely_dy;
                vv[5] = -0.5f * gdx * dx * dL_dG;
                vv[6] = -0.5f * gdx * dy * dL_dG;
                vv[7] = -0.5f * gdy * dy * dL_dG;
                vv[8] = G * dL_dalpha;
            }
            warp_reduce9(vv, lane);
            if (owner) atomicAdd(&sAcc[j * 9 + slot], vv[0]);
        }
        __syncthreads();
        if (tid < cnt) {
            float r[9];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 9; k++) { r[k] = sAcc[tid * 9 + k]; any = any || (r[k] != 0.f); }
            if (any) {
                float4* dst = acc + (size_t)3 * sId[tid];
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
    k_blend_fwd<<<grid, kBlock, 0, s>>>(v, tile_off, list, rec, status, capacity, final_T, n_contrib, out_color,
                                        out_depth);
}
void gs_launch_blend_bwd(const GsView& v, const uint32_t* tile_off, const uint32_t* list, const float4* rec,
                         const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, float4* acc,
                         cudaStream_t s) {
    dim3 grid(v.gx, v.gy);
    k_blend_bwd<<<grid, kBlock, 0, s>>>(v, tile_off, list, rec, final_T, n_contrib, dL_dpix, acc);
}
