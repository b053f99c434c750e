// Per-Gaussian forward pass, tile histogram scan and (tile, Gaussian) pair emission for sm_100a.
//
// Replaces (RAST = reference rasterizer):
//   k_preprocess   RAST/cuda_rasterizer/forward.cu:155-256 (preprocessCUDA) + auxiliary.h:139-164 (in_frustum)
//   k_tile_scan    cub::DeviceScan::InclusiveSum over P Gaussians + blocking D2H of num_rendered
//                  (rasterizer_impl.cu:278-282): here a scan over the G tiles only, result mirrored to pinned
//                  host memory so the host never blocks the stream
//   k_emit         duplicateWithKeys (rasterizer_impl.cu:70-111): pairs go straight into their tile's bucket
//                  (per-tile cursors), keyed by (depth bits, Gaussian index) for the per-tile sort
//   k_mark_visible checkFrustum (rasterizer_impl.cu:54-66)
#include "gs_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBigRect = 32;   // rects with more tiles than this are walked by the whole warp

// Exact tile culling (parity-safe): a (tile, splat) pair is binned only if the splat can
// reach alpha >= 1/255 somewhere in the tile.  power(d) = -0.5 (A dx^2 + C dy^2) - B dx dy is concave with its
// maximum (0) at the mean, so its maximum over the tile's pixel box [x0, x0+15] x [y0, y0+15] is 0 if the mean is
// inside and otherwise lies on an edge facing the mean; each facing edge is a 1-D concave maximisation (clamp
// the stationary point).  The pair is kept iff max power >= thr, thr = -ln(255 * opacity) - slack: pairs that are
// dropped are skipped by every pixel of the tile in the reference too (forward.cu:336-346), so no output changes.
// __noinline__: the histogram pass (k_preprocess) and the emission pass (k_emit) must take bit-identical
// decisions, so both call the same machine code on the same stored floats.
__device__ __noinline__ bool gs_tile_hit(float mx, float my, float A, float B, float C, float thr, int tx, int ty) {
    const float x0 = (float)(tx * GS_TILE), y0 = (float)(ty * GS_TILE);
    return gs_box_hit(mx, my, A, B, C, thr, x0, y0, x0 + (float)(GS_TILE - 1), y0 + (float)(GS_TILE - 1));
}

struct GsCullArgs {
    float mx, my, A, B, C, thr;
};

// Calls f(tile_id, payload...) for every tile of every lane's rect that passes gs_tile_hit.  Small rects: each lane on its own.
// Large rects: the warp takes them one at a time, lanes striding over the tiles (avoids one lane looping over
// hundreds of tiles while 31 wait).  Must be called by all 32 lanes.
template <typename F>
__device__ __forceinline__ void gs_for_each_tile(bool vis, int4 rect, int gx, GsCullArgs c, uint32_t p0, uint32_t p1,
                                                 F f) {
    const int w = rect.z - rect.x, h = rect.w - rect.y;
    const int area = vis ? w * h : 0;
    const bool big = area > kBigRect;
    if (area > 0 && !big) {
        for (int y = rect.y; y < rect.w; y++)
            for (int x = rect.x; x < rect.z; x++)
                if (gs_tile_hit(c.mx, c.my, c.A, c.B, c.C, c.thr, x, y)) f((uint32_t)(y * gx + x), p0, p1);
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    const int lane = threadIdx.x & 31;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int rx = __shfl_sync(0xffffffffu, rect.x, src), ry = __shfl_sync(0xffffffffu, rect.y, src);
        const int rw = __shfl_sync(0xffffffffu, w, src), n = __shfl_sync(0xffffffffu, area, src);
        const uint32_t q0 = __shfl_sync(0xffffffffu, p0, src), q1 = __shfl_sync(0xffffffffu, p1, src);
        GsCullArgs d;
        d.mx = __shfl_sync(0xffffffffu, c.mx, src); d.my = __shfl_sync(0xffffffffu, c.my, src);
        d.A = __shfl_sync(0xffffffffu, c.A, src); d.B = __shfl_sync(0xffffffffu, c.B, src);
        d.C = __shfl_sync(0xffffffffu, c.C, src); d.thr = __shfl_sync(0xffffffffu, c.thr, src);
        for (int k = lane; k < n; k += 32) {
            const int yy = k / rw, xx = k - yy * rw;
            if (gs_tile_hit(d.mx, d.my, d.A, d.B, d.C, d.thr, rx + xx, ry + yy))
                f((uint32_t)((ry + yy) * gx + rx + xx), q0, q1);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
k_preprocess(const GsView v, const float* __restrict__ means3D, const float* __restrict__ shs,
             const float* __restrict__ colors_precomp, const float* __restrict__ opacities,
             const float* __restrict__ scales, const float* __restrict__ rotations,
             const float* __restrict__ cov3D_precomp, int* __restrict__ radii, float4* __restrict__ rec,
             float4* __restrict__ acc, uint32_t* __restrict__ tile_cnt, GsDevStatus* __restrict__ status) {
    __shared__ GsCam cam;
    gs_load_cam(v, &cam);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    bool vis = false;
    int4 rect = make_int4(0, 0, 0, 0);
    int my_radius = 0;
    GsCullArgs cull = {0.f, 0.f, 1.f, 0.f, 1.f, 0.f};

    if (i < v.P) {
        const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        const float3 p_view = gs_xf4x3(p, cam.vm);
        if (p_view.z > GS_NEAR) {                       // only the near plane culls (auxiliary.h:154)
            const float4 p_hom = gs_xf4x4(p, cam.pm);
            const float p_w = 1.0f / (p_hom.w + 0.0000001f);
            const float ppx = p_hom.x * p_w, ppy = p_hom.y * p_w;
            float c6[6];
            if (cov3D_precomp) {
#pragma unroll
                for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * i + k];
            } else {
                const float3 s = make_float3(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]);
                const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
                gs_cov3d(s, v.scale_modifier, q, c6);
            }
            GsCov2D cc;
            gs_cov2d(p, v, cam.vm, c6, cc);
            const float det = cc.a * cc.c - cc.b * cc.b;
            if (det != 0.0f) {
                const float det_inv = 1.f / det;
                const float3 conic = make_float3(cc.c * det_inv, -cc.b * det_inv, cc.a * det_inv);
                const float mid = 0.5f * (cc.a + cc.c);
                const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
                const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
                const float rad = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
                const float px = gs_ndc2pix(ppx, v.W), py = gs_ndc2pix(ppy, v.H);
                rect = gs_rect(px, py, (int)rad, v.gx, v.gy);
                if ((rect.z - rect.x) * (rect.w - rect.y) != 0) {
                    vis = true;
                    my_radius = (int)rad;
                    float r_, g_, b_;
                    uint32_t clamped = 0;
                    if (colors_precomp) {
                        r_ = colors_precomp[3 * i]; g_ = colors_precomp[3 * i + 1]; b_ = colors_precomp[3 * i + 2];
                    } else {                           // forward.cu:20-71
                        float3 d = make_float3(p.x - cam.campos[0], p.y - cam.campos[1], p.z - cam.campos[2]);
                        const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
                        d.x = d.x / len; d.y = d.y / len; d.z = d.z / len;
                        float bs[16];
                        gs_sh_basis(v.D, d.x, d.y, d.z, bs);
                        const float* sh = shs + (size_t)i * v.M * 3;
                        float cr = bs[0] * sh[0], cg = bs[0] * sh[1], cb = bs[0] * sh[2];
                        const int na = (v.D + 1) * (v.D + 1);
                        for (int k = 1; k < na; k++) {
                            cr = cr + bs[k] * sh[3 * k]; cg = cg + bs[k] * sh[3 * k + 1]; cb = cb + bs[k] * sh[3 * k + 2];
                        }
                        cr += 0.5f; cg += 0.5f; cb += 0.5f;
                        clamped = (cr < 0.f ? 1u : 0u) | (cg < 0.f ? 2u : 0u) | (cb < 0.f ? 4u : 0u);
                        r_ = fmaxf(cr, 0.f); g_ = fmaxf(cg, 0.f); b_ = fmaxf(cb, 0.f);
                    }
                    const float opac = opacities[i];
                    // alpha >= 1/255  <=>  power >= -ln(255 * opacity); slack keeps the test conservative
                    const float thr = -logf(255.0f * opac) - GS_CULL_SLACK;
                    cull.mx = px; cull.my = py; cull.A = conic.x; cull.B = conic.y; cull.C = conic.z; cull.thr = thr;
                    float4* rr = rec + (size_t)3 * i;
                    rr[0] = make_float4(px, py, conic.x, conic.y);
                    rr[1] = make_float4(conic.z, opac, r_, g_);
                    rr[2] = make_float4(b_, p_view.z, __uint_as_float(clamped), thr);
                    float4* aa = acc + (size_t)3 * i;
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    aa[0] = z4; aa[1] = z4; aa[2] = z4;
                }
            }
        }
        radii[i] = my_radius;
    }

    // tile histogram (one RED per pair)
    gs_for_each_tile(vis, rect, v.gx, cull, 0u, 0u,
                     [&](uint32_t tile, uint32_t, uint32_t) { atomicAdd(&tile_cnt[tile], 1u); });

    // block-level totals -> two atomics per CTA
    const unsigned area = vis ? (unsigned)((rect.z - rect.x) * (rect.w - rect.y)) : 0u;
    unsigned a = area, c = vis ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    __shared__ unsigned s_a[kThreads / 32], s_c[kThreads / 32];
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_c[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ta = 0, tc = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) { ta += s_a[w]; tc += s_c[w]; }
        if (ta) atomicAdd(&status->num_rendered, ta);
        if (tc) atomicAdd(&status->num_visible, tc);
    }
}

// Exclusive scan of the G tile counts (single CTA: G is ~8k at 1080p).  Resets the counts to zero so the same
// array serves as the emission cursors, and mirrors the totals to a pinned host slot.
__global__ void __launch_bounds__(1024)
k_tile_scan(int G, uint32_t* __restrict__ tile_cnt, uint32_t* __restrict__ tile_off, GsDevStatus* __restrict__ status,
            GsDevStatus* host_slot) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < G; base += 1024) {
        const int t = base + tid;
        const uint32_t c = t < G ? tile_cnt[t] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;                           // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t incl = x + (wid ? s_warp[wid - 1] : 0u);
        if (t < G) { tile_off[t] = carry + incl - c; tile_cnt[t] = 0u; }
        __syncthreads();
        if (tid == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t total = s_carry;
        tile_off[G] = total;
        status->num_pairs = total;
        if (host_slot) {
            host_slot->num_rendered = status->num_rendered;
            host_slot->num_pairs = total;
            host_slot->num_visible = status->num_visible;
            __threadfence_system();
            host_slot->overflow = 0xC0FFEEu;            // "written" marker, checked by gs_forward_counts
        }
    }
}

__global__ void __launch_bounds__(kThreads)
k_emit(const GsView v, const int* __restrict__ radii, const float4* __restrict__ rec,
       const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ tile_cur, GsDevStatus* __restrict__ status,
       unsigned long long* __restrict__ keys, long long capacity) {
    if ((long long)status->num_pairs > capacity) {      // device-side guard: nothing is binned, host re-renders
        if (blockIdx.x == 0 && threadIdx.x == 0) status->overflow = 1u;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) status->n_big = 0u;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    bool vis = false;
    int4 rect = make_int4(0, 0, 0, 0);
    uint32_t dbits = 0;
    GsCullArgs cull = {0.f, 0.f, 1.f, 0.f, 1.f, 0.f};
    if (i < v.P) {
        const int r = radii[i];
        if (r > 0) {
            const float4 q0 = __ldg(rec + (size_t)3 * i);
            const float4 q1 = __ldg(rec + (size_t)3 * i + 1);
            const float4 q2 = __ldg(rec + (size_t)3 * i + 2);
            rect = gs_rect(q0.x, q0.y, r, v.gx, v.gy);  // same recomputation as rasterizer_impl.cu:91
            dbits = __float_as_uint(q2.y);
            cull.mx = q0.x; cull.my = q0.y; cull.A = q0.z; cull.B = q0.w; cull.C = q1.x; cull.thr = q2.w;
            vis = true;
        }
    }
    gs_for_each_tile(vis, rect, v.gx, cull, dbits, (uint32_t)i, [&](uint32_t tile, uint32_t d, uint32_t idx) {
        const uint32_t pos = tile_off[tile] + atomicAdd(&tile_cur[tile], 1u);
        keys[pos] = ((unsigned long long)d << 32) | idx;
    });
}

__global__ void k_mark_visible(int P, const float* __restrict__ means3D, const float* __restrict__ vm,
                               uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    const float z = vm[2] * p.x + vm[6] * p.y + vm[10] * p.z + vm[14];
    present[i] = z > GS_NEAR ? 1 : 0;
}

}  // namespace

void gs_launch_preprocess(const GsView& v, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, const float* rotations,
                          const float* cov3D_precomp, int* radii, float4* rec, float4* acc, uint32_t* tile_cnt,
                          GsDevStatus* status, cudaStream_t s) {
    const int grid = (v.P + kThreads - 1) / kThreads;
    k_preprocess<<<grid, kThreads, 0, s>>>(v, means3D, shs, colors_precomp, opacities, scales, rotations,
                                          cov3D_precomp, radii, rec, acc, tile_cnt, status);
}
void gs_launch_tile_scan(int G, uint32_t* tile_cnt, uint32_t* tile_off, GsDevStatus* status, GsDevStatus* host_slot,
                         cudaStream_t s) {
    k_tile_scan<<<1, 1024, 0, s>>>(G, tile_cnt, tile_off, status, host_slot);
}
void gs_launch_emit(const GsView& v, const int* radii, const float4* rec, const uint32_t* tile_off,
                    uint32_t* tile_cur, GsDevStatus* status, unsigned long long* keys, long long capacity,
                    cudaStream_t s) {
    const int grid = (v.P + kThreads - 1) / kThreads;
    k_emit<<<grid, kThreads, 0, s>>>(v, radii, rec, tile_off, tile_cur, status, keys, capacity);
}
void gs_launch_mark_visible(int P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t s) {
    k_mark_visible<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, vm, present);
}
