// Fused optimiser step for the Gaussian parameters (sm_100a) -- the "next" row SURVEY.md 8f-2.
//
// Replaces, for one training iteration of the reference (luciddreamer.py:304,325-327):
//   * the autograd backward through the parameter activations of scene/gaussian_model.py:97-117
//     (sigmoid for opacity, exp for scaling, F.normalize for rotation, cat for the SH features) and
//   * torch.optim.Adam(eps=1e-15) over the six parameter groups (gaussian_model.py:156-165)
// i.e. ~30 elementwise / foreach kernels and their temporaries, by ONE launch that reads the rasterizer's flat
// gradient bucket (gradients w.r.t. the ACTIVATED values, exactly what gs_backward writes), applies the activation
// chain rule in registers and performs the Adam update in place.  HBM-bound by construction: 4 (grad) + 8 (param rw)
// + 16 (two moments rw) bytes per parameter, every access coalesced.
#include "gs_common.cuh"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, const float g, const float step_size,
                                            const GsAdamArgs& a) {
    // same operation order as torch.optim.Adam (lerp, mul+addcmul, sqrt/bias2 + eps, addcdiv); the scalar factors
    // (1-beta, bias corrections, lr/bias1) are formed in double on the host like torch forms them in Python
    m = m + (g - m) * a.om_beta1;
    v = v * a.beta2 + a.om_beta2 * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kT)
k_gaussian_adam(const GsAdamArgs a) {
    // which segment does this block belong to?
    int seg = 0;
    int b = blockIdx.x;
    while (seg < a.nseg - 1 && b >= a.seg[seg].blocks) { b -= a.seg[seg].blocks; seg++; }
    const GsAdamSeg s = a.seg[seg];
    const long long base = (long long)b * kT + threadIdx.x;
    if (s.type == 3) {                                   // rotation: one thread per quaternion (F.normalize chain rule)
        const long long i = base;
        if (i >= s.rows) return;
        const float4 q = reinterpret_cast<const float4*>(s.p)[i];
        const float* gp = s.g + i * s.g_row_w + s.g_off;
        const float gx = gp[0], gy = gp[1], gz = gp[2], gw = gp[3];
        const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);   // F.normalize eps
        const float inv = 1.f / n;
        const float hx = q.x * inv, hy = q.y * inv, hz = q.z * inv, hw = q.w * inv;
        const float dot = hx * gx + hy * gy + hz * gz + hw * gw;
        float gr[4] = {(gx - hx * dot) * inv, (gy - hy * dot) * inv, (gz - hz * dot) * inv, (gw - hw * dot) * inv};
        float pv[4] = {q.x, q.y, q.z, q.w};
        float4 m4 = reinterpret_cast<float4*>(s.m)[i], v4 = reinterpret_cast<float4*>(s.v)[i];
        float mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; k++) adam_update(pv[k], mv[k], vv[k], gr[k], s.step_size, a);
        reinterpret_cast<float4*>(s.p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        reinterpret_cast<float4*>(s.m)[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        reinterpret_cast<float4*>(s.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        return;
    }
    // identity / sigmoid / exp groups: 4 consecutive parameters per thread, 128-bit accesses on param and moments
    const long long n = s.rows * s.row_w;
    const long long e0 = base * 4;
    if (e0 >= n) return;
    const int cnt = (int)min(4LL, n - e0);
    float p[4], m[4], v[4], g[4];
    if (cnt == 4) {
        const float4 p4 = reinterpret_cast<const float4*>(s.p)[base], m4 = reinterpret_cast<const float4*>(s.m)[base],
                     v4 = reinterpret_cast<const float4*>(s.v)[base];
        p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
        m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
        v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in = k < cnt;
            p[k] = in ? s.p[e0 + k] : 0.f; m[k] = in ? s.m[e0 + k] : 0.f; v[k] = in ? s.v[e0 + k] : 0.f;
        }
    }
    if (s.g_row_w == s.row_w && s.g_off == 0 && cnt == 4 && (reinterpret_cast<uintptr_t>(s.g) & 15) == 0) {
        const float4 g4 = reinterpret_cast<const float4*>(s.g)[base];
        g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    } else {
        long long row = e0 / s.row_w;
        int c = (int)(e0 - row * s.row_w);
        const float* gp = s.g + row * s.g_row_w + s.g_off;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            g[k] = k < cnt ? gp[c] : 0.f;
            if (++c == s.row_w) { c = 0; gp += s.g_row_w; }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (s.type == 1) { const float sg = 1.f / (1.f + expf(-p[k])); g[k] = g[k] * sg * (1.f - sg); }   // sigmoid'
        else if (s.type == 2) g[k] = g[k] * expf(p[k]);                                                 // exp'
        adam_update(p[k], m[k], v[k], g[k], s.step_size, a);
    }
    if (cnt == 4) {
        reinterpret_cast<float4*>(s.p)[base] = make_float4(p[0], p[1], p[2], p[3]);
        reinterpret_cast<float4*>(s.m)[base] = make_float4(m[0], m[1], m[2], m[3]);
        reinterpret_cast<float4*>(s.v)[base] = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < cnt) { s.p[e0 + k] = p[k]; s.m[e0 + k] = m[k]; s.v[e0 + k] = v[k]; }
    }
}

// luciddreamer.py:308-312 + scene/gaussian_model.py:405-407 in one pass: for every visible Gaussian (radii > 0)
//   max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += |dL_dmeans2D.xy|;  denom += 1
__global__ void __launch_bounds__(kT)
k_densify_stats(const int P, const int* __restrict__ radii, const float* __restrict__ dm2, float* __restrict__ accum,
                float* __restrict__ denom, float* __restrict__ max_radii) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = dm2[3 * (size_t)i], gy = dm2[3 * (size_t)i + 1];
    max_radii[i] = fmaxf(max_radii[i], (float)r);
    accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.f;
}

}  // namespace

void gs_launch_densify_stats(int P, const int* radii, const float* dm2, float* accum, float* denom, float* max_radii,
                             cudaStream_t s) {
    if (P > 0) k_densify_stats<<<(P + kT - 1) / kT, kT, 0, s>>>(P, radii, dm2, accum, denom, max_radii);
}

int gs_launch_gaussian_adam(GsAdamArgs a, cudaStream_t s) {
    long long total = 0;
    for (int k = 0; k < a.nseg; k++) {
        const long long n = a.seg[k].type == 3 ? a.seg[k].rows : (a.seg[k].rows * a.seg[k].row_w + 3) / 4;
        const long long blocks = (n + kT - 1) / kT;
        if (blocks > 0x7fffffffLL) return -1;
        a.seg[k].blocks = (int)blocks;
        total += blocks;
    }
    if (total == 0) return 0;
    if (total > 0x7fffffffLL) return -1;
    k_gaussian_adam<<<(unsigned)total, kT, 0, s>>>(a);
    return 0;
}
