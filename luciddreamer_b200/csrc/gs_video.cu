// Frame packing for the batched video-render path (sm_100a) -- the "next" row SURVEY.md 8f-4.
//
// The reference's render_video loop (luciddreamer.py:250-262) does, PER FRAME, two device->host copies of float images
// followed by numpy work on the host:
//     frame8 = np.round(frame.permute(1,2,0).cpu().numpy().clip(0,1) * 255.).astype(np.uint8)
//     depth  = -(depth * (depth > 0)).cpu().numpy();  dmin = min(dmin, depth.min());  dmax = max(dmax, depth.max())
// i.e. a stream sync per frame and 16 bytes/pixel over PCIe.  This kernel does the same arithmetic on the device
// straight into slot `frame` of a batch buffer: uint8 HWC colour (round-half-even like np.round, 3 bytes/pixel),
// the negated masked depth, and running min/max -- so a whole video needs ONE host copy at the end.
#include <cfloat>

#include "gs_common.cuh"

namespace {

__device__ __forceinline__ unsigned v_f2ord(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one thread = 4 consecutive pixels of a row-major image: 3 x float4 loads (planar CHW), 12 packed bytes out
__global__ void __launch_bounds__(256)
k_pack_frame(const int N, const float* __restrict__ color, const float* __restrict__ depth, uint8_t* __restrict__ rgb8,
             float* __restrict__ depth_out, unsigned* __restrict__ minmax) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = q * 4;
    float dmin = FLT_MAX, dmax = -FLT_MAX;
    if (p0 < N) {
        const int cnt = min(4, N - p0);
        float c[3][4], d[4];
        if (cnt == 4 && (N & 3) == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float4 v = *reinterpret_cast<const float4*>(color + (size_t)ch * N + p0);
                c[ch][0] = v.x; c[ch][1] = v.y; c[ch][2] = v.z; c[ch][3] = v.w;
            }
            if (depth) { const float4 v = *reinterpret_cast<const float4*>(depth + p0); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = min(p0 + k, N - 1);
#pragma unroll
                for (int ch = 0; ch < 3; ch++) c[ch][k] = color[(size_t)ch * N + p];
                if (depth) d[k] = depth[p];
            }
        }
        uint8_t out[12];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                out[3 * k + ch] = (uint8_t)__float2int_rn(fminf(fmaxf(c[ch][k], 0.f), 1.f) * 255.f);   // np.round: half to even
        if (cnt == 4 && (((size_t)rgb8 + 3 * (size_t)p0) & 3) == 0) {
            uint32_t* o = reinterpret_cast<uint32_t*>(rgb8 + 3 * (size_t)p0);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(out);
            o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        } else {
            for (int k = 0; k < 3 * cnt; k++) rgb8[3 * (size_t)p0 + k] = out[k];
        }
        if (depth) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < cnt) {
                    // -(depth * (depth > 0)) with IEEE signs of zero (0 -> -0, negative -> +0) and NaN for -inf like
                    // numpy: the product is formed for real and negated by flipping the sign bit (a plain "-"
                    // lets the compiler fold the negation into the 0/1 factor and lose the sign of zero)
                    const float t = __fmul_rn(d[k], d[k] > 0.f ? 1.f : 0.f);
                    const float v = __uint_as_float(__float_as_uint(t) ^ 0x80000000u);
                    depth_out[p0 + k] = v;
                    dmin = fminf(dmin, v); dmax = fmaxf(dmax, v);
                }
            }
        }
    }
    if (depth && minmax) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            dmin = fminf(dmin, __shfl_xor_sync(~0u, dmin, o));
            dmax = fmaxf(dmax, __shfl_xor_sync(~0u, dmax, o));
        }
        if ((threadIdx.x & 31) == 0 && dmin <= dmax) { atomicMin(&minmax[0], v_f2ord(dmin)); atomicMax(&minmax[1], v_f2ord(dmax)); }
    }
}

__global__ void k_minmax_decode(unsigned* mm, float* out2) {      // ordered-uint -> float, in place semantics kept simple
    if (threadIdx.x < 2) {
        const unsigned u = mm[threadIdx.x];
        out2[threadIdx.x] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
    }
}

}  // namespace

void gs_launch_pack_frame(int N, const float* color, const float* depth, uint8_t* rgb8, float* depth_out, unsigned* minmax,
                          cudaStream_t s) {
    if (N <= 0) return;
    const int q = (N + 3) / 4;
    k_pack_frame<<<(q + 255) / 256, 256, 0, s>>>(N, color, depth, rgb8, depth_out, minmax);
}

void gs_launch_minmax_decode(unsigned* mm, float* out2, cudaStream_t s) { k_minmax_decode<<<1, 32, 0, s>>>(mm, out2); }
