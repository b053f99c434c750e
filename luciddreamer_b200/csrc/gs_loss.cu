// Fused photometric L1 loss + gradient (sm_100a) -- first slice of the "next" row SURVEY.md 8f-1.
//
// Replaces, for the L1 term of LucidDreamer's training loss (luciddreamer.py:301-303, utils/loss.py:18
// `l1_loss = |network_output - gt|.mean()`), the chain  target.float()/255 -> sub -> abs -> mean  and its autograd
// backward (five elementwise kernels and four image-sized temporaries) by ONE pass: reads the rendered image and
// the uint8 target once, writes dL/dcolor = weight * sign(color - gt) / N (the rasterizer's incoming cotangent)
// and accumulates the scalar loss.  HBM-bound: 4 + 1 + 4 bytes per element.
#include "gs_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_l1_loss_grad(const float* __restrict__ color /*[3,H,W]*/, const uint8_t* __restrict__ target /*[H,W,3] u8*/,
               int H, int W, float weight, float* __restrict__ dL_dcolor /*[3,H,W]*/, float* __restrict__ loss) {
    const long long HW = (long long)H * W;
    const float inv_n = 1.0f / (3.0f * (float)HW);
    float part = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float c = color[ch * HW + p];
            const float g = (float)target[p * 3 + ch] * (1.0f / 255.0f);
            const float d = c - g;
            part += fabsf(d);
            dL_dcolor[ch * HW + p] = d > 0.f ? weight * inv_n : (d < 0.f ? -weight * inv_n : 0.f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s[w];
        atomicAdd(loss, t * inv_n * weight);
    }
}

}  // namespace

void gs_launch_l1_loss_grad(const float* color, const uint8_t* target, int H, int W, float weight, float* dL_dcolor,
                            float* loss, int num_sms, cudaStream_t s) {
    cudaMemsetAsync(loss, 0, sizeof(float), s);
    k_l1_loss_grad<<<num_sms * 8, 256, 0, s>>>(color, target, H, W, weight, dL_dcolor, loss);
}
