// Fused photometric loss + gradient (sm_100a) -- the "next" row SURVEY.md 8f-1: the step right after the
// rasterizer in every training iteration of the reference (luciddreamer.py:301-304):
//     loss = (1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))           lambda_dssim = 0.2
// utils/loss.py:18 (l1_loss) and :38-69 (ssim: 11x11 Gaussian window sigma 1.5, zero padding 5, per-channel
// conv2d of x, y, x^2, y^2, xy).  The reference spends 5 grouped 11x11 convolutions, ~15 elementwise kernels and as
// many image-sized temporaries in the forward, and the autograd transpose of all of it in the backward.  Here:
//   k_photo_fwd   one pass: separable window in shared memory for the five moments of a 16x16 tile (+5 halo),
//                 SSIM value, L1 value, and the three partial-derivative maps d ssim / d(mu1, E[x^2], E[xy])
//   k_photo_bwd   one pass: separable window over the three maps (the window is symmetric, so the transposed
//                 convolution is the same convolution), combined with x, y and the L1 sign into dL/dimage
// -- the cotangent the rasterizer's backward consumes.  HBM-bound: ~ (2 + 3) + (3 + 2 + 1) floats per element.
// k_l1_loss_grad is the L1-only variant on a uint8 target (what bench.py's e2e leg uploads per step).
#include "gs_common.cuh"

namespace {

constexpr int kWin = 11, kHalo = 5, kTile = 16, kIn = kTile + 2 * kHalo;   // 26

// utils/loss.py:26-28: exp(-(i-5)^2 / (2 * 1.5^2)) in double, stored as float32, normalised in float32 -- the
// eleven resulting float32 values, so that the window is bit-identical to the reference's and costs no expf here
__device__ __forceinline__ void gauss_window(float* w) {
    w[0] = 1.028380124e-03f; w[1] = 7.598758209e-03f; w[2] = 3.600077331e-02f; w[3] = 1.093606874e-01f;
    w[4] = 2.130055279e-01f; w[5] = 2.660117149e-01f; w[6] = 2.130055279e-01f; w[7] = 1.093606874e-01f;
    w[8] = 3.600077331e-02f; w[9] = 7.598758209e-03f; w[10] = 1.028380124e-03f;
}

__global__ void __launch_bounds__(kTile * kTile)
k_photo_fwd(const float* __restrict__ img, const float* __restrict__ gt, int H, int W, float* __restrict__ maps /*[3 maps][3][H][W]*/,
            float* __restrict__ sums /*[0] = sum |x-y|, [1] = sum ssim*/) {
    __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
    __shared__ float h[5][kIn][kTile + 1];
    __shared__ float red[2][8];
    float w[kWin];
    gauss_window(w);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, tid = threadIdx.x;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const float* X = img + ch * plane;
    const float* Y = gt + ch * plane;
    for (int e = tid; e < kIn * kIn; e += kTile * kTile) {
        const int r = e / kIn, c = e - r * kIn;
        const int gy = y0 + r - kHalo, gx = x0 + c - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;        // zero padding (conv2d padding = 5)
        sx[r][c] = in ? X[(size_t)gy * W + gx] : 0.f;
        sy[r][c] = in ? Y[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < kIn * kTile; e += kTile * kTile) {             // horizontal pass: 26 rows x 16 columns
        const int r = e / kTile, c = e - r * kTile;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float xv = sx[r][c + k], yv = sy[r][c + k], wk = w[k];
            a0 += wk * xv; a1 += wk * yv; a2 += wk * xv * xv; a3 += wk * yv * yv; a4 += wk * xv * yv;
        }
        h[0][r][c] = a0; h[1][r][c] = a1; h[2][r][c] = a2; h[3][r][c] = a3; h[4][r][c] = a4;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; k++) {
        const float wk = w[k];
        mu1 += wk * h[0][ty + k][tx]; mu2 += wk * h[1][ty + k][tx]; e11 += wk * h[2][ty + k][tx];
        e22 += wk * h[3][ty + k][tx]; e12 += wk * h[4][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    float l1 = 0.f, ss = 0.f;
    if (gx < W && gy < H) {
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;              // utils/loss.py:60-61
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu1_mu2;
        const float A1 = 2.f * mu1_mu2 + C1, A2 = 2.f * s12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
        const float inv = 1.f / (B1 * B2);
        ss = A1 * A2 * inv;
        // partial derivatives with (mu1, E[x^2], E[xy]) as the independent windowed moments of x
        const float d_mu = ((2.f * mu2 * A2 - 2.f * mu2 * A1) - ss * (2.f * mu1 * B2 - 2.f * mu1 * B1)) * inv;
        const float d_11 = -ss / B2;
        const float d_12 = 2.f * A1 * inv;
        const size_t o = (size_t)ch * plane + (size_t)gy * W + gx;
        maps[o] = d_mu; maps[3 * plane + o] = d_11; maps[6 * plane + o] = d_12;
        l1 = fabsf(sx[ty + kHalo][tx + kHalo] - sy[ty + kHalo][tx + kHalo]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { l1 += __shfl_xor_sync(0xffffffffu, l1, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = l1; red[1][tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) { a += red[0][k]; b += red[1][k]; }
        atomicAdd(&sums[0], a); atomicAdd(&sums[1], b);
    }
}

__global__ void __launch_bounds__(kTile * kTile)
k_photo_bwd(const float* __restrict__ img, const float* __restrict__ gt, const float* __restrict__ maps, int H, int W,
            float lambda_dssim, float* __restrict__ dL_dimg) {
    __shared__ float sm[3][kIn][kIn + 1];
    __shared__ float h[3][kIn][kTile + 1];
    float w[kWin];
    gauss_window(w);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, tid = threadIdx.x;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    for (int e = tid; e < kIn * kIn; e += kTile * kTile) {
        const int r = e / kIn, c = e - r * kIn;
        const int gy = y0 + r - kHalo, gx = x0 + c - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = (size_t)ch * plane + (size_t)gy * W + gx;
        sm[0][r][c] = in ? maps[o] : 0.f;
        sm[1][r][c] = in ? maps[3 * plane + o] : 0.f;
        sm[2][r][c] = in ? maps[6 * plane + o] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < kIn * kTile; e += kTile * kTile) {
        const int r = e / kTile, c = e - r * kTile;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; k++) { const float wk = w[k]; a0 += wk * sm[0][r][c + k]; a1 += wk * sm[1][r][c + k]; a2 += wk * sm[2][r][c + k]; }
        h[0][r][c] = a0; h[1][r][c] = a1; h[2][r][c] = a2;
    }
    __syncthreads();
    float c_mu = 0.f, c_11 = 0.f, c_12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; k++) { const float wk = w[k]; c_mu += wk * h[0][ty + k][tx]; c_11 += wk * h[1][ty + k][tx]; c_12 += wk * h[2][ty + k][tx]; }
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
        const size_t o = (size_t)ch * plane + (size_t)gy * W + gx;
        const float x = img[o], y = gt[o];
        const float inv_n = 1.f / (3.f * (float)plane);
        const float d = x - y;
        const float g_l1 = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        const float g_ssim = c_mu + 2.f * x * c_11 + y * c_12;            // d(sum ssim)/dx
        dL_dimg[o] = ((1.f - lambda_dssim) * g_l1 - lambda_dssim * g_ssim) * inv_n;
    }
}

__global__ void k_photo_finish(const float* __restrict__ sums, int H, int W, float lambda_dssim, float* __restrict__ out) {
    const float inv_n = 1.f / (3.f * (float)H * (float)W);
    const float l1 = sums[0] * inv_n, ssim = sums[1] * inv_n;
    out[0] = (1.f - lambda_dssim) * l1 + lambda_dssim * (1.f - ssim);
    out[1] = l1;
    out[2] = ssim;
}

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
    gs_launch_clear_words(reinterpret_cast<uint32_t*>(loss), 1, s);
    k_l1_loss_grad<<<num_sms * 8, 256, 0, s>>>(color, target, H, W, weight, dL_dcolor, loss);
}

// scratch: [2 floats sums | pad to 256 B | 9 planes of H*W floats]
void gs_launch_photometric(const float* img, const float* gt, int H, int W, float lambda_dssim, void* scratch,
                           float* dL_dimg, float* loss3, cudaStream_t s) {
    float* sums = (float*)scratch;
    float* maps = (float*)((char*)scratch + 256);
    gs_launch_clear_words(reinterpret_cast<uint32_t*>(sums), 2, s);
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, 3);
    k_photo_fwd<<<grid, kTile * kTile, 0, s>>>(img, gt, H, W, maps, sums);
    k_photo_finish<<<1, 1, 0, s>>>(sums, H, W, lambda_dssim, loss3);
    if (dL_dimg) k_photo_bwd<<<grid, kTile * kTile, 0, s>>>(img, gt, maps, H, W, lambda_dssim, dL_dimg);
}
