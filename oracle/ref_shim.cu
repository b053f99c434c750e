// C-ABI shim over the UNMODIFIED reference CUDA rasterizer (compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/libref_rasterizer.so).  Our code; calls the reference's public static
// C++ API CudaRasterizer::Rasterizer::{forward,backward,markVisible} (cuda_rasterizer/rasterizer.h:20-87)
// and does what its torch binding does around it (rasterize_points.cu:57-117,144-199): zero-filled
// outputs, three growable scratch buffers, P == 0 short-circuit.  Test / baseline infrastructure only.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <functional>
#include "cuda_rasterizer/rasterizer.h"

namespace {
struct Buf {
    char* p = nullptr;
    size_t cap = 0;
    char* get(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);
            size_t want = n + n / 4 + 256;
            if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return p;
    }
    ~Buf() { if (p) cudaFree(p); }
};
}  // namespace

struct RefCtx {
    Buf geom, binning, img;
    int P = 0, W = 0, H = 0, R = 0;
};

extern "C" {

RefCtx* ref_create() { return new RefCtx(); }
void ref_destroy(RefCtx* c) { delete c; }

// All pointers are DEVICE pointers (nullptr = absent, like the empty-tensor convention of the binding).
// Outputs out_color[3HW], out_depth[HW], radii[P] are zero-filled here like torch::full(..., 0).
int ref_forward(RefCtx* c, int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_depth, int* radii, int debug) {
    cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * (size_t)W * H, 0);
    cudaMemsetAsync(out_depth, 0, sizeof(float) * (size_t)W * H, 0);
    if (P > 0) cudaMemsetAsync(radii, 0, sizeof(int) * (size_t)P, 0);
    c->P = P; c->W = W; c->H = H; c->R = 0;
    if (P == 0) return 0;
    std::function<char*(size_t)> g = [c](size_t n) { return c->geom.get(n); };
    std::function<char*(size_t)> b = [c](size_t n) { return c->binning.get(n); };
    std::function<char*(size_t)> i = [c](size_t n) { return c->img.get(n); };
    c->R = CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities,
                                               scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                                               projmatrix, campos, tan_fovx, tan_fovy, prefiltered != 0, out_color,
                                               out_depth, radii, debug != 0);
    return c->R;
}

// Gradient outputs (device): dL_dmeans2D[3P] dL_dcolors[3P] dL_dopacity[P] dL_dmeans3D[3P] dL_dcov3D[6P]
// dL_dsh[P*M*3] dL_dscales[3P] dL_drots[4P] dL_dconic[4P]; zero-filled here like torch::zeros.
void ref_backward(RefCtx* c, int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                  const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                  const float* dL_dpix, const float* dL_depths, float* dL_dmeans2D, float* dL_dconic,
                  float* dL_dopacity, float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh,
                  float* dL_dscales, float* dL_drots, int debug) {
    size_t Ps = (size_t)P;
    cudaMemsetAsync(dL_dmeans3D, 0, 12 * Ps, 0); cudaMemsetAsync(dL_dmeans2D, 0, 12 * Ps, 0);
    cudaMemsetAsync(dL_dcolors, 0, 12 * Ps, 0);  cudaMemsetAsync(dL_dconic, 0, 16 * Ps, 0);
    cudaMemsetAsync(dL_dopacity, 0, 4 * Ps, 0);  cudaMemsetAsync(dL_dcov3D, 0, 24 * Ps, 0);
    if (M > 0) cudaMemsetAsync(dL_dsh, 0, 12 * Ps * M, 0);
    cudaMemsetAsync(dL_dscales, 0, 12 * Ps, 0);  cudaMemsetAsync(dL_drots, 0, 16 * Ps, 0);
    if (P == 0) return;
    CudaRasterizer::Rasterizer::backward(P, D, M, c->R, bg, W, H, means3D, shs, colors_precomp, scales,
                                         scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
                                         tan_fovx, tan_fovy, radii, c->geom.p, c->binning.p, c->img.p, dL_dpix,
                                         dL_depths, dL_dmeans2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dmeans3D,
                                         dL_dcov3D, dL_dsh, dL_dscales, dL_drots, debug != 0);
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    if (P > 0) CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

}  // extern "C"
