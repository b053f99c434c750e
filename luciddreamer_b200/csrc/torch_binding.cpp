// Torch binding over the C ABI (include/gsraster.h) -- what RAST/rasterize_points.cu:35-200 (RasterizeGaussiansCUDA /
// RasterizeGaussiansBackwardCUDA, bound by RAST/ext.cpp:15-19) becomes when the rasterizer lives behind gsraster.h:
// torch only allocates, finds the current stream and carries tensors; every kernel is in libgsraster_b200.so.
// It is the native twin of luciddreamer_b200/rasterizer.py:_forward_impl/_backward_impl (same calls in the same order,
// incl. the speculative render before the counts wait) and exists to take ~100 us of Python per step off the host.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>
#include <vector>

#include "gsraster.h"

namespace {

// Reference convention: an empty tensor means "absent" (RAST/.../__init__.py:198-208); otherwise float32, contiguous,
// on the compute device (the reference calls .contiguous() on everything, rasterize_points.cu:95-113).
at::Tensor prep(const c10::optional<at::Tensor>& t, const at::Device& dev, std::vector<at::Tensor>& keep) {
    if (!t.has_value() || !t->defined() || t->numel() == 0) return at::Tensor();
    at::Tensor x = *t;
    if (x.device() != dev || x.scalar_type() != at::kFloat) x = x.to(dev, at::kFloat);
    x = x.contiguous();
    if (reinterpret_cast<uintptr_t>(x.data_ptr()) & 15) x = x.clone();   // 128-bit loads in the kernels (gsraster.h)
    keep.push_back(x);
    return x;
}
const float* fp(const at::Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

int64_t round_cap(double n) {
    int64_t v = (int64_t)n;
    v = (v + 63) / 64 * 64;
    return v < 64 ? 64 : v;
}

struct Frame {
    GsFrame f;
    std::vector<at::Tensor> keep;
    at::Device dev{at::kCUDA, 0};
    int64_t P = 0, M = 0;
};

void fill_frame(Frame& fr, const c10::optional<at::Tensor>& bg, const at::Tensor& means3D,
                const c10::optional<at::Tensor>& colors, const c10::optional<at::Tensor>& opacity,
                const c10::optional<at::Tensor>& scales, const c10::optional<at::Tensor>& rotations, double scale_modifier,
                const c10::optional<at::Tensor>& cov3D, const c10::optional<at::Tensor>& viewmatrix,
                const c10::optional<at::Tensor>& projmatrix, double tan_fovx, double tan_fovy, int64_t H, int64_t W,
                const c10::optional<at::Tensor>& sh, int64_t degree, const c10::optional<at::Tensor>& campos,
                bool prefiltered, bool debug) {
    TORCH_CHECK(means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)");
    fr.dev = means3D.device();
    fr.P = means3D.size(0);
    auto& k = fr.keep;
    const at::Tensor bg_ = prep(bg, fr.dev, k), m3 = prep(means3D, fr.dev, k), sh_ = prep(sh, fr.dev, k),
                     col = prep(colors, fr.dev, k), op = prep(opacity, fr.dev, k), sc = prep(scales, fr.dev, k),
                     rot = prep(rotations, fr.dev, k), cov = prep(cov3D, fr.dev, k), vm = prep(viewmatrix, fr.dev, k),
                     pm = prep(projmatrix, fr.dev, k), cp = prep(campos, fr.dev, k);
    fr.M = sh_.defined() ? sh_.size(1) : 0;
    GsFrame& f = fr.f;
    f.P = (int32_t)fr.P; f.D = (int32_t)degree; f.M = (int32_t)fr.M; f.W = (int32_t)W; f.H = (int32_t)H;
    f.tan_fovx = (float)tan_fovx; f.tan_fovy = (float)tan_fovy; f.scale_modifier = (float)scale_modifier;
    f.prefiltered = prefiltered ? 1 : 0; f.debug = debug ? 1 : 0;
    f.bg = fp(bg_); f.means3D = fp(m3); f.shs = fp(sh_); f.colors_precomp = fp(col);
    f.opacities = fp(op); f.scales = fp(sc); f.rotations = fp(rot); f.cov3D_precomp = fp(cov);
    f.viewmatrix = fp(vm); f.projmatrix = fp(pm); f.campos = fp(cp);
}

#define GS_OK_OR_THROW(call) TORCH_CHECK((call) == 0, "gsraster: ", gs_last_error())

// -> (num_rendered, color, depth, radii, geom, binning, img, pair_capacity, num_visible, num_pairs, ticket)
// static_cap >= 0: sync-free forward (gsraster.h: gs_forward_counts_peek) -- the render is enqueued with exactly that pair
// capacity and the host never waits; num_rendered / num_visible / num_pairs come back as -1 and the caller checks the
// ticket later.  This is also the only mode that can be captured into a CUDA graph.
std::tuple<int64_t, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, int64_t, int64_t, int64_t, int64_t>
forward(int64_t ctx_ptr, const c10::optional<at::Tensor>& bg, const at::Tensor& means3D,
        const c10::optional<at::Tensor>& colors, const c10::optional<at::Tensor>& opacity,
        const c10::optional<at::Tensor>& scales, const c10::optional<at::Tensor>& rotations, double scale_modifier,
        const c10::optional<at::Tensor>& cov3D, const c10::optional<at::Tensor>& viewmatrix,
        const c10::optional<at::Tensor>& projmatrix, double tan_fovx, double tan_fovy, int64_t H, int64_t W,
        const c10::optional<at::Tensor>& sh, int64_t degree, const c10::optional<at::Tensor>& campos, bool prefiltered,
        bool debug, int64_t pair_hint, int64_t static_cap) {
    Frame fr;
    fill_frame(fr, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, viewmatrix, projmatrix, tan_fovx,
               tan_fovy, H, W, sh, degree, campos, prefiltered, debug);
    GsContext* ctx = reinterpret_cast<GsContext*>(ctx_ptr);
    const c10::cuda::CUDAGuard guard(fr.dev);
    gs_stream_t s = (gs_stream_t)c10::cuda::getCurrentCUDAStream(fr.dev.index()).stream();
    const auto f32 = at::TensorOptions().dtype(at::kFloat).device(fr.dev);
    const auto u8 = f32.dtype(at::kByte);
    const int64_t P = fr.P;
    at::Tensor color = at::empty({3, H, W}, f32), depth = at::empty({1, H, W}, f32);
    at::Tensor radii = at::empty({P}, f32.dtype(at::kInt));
    const int64_t ng = (int64_t)gs_geom_bytes((int32_t)P), ni = (int64_t)gs_image_bytes((int32_t)W, (int32_t)H);
    at::Tensor scratch = at::empty({ng + ni}, u8);           // one allocation; both sizes are multiples of 256 B
    at::Tensor geom = scratch.narrow(0, 0, ng), img = scratch.narrow(0, ng, ni);
    int32_t* rad = radii.data_ptr<int32_t>();
    int32_t ticket = -1;
    GS_OK_OR_THROW(gs_forward_preprocess(ctx, &fr.f, geom.data_ptr(), img.data_ptr(), rad, s, &ticket));
    GsCounts n{};
    at::Tensor binning;
    int64_t cap = 0;
    if (static_cap >= 0) {
        cap = round_cap((double)static_cap);
        binning = at::empty({(int64_t)gs_binning_bytes(cap)}, u8);
        GS_OK_OR_THROW(gs_forward_render(ctx, &fr.f, rad, geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                         color.data_ptr<float>(), depth.data_ptr<float>(), 0, s));
        return {-1, color, depth, radii, geom, binning, img, cap, -1, -1, (int64_t)ticket};
    }
    if (pair_hint < 0) {
        // first frame on this device: learn the pair count (one event wait), then render
        GS_OK_OR_THROW(gs_forward_counts(ctx, ticket, &n));
        cap = round_cap((double)n.num_pairs * 1.25 + 4096);
        binning = at::empty({(int64_t)gs_binning_bytes(cap)}, u8);
        GS_OK_OR_THROW(gs_forward_render(ctx, &fr.f, rad, geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                         color.data_ptr<float>(), depth.data_ptr<float>(), 0, s));
    } else {
        // speculative: enqueue the render with the previous capacity, *then* wait for the count; the GPU never idles on
        // the host (reference: blocking cudaMemcpy, rasterizer_impl.cu:282)
        cap = round_cap((double)pair_hint * 1.25 + 4096);
        binning = at::empty({(int64_t)gs_binning_bytes(cap)}, u8);
        GS_OK_OR_THROW(gs_forward_render(ctx, &fr.f, rad, geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                         color.data_ptr<float>(), depth.data_ptr<float>(), 0, s));
        GS_OK_OR_THROW(gs_forward_counts(ctx, ticket, &n));
        if (n.num_pairs > cap) {                             // device-side guard skipped the render: grow and redo it
            cap = round_cap((double)n.num_pairs * 1.25 + 4096);
            binning = at::empty({(int64_t)gs_binning_bytes(cap)}, u8);
            GS_OK_OR_THROW(gs_forward_render(ctx, &fr.f, rad, geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                             color.data_ptr<float>(), depth.data_ptr<float>(), 1, s));
        }
    }
    return {n.num_rendered, color, depth, radii, geom, binning, img, cap, n.num_visible, n.num_pairs, (int64_t)ticket};
}

// -> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
// (order of rasterize_points.h:40-63; undefined tensors for absent inputs / gradients nobody asked for)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>
backward(int64_t ctx_ptr, const c10::optional<at::Tensor>& bg, const at::Tensor& means3D,
         const c10::optional<at::Tensor>& colors, const c10::optional<at::Tensor>& opacity,
         const c10::optional<at::Tensor>& scales, const c10::optional<at::Tensor>& rotations, double scale_modifier,
         const c10::optional<at::Tensor>& cov3D, const c10::optional<at::Tensor>& viewmatrix,
         const c10::optional<at::Tensor>& projmatrix, double tan_fovx, double tan_fovy, int64_t H, int64_t W,
         const c10::optional<at::Tensor>& sh, int64_t degree, const c10::optional<at::Tensor>& campos, bool debug,
         const at::Tensor& radii, const at::Tensor& geom, const at::Tensor& binning, const at::Tensor& img, int64_t cap,
         int64_t num_visible, const at::Tensor& grad_color, bool want_colors, bool want_cov) {
    Frame fr;
    fill_frame(fr, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, viewmatrix, projmatrix, tan_fovx,
               tan_fovy, H, W, sh, degree, campos, false, debug);
    GsContext* ctx = reinterpret_cast<GsContext*>(ctx_ptr);
    const c10::cuda::CUDAGuard guard(fr.dev);
    gs_stream_t s = (gs_stream_t)c10::cuda::getCurrentCUDAStream(fr.dev.index()).stream();
    const auto f32 = at::TensorOptions().dtype(at::kFloat).device(fr.dev);
    const int64_t P = fr.P, M = fr.M;
    const at::Tensor gc = prep(grad_color, fr.dev, fr.keep);
    TORCH_CHECK(gc.defined() && gc.numel() == 3 * H * W, "dL_dout_color must be [3, H, W]");
    // all gradients are views of ONE flat allocation, in bucket order [means3D | sh | opacity | scales | rotations |
    // means2D]; segment starts stay 256-byte aligned (128-bit stores in the kernels); every row is overwritten
    const int64_t w[6] = {3, 3 * M, 1, 3, 4, 3};
    int64_t off[6], o = 0;
    for (int k = 0; k < 6; k++) { off[k] = o; o += (P * w[k] + 63) / 64 * 64; }
    at::Tensor flat = at::empty({o}, f32);
    auto seg = [&](int k, at::IntArrayRef shape) { return flat.narrow(0, off[k], P * w[k]).view(shape); };
    at::Tensor dm3 = seg(0, {P, 3}), dsh = seg(1, {P, M, 3}), dop = seg(2, {P, 1}), dsc = seg(3, {P, 3}),
               drot = seg(4, {P, 4}), dm2 = seg(5, {P, 3});
    if (!fr.f.scales) dsc.zero_();
    if (!fr.f.rotations) drot.zero_();
    at::Tensor dcol = want_colors ? at::empty({P, 3}, f32) : at::Tensor();
    at::Tensor dcov = want_cov ? at::empty({P, 6}, f32) : at::Tensor();
    GsGrads g{};
    g.dL_dmeans3D = dm3.data_ptr<float>(); g.dL_dmeans2D = dm2.data_ptr<float>(); g.dL_dopacity = dop.data_ptr<float>();
    g.dL_dsh = M > 0 ? dsh.data_ptr<float>() : nullptr;
    g.dL_dscales = fr.f.scales ? dsc.data_ptr<float>() : nullptr;
    g.dL_drotations = fr.f.rotations ? drot.data_ptr<float>() : nullptr;
    g.dL_dcolors = dcol.defined() ? dcol.data_ptr<float>() : nullptr;
    g.dL_dcov3D = dcov.defined() ? dcov.data_ptr<float>() : nullptr;
    // zero-fill of the outputs on the context's side stream, beside the tile pass (no-op when they do not qualify)
    GS_OK_OR_THROW(gs_backward_prefill(ctx, &fr.f, img.data_ptr(), &g, s));
    GS_OK_OR_THROW(gs_backward_blend(ctx, &fr.f, geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(), fp(gc), s));
    const int64_t nscr = (int64_t)gs_backward_scratch_bytes(num_visible >= 0 ? num_visible : P);
    at::Tensor scratch = at::empty({nscr}, f32.dtype(at::kByte));
    GS_OK_OR_THROW(gs_backward_gradients(ctx, &fr.f, radii.data_ptr<int32_t>(), geom.data_ptr(), img.data_ptr(),
                                         scratch.data_ptr(), (size_t)nscr, &g, s));
    return {dm2, dcol, dop, dm3, dcov, dsh, dsc, drot};
}

// ---- the autograd node itself in C++ (RAST/depth_diff_gaussian_rasterization_min/__init__.py:44-156): same inputs,
// outputs (color, radii, depth), gradient order and None/zero conventions as the Python _RasterizeGaussians
int64_t g_last_pairs[64] = {0};
int64_t g_last_ticket[64] = {0};

struct RasterizeFn : public torch::autograd::Function<RasterizeFn> {
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, at::Tensor means3D,
                                                  at::Tensor means2D, at::Tensor sh, at::Tensor colors, at::Tensor opacity,
                                                  at::Tensor scales, at::Tensor rotations, at::Tensor cov3D, at::Tensor bg,
                                                  at::Tensor viewmatrix, at::Tensor projmatrix, at::Tensor campos,
                                                  int64_t ctx_ptr, double scale_modifier, double tan_fovx, double tan_fovy,
                                                  int64_t H, int64_t W, int64_t degree, bool prefiltered, bool debug,
                                                  int64_t pair_hint, int64_t static_cap) {
        auto r = ::forward(ctx_ptr, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, viewmatrix,
                           projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, prefiltered, debug, pair_hint,
                           static_cap);
        at::Tensor color = std::get<1>(r), depth = std::get<2>(r), radii = std::get<3>(r);
        const int dev = means3D.device().index();
        if (dev >= 0 && dev < 64) { g_last_pairs[dev] = std::get<9>(r); g_last_ticket[dev] = std::get<10>(r); }
        ctx->save_for_backward({radii, std::get<4>(r), std::get<5>(r), std::get<6>(r), means3D, sh, colors, opacity, scales,
                                rotations, cov3D, bg, viewmatrix, projmatrix, campos});
        ctx->saved_data["ctx_ptr"] = ctx_ptr;
        ctx->saved_data["scale_modifier"] = scale_modifier;
        ctx->saved_data["tan_fovx"] = tan_fovx;
        ctx->saved_data["tan_fovy"] = tan_fovy;
        ctx->saved_data["H"] = H;
        ctx->saved_data["W"] = W;
        ctx->saved_data["degree"] = degree;
        ctx->saved_data["debug"] = debug;
        ctx->saved_data["cap"] = std::get<7>(r);
        ctx->saved_data["nvis"] = std::get<8>(r);
        ctx->mark_non_differentiable({radii});
        ctx->set_materialize_grads(false);       // no zero-fill kernels for the unused grad_radii / grad_depth
        return {color, radii, depth};
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_out) {
        const auto sv = ctx->get_saved_variables();
        const at::Tensor &radii = sv[0], &geom = sv[1], &binning = sv[2], &img = sv[3], &means3D = sv[4], &sh = sv[5],
                         &colors = sv[6], &opacity = sv[7], &scales = sv[8], &rotations = sv[9], &cov3D = sv[10],
                         &bg = sv[11], &viewmatrix = sv[12], &projmatrix = sv[13], &campos = sv[14];
        const int64_t H = ctx->saved_data["H"].toInt(), W = ctx->saved_data["W"].toInt();
        at::Tensor gc = grad_out[0];
        if (!gc.defined())                       // only depth was used downstream: depth carries no gradient
            gc = at::zeros({3, H, W}, at::TensorOptions().dtype(at::kFloat).device(means3D.device()));
        auto has = [](const at::Tensor& t) { return t.defined() && t.numel() > 0; };
        auto g = ::backward(ctx->saved_data["ctx_ptr"].toInt(), bg, means3D, colors, opacity, scales, rotations,
                            ctx->saved_data["scale_modifier"].toDouble(), cov3D, viewmatrix, projmatrix,
                            ctx->saved_data["tan_fovx"].toDouble(), ctx->saved_data["tan_fovy"].toDouble(), H, W, sh,
                            ctx->saved_data["degree"].toInt(), campos, ctx->saved_data["debug"].toBool(), radii, geom,
                            binning, img, ctx->saved_data["cap"].toInt(), ctx->saved_data["nvis"].toInt(), gc, has(colors),
                            has(cov3D));
        const at::Tensor none;
        // (dm2, dcol, dop, dm3, dcov, dsh, dsc, drot) -> the order of the forward arguments (__init__.py:144-156)
        return {std::get<3>(g), std::get<0>(g), has(sh) ? std::get<5>(g) : none, std::get<1>(g), std::get<2>(g),
                has(scales) ? std::get<6>(g) : none, has(rotations) ? std::get<7>(g) : none, std::get<4>(g),
                none, none, none, none,                                        // bg, viewmatrix, projmatrix, campos
                none, none, none, none, none, none, none, none, none, none, none};   // the eleven non-tensor arguments
    }
};

std::tuple<at::Tensor, at::Tensor, at::Tensor> rasterize(at::Tensor means3D, at::Tensor means2D, at::Tensor sh,
                                                         at::Tensor colors, at::Tensor opacity, at::Tensor scales,
                                                         at::Tensor rotations, at::Tensor cov3D, at::Tensor bg,
                                                         at::Tensor viewmatrix, at::Tensor projmatrix, at::Tensor campos,
                                                         int64_t ctx_ptr, double scale_modifier, double tan_fovx,
                                                         double tan_fovy, int64_t H, int64_t W, int64_t degree,
                                                         bool prefiltered, bool debug, int64_t pair_hint,
                                                         int64_t static_cap) {
    auto out = RasterizeFn::apply(means3D, means2D, sh, colors, opacity, scales, rotations, cov3D, bg, viewmatrix,
                                  projmatrix, campos, ctx_ptr, scale_modifier, tan_fovx, tan_fovy, H, W, degree, prefiltered,
                                  debug, pair_hint, static_cap);
    return {out[0], out[1], out[2]};
}

int64_t last_pairs(int64_t dev) { return dev >= 0 && dev < 64 ? g_last_pairs[dev] : 0; }
int64_t last_ticket(int64_t dev) { return dev >= 0 && dev < 64 ? g_last_ticket[dev] : -1; }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "torch binding of libgsraster_b200 (include/gsraster.h)";
    m.def("forward", &forward);
    m.def("backward", &backward);
    m.def("rasterize", &rasterize);
    m.def("last_pairs", &last_pairs);
    m.def("last_ticket", &last_ticket);
}
