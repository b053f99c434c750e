// C ABI of the B200 rasterizer (include/gsraster.h).  Host-side orchestration only: argument checks, scratch
// carving, kernel launches on the caller's stream, pinned status slots.  Mirrors the structure of the
// reference's CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (RAST/cuda_rasterizer/rasterizer_impl.cu:141-153,198-339,343-444) and of its torch binding
// (RAST/rasterize_points.cu:35-221) without any torch types.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/gsraster.h"
#include "gs_common.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define GS_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(GS_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));   \
    } while (0)

constexpr int kSlots = 64;
constexpr int kGraphSlots = 256;       // status slots handed out to forwards that are being stream-captured (never recycled)
constexpr int kGraphTicket = 0x40000000;
constexpr int kMaxViewStreams = 8;

int check_frame(const GsFrame* f) {
    if (!f) return fail(GS_EINVAL, "frame is NULL");
    if (f->P < 0 || f->W <= 0 || f->H <= 0) return fail(GS_EINVAL, "bad sizes P=%d W=%d H=%d", f->P, f->W, f->H);
    if (f->P == 0) return GS_OK;
    if (!f->means3D || !f->opacities || !f->viewmatrix || !f->projmatrix || !f->campos || !f->bg)
        return fail(GS_EINVAL, "means3D/opacities/viewmatrix/projmatrix/campos/bg must be non-NULL");
    // RAST/depth_diff_gaussian_rasterization_min/__init__.py:192-196
    if ((f->shs == nullptr) == (f->colors_precomp == nullptr))
        return fail(GS_EINVAL, "Please provide excatly one of either SHs or precomputed colors!");
    const bool has_sr = f->scales != nullptr && f->rotations != nullptr;
    if (((f->scales == nullptr || f->rotations == nullptr) && f->cov3D_precomp == nullptr) ||
        ((f->scales != nullptr || f->rotations != nullptr) && f->cov3D_precomp != nullptr))
        return fail(GS_EINVAL, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    (void)has_sr;
    // k_project / k_grad_vis / sh_to_rgb read rotations and SH rows with 128-bit loads
    if ((reinterpret_cast<uintptr_t>(f->rotations) & 15) || (reinterpret_cast<uintptr_t>(f->shs) & 15))
        return fail(GS_EINVAL, "rotations and shs must be 16-byte aligned (see gsraster.h)");
    if (f->shs && (f->M <= 0 || f->D < 0 || f->D > 3 || (f->D + 1) * (f->D + 1) > f->M))
        return fail(GS_EINVAL, "SH degree %d needs %d coefficients, M=%d", f->D, (f->D + 1) * (f->D + 1), f->M);
    return GS_OK;
}

GsView make_view(const GsFrame* f) {
    GsView v;
    v.vm = f->viewmatrix; v.pm = f->projmatrix; v.campos = f->campos; v.bg = f->bg;
    v.tan_fovx = f->tan_fovx; v.tan_fovy = f->tan_fovy;
    v.focal_y = f->H / (2.0f * f->tan_fovy);            // rasterizer_impl.cu:223-224
    v.focal_x = f->W / (2.0f * f->tan_fovx);
    v.scale_modifier = f->scale_modifier;
    v.W = f->W; v.H = f->H;
    v.gx = (f->W + GS_TILE - 1) / GS_TILE; v.gy = (f->H + GS_TILE - 1) / GS_TILE;
    v.P = f->P; v.D = f->D; v.M = f->shs ? f->M : 0;
    return v;
}

int debug_sync(const GsFrame* f, cudaStream_t s, const char* what) {
    if (!f->debug) {
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) return fail(GS_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
        return GS_OK;
    }
    cudaError_t e = cudaStreamSynchronize(s);           // CHECK_CUDA(., debug): auxiliary.h:166-173
    if (e != cudaSuccess) return fail(GS_ECUDA, "[CUDA ERROR] in %s: %s", what, cudaGetErrorString(e));
    return GS_OK;
}

}  // namespace

__global__ void k_clear_words(uint32_t* __restrict__ p, size_t words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
void gs_launch_clear_words(uint32_t* p, size_t words, cudaStream_t s) {
    if (!words) return;
    const int grid = (int)((words + 255) / 256 < 296 ? (words + 255) / 256 : 296);
    k_clear_words<<<grid, 256, 0, s>>>(p, words);
}

constexpr int kNumKernels = GS_NUM_KERNELS;
static const char* const kKernelNames[kNumKernels] = {"k_project", "k_tile_scan", "k_shade_emit", "k_tile_sort",
                                                      "k_tile_sort_big", "k_blend_fwd", "k_blend_bwd", "k_grad_vis",
                                                      "k_count_tiles", "k_grad_write",   // slot 9 also times
                                                      "k_grad_dense", "k_fill_zero"};    // k_grad_reduce_peers

struct GsContext {
    int device;
    int num_sms;
    int profile;                   // != 0: bracket every kernel with timing events (gs_profile_*)
    cudaEvent_t pev[2 * kNumKernels];
    bool pev_used[kNumKernels];
    GsDevStatus* slots;            // pinned, mapped: kSlots recycled slots, then kGraphSlots persistent ones
    std::atomic<int> next_graph_slot;
    std::atomic<long long> last_visible;   // num_visible of the most recent forward whose counts were read (-1: none yet)
    cudaEvent_t events[kSlots];
    unsigned slot_gen[kSlots];     // generation of the ticket that currently owns the slot
    std::atomic<unsigned> next;
    cudaStream_t aux;              // side stream of gs_backward_prefill (zero-fill beside the tile pass)
    cudaEvent_t aux_fork, aux_join;
    bool aux_ready;
    const void* prefilled;         // dL_dsh pointer of the outputs the most recent prefill was issued for
    bool peer_barrier_pending;     // the "every bucket is cleared" barrier of this step runs on `aux`
    cudaStream_t vstreams[kMaxViewStreams];   // internal streams of gs_forward_views (created on first use)
    cudaEvent_t vfork, vjoin[kMaxViewStreams];
    bool vstreams_ready;
};

// Brackets a launch with timing events on the launching stream when profiling is on (bench.py's roofline leg).
#define GS_TIMED(ctx, k, s, launch)                                   \
    do {                                                              \
        if ((ctx) && (ctx)->profile) cudaEventRecord((ctx)->pev[2 * (k)], s);     \
        launch;                                                       \
        if ((ctx) && (ctx)->profile) { cudaEventRecord((ctx)->pev[2 * (k) + 1], s); (ctx)->pev_used[k] = true; } \
    } while (0)

extern "C" {

int gs_profile_enable(GsContext* ctx, int on) {
    if (!ctx) return fail(GS_EINVAL, "ctx is NULL");
    ctx->profile = on;
    for (int i = 0; i < kNumKernels; i++) ctx->pev_used[i] = false;
    return GS_OK;
}
int gs_profile_num_kernels(void) { return kNumKernels; }
const char* gs_profile_kernel_name(int i) { return (i >= 0 && i < kNumKernels) ? kKernelNames[i] : ""; }
int gs_profile_read(GsContext* ctx, float* ms) {
    if (!ctx || !ms) return fail(GS_EINVAL, "NULL argument");
    for (int k = 0; k < kNumKernels; k++) {
        ms[k] = -1.f;
        if (!ctx->pev_used[k]) continue;
        GS_CUDA(cudaEventSynchronize(ctx->pev[2 * k + 1]));
        GS_CUDA(cudaEventElapsedTime(&ms[k], ctx->pev[2 * k], ctx->pev[2 * k + 1]));
    }
    return GS_OK;
}

int gs_abi_version(void) { return GS_ABI_VERSION; }
const char* gs_last_error(void) { return g_err; }

int gs_context_create(int device, GsContext** out) {
    if (!out) return fail(GS_EINVAL, "out is NULL");
    int prev = 0;
    GS_CUDA(cudaGetDevice(&prev));
    GS_CUDA(cudaSetDevice(device));
    GsContext* c = new GsContext();
    c->device = device;
    c->next = 0;
    c->next_graph_slot = 0;
    c->last_visible = -1;
    c->profile = 0;
    c->vstreams_ready = false;
    c->aux_ready = false;
    c->prefilled = nullptr;
    c->peer_barrier_pending = false;
    for (int i = 0; i < kSlots; i++) c->slot_gen[i] = 0;
    for (int i = 0; i < kNumKernels; i++) c->pev_used[i] = false;
    for (int i = 0; i < 2 * kNumKernels; i++) cudaEventCreate(&c->pev[i]);
    c->num_sms = 148;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    g_gs_num_sms = c->num_sms;
    gs_tile_sort_init();
    gs_grad_write_init();
    gs_preprocess_init();
    if (const char* e = getenv("GS_BLEND_VARIANT")) g_gs_blend_variant = atoi(e);
    cudaError_t e = cudaHostAlloc((void**)&c->slots, sizeof(GsDevStatus) * (kSlots + kGraphSlots),
                                cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {
        for (int j = 0; j < 2 * kNumKernels; j++) cudaEventDestroy(c->pev[j]);
        delete c;
        cudaSetDevice(prev);
        return fail(GS_ECUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
    }
    memset(c->slots, 0, sizeof(GsDevStatus) * (kSlots + kGraphSlots));
    for (int i = 0; i < kSlots; i++) {
        e = cudaEventCreateWithFlags(&c->events[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; j++) cudaEventDestroy(c->events[j]);
            for (int j = 0; j < 2 * kNumKernels; j++) cudaEventDestroy(c->pev[j]);
            cudaFreeHost(c->slots);
            delete c;
            cudaSetDevice(prev);
            return fail(GS_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
        }
    }
    cudaSetDevice(prev);
    *out = c;
    return GS_OK;
}

void gs_context_destroy(GsContext* c) {
    if (!c) return;
    for (int i = 0; i < kSlots; i++) cudaEventDestroy(c->events[i]);
    for (int i = 0; i < 2 * kNumKernels; i++) cudaEventDestroy(c->pev[i]);
    if (c->aux_ready) { cudaEventDestroy(c->aux_fork); cudaEventDestroy(c->aux_join); cudaStreamDestroy(c->aux); }
    if (c->vstreams_ready) {
        cudaEventDestroy(c->vfork);
        for (int i = 0; i < kMaxViewStreams; i++) { cudaEventDestroy(c->vjoin[i]); cudaStreamDestroy(c->vstreams[i]); }
    }
    cudaFreeHost(c->slots);
    delete c;
}

size_t gs_geom_bytes(int32_t P) { return gs_geom_layout(nullptr, P > 0 ? P : 1).bytes; }
size_t gs_image_bytes(int32_t W, int32_t H) { return gs_image_layout(nullptr, W, H).bytes; }
size_t gs_binning_bytes(int64_t cap) { return gs_bin_layout(nullptr, cap > 0 ? cap : 1).bytes; }

int gs_forward_preprocess(GsContext* ctx, const GsFrame* f, void* geom_buffer, void* image_buffer, int32_t* radii,
                          gs_stream_t stream, int32_t* ticket) {
    if (!ctx || !ticket) return fail(GS_EINVAL, "ctx/ticket is NULL");
    int rc = check_frame(f);
    if (rc) return rc;
    if (!image_buffer || (f->P > 0 && (!geom_buffer || !radii))) return fail(GS_EINVAL, "scratch/radii is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const GsView v = make_view(f);
    const int G = v.gx * v.gy;
    GsImageLayout il = gs_image_layout(image_buffer, f->W, f->H);
    // A forward that is being captured into a CUDA graph gets a PERSISTENT status slot (the graph bakes its address in
    // and rewrites it on every replay); no event is involved, nothing may block.  Eager forwards recycle kSlots slots.
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    GS_CUDA(cudaStreamIsCapturing(s, &cap_st));
    const bool capturing = cap_st != cudaStreamCaptureStatusNone;
    int slot = -1;
    GsDevStatus* host_slot = nullptr;
    if (capturing) {
        const int gsl = ctx->next_graph_slot.fetch_add(1);
        if (gsl >= kGraphSlots) return fail(GS_ENOMEM, "more than %d forwards captured into CUDA graphs on this context", kGraphSlots);
        host_slot = ctx->slots + kSlots + gsl;
        *ticket = kGraphTicket | gsl;
    } else {
        const unsigned seq = ctx->next.fetch_add(1u);
        slot = (int)(seq % (unsigned)kSlots);
        const unsigned gen = (seq / (unsigned)kSlots) & 0x00ffffffu;
        host_slot = ctx->slots + slot;
        // the slot's previous owner (64 forwards ago) must have written its status before the slot is cleared
        if (seq >= (unsigned)kSlots) GS_CUDA(cudaEventSynchronize(ctx->events[slot]));
        ctx->slot_gen[slot] = gen;
        *ticket = (int32_t)((gen << 6) | (unsigned)slot);     // slot in the low 6 bits, generation above
    }
    host_slot->overflow = 0;                             // cleared; kernel writes 0xC0FFEE when done
    // zero tile histogram + status in one memset (they are adjacent)
    gs_launch_clear_words(il.tile_cnt, ((size_t)((char*)il.status - (char*)il.tile_cnt) + sizeof(GsDevStatus)) / 4, s);
    GsDevStatus* dev_slot = nullptr;
    GS_CUDA(cudaHostGetDevicePointer((void**)&dev_slot, host_slot, 0));
    bool scanned = false;
    if (f->P > 0) {
        GsGeomLayout gl = gs_geom_layout(geom_buffer, f->P);
        GS_TIMED(ctx, 0, s, gs_launch_project(v, f->means3D, f->opacities, f->scales, f->rotations, f->cov3D_precomp,
                                              radii, gl.rec, gl.vis_list, il.status,
                                              2 * ctx->last_visible.load() > (long long)f->P, s));
        if ((rc = debug_sync(f, s, "project"))) return rc;
        // the tile histogram; its last CTA also scans it and mirrors the totals to the host slot
        GS_TIMED(ctx, 8, s, gs_launch_count_tiles(v, ctx->num_sms, radii, gl.rec, gl.vis_list, gl.hitmask, il.tile_cnt, il.status,
                                                  il.tile_off, dev_slot, s));
        if ((rc = debug_sync(f, s, "count_tiles"))) return rc;
        scanned = gs_scan_folded();
    }
    if (!scanned) {
        GS_TIMED(ctx, 1, s, gs_launch_tile_scan(G, il.tile_cnt, il.tile_off, il.status, dev_slot, s));
        if ((rc = debug_sync(f, s, "tile_scan"))) return rc;
    }
    if (!capturing) GS_CUDA(cudaEventRecord(ctx->events[slot], s));
    return GS_OK;
}

static int read_slot(GsContext* ctx, const volatile GsDevStatus* h, int slot, GsCounts* out, bool must_be_ready) {
    if (h->overflow != 0xC0FFEEu) {
        if (must_be_ready) return fail(GS_ECUDA, "status slot %d was not written by the device", slot);
        return GS_ENOTREADY;
    }
    out->num_rendered = (int64_t)h->num_rendered;
    out->num_pairs = (int64_t)h->num_pairs;
    out->num_visible = (int64_t)h->num_visible;
    ctx->last_visible = out->num_visible;
    return GS_OK;
}

int gs_forward_counts(GsContext* ctx, int32_t ticket, GsCounts* out) {
    if (!ctx || !out || ticket < 0) return fail(GS_EINVAL, "bad ctx/ticket/out");
    if (ticket & kGraphTicket)
        return fail(GS_EINVAL, "ticket of a captured forward: replay the graph, synchronise, then gs_forward_counts_peek");
    const int slot = ticket & (kSlots - 1);
    const unsigned gen = (unsigned)ticket >> 6;
    if (ctx->slot_gen[slot] != gen)
        return fail(GS_EINVAL, "ticket expired: more than %d forwards were enqueued on this context since", kSlots);
    GS_CUDA(cudaEventSynchronize(ctx->events[slot]));
    if (ctx->slot_gen[slot] != gen)
        return fail(GS_EINVAL, "ticket expired: more than %d forwards were enqueued on this context since", kSlots);
    return read_slot(ctx, ctx->slots + slot, slot, out, true);
}

int gs_forward_counts_peek(GsContext* ctx, int32_t ticket, GsCounts* out) {
    if (!ctx || !out || ticket < 0) return fail(GS_EINVAL, "bad ctx/ticket/out");
    if (ticket & kGraphTicket) {
        const int gsl = ticket & (kGraphTicket - 1);
        if (gsl >= kGraphSlots) return fail(GS_EINVAL, "bad graph ticket");
        return read_slot(ctx, ctx->slots + kSlots + gsl, kSlots + gsl, out, false);
    }
    const int slot = ticket & (kSlots - 1);
    const unsigned gen = (unsigned)ticket >> 6;
    if (ctx->slot_gen[slot] != gen)
        return fail(GS_EINVAL, "ticket expired: more than %d forwards were enqueued on this context since", kSlots);
    return read_slot(ctx, ctx->slots + slot, slot, out, false);
}

int gs_forward_render(GsContext* ctx, const GsFrame* f, const int32_t* radii, void* geom_buffer, void* binning_buffer,
                      int64_t pair_capacity, void* image_buffer, float* out_color, float* out_depth,
                      int32_t rerender, gs_stream_t stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    if (!image_buffer || !out_color || !out_depth) return fail(GS_EINVAL, "image buffer / outputs are NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t N = (size_t)f->W * f->H;
    if (f->P == 0) {                                     // rasterize_points.cu:68-70,82: zero images, no kernels
        GS_CUDA(cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), s));
        GS_CUDA(cudaMemsetAsync(out_depth, 0, N * sizeof(float), s));
        return GS_OK;
    }
    if (!geom_buffer || !radii || !binning_buffer || pair_capacity < 0) return fail(GS_EINVAL, "scratch/radii is NULL");
    const GsView v = make_view(f);
    const int G = v.gx * v.gy;
    GsImageLayout il = gs_image_layout(image_buffer, f->W, f->H);
    GsGeomLayout gl = gs_geom_layout(geom_buffer, f->P);
    GsBinLayout bl = gs_bin_layout(binning_buffer, pair_capacity > 0 ? pair_capacity : 1);
    GS_TIMED(ctx, 2, s, gs_launch_shade_emit(v, ctx ? ctx->num_sms : 148, f->means3D, f->shs, f->colors_precomp, radii,
                                             gl.rec, gl.acc, gl.vis_list, gl.hitmask, il.tile_off, il.tile_cnt, il.status,
                                             bl.keys, pair_capacity, rerender != 0,
                                             ctx && 2 * ctx->last_visible.load() > (long long)f->P, s));
    if ((rc = debug_sync(f, s, "emit"))) return rc;
    {
        const bool prof = ctx && ctx->profile;
        cudaEvent_t* pe = prof ? ctx->pev + 2 * 3 : nullptr;   // [start sort, end sort, start big, end big]
        gs_launch_tile_sort(G, ctx ? ctx->num_sms : 148, il.tile_off, il.tile_cnt, il.status, il.big_list, bl.keys,
                            bl.list, pair_capacity, s, pe);
        if (prof) ctx->pev_used[3] = ctx->pev_used[4] = true;
    }
    if ((rc = debug_sync(f, s, "tile_sort"))) return rc;
    GS_TIMED(ctx, 5, s, gs_launch_blend_fwd(v, il.tile_off, bl.list, gl.rec, il.status, pair_capacity, il.final_T,
                                                il.n_contrib, out_color, out_depth, s));
    if ((rc = debug_sync(f, s, "blend_fwd"))) return rc;
    return GS_OK;
}

int gs_forward_views(GsContext* ctx, const GsFrame* frames, int32_t n_views, const GsViewScratch* scratch,
                     int32_t n_streams, GsViewResult* results, gs_stream_t stream) {
    if (!ctx || !frames || !scratch || !results) return fail(GS_EINVAL, "NULL argument");
    if (n_views < 0 || n_views > kSlots) return fail(GS_EINVAL, "n_views must be in [0, %d]", kSlots);
    if (n_streams < 1 || n_streams > kMaxViewStreams) return fail(GS_EINVAL, "n_streams must be in [1, %d]", kMaxViewStreams);
    if (n_views == 0) return GS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int prev = 0;
    GS_CUDA(cudaGetDevice(&prev));
    GS_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->vstreams_ready) {
        GS_CUDA(cudaEventCreateWithFlags(&ctx->vfork, cudaEventDisableTiming));
        for (int i = 0; i < kMaxViewStreams; i++) {
            GS_CUDA(cudaStreamCreateWithFlags(&ctx->vstreams[i], cudaStreamNonBlocking));
            GS_CUDA(cudaEventCreateWithFlags(&ctx->vjoin[i], cudaEventDisableTiming));
        }
        ctx->vstreams_ready = true;
    }
    const int ns = n_streams < n_views ? n_streams : n_views;
    GS_CUDA(cudaEventRecord(ctx->vfork, s));                       // fork: the views start after the caller's work
    for (int i = 0; i < ns; i++) GS_CUDA(cudaStreamWaitEvent(ctx->vstreams[i], ctx->vfork, 0));
    int32_t tickets[kSlots];
    int rc = GS_OK;
    for (int k = 0; k < n_views && rc == GS_OK; k++) {
        const GsViewScratch& sc = scratch[k % ns];
        GsViewResult& r = results[k];
        cudaStream_t vs = ctx->vstreams[k % ns];
        int32_t* radii = r.radii ? r.radii : sc.radii;
        r.status = GS_OK;
        rc = gs_forward_preprocess(ctx, &frames[k], sc.geom_buffer, sc.image_buffer, radii, vs, &tickets[k]);
        if (rc == GS_OK)       // speculative: the device-side capacity guard skips the render if the buffer is too small
            rc = gs_forward_render(ctx, &frames[k], radii, sc.geom_buffer, sc.binning_buffer, sc.pair_capacity,
                                   sc.image_buffer, r.out_color, r.out_depth, 0, vs);
    }
    for (int i = 0; i < ns; i++) {                                 // join, also on the error path
        cudaEventRecord(ctx->vjoin[i], ctx->vstreams[i]);
        cudaStreamWaitEvent(s, ctx->vjoin[i], 0);
    }
    cudaSetDevice(prev);
    if (rc != GS_OK) return rc;
    for (int k = 0; k < n_views; k++) {                            // only now does the host wait, once per view
        const int rc2 = gs_forward_counts(ctx, tickets[k], &results[k].counts);
        if (rc2 != GS_OK) return rc2;
        if (results[k].counts.num_pairs > scratch[k % ns].pair_capacity) { results[k].status = GS_ECAPACITY; rc = GS_ECAPACITY; }
    }
    if (rc == GS_ECAPACITY) fail(GS_ECAPACITY, "at least one view needs a larger binning buffer (see results[].status)");
    return rc;
}

size_t gs_backward_scratch_bytes(int64_t num_visible) {
    return gs_align_up((size_t)(num_visible > 0 ? num_visible : 1) * GS_GOUT_FLOATS * sizeof(float), 256);
}

static GsGradPtrs grad_ptrs(const GsFrame* f, const GsGrads* grads) {
    GsGradPtrs g;
    g.dmeans3D = grads->dL_dmeans3D; g.dmeans2D = grads->dL_dmeans2D; g.dsh = f->shs ? grads->dL_dsh : nullptr;
    g.dcolors = grads->dL_dcolors; g.dopacity = grads->dL_dopacity;
    g.dscales = grads->dL_dscales; g.drots = grads->dL_drotations; g.dcov3D = grads->dL_dcov3D;
    return g;
}

// the dense-regime kernel (k_grad_dense) covers the reference's own input mode: SH degree <= 3 stored as 16 coefficients,
// scales + rotations, all six gradients wanted; its bulk copies need 16-byte aligned bases (gsraster.h)
static bool dense_eligible(const GsFrame* f, const GsGrads* grads, const GsGradPtrs& g) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return grads->peer_world <= 0 && f->shs && f->M == 16 && !f->cov3D_precomp && f->scales && f->rotations &&
           !grads->dL_dcolors && !grads->dL_dcov3D && g.dmeans3D && g.dmeans2D && g.dsh && g.dopacity && g.dscales &&
           g.drots && al16(f->means3D) && al16(f->scales) && al16(g.dmeans3D) && al16(g.dmeans2D) && al16(g.dsh) &&
           al16(g.dopacity) && al16(g.dscales) && al16(g.drots) && !getenv("GS_NO_DENSE");
}

static int ensure_aux(GsContext* ctx) {
    if (ctx->aux_ready) return GS_OK;
    int prev = 0;
    GS_CUDA(cudaGetDevice(&prev));
    GS_CUDA(cudaSetDevice(ctx->device));
    // highest priority: the block scheduler hands freed SM resources to the side stream's small CTAs first; at equal
    // priority they would wait until the tile pass (8 160 CTAs, launched into the caller's stream at the same moment)
    // has nothing left to dispatch, i.e. run AFTER it (measured: +0.12 ms per step in the eager loop)
    int pr_least = 0, pr_greatest = 0;
    GS_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
    GS_CUDA(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, pr_greatest));
    GS_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
    GS_CUDA(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
    GS_CUDA(cudaSetDevice(prev));
    ctx->aux_ready = true;
    return GS_OK;
}

// Optional, before gs_backward_blend: zero-fills the dense gradient outputs on a side stream of the context, BESIDE the
// tile pass (which is issue bound and leaves the memory system idle).  gs_backward_gradients then waits for the fill and
// lets the per-Gaussian kernel write the visible rows straight into the outputs -- the streaming writer of the sparse
// regime disappears from the critical path.  Returns GS_OK without doing anything when the outputs do not qualify
// (other input modes, misaligned bases, peer reduce): gs_backward_gradients falls back to its own writer.
int gs_backward_prefill(GsContext* ctx, const GsFrame* f, const void* image_buffer, const GsGrads* grads, gs_stream_t stream) {
    if (!ctx || !grads) return fail(GS_EINVAL, "ctx / grads is NULL");
    int rc = check_frame(f);
    if (rc) return rc;
    ctx->prefilled = nullptr;
    if (f->P == 0 || !image_buffer || grads->peer_world > 0 || (f->P & 3) || f->cov3D_precomp || !f->scales || !f->rotations ||
        getenv("GS_NO_PREFILL"))
        return GS_OK;
    const GsGradPtrs g = grad_ptrs(f, grads);
    if (!gs_grads_tma_ok(f->shs ? f->M : 0, g)) return GS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_aux(ctx))) return rc;
    GsImageLayout il = gs_image_layout(const_cast<void*>(image_buffer), f->W, f->H);
    GS_CUDA(cudaEventRecord(ctx->aux_fork, s));            // the outputs were allocated in `stream` order before this call
    GS_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->aux_fork, 0));
    // in the dense regime k_grad_dense writes every row itself: the fill exits on the device (only if that kernel will run)
    GS_TIMED(ctx, 11, ctx->aux, gs_launch_fill_zero(f->P, ctx->num_sms, g, il.status, dense_eligible(f, grads, g), ctx->aux));
    GS_CUDA(cudaEventRecord(ctx->aux_join, ctx->aux));
    ctx->prefilled = grads->dL_dsh;
    return GS_OK;
}

// Backward, first half: reverse tile traversal into the per-Gaussian accumulators.  Needs no gradient outputs,
// so a host can enqueue it before it has allocated them.
int gs_backward_blend(GsContext* ctx, const GsFrame* f, const void* geom_buffer, const void* binning_buffer,
                      int64_t pair_capacity, const void* image_buffer, const float* dL_dout_color,
                      gs_stream_t stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    if (f->P == 0) return GS_OK;
    if (!geom_buffer || !binning_buffer || !image_buffer || !dL_dout_color)
        return fail(GS_EINVAL, "scratch / dL_dout_color is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const GsView v = make_view(f);
    GsImageLayout il = gs_image_layout(const_cast<void*>(image_buffer), f->W, f->H);
    GsGeomLayout gl = gs_geom_layout(const_cast<void*>(geom_buffer), f->P);
    GsBinLayout bl = gs_bin_layout(const_cast<void*>(binning_buffer), pair_capacity > 0 ? pair_capacity : 1);
    GS_TIMED(ctx, 6, s, gs_launch_blend_bwd(v, il.tile_off, bl.list, gl.rec, il.final_T, il.n_contrib, dL_dout_color,
                                            gl.acc, il.status, pair_capacity, s));
    return debug_sync(f, s, "blend_bwd");
}

// Backward, second half: per-Gaussian gradients from the accumulators into the caller's tensors (or peers' buckets).
int gs_backward_gradients(GsContext* ctx, const GsFrame* f, const int32_t* radii, const void* geom_buffer,
                          const void* image_buffer, void* grad_scratch, size_t grad_scratch_bytes,
                          const GsGrads* grads, gs_stream_t stream) {
    int rc = check_frame(f);
    if (rc) return rc;
    if (!grads) return fail(GS_EINVAL, "grads is NULL");
    if (f->P == 0) return GS_OK;
    if (!radii || !geom_buffer || !image_buffer || !grad_scratch) return fail(GS_EINVAL, "radii / scratch is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    const GsView v = make_view(f);
    GsImageLayout il = gs_image_layout(const_cast<void*>(image_buffer), f->W, f->H);
    GsGeomLayout gl = gs_geom_layout(const_cast<void*>(geom_buffer), f->P);
    const GsGradPtrs g = grad_ptrs(f, grads);
    float* gout = (float*)grad_scratch;
    // gs_backward_prefill ran for these outputs: join the side stream, the per-Gaussian kernel scatters final rows
    const bool prefilled = ctx && ctx->prefilled && ctx->prefilled == (const void*)grads->dL_dsh && grads->peer_world <= 0;
    if (prefilled) {
        GS_CUDA(cudaStreamWaitEvent(s, ctx->aux_join, 0));
        ctx->prefilled = nullptr;
    }
    // When most Gaussians are visible (decided on the device from num_visible) ONE dense kernel does the whole
    // per-Gaussian backward and the compact pair below returns at once; otherwise the other way round.  The dense
    // kernel covers the reference's own input mode (SH degree <= 3 stored as 16 coefficients, scales + rotations).
    const bool dense_ok = dense_eligible(f, grads, g);
    GS_TIMED(ctx, 7, s, gs_launch_grad_vis(v, ctx ? ctx->num_sms : 148, f->means3D, f->shs,
                                           f->cov3D_precomp ? nullptr : f->scales,
                                           f->cov3D_precomp ? nullptr : f->rotations, f->cov3D_precomp, gl.rec, gl.acc,
                                           gl.vis_list, il.status, gout, dense_ok, prefilled, g, s));
    if ((rc = debug_sync(f, s, "grad_vis"))) return rc;
    if (grads->peer_world > 0) {
        // data-parallel shared-model step: the five parameter gradients are reduced straight into every rank's
        // bucket (peer stores / NVSwitch multicast); only dL_dmeans2D is written locally
        if (grads->peer_world > GS_MAX_PEERS || !grads->peer_buckets) return fail(GS_EINVAL, "bad peer arguments");
        const bool sync = ctx && grads->peer_signals;
        if (sync && grads->peer_epoch_begin) {
            // nobody adds into a bucket before every rank has cleared its own: one-warp barrier kernel, normally already
            // done -- gs_backward put it on the side stream beside the tile pass
            if (ctx->peer_barrier_pending) GS_CUDA(cudaStreamWaitEvent(s, ctx->aux_join, 0));
            else gs_launch_peer_barrier((uint32_t* const*)grads->peer_signals, grads->peer_world, grads->peer_rank, 0,
                                        grads->peer_epoch_begin, s);
            ctx->peer_barrier_pending = false;
        }
        GS_TIMED(ctx, 9, s, gs_launch_grad_reduce_peers(f->P, v.M, radii, gl.acc, gout, grads->dL_dmeans2D,
                                                        (float* const*)grads->peer_buckets, grads->peer_world,
                                                        (float*)grads->peer_multicast, (const long long*)grads->peer_seg_off, s));
        // the step's result is complete on this rank when every rank's adds have landed
        if (sync && grads->peer_epoch_end)
            gs_launch_peer_barrier((uint32_t* const*)grads->peer_signals, grads->peer_world, grads->peer_rank, 1,
                                   grads->peer_epoch_end, s);
        return debug_sync(f, s, "grad_reduce_peers");
    }
    if (!prefilled) {
        GS_TIMED(ctx, 9, s, gs_launch_grad_write(f->P, v.M, radii, gl.acc, gout, g, il.status, dense_ok, s));
        if ((rc = debug_sync(f, s, "grad_write"))) return rc;
    }
    // a scratch buffer sized from the true visible count (synchronising hosts) can prove the sparse regime: no launch
    const long long nvis_bound = (long long)(grad_scratch_bytes / (GS_GOUT_FLOATS * sizeof(float)));
    if (dense_ok && 2 * nvis_bound > (long long)f->P) {
        GS_TIMED(ctx, 10, s, gs_launch_grad_dense(v, ctx ? ctx->num_sms : 148, f->means3D, f->shs, f->scales, f->rotations,
                                                  radii, gl.acc, il.status, g, s));
        if ((rc = debug_sync(f, s, "grad_dense"))) return rc;
    }
    // outputs the fused kernel does not produce in this input mode are defined as zeros (reference: torch::zeros)
    const size_t Ps = (size_t)f->P;
    if (!f->shs && grads->dL_dsh && f->M > 0) GS_CUDA(cudaMemsetAsync(grads->dL_dsh, 0, Ps * f->M * 3 * sizeof(float), s));
    return GS_OK;
}

int gs_backward(GsContext* ctx, const GsFrame* f, const int32_t* radii, const void* geom_buffer,
                const void* binning_buffer, int64_t pair_capacity, const void* image_buffer, void* grad_scratch,
                size_t grad_scratch_bytes, const float* dL_dout_color, const float* dL_dout_depth,
                const GsGrads* grads, gs_stream_t stream) {
    (void)dL_dout_depth;                                 // depth gradient disabled in the reference
    if (!grads) return fail(GS_EINVAL, "grads is NULL");
    int rc = ctx ? gs_backward_prefill(ctx, f, image_buffer, grads, stream) : GS_OK;
    if (rc) return rc;
    if (ctx && grads->peer_world > 0 && grads->peer_signals && grads->peer_epoch_begin) {
        // shared-model step: the "every bucket is cleared" barrier waits on the side stream while the tile pass runs
        if ((rc = ensure_aux(ctx))) return rc;
        cudaStream_t s = (cudaStream_t)stream;
        GS_CUDA(cudaEventRecord(ctx->aux_fork, s));
        GS_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->aux_fork, 0));
        gs_launch_peer_barrier((uint32_t* const*)grads->peer_signals, grads->peer_world, grads->peer_rank, 0,
                               grads->peer_epoch_begin, ctx->aux);
        GS_CUDA(cudaEventRecord(ctx->aux_join, ctx->aux));
        ctx->peer_barrier_pending = true;
    }
    rc = gs_backward_blend(ctx, f, geom_buffer, binning_buffer, pair_capacity, image_buffer, dL_dout_color, stream);
    if (rc) return rc;
    return gs_backward_gradients(ctx, f, radii, geom_buffer, image_buffer, grad_scratch, grad_scratch_bytes, grads, stream);
}

int gs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, gs_stream_t stream) {
    (void)projmatrix;
    if (P < 0) return fail(GS_EINVAL, "P < 0");
    if (P == 0) return GS_OK;
    if (!means3D || !viewmatrix || !present) return fail(GS_EINVAL, "NULL argument");
    gs_launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "mark_visible launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_l1_loss_backward(GsContext* ctx, const float* color, const uint8_t* target_u8, int32_t H, int32_t W, float weight,
                        float* dL_dcolor, float* loss, gs_stream_t stream) {
    if (!color || !target_u8 || !dL_dcolor || !loss || H <= 0 || W <= 0) return fail(GS_EINVAL, "bad argument");
    gs_launch_l1_loss_grad(color, target_u8, H, W, weight, dL_dcolor, loss, ctx ? ctx->num_sms : 148, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "l1_loss launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_gaussian_adam_step(GsContext* ctx, const GsAdamGroup* groups, int32_t ngroups, double beta1, double beta2,
                          double eps, int32_t step, gs_stream_t stream) {
    (void)ctx;
    if (!groups || ngroups <= 0 || ngroups > 8 || step < 1) return fail(GS_EINVAL, "bad optimiser arguments");
    GsAdamArgs a;
    a.nseg = ngroups; a.om_beta1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.om_beta2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    for (int k = 0; k < ngroups; k++) {
        const GsAdamGroup& g = groups[k];
        if (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq || g.rows < 0 || g.activation < 0 || g.activation > 3)
            return fail(GS_EINVAL, "optimiser group %d is malformed", k);
        if ((reinterpret_cast<uintptr_t>(g.param) | reinterpret_cast<uintptr_t>(g.exp_avg) |
             reinterpret_cast<uintptr_t>(g.exp_avg_sq)) & 15)
            return fail(GS_EINVAL, "optimiser group %d: param / moments must be 16-byte aligned", k);
        GsAdamSeg& s = a.seg[k];
        s.type = g.activation; s.row_w = g.activation == 3 ? 4 : g.row_width; s.g_row_w = g.grad_row_width;
        s.g_off = g.grad_offset; s.rows = g.rows; s.p = g.param; s.g = g.grad; s.m = g.exp_avg; s.v = g.exp_avg_sq;
        s.step_size = (float)(g.lr / bc1); s.blocks = 0;
        if (s.row_w <= 0 || s.g_row_w < s.row_w + s.g_off) return fail(GS_EINVAL, "optimiser group %d: bad row widths", k);
    }
    if (gs_launch_gaussian_adam(a, (cudaStream_t)stream) != 0) return fail(GS_EINVAL, "optimiser problem too large");
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "adam launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_pack_frame(GsContext* ctx, int32_t H, int32_t W, const float* color, const float* depth, uint8_t* rgb8,
                  float* neg_depth, uint32_t* minmax_state, gs_stream_t stream) {
    (void)ctx;
    if (H < 0 || W < 0 || (H * (int64_t)W > 0 && (!color || !rgb8 || (depth && !neg_depth))))
        return fail(GS_EINVAL, "bad argument");
    if ((int64_t)H * W > 0x7fffffffLL / 4) return fail(GS_EINVAL, "image too large");
    gs_launch_pack_frame(H * W, color, depth, rgb8, neg_depth, minmax_state, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "pack frame launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_minmax_read(GsContext* ctx, uint32_t* minmax_state, float* minmax2, gs_stream_t stream) {
    (void)ctx;
    if (!minmax_state || !minmax2) return fail(GS_EINVAL, "NULL argument");
    gs_launch_minmax_decode(minmax_state, minmax2, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "minmax decode launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

size_t gs_knn_scratch_bytes(int32_t P) { return P > 0 ? gs_knn_scratch_bytes_impl(P) : 256; }

int gs_knn_mean_dist2(GsContext* ctx, int32_t P, const float* points, void* scratch, float* mean_dist2, gs_stream_t stream) {
    (void)ctx;
    if (P < 0 || (P > 0 && (!points || !scratch || !mean_dist2))) return fail(GS_EINVAL, "bad argument");
    const int rc = gs_launch_knn(P, points, scratch, mean_dist2, (cudaStream_t)stream);
    if (rc == -1) return fail(GS_EINVAL, "knn: radix-sort temporary storage exceeds the scratch allowance");
    cudaError_t e = cudaPeekAtLastError();
    if (rc != 0 || e != cudaSuccess) return fail(GS_ECUDA, "knn launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_densify_stats(GsContext* ctx, int32_t P, const int32_t* radii, const float* dL_dmeans2D, float* xyz_gradient_accum,
                     float* denom, float* max_radii2D, gs_stream_t stream) {
    (void)ctx;
    if (P < 0 || (P > 0 && (!radii || !dL_dmeans2D || !xyz_gradient_accum || !denom || !max_radii2D)))
        return fail(GS_EINVAL, "bad argument");
    gs_launch_densify_stats(P, radii, dL_dmeans2D, xyz_gradient_accum, denom, max_radii2D, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "densify stats launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

size_t gs_photometric_scratch_bytes(int32_t H, int32_t W) { return 256 + (size_t)9 * H * W * sizeof(float); }

int gs_photometric_loss_backward(GsContext* ctx, const float* image, const float* gt, int32_t H, int32_t W,
                                 float lambda_dssim, void* scratch, float* dL_dimage, float* loss3, gs_stream_t stream) {
    (void)ctx;
    if (!image || !gt || !scratch || !loss3 || H <= 0 || W <= 0) return fail(GS_EINVAL, "bad argument");
    gs_launch_photometric(image, gt, H, W, lambda_dssim, scratch, dL_dimage, loss3, (cudaStream_t)stream);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(GS_ECUDA, "photometric loss launch: %s", cudaGetErrorString(e));
    return GS_OK;
}

int gs_debug_export_binning(const GsFrame* f, const void* binning_buffer, int64_t pair_capacity,
                            const void* image_buffer, uint32_t* ranges, uint32_t* point_list, int64_t max_pairs,
                            gs_stream_t stream) {
    if (!f || !binning_buffer || !image_buffer) return fail(GS_EINVAL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    GsImageLayout il = gs_image_layout(const_cast<void*>(image_buffer), f->W, f->H);
    GsBinLayout bl = gs_bin_layout(const_cast<void*>(binning_buffer), pair_capacity > 0 ? pair_capacity : 1);
    const int G = ((f->W + GS_TILE - 1) / GS_TILE) * ((f->H + GS_TILE - 1) / GS_TILE);
    if (ranges) GS_CUDA(cudaMemcpyAsync(ranges, il.tile_off, (size_t)(G + 1) * 4, cudaMemcpyDeviceToDevice, s));
    if (point_list && max_pairs > 0)
        GS_CUDA(cudaMemcpyAsync(point_list, bl.list, (size_t)max_pairs * 4, cudaMemcpyDeviceToDevice, s));
    return GS_OK;
}

}  // extern "C"
