"""Drop-in host side of the B200 rasterizer: the reference's Python operator API, name for name.

Mirrors RAST/depth_diff_gaussian_rasterization_min/__init__.py (RAST = submodules/depth-diff-gaussian-
rasterization-min of LucidDreamer):

    GaussianRasterizationSettings   NamedTuple, 12 fields in the reference order      (__init__.py:158-170)
    GaussianRasterizer              nn.Module: forward(...) -> (color, radii, depth), markVisible(...)  (:172-221)
    rasterize_gaussians             functional form                                   (:21-42)
    _RasterizeGaussians             torch.autograd.Function, 8 gradients + None       (:44-156)
    _C                              rasterize_gaussians / rasterize_gaussians_backward / mark_visible with the
                                    signatures of the reference's pybind module       (ext.cpp:15-19)

so `gaussian_renderer.render()` (gaussian_renderer/__init__.py:18-104) and `scene.GaussianModel` run unchanged.
Everything below the argument checks is the C ABI in include/gsraster.h (hand-written sm_100a CUDA); PyTorch is
used for device memory, streams and autograd only.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _native as N

# ---------------------------------------------------------------------------------------------- plumbing

_contexts: dict = {}      # device index -> GsContext*
_cap_hint: dict = {}      # device index -> decayed maximum of the (tile, Gaussian) pair counts seen
_last_pairs: dict = {}    # device index -> pair count of the most recent forward (statistics only)
_HINT_DECAY = 0.97        # per forward; a training loop that draws a random camera per iteration (luciddreamer.py:
                          # 291-292) keeps the capacity of its largest view instead of re-rendering whenever a view
                          # has > 25 % more pairs than the previous one


def _note_pairs(idx: int, pairs: int) -> None:
    prev = _cap_hint.get(idx, 0)
    _cap_hint[idx] = max(int(pairs), int(prev * _HINT_DECAY))
    _last_pairs[idx] = int(pairs)


# ---- sync-free ("static capacity") mode ------------------------------------------------------------------------------
# The reference blocks the host once per forward (cudaMemcpy D2H of num_rendered, rasterizer_impl.cu:282); the default
# path here waits once too (an event, for the pair count that sizes the binning buffer).  With a STATIC pair capacity
# the forward is pure stream work: nothing waits, the call can be captured into a CUDA graph (graphs.GraphedStep), and
# the host runs ahead of the GPU.  The price is a deferred check: a frame with more (tile, Gaussian) pairs than the
# capacity renders NOTHING (device-side guard) and the error surfaces at the next rasterizer call / check_static().
_static: dict = {}        # device index -> {"cap": int, "pending": deque[(ticket, cap)], "graph": [(ticket, cap)]}
_last_counts: dict = {}   # device index -> dict(num_rendered, num_pairs, num_visible) of the last CHECKED static forward


def set_static_capacity(pairs: Optional[int], device=None) -> None:
    """pairs = N: every forward on `device` renders with room for exactly N (tile, Gaussian) pairs and never waits for
    the GPU; pairs = None: back to the default (one event wait per forward, buffer sized from the frame's own count)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if pairs is None:
        if idx in _static:
            _static[idx]["cap"] = None            # forwards already issued stay on the list until they are checked
        return
    st = _static.get(idx)
    if st is None:
        from collections import deque
        st = _static[idx] = dict(cap=None, pending=deque(), graph=[])
    st["cap"] = _round_cap(int(pairs))


class static_capacity:
    """`with static_capacity(N): ...` -- see set_static_capacity; restores the previous mode and checks on exit."""

    def __init__(self, pairs: int, device=None):
        self.pairs, self.device = pairs, device

    def __enter__(self):
        idx = torch.cuda.current_device() if self.device is None else torch.device(self.device).index
        self.idx = idx
        self.prev = (_static.get(idx) or {}).get("cap")
        set_static_capacity(self.pairs, idx)
        return self

    def __exit__(self, *a):
        set_static_capacity(self.prev, self.idx)
        return False


def _static_cap(idx: int) -> int:
    st = _static.get(idx)
    return -1 if st is None or st["cap"] is None else st["cap"]


def _check_ticket(idx: int, ticket: int, cap: int, wait: bool) -> bool:
    """True when the ticket has been checked (and was fine); False when the device has not written it yet."""
    L = N.lib()
    counts = N.GsCounts()
    rc = L.gs_forward_counts_peek(_ctx(idx), int(ticket), C.byref(counts))
    if rc == N.GS_ENOTREADY and wait and not (int(ticket) & 0x40000000):
        rc = L.gs_forward_counts(_ctx(idx), int(ticket), C.byref(counts))
    if rc == N.GS_ENOTREADY:
        return False
    N.check(rc)
    _last_counts[idx] = dict(num_rendered=int(counts.num_rendered), num_pairs=int(counts.num_pairs),
                             num_visible=int(counts.num_visible))
    _note_pairs(idx, counts.num_pairs)
    if counts.num_pairs > cap:
        raise RuntimeError(
            f"luciddreamer_b200: a forward rendered with static pair capacity {cap} had {int(counts.num_pairs)} (tile, "
            "Gaussian) pairs; its outputs (and everything computed from them) are invalid. Raise the capacity "
            "(set_static_capacity / GraphedStep(pair_capacity=...)) or use the default synchronising mode.")
    return True


def _poll_static(idx: int, wait: bool = False) -> None:
    st = _static.get(idx)
    if st is None:
        return
    pend = st["pending"]
    while pend:
        ticket, cap = pend[0]
        try:
            if not _check_ticket(idx, ticket, cap, wait or len(pend) > 48):   # 64 status slots per context: never lap them
                break
        except RuntimeError:
            pend.popleft()                # an overflow is reported once
            raise
        pend.popleft()


def _note_static(idx: int, ticket: int, cap: int) -> None:
    st = _static[idx]
    if int(ticket) & 0x40000000:          # captured into a CUDA graph: the slot is rewritten by every replay
        st["graph"].append((int(ticket), cap))
    else:
        st["pending"].append((int(ticket), cap))


def check_static(device=None, synchronize: bool = True) -> dict:
    """Checks every static-capacity forward issued so far on `device` (eager ones and the most recent replay of every
    captured one); raises RuntimeError if one of them overflowed its capacity.  Returns the counts of the last one."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if synchronize:
        torch.cuda.synchronize(idx)
    _poll_static(idx, wait=True)
    for ticket, cap in (_static.get(idx) or {}).get("graph", []):
        _check_ticket(idx, ticket, cap, False)
    return dict(_last_counts.get(idx, {}))


def _ctx(dev_index: int) -> C.c_void_p:
    c = _contexts.get(dev_index)
    if c is None:
        h = C.c_void_p()
        N.check(N.lib().gs_context_create(int(dev_index), C.byref(h)))
        _contexts[dev_index] = c = h
    return c


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def _guard(idx: int):
    """Device guard only when the tensor's device is not already current (the context manager costs ~10 us)."""
    return _NO_GUARD if torch.cuda.current_device() == idx else torch.cuda.device(idx)


def _raw_stream(idx: int) -> int:
    """cudaStream_t of torch's current stream on device idx, without building a torch.cuda.Stream object."""
    return torch._C._cuda_getCurrentRawStream(idx)


_fast_mod = None          # the torch binding (csrc/_gsraster_torch.so); False = not built / disabled
_CPP_AUTOGRAD = __import__("os").environ.get("GS_CPP_AUTOGRAD", "1") != "0"


def _fast():
    """Native twin of _forward_impl / _backward_impl (csrc/torch_binding.cpp over the same C ABI): same calls in the same
    order, ~100 us less Python per step.  Built by __graft_entry__.build(); GS_NO_TORCH_BINDING=1 forces the ctypes path."""
    global _fast_mod
    if _fast_mod is None:
        _fast_mod = False
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "_gsraster_torch.so")
        if os.environ.get("GS_NO_TORCH_BINDING", "0") != "1" and os.path.exists(path):
            import importlib.util
            N.lib()                                   # libgsraster_b200.so first (fails loudly if missing)
            try:
                spec = importlib.util.spec_from_file_location("_gsraster_torch", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _fast_mod = mod
            except Exception as ex:                   # host glue only: the ctypes path drives the same CUDA library
                import warnings
                warnings.warn(f"luciddreamer_b200: torch binding {path} failed to load ({ex}); using the ctypes host path")
    return _fast_mod


def _round_cap(n: int) -> int:
    return max(64, (int(n) + 63) // 64 * 64)      # multiples of 64 keep gs_binning_bytes(cap) == 12 * cap


def _dev_f32(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    """Reference convention: empty tensor == absent (RAST/.../__init__.py:198-208). Otherwise f32, contiguous,
    on the compute device (the reference calls .contiguous() on everything, rasterize_points.cu:95-113)."""
    if t is None or t.numel() == 0:
        return None
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    t = t.contiguous()
    if t.data_ptr() & 15:            # a view at an odd storage offset: the kernels use 128-bit loads (gsraster.h)
        t = t.clone()
    return t


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _Prepared(NamedTuple):
    frame: N.GsFrame
    keep: tuple
    device: torch.device
    P: int
    M: int


def _prepare(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
             projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug
             ) -> _Prepared:
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")      # rasterize_points.cu:57-59
    if not means3D.is_cuda:
        raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
    dev = means3D.device
    P = means3D.size(0)
    t = [_dev_f32(x, dev) for x in (bg, means3D, sh, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                    viewmatrix, projmatrix, campos)]
    bg_, m3, sh_, col, op, sc, rot, cov, vm, pm, cp = t
    M = 0 if sh_ is None else sh_.size(1)
    f = N.GsFrame()
    f.P, f.D, f.M, f.W, f.H = P, int(degree), int(M), int(image_width), int(image_height)
    f.tan_fovx, f.tan_fovy, f.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
    f.prefiltered, f.debug = int(bool(prefiltered)), int(bool(debug))
    f.bg, f.means3D, f.shs, f.colors_precomp = _p(bg_), _p(m3), _p(sh_), _p(col)
    f.opacities, f.scales, f.rotations, f.cov3D_precomp = _p(op), _p(sc), _p(rot), _p(cov)
    f.viewmatrix, f.projmatrix, f.campos = _p(vm), _p(pm), _p(cp)
    return _Prepared(f, tuple(t), dev, P, M)


def _forward_impl(prep: _Prepared):
    """Returns (num_rendered, color, depth, radii, geom, binning, img, pair_capacity)."""
    L = N.lib()
    f, dev, P = prep.frame, prep.device, prep.P
    H, W = f.H, f.W
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    _poll_static(idx)
    with _guard(idx):
        ctx = _ctx(idx)
        stream = _raw_stream(idx)
        u8 = dict(dtype=torch.uint8, device=dev)
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        ng, ni = L.gs_geom_bytes(P), L.gs_image_bytes(W, H)
        scratch = torch.empty((ng + ni,), **u8)          # one allocation; both sizes are multiples of 256 B
        geom, img = scratch[:ng], scratch[ng:]
        ticket = C.c_int32(-1)
        N.check(L.gs_forward_preprocess(ctx, C.byref(f), geom.data_ptr(), img.data_ptr(), radii.data_ptr(), stream,
                                        C.byref(ticket)))
        counts = N.GsCounts()
        hint = _cap_hint.get(idx)
        scap = _static_cap(idx)
        if scap >= 0:
            # sync-free: exactly the caller's capacity, no wait; the ticket is checked at a later call
            cap = scap
            binning = torch.empty((L.gs_binning_bytes(cap),), **u8)
            N.check(L.gs_forward_render(ctx, C.byref(f), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(), cap,
                                        img.data_ptr(), color.data_ptr(), depth.data_ptr(), 0, stream))
            _note_static(idx, ticket.value, cap)
            return -1, color, depth, radii, geom, binning, img, (cap, P)
        if hint is None:
            # first frame on this device: learn the pair count (one event wait), then render
            N.check(L.gs_forward_counts(ctx, ticket, C.byref(counts)))
            cap = _round_cap(counts.num_pairs * 1.25 + 4096)
            binning = torch.empty((L.gs_binning_bytes(cap),), **u8)
            N.check(L.gs_forward_render(ctx, C.byref(f), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(), cap,
                                        img.data_ptr(), color.data_ptr(), depth.data_ptr(), 0, stream))
        else:
            # speculative: enqueue the render with the previous capacity, *then* wait for the count; the GPU
            # never idles on the host (reference: blocking cudaMemcpy, rasterizer_impl.cu:282)
            cap = _round_cap(hint * 1.25 + 4096)
            binning = torch.empty((L.gs_binning_bytes(cap),), **u8)
            N.check(L.gs_forward_render(ctx, C.byref(f), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(), cap,
                                        img.data_ptr(), color.data_ptr(), depth.data_ptr(), 0, stream))
            N.check(L.gs_forward_counts(ctx, ticket, C.byref(counts)))
            if counts.num_pairs > cap:       # device-side guard skipped the render: grow and redo it
                cap = _round_cap(counts.num_pairs * 1.25 + 4096)
                binning = torch.empty((L.gs_binning_bytes(cap),), **u8)
                N.check(L.gs_forward_render(ctx, C.byref(f), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(),
                                            cap, img.data_ptr(), color.data_ptr(), depth.data_ptr(), 1, stream))
        _note_pairs(idx, counts.num_pairs)
    return int(counts.num_rendered), color, depth, radii, geom, binning, img, (cap, int(counts.num_visible))


def _backward_impl(prep: _Prepared, radii, geom, binning, img, cap, grad_color, want_colors: bool, want_cov: bool,
                   out: Optional[dict] = None, peers: Optional[dict] = None):
    """`out` (optional): pre-allocated contiguous f32 destinations keyed dm3/dm2/dop/dsh/dsc/drot -- e.g. views of
    a flat multiview.GradBucket -- that the kernels write instead of fresh tensors (every row is overwritten)."""
    L = N.lib()
    out = out or {}
    f, dev, P, M = prep.frame, prep.device, prep.P, prep.M
    if peers is not None:
        return _backward_peers(prep, radii, geom, binning, img, cap, grad_color, out, peers)
    cap, nvis = cap if isinstance(cap, tuple) else (cap, P)   # `_C` path: num_visible unknown -> bound by P
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _guard(idx):
        stream = _raw_stream(idx)
        f32 = dict(dtype=torch.float32, device=dev)
        # the tile pass goes to the GPU first; allocations below overlap with it
        gc = _dev_f32(grad_color, dev)
        if gc is None or gc.numel() != 3 * f.H * f.W:
            raise RuntimeError(f"dL_dout_color must hold 3 x {f.H} x {f.W} values")   # k_blend_bwd reads exactly that many
        N.check(L.gs_backward_blend(_ctx(idx), C.byref(f), geom.data_ptr(), binning.data_ptr(), cap, img.data_ptr(),
                                    _p(gc), stream))
        g = N.GsGrads()
        # every gradient without a caller-provided sink is a view of ONE flat allocation, in bucket order
        # [means3D | sh | opacity | scales | rotations | means2D]
        need = {"dm3": 3, "dsh": 3 * M, "dop": 1, "dsc": 3, "drot": 4, "dm2": 3}
        missing = [k for k in need if out.get(k) is None]
        offs, o = {}, 0
        for k in missing:                      # segment starts stay 256-byte aligned (128-bit stores in the kernels)
            offs[k] = o
            o += (P * need[k] + 63) // 64 * 64
        flat = torch.empty((o,), **f32) if missing else None

        def dst(key, shape, zero=False):
            t = out.get(key)
            if t is not None:
                if tuple(t.shape) != tuple(shape) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                    raise RuntimeError(f"gradient sink '{key}' must be a contiguous f32 {tuple(shape)} tensor on {dev}")
            else:
                t = flat.narrow(0, offs[key], P * need[key]).view(shape)
            return t.zero_() if zero else t

        dm3 = dst("dm3", (P, 3)); dm2 = dst("dm2", (P, 3))
        dop = dst("dop", (P, 1))
        dsh = dst("dsh", (P, M, 3))
        dsc = dst("dsc", (P, 3), zero=not f.scales)
        drot = dst("drot", (P, 4), zero=not f.rotations)
        dcol = torch.empty((P, 3), **f32) if want_colors else None
        dcov = torch.empty((P, 6), **f32) if want_cov else None
        g.dL_dmeans3D, g.dL_dmeans2D, g.dL_dopacity = dm3.data_ptr(), dm2.data_ptr(), dop.data_ptr()
        g.dL_dsh = dsh.data_ptr() if M > 0 else None
        g.dL_dscales = dsc.data_ptr() if f.scales else None
        g.dL_drotations = drot.data_ptr() if f.rotations else None
        g.dL_dcolors = _p(dcol)
        g.dL_dcov3D = _p(dcov)
        nscr = L.gs_backward_scratch_bytes(nvis)
        scratch = torch.empty((nscr,), dtype=torch.uint8, device=dev)
        N.check(L.gs_backward_gradients(_ctx(idx), C.byref(f), radii.data_ptr(), geom.data_ptr(), img.data_ptr(),
                                        scratch.data_ptr(), nscr, C.byref(g), stream))
    return dm2, dcol, dop, dm3, dcov, dsh, dsc, drot


def _backward_peers(prep: _Prepared, radii, geom, binning, img, cap, grad_color, out, peers):
    """Shared-model data-parallel backward: the parameter gradients of this view are ADDED into every rank's
    symmetric gradient bucket by the final kernel (multiview.SymmGradBucket); only dL_dmeans2D is returned."""
    L = N.lib()
    f, dev, P = prep.frame, prep.device, prep.P
    cap, nvis = cap if isinstance(cap, tuple) else (cap, P)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _guard(idx):
        stream = _raw_stream(idx)
        dm2 = out.get("dm2")
        if dm2 is None:
            dm2 = torch.empty((P, 3), dtype=torch.float32, device=dev)
        elif tuple(dm2.shape) != (P, 3) or dm2.dtype != torch.float32 or not dm2.is_contiguous() or dm2.device != dev:
            raise RuntimeError(f"gradient sink 'dm2' must be a contiguous f32 {(P, 3)} tensor on {dev}")
        g = N.GsGrads()
        g.dL_dmeans2D = dm2.data_ptr()
        world = int(peers["world"])
        ptrs = (C.c_void_p * world)(*[int(p) for p in peers["ptrs"]])
        seg = (C.c_int64 * 5)(*[int(o) for o in peers["seg_off"]])
        g.peer_world = world
        g.peer_buckets = C.cast(ptrs, C.POINTER(C.c_void_p))
        g.peer_multicast = int(peers.get("mc") or 0) or None
        g.peer_seg_off = C.cast(seg, C.POINTER(C.c_int64))
        sig = None
        if peers.get("signals"):          # cross-rank barriers folded into the kernel (multiview.SymmGradBucket)
            sig = (C.c_void_p * world)(*[int(p) for p in peers["signals"]])
            g.peer_signals = C.cast(sig, C.POINTER(C.c_void_p))
            g.peer_rank = int(peers["rank"])
            g.peer_epoch_begin = int(peers.get("epoch_begin", 0))
            g.peer_epoch_end = int(peers.get("epoch_end", 0))
        gc = _dev_f32(grad_color, dev)
        if gc is None or gc.numel() != 3 * f.H * f.W:
            raise RuntimeError(f"dL_dout_color must hold 3 x {f.H} x {f.W} values")
        nscr = L.gs_backward_scratch_bytes(nvis)
        scratch = torch.empty((nscr,), dtype=torch.uint8, device=dev)
        N.check(L.gs_backward(_ctx(idx), C.byref(f), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(), cap,
                              img.data_ptr(), scratch.data_ptr(), nscr, _p(gc), None, C.byref(g), stream))
    return dm2


# ------------------------------------------------------------------------- `_C`-compatible module surface

class _C:
    """Same three entry points, argument order and return tuples as the reference's pybind module
    (RAST/ext.cpp:15-19, rasterize_points.h:18-68).  The three byte tensors are opaque scratch; capacity of the
    binning buffer is recovered from its size, so the reference's own __init__.py could sit on top verbatim."""

    @staticmethod
    def rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                            viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                            campos, prefiltered, debug):
        prep = _prepare(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                        projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
                        debug)
        nr, color, depth, radii, geom, binning, img, _capnv = _forward_impl(prep)
        return nr, color, depth, radii, geom, binning, img

    @staticmethod
    def rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                     viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, sh,
                                     degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
        H, W = dL_dout_color.size(1), dL_dout_color.size(2)
        prep = _prepare(bg, means3D, colors, torch.empty(0), scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, False, debug)
        # opacities are not an input of the reference backward; the C ABI only checks non-NULL
        prep.frame.opacities = means3D.data_ptr()
        cap = binningBuffer.numel() // 12
        dm2, dcol, dop, dm3, dcov, dsh, dsc, drot = _backward_impl(prep, radii, geomBuffer, binningBuffer,
                                                                   imageBuffer, cap, dL_dout_color, True, True)
        return dm2, dcol, dop, dm3, dcov, dsh, dsc, drot

    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        dev = means3D.device
        if not means3D.is_cuda:
            raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
        P = means3D.size(0)
        present = torch.empty((P,), dtype=torch.bool, device=dev)
        if P:
            m3, vm, pm = _dev_f32(means3D, dev), _dev_f32(viewmatrix, dev), _dev_f32(projmatrix, dev)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            with torch.cuda.device(idx):
                N.check(N.lib().gs_mark_visible(P, m3.data_ptr(), vm.data_ptr(), _p(pm), present.data_ptr(),
                                                torch.cuda.current_stream(idx).cuda_stream))
        return present


# -------------------------------------------------------------------------------- reference Python surface

def _t(x):
    return x if x is not None else _EMPTY


_EMPTY = torch.empty(0)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    """__init__.py:21-42.  With the torch binding built, the autograd node itself is native (torch_binding.cpp:
    RasterizeFn -- same inputs, outputs, gradient order and None conventions as _RasterizeGaussians below, which stays
    the implementation when the binding is absent or GS_CPP_AUTOGRAD=0)."""
    fast = _fast()
    if fast and _CPP_AUTOGRAD and means3D.is_cuda:
        rs = raster_settings
        idx = means3D.device.index
        hint = _cap_hint.get(idx)
        scap = -1
        if _static:
            _poll_static(idx)
            scap = _static_cap(idx)
        color, radii, depth = fast.rasterize(
            means3D, _t(means2D), _t(sh), _t(colors_precomp), _t(opacities), _t(scales), _t(rotations), _t(cov3Ds_precomp),
            rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos, _ctx(idx).value, rs.scale_modifier, rs.tanfovx, rs.tanfovy,
            rs.image_height, rs.image_width, rs.sh_degree, bool(rs.prefiltered), bool(rs.debug), -1 if hint is None else hint,
            scap)
        if scap >= 0:
            _note_static(idx, fast.last_ticket(idx), scap)
        else:
            _note_pairs(idx, fast.last_pairs(idx))
        return color, radii, depth
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        fast = _fast()
        ctx.raster_settings = rs
        ctx.set_materialize_grads(False)      # no zero-fill kernels for the unused grad_radii / grad_depth
        if fast:
            if means3D.dim() != 2 or means3D.size(1) != 3:
                raise RuntimeError("means3D must have dimensions (num_points, 3)")      # rasterize_points.cu:57-59
            if not means3D.is_cuda:
                raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
            idx = means3D.device.index
            hint = _cap_hint.get(idx)
            _poll_static(idx)
            scap = _static_cap(idx)
            (num_rendered, color, depth, radii, geom, binning, img, cap, nvis, npairs, ticket) = fast.forward(
                _ctx(idx).value, rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh,
                rs.sh_degree, rs.campos, bool(rs.prefiltered), bool(rs.debug), -1 if hint is None else hint, scap)
            if scap >= 0:
                _note_static(idx, ticket, scap)
            else:
                _note_pairs(idx, npairs)
            ctx.num_rendered = num_rendered
            ctx.pair_capacity = (cap, nvis)
            ctx.prep = None
            ctx.save_for_backward(radii, geom, binning, img, means3D, sh, colors_precomp, opacities, scales, rotations,
                                  cov3Ds_precomp)
            ctx.mark_non_differentiable(radii)
            return color, radii, depth
        prep = _prepare(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                        cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                        rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        num_rendered, color, depth, radii, geom, binning, img, cap = _forward_impl(prep)
        ctx.num_rendered = num_rendered
        ctx.pair_capacity = cap
        ctx.prep = prep                       # keeps the contiguous f32 inputs alive + the filled GsFrame
        ctx.save_for_backward(radii, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        if ctx.prep is None:                  # torch-binding path
            (radii, geom, binning, img, means3D, sh, colors_precomp, opacities, scales, rotations,
             cov3Ds_precomp) = ctx.saved_tensors
            H, W = rs.image_height, rs.image_width
            if grad_out_color is None:        # only depth was used downstream: depth carries no gradient
                grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=means3D.device)

            def has(t):
                return t is not None and t.numel() > 0
            cap, nvis = ctx.pair_capacity
            dm2, dcol, dop, dm3, dcov, dsh, dsc, drot = _fast().backward(
                _ctx(means3D.device.index).value, rs.bg, means3D, colors_precomp, opacities, scales, rotations,
                rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W, sh,
                rs.sh_degree, rs.campos, bool(rs.debug), radii, geom, binning, img, cap, nvis, grad_out_color,
                has(colors_precomp), has(cov3Ds_precomp))
            return (dm3, dm2, dsh if has(sh) else None, dcol, dop, dsc if has(scales) else None,
                    drot if has(rotations) else None, dcov, None)
        radii, geom, binning, img = ctx.saved_tensors
        prep = ctx.prep
        if grad_out_color is None:            # only depth was used downstream: depth carries no gradient
            grad_out_color = torch.zeros((3, prep.frame.H, prep.frame.W), dtype=torch.float32, device=prep.device)
        f = prep.frame
        want_colors = bool(f.colors_precomp)
        want_cov = bool(f.cov3D_precomp)
        dm2, dcol, dop, dm3, dcov, dsh, dsc, drot = _backward_impl(prep, radii, geom, binning, img,
                                                                   ctx.pair_capacity, grad_out_color, want_colors,
                                                                   want_cov)
        # order of __init__.py:144-156; absent inputs get None (the reference returns zero tensors for them,
        # which autograd discards for inputs that do not require grad)
        return (dm3, dm2, dsh if f.shs else None, dcol, dop, dsc if f.scales else None,
                drot if f.rotations else None, dcov, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        shs = e if shs is None else shs
        colors_precomp = e if colors_precomp is None else colors_precomp
        scales = e if scales is None else scales
        rotations = e if rotations is None else rotations
        cov3D_precomp = e if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs)
