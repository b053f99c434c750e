"""Runs the UNMODIFIED reference CUDA rasterizer (oracle/_ref/libref_rasterizer.so, built by oracle/Makefile from
the sources under /root/reference) on torch CUDA tensors.  TEST / BASELINE INFRASTRUCTURE ONLY -- used by the GPU
parity tests, tests/golden/make_golden.py and `bench.py --impl reference`; never by the product path.

Same argument order and return tuples as the reference binding (RAST/rasterize_points.h:18-68)."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_rasterizer.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_forward.restype = C.c_int
        L.ref_forward.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 +
                                  [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 3 + [C.c_int])
        L.ref_backward.restype = None
        L.ref_backward.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 +
                                   [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float] + [C.c_void_p] * 12 + [C.c_int])
        L.ref_mark_visible.argtypes = [C.c_int] + [C.c_void_p] * 4
        _lib = L
    return _lib


def _f(t, dev):
    if t is None or t.numel() == 0:
        return None
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


class RefContext:
    """Holds the reference's three scratch buffers between forward and backward (what its torch binding keeps in
    geomBuffer / binningBuffer / imgBuffer)."""

    def __init__(self):
        self.h = lib().ref_create()

    def __del__(self):
        try:
            lib().ref_destroy(self.h)
        except Exception:
            pass


def rasterize_gaussians(ctx: RefContext, bg, means3D, colors, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                        degree, campos, prefiltered=False, debug=False):
    """All work runs on the legacy default stream like the reference; torch's current stream must be the default."""
    dev = means3D.device
    P = means3D.size(0)
    H, W = int(image_height), int(image_width)
    t = [_f(x, dev) for x in (bg, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix,
                              projmatrix, campos)]
    bg_, m3, sh_, col, op, sc, rot, cov, vm, pm, cp = t
    M = 0 if sh_ is None else sh_.size(1)
    color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    R = lib().ref_forward(ctx.h, P, int(degree), M, _p(bg_), W, H, _p(m3), _p(sh_), _p(col), _p(op), _p(sc),
                          float(scale_modifier), _p(rot), _p(cov), _p(vm), _p(pm), _p(cp), float(tan_fovx),
                          float(tan_fovy), int(prefiltered), color.data_ptr(), depth.data_ptr(), radii.data_ptr(),
                          int(debug))
    ctx.keep = t
    ctx.meta = (P, int(degree), M, W, H, float(scale_modifier), float(tan_fovx), float(tan_fovy))
    return R, color, depth, radii


def rasterize_gaussians_backward(ctx: RefContext, radii, dL_dout_color, debug=False):
    """Returns (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
    dL_dconic) -- the binding's 8-tuple plus the intermediate dL_dconic."""
    P, D, M, W, H, mod, tfx, tfy = ctx.meta
    bg_, m3, sh_, col, op, sc, rot, cov, vm, pm, cp = ctx.keep
    dev = m3.device
    f32 = dict(dtype=torch.float32, device=dev)
    dm2 = torch.empty((P, 3), **f32); dcol = torch.empty((P, 3), **f32); dop = torch.empty((P, 1), **f32)
    dm3 = torch.empty((P, 3), **f32); dcov = torch.empty((P, 6), **f32); dsh = torch.empty((P, max(M, 0), 3), **f32)
    dsc = torch.empty((P, 3), **f32); drot = torch.empty((P, 4), **f32); dconic = torch.empty((P, 2, 2), **f32)
    g = dL_dout_color.to(device=dev, dtype=torch.float32).contiguous()
    gd = torch.zeros((1, H, W), **f32)
    lib().ref_backward(ctx.h, P, D, M, _p(bg_), W, H, _p(m3), _p(sh_), _p(col), _p(sc), mod, _p(rot), _p(cov), _p(vm),
                       _p(pm), _p(cp), tfx, tfy, radii.data_ptr(), g.data_ptr(), gd.data_ptr(), dm2.data_ptr(),
                       dconic.data_ptr(), dop.data_ptr(), dcol.data_ptr(), dm3.data_ptr(), dcov.data_ptr(),
                       dsh.data_ptr() if M > 0 else None, dsc.data_ptr(), drot.data_ptr(), int(debug))
    return dm2, dcol, dop, dm3, dcov, dsh, dsc, drot, dconic


# ---------------------------------------------------------------- simple_knn (the reference's second native extension)
KNN_LIB_PATH = os.path.join(_HERE, "_ref", "libref_simpleknn.so")
_knn = None


def knn_available() -> bool:
    return os.path.exists(KNN_LIB_PATH)


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """The reference's distCUDA2 (spatial.cu:15-26): torch.full({P}, 0) then SimpleKNN::knn."""
    global _knn
    if _knn is None:
        _knn = C.CDLL(KNN_LIB_PATH)
        _knn.ref_knn.restype = C.c_int
        _knn.ref_knn.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    pts = points.contiguous().float()
    out = torch.zeros(pts.shape[0], device=pts.device)
    with torch.cuda.device(pts.device):
        torch.cuda.synchronize()
        rc = _knn.ref_knn(pts.shape[0], pts.data_ptr(), out.data_ptr())
    if rc != 0:
        raise RuntimeError(f"reference simple_knn failed: cuda error {rc}")
    return out
