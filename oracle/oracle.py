"""ctypes front-end of the CPU oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.  The product path
(luciddreamer_b200/) never imports this module.

The argument names mirror the reference `_C.rasterize_gaussians` binding
(RAST/rasterize_points.h:18-63) so parity tests read like calls to the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/gs_oracle.c -> oracle/liboracle.so (gcc, OpenMP, no FMA contraction)."""
    src = [os.path.join(_HERE, "gs_oracle.c"), os.path.join(_HERE, "gs_oracle_impl.h"), os.path.join(_HERE, "knn_oracle.c")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src if os.path.exists(s))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gso_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().gso_max_threads())


def set_threads(n: int) -> None:
    lib().gso_set_threads(int(n))


def _mk_params(real):
    class Params(C.Structure):
        _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                    ("tanfovx", real), ("tanfovy", real), ("scale_modifier", real),
                    ("bg", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
                    ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
                    ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p),
                    ("projmatrix", C.c_void_p), ("campos", C.c_void_p)]
    return Params


_PARAMS = {"f32": _mk_params(C.c_float), "f64": _mk_params(C.c_double)}
_NP = {"f32": np.float32, "f64": np.float64}


def _arr(x, dt):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(x), dtype=dt)
    return a if a.size else None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class OracleForward:
    mode: str
    color: np.ndarray
    depth: np.ndarray
    radii: np.ndarray
    num_rendered: int
    _state: int = 0
    _params: object = None
    _keep: dict = field(default_factory=dict)

    # ---- internals exposed for unit tests of the integer pipeline
    def _get(self, name, ctype, n, dt):
        fn = getattr(lib(), f"gso_{name}_{self.mode}")
        fn.restype = C.POINTER(ctype)
        fn.argtypes = [C.c_void_p]
        p = fn(self._state)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.zeros((0,), dt)

    @property
    def point_list(self):
        return self._get("point_list", C.c_uint32, self.num_rendered, np.uint32)

    @property
    def keys(self):
        return self._get("keys", C.c_uint64, self.num_rendered, np.uint64)

    @property
    def ranges(self):
        H, W = self.depth.shape[-2:]
        G = ((W + 15) // 16) * ((H + 15) // 16)
        return self._get("ranges", C.c_uint32, 2 * G, np.uint32).reshape(G, 2)

    @property
    def n_contrib(self):
        H, W = self.depth.shape[-2:]
        return self._get("n_contrib", C.c_uint32, H * W, np.uint32).reshape(H, W)

    @property
    def final_T(self):
        H, W = self.depth.shape[-2:]
        rt = C.c_float if self.mode == "f32" else C.c_double
        return self._get("final_T", rt, H * W, _NP[self.mode]).reshape(H, W)

    def geom(self, name, width):
        rt = C.c_float if self.mode == "f32" else C.c_double
        P = self.radii.shape[0]
        return self._get(name, rt, P * width, _NP[self.mode]).reshape(P, width)

    @property
    def tiles_touched(self):
        return self._get("tiles_touched", C.c_uint32, self.radii.shape[0], np.uint32)

    def free(self):
        if self._state:
            fn = getattr(lib(), f"gso_free_{self.mode}")
            fn.argtypes = [C.c_void_p]
            fn(self._state)
            self._state = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def rasterize_gaussians(bg, means3D, colors_precomp, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered=False, debug=False, mode="f32") -> OracleForward:
    """Oracle forward. Same argument order as the reference binding (rasterize_points.h:18-38)."""
    dt = _NP[mode]
    keep = dict(bg=_arr(bg, dt), means3D=_arr(means3D, dt), colors=_arr(colors_precomp, dt),
                opac=_arr(opacity, dt), scales=_arr(scales, dt), rots=_arr(rotations, dt),
                cov=_arr(cov3D_precomp, dt), vm=_arr(viewmatrix, dt), pm=_arr(projmatrix, dt),
                sh=_arr(sh, dt), campos=_arr(campos, dt))
    m3 = keep["means3D"]
    P = 0 if m3 is None else m3.reshape(-1, 3).shape[0]
    M = 0 if keep["sh"] is None else keep["sh"].reshape(P, -1, 3).shape[1]
    H, W = int(image_height), int(image_width)
    prm = _PARAMS[mode]()
    prm.P, prm.D, prm.M, prm.W, prm.H = P, int(degree), M, W, H
    prm.tanfovx, prm.tanfovy, prm.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
    prm.bg = _ptr(keep["bg"]); prm.means3D = _ptr(m3); prm.shs = _ptr(keep["sh"])
    prm.colors_precomp = _ptr(keep["colors"]); prm.opacities = _ptr(keep["opac"])
    prm.scales = _ptr(keep["scales"]); prm.rotations = _ptr(keep["rots"]); prm.cov3D_precomp = _ptr(keep["cov"])
    prm.viewmatrix = _ptr(keep["vm"]); prm.projmatrix = _ptr(keep["pm"]); prm.campos = _ptr(keep["campos"])
    color = np.zeros((3, H, W), dt)
    depth = np.zeros((1, H, W), dt)
    radii = np.zeros((P,), np.int32)
    nr = C.c_longlong(0)
    fn = getattr(lib(), f"gso_forward_{mode}")
    fn.restype = C.c_void_p
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong)]
    st = fn(C.byref(prm), _ptr(color), _ptr(depth), _ptr(radii), C.byref(nr))
    return OracleForward(mode, color, depth, radii, int(nr.value), st, prm, keep)


def rasterize_gaussians_backward(fwd: OracleForward, dL_dout_color, dL_dout_depth=None):
    """Oracle backward. Returns the 8-tuple of the reference binding (rasterize_points.h:40-63):
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    plus dL_dconic as a 9th element (intermediate, handy for debugging)."""
    dt = _NP[fwd.mode]
    prm = fwd._params
    P, M = prm.P, prm.M
    g = _arr(dL_dout_color, dt)
    outs = dict(m2=np.zeros((P, 3), dt), col=np.zeros((P, 3), dt), op=np.zeros((P, 1), dt), m3=np.zeros((P, 3), dt),
                cov=np.zeros((P, 6), dt), sh=np.zeros((P, M, 3), dt), sc=np.zeros((P, 3), dt), rot=np.zeros((P, 4), dt),
                conic=np.zeros((P, 2, 2), dt))
    fn = getattr(lib(), f"gso_backward_{fwd.mode}")
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 12
    fn(C.byref(prm), fwd._state, _ptr(g) if g is not None else None, _ptr(outs["m2"]), _ptr(outs["col"]),
       _ptr(outs["op"]), _ptr(outs["m3"]), _ptr(outs["cov"]), _ptr(outs["sh"]), _ptr(outs["sc"]), _ptr(outs["rot"]),
       _ptr(outs["conic"]))
    return (outs["m2"], outs["col"], outs["op"], outs["m3"], outs["cov"], outs["sh"], outs["sc"], outs["rot"],
            outs["conic"])


def mark_visible(means3D, viewmatrix, projmatrix=None, mode="f32"):
    dt = _NP[mode]
    m = _arr(means3D, dt)
    P = 0 if m is None else m.reshape(-1, 3).shape[0]
    out = np.zeros((P,), np.uint8)
    if P:
        fn = getattr(lib(), f"gso_mark_visible_{mode}")
        fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        vm = _arr(viewmatrix, dt)
        fn(P, _ptr(m), _ptr(vm), _ptr(out))
    return out.astype(bool)
