"""CPU oracle of simple_knn.distCUDA2 (TEST INFRASTRUCTURE ONLY): ctypes front-end of oracle/knn_oracle.c (brute force,
the reference's float arithmetic, bit-exact) plus an independent float64 k-d tree cross-check for sizes where O(P^2)
is too slow."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o


def dist2(points: np.ndarray) -> np.ndarray:
    """mean squared distance to the 3 nearest other points, float32, like distCUDA2 (spatial.cu:15-26)."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    f = _o.lib().gso_knn_mean_dist2
    f.restype = None
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    f(pts.shape[0], pts.ctypes.data, out.ctypes.data)
    return out


def dist2_kdtree(points: np.ndarray) -> np.ndarray:
    """Independent restatement in float64 (scipy cKDTree, k = 4 incl. the query): for sizes the brute force cannot do.
    Agrees with the float32 result to rounding (not bit-exact)."""
    from scipy.spatial import cKDTree
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3).astype(np.float64)
    d, _ = cKDTree(pts).query(pts, k=4)
    return (d[:, 1:] ** 2).mean(axis=1)
