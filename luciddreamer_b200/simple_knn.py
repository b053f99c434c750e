"""simple_knn.distCUDA2 on the B200 kernels.

`distCUDA2(points)` has the reference binding's contract (submodules/simple-knn/spatial.cu:15-26; used by
scene/gaussian_model.py:136 to initialise the Gaussian scales): points [P,3] float CUDA tensor -> [P] float32, the mean
of the squared distances to the 3 nearest other points.  Runs on the caller's current stream, no host sync."""
from __future__ import annotations

import torch

from . import _native as N
from . import rasterizer as R


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (there is no CPU fallback)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("distCUDA2: points must have shape [P, 3]")
    pts = points.detach().contiguous().float()
    P = pts.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    idx = pts.device.index if pts.device.index is not None else torch.cuda.current_device()
    with R._guard(idx):
        scratch = torch.empty(N.lib().gs_knn_scratch_bytes(P), dtype=torch.uint8, device=pts.device)
        N.check(N.lib().gs_knn_mean_dist2(R._ctx(idx), P, pts.data_ptr(), scratch.data_ptr(), out.data_ptr(),
                                          R._raw_stream(idx)))
    return out
