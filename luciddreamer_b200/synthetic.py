"""Deterministic synthetic scenes and cameras for parity tests and bench.py.

Host-side only (torch CPU generators -> bit-reproducible on every machine).  Camera matrix
conventions follow the reference:
  * `world_view_transform` = W2C^T stored row-major   (scene/cameras.py:58)
  * `full_proj_transform`  = (Proj . W2C)^T           (scene/cameras.py:60, utils/graphics.py:55-75)
  * `camera_center`        = inverse(world_view_transform)[3, :3]   (scene/cameras.py:61,74-75)
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple, Optional

import numpy as np
import torch

REF_FOCAL = 582.69          # arguments.py:45
REF_SIZE = 512              # arguments.py:42
REF_FOVX = 2.0 * math.atan(REF_SIZE / (2.0 * REF_FOCAL))   # 0.8279 rad (== cameras/*.json camera_angle_x)


class CameraView(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor      # [4,4] f32  (W2C^T)
    projmatrix: torch.Tensor      # [4,4] f32  ((Proj.W2C)^T)
    campos: torch.Tensor          # [3]   f32


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """Same entries as utils/graphics.py:55-75 (getProjectionMatrix)."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_camera(W: int, H: int, c2w: Optional[np.ndarray] = None, fovx: float = REF_FOVX,
                znear: float = 0.01, zfar: float = 100.0, swap_fov_like_load_json: bool = False) -> CameraView:
    """Camera at pose `c2w` (COLMAP axes: x right, y down, z forward; identity = at origin looking +z).

    swap_fov_like_load_json reproduces a quirk of the reference's utils/camera.py:48
    (`MiniCam(W, H, FoVx, FoVy, ...)` against the signature `(width, height, fovy, fovx, ...)`): the
    projection matrix is built from the true FoVs but render() derives tanfovx/tanfovy from the swapped ones.
    """
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)       # focal2fov(fov2focal(FoVx, W), H)
    if c2w is None:
        c2w = np.eye(4)
    c2w = np.asarray(c2w, dtype=np.float64)
    if c2w.shape[0] == 3:
        c2w = np.concatenate([c2w, np.array([[0, 0, 0, 1.0]])], 0)
    w2c = np.float32(np.linalg.inv(c2w))
    wvt = torch.from_numpy(w2c).T.contiguous()                # world_view_transform
    proj = projection_matrix(znear, zfar, fovx, fovy).T       # projection_matrix (transposed)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = torch.inverse(wvt)[3, :3].contiguous()
    fx_, fy_ = (fovy, fovx) if swap_fov_like_load_json else (fovx, fovy)
    return CameraView(H, W, math.tan(fx_ * 0.5), math.tan(fy_ * 0.5), wvt, full, campos)


def nerf_c2w_to_colmap(c2w_gl: np.ndarray) -> np.ndarray:
    """NeRF/Blender camera-to-world (y up, z back) -> COLMAP axes (utils/camera.py:38-40)."""
    c2w = np.array(c2w_gl, dtype=np.float64, copy=True)
    if c2w.shape[0] == 3:
        c2w = np.concatenate([c2w, np.array([[0, 0, 0, 1.0]])], 0)
    c2w[:3, 1:3] *= -1
    return c2w


def rotate360_poses(n_frames: int = 720) -> np.ndarray:
    """Our own generator of a `rotate360`-style path: camera fixed at the origin, yawing a full turn about
    the vertical axis (the reference preset cameras/rotate360.json has 720 frames with zero translation;
    utils/trajectory.py:168-176).  Returns [n,4,4] COLMAP-convention c2w matrices."""
    out = np.zeros((n_frames, 4, 4))
    for i in range(n_frames):
        th = 2.0 * math.pi * i / n_frames
        c, s = math.cos(th), math.sin(th)
        out[i] = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1.0]])
    return out


def llff_poses(n_frames: int = 400, max_deg: float = 5.0, radius: float = 0.15) -> np.ndarray:
    """`llff`-style path: small spiral in front of the scene, camera orbiting within +-max_deg while
    translating on a circle of `radius` (utils/trajectory.py:431-446 orbits within +-5 degrees)."""
    out = np.zeros((n_frames, 4, 4))
    for i in range(n_frames):
        ph = 2.0 * math.pi * i / n_frames
        yaw = math.radians(max_deg) * math.cos(ph)
        pitch = math.radians(max_deg) * math.sin(ph)
        cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        M = np.eye(4)
        M[:3, :3] = Ry @ Rx
        M[:3, 3] = [radius * math.cos(ph), radius * math.sin(ph), 0.0]
        out[i] = M
    return out


def make_scene(P: int, seed: int, M: int = 16, scale_mult: float = 1.0, opacity_shift: float = 0.0,
               aniso_sigma: float = 0.3) -> Dict[str, torch.Tensor]:
    """Shell of Gaussians around the origin.  All tensors CPU float32, contiguous.

    means3D   direction = normalise(N(0,I)); radius rho = exp(U(ln1, ln8))
    scales    rho * (1.5/582.69) * exp(N(0,.5^2)) * exp(N(0,.3^2) per axis)   (post-exp, as get_scaling)
    rotations normalise(N(0,I4))       opacities sigmoid(N(0,2^2))  [P,1]
    shs       [P,M,3]: DC (U(0,1)-.5)/0.28209479, rest N(0,.05^2)
    opacity_shift (added to the logit) and aniso_sigma (per-axis log-normal spread) only rescale the same random
    draws: the defaults reproduce the scenes the golden fixtures were generated from.
    """
    g = torch.Generator().manual_seed(int(seed))
    d = torch.randn(P, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-12)
    rho = torch.exp(torch.rand(P, 1, generator=g) * math.log(8.0))
    means = (d * rho).contiguous()
    iso = torch.exp(torch.randn(P, 1, generator=g) * 0.5)
    aniso = torch.exp(torch.randn(P, 3, generator=g) * aniso_sigma)
    scales = (rho * (1.5 / REF_FOCAL) * iso * aniso * scale_mult).contiguous()
    q = torch.randn(P, 4, generator=g)
    rots = (q / q.norm(dim=1, keepdim=True).clamp_min(1e-12)).contiguous()
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0 + opacity_shift).contiguous()
    shs = torch.randn(P, M, 3, generator=g) * 0.05
    shs[:, 0, :] = (torch.rand(P, 3, generator=g) - 0.5) / 0.28209479177387814
    return {"means3D": means, "scales": scales, "rotations": rots, "opacities": opac, "shs": shs.contiguous()}


def make_frustum_scene(P: int, seed: int, W: int, H: int, M: int = 16, sigma_px: float = 1.0, fovx: float = REF_FOVX,
                       margin: float = 0.95, raster: bool = False) -> Dict[str, torch.Tensor]:
    """LucidDreamer-shaped population: what the reference's optimisation loop actually renders (luciddreamer.py:
    283-327).  Its Gaussians come from depth maps re-projected through the training cameras (luciddreamer.py:370-374,
    492; one point per pixel and view), so nearly ALL of them are inside the frustum of a training view, about one
    pixel in size (initial scale = nearest-neighbour spacing, scene/gaussian_model.py:136-137), at 512x512
    (arguments.py:42).  Here: image-plane position uniform over `margin` of the identity camera's view, depth
    exp(U(ln 1, ln 8)), world sigma = sigma_px pixels at that depth times exp(N(0,.3^2)) (and per axis), everything else
    as make_scene.  CPU float32, contiguous, bit-reproducible from the seed."""
    g = torch.Generator().manual_seed(int(seed))
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    uv = (torch.rand(P, 2, generator=g) * 2.0 - 1.0) * margin
    z = torch.exp(torch.rand(P, 1, generator=g) * math.log(8.0))
    means = torch.cat([uv[:, :1] * tx * z, uv[:, 1:] * ty * z, z], 1).contiguous()
    focal_px = W / (2.0 * tx)
    iso = torch.exp(torch.randn(P, 1, generator=g) * 0.3)
    aniso = torch.exp(torch.randn(P, 3, generator=g) * 0.3)
    scales = (z * (sigma_px / focal_px) * iso * aniso).contiguous()
    q = torch.randn(P, 4, generator=g)
    rots = (q / q.norm(dim=1, keepdim=True).clamp_min(1e-12)).contiguous()
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0).contiguous()
    shs = torch.randn(P, M, 3, generator=g) * 0.05
    shs[:, 0, :] = (torch.rand(P, 3, generator=g) - 0.5) / 0.28209479177387814
    out = {"means3D": means, "scales": scales, "rotations": rots, "opacities": opac, "shs": shs.contiguous()}
    if raster:
        # LucidDreamer's own memory order: the point cloud is the concatenation, view after view, of one point per
        # pixel in row-major pixel order (luciddreamer.py:363-372: meshgrid(x, y, indexing='xy') ... reshape(3, -1), then
        # np.concatenate per view :492-495).  Same Gaussians as above, permuted: consecutive blocks of W*H points play the
        # views, each sorted by (pixel row, pixel column) of the identity camera.
        px = ((uv[:, 0] * 0.5 + 0.5) * W).clamp(0, W - 1).floor().long()
        py = ((uv[:, 1] * 0.5 + 0.5) * H).clamp(0, H - 1).floor().long()
        layer = torch.arange(P) // (W * H)
        perm = torch.argsort(layer * (W * H) + py * W + px, stable=True)
        out = {k: v[perm].contiguous() for k, v in out.items()}
    return out


def jitter_poses(n: int, seed: int, max_deg: float = 3.0, max_shift: float = 0.05) -> np.ndarray:
    """n camera poses scattered around the identity pose (small yaw / pitch and translation), the way the reference's
    training views scatter around each generation pose (5 hemisphere jitters per pose, luciddreamer.py:516-518;
    one of them is drawn at random per iteration, luciddreamer.py:291-292).  [n,4,4] COLMAP-convention c2w."""
    rng = np.random.RandomState(int(seed))
    out = np.zeros((n, 4, 4))
    for i in range(n):
        yaw, pitch = np.radians(rng.uniform(-max_deg, max_deg, 2))
        cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        M = np.eye(4)
        M[:3, :3] = Ry @ Rx
        M[:3, 3] = rng.uniform(-max_shift, max_shift, 3)
        out[i] = M
    out[0] = np.eye(4)
    return out


def make_cotangent(H: int, W: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(int(seed) + 7919)
    return (torch.rand(3, H, W, generator=g) * 2.0 - 1.0).contiguous()


# BASELINE.json configs: id -> (P, W, H, active SH degree)
CONFIGS = {
    1: dict(P=10_000, W=256, H=256, sh_degree=0),
    2: dict(P=100_000, W=512, H=512, sh_degree=3),
    3: dict(P=1_000_000, W=1920, H=1080, sh_degree=3),
    4: dict(P=1_000_000, W=1920, H=1080, sh_degree=3, views=64, path="rotate360"),
    5: dict(P=2_000_000, W=1920, H=1080, sh_degree=3, views=8, path="llff"),
}


def config_scene(cfg_id: int, scale_mult: float = 1.0):
    c = CONFIGS[cfg_id]
    return make_scene(c["P"], 1000 + cfg_id, scale_mult=scale_mult), c
