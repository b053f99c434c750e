"""View-parallel host logic: the rasterizer path shards over independent camera views.

The reference renders one camera per iteration on one GPU (luciddreamer.py:291-296) and has no multi-GPU code on
this path; what is here is the natural data-parallel generalisation:

  * `shard_views`      round-robin partition of a view list over ranks (no data-path collective)
  * `GradBucket`       ONE flat per-Gaussian gradient buffer [xyz 3 | sh 3M | opacity 1 | scaling 3 | rotation 4]
                       (59 floats at M = 16); the C ABI writes final gradients straight into views of it, so the
                       shared-model step needs exactly one all-reduce and no packing copy
  * `view_step`        forward+backward of one view through the C ABI, gradients into a bucket (no autograd)
  * `allreduce_bucket` the single exchange step of a shared-model optimisation step (NCCL on GPUs, gloo in tests)
  * `DensifyStats`     the statistics scene/gaussian_model.py:405-407 accumulates per view; norms do not commute
                       with a gradient all-reduce, so the per-view norm is taken locally and only the [P,1]
                       accumulators are reduced (sum / sum / max)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Views {v : v mod world == rank} -- GPU g renders views g, g+world, ..."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_views, world))


class GradBucket:
    """Flat gradient bucket; `.flat` is what gets all-reduced, the named views are what the kernels write."""

    def __init__(self, P: int, M: int, device, dtype=torch.float32):
        self.P, self.M = int(P), int(M)
        self.width = 3 + 3 * self.M + 1 + 3 + 4
        widths = (3, 3 * self.M, 1, 3, 4)
        seg = [(self.P * w + 63) // 64 * 64 for w in widths]      # segment starts stay 256-byte aligned
        self.flat = torch.zeros(sum(seg), dtype=dtype, device=device)
        o = 0
        it = iter(seg)

        def take(w, shape):
            nonlocal o
            v = self.flat[o:o + self.P * w].view(*shape)
            o += next(it)
            return v

        self.means3D = take(3, (self.P, 3))
        self.shs = take(3 * self.M, (self.P, self.M, 3))
        self.opacities = take(1, (self.P, 1))
        self.scales = take(3, (self.P, 3))
        self.rotations = take(4, (self.P, 4))

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


class SymmGradBucket(GradBucket):
    """GradBucket whose storage is symmetric memory (one peer-mapped allocation per rank, same layout everywhere), so
    the final backward kernel of every rank can add its visible rows straight into EVERY rank's bucket over NVLink
    (`red.global.add` on peer pointers, or one `multimem.red` per element through the NVSwitch multicast address).
    This replaces the dense 59-floats-per-Gaussian all-reduce of the shared-model step by traffic proportional to
    the number of VISIBLE Gaussians.  Usage per optimisation step, on every rank:

        bucket.begin_step()                 # zero the local bucket, then a cross-rank barrier
        view_step(params, settings, cot, bucket=bucket)     # forward + backward, gradients land in all buckets
        bucket.end_step()                   # cross-rank barrier: bucket.flat now holds the sum over all views
    """

    def __init__(self, P: int, M: int, device, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.P, self.M = int(P), int(M)
        self.width = 3 + 3 * self.M + 1 + 3 + 4
        widths = (3, 3 * self.M, 1, 3, 4)
        seg = [(self.P * w + 63) // 64 * 64 for w in widths]
        self.group = group if group is not None else dist.group.WORLD
        self.flat = symm_mem.empty(sum(seg), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.flat, self.group)
        self.flat.zero_()
        offs = [0]
        for x in seg[:-1]:
            offs.append(offs[-1] + x)
        self.seg_off = offs
        shapes = ((self.P, 3), (self.P, self.M, 3), (self.P, 1), (self.P, 3), (self.P, 4))
        views = [self.flat[o:o + self.P * w].view(*sh) for o, w, sh in zip(offs, widths, shapes)]
        self.means3D, self.shs, self.opacities, self.scales, self.rotations = views
        mc = 0
        if use_multicast and getattr(self.handle, "has_multicast_support", False):
            mc = int(self.handle.multicast_ptr or 0)
        self.peers = dict(world=self.handle.world_size, ptrs=[int(p) for p in self.handle.buffer_ptrs], mc=mc,
                          seg_off=offs)

    def begin_step(self):
        self.flat.zero_()
        self.handle.barrier(channel=0)

    def end_step(self):
        self.handle.barrier(channel=1)


class DensifyStats:
    """xyz_gradient_accum / denom / max_radii2D of scene/gaussian_model.py:151-155,405-407 and
    luciddreamer.py:306-312, kept so that they can be reduced across ranks."""

    def __init__(self, P: int, device):
        self.xyz_gradient_accum = torch.zeros(P, 1, device=device)
        self.denom = torch.zeros(P, 1, device=device)
        self.max_radii2D = torch.zeros(P, device=device)

    def add_view(self, dL_dmeans2D: torch.Tensor, radii: torch.Tensor) -> None:
        if dL_dmeans2D.is_cuda:               # one fused pass on the GPU (gs_densify_stats); torch ops only for CPU tests
            from . import _native as N
            from . import rasterizer as R
            if not (dL_dmeans2D.is_contiguous() and radii.is_contiguous() and radii.dtype == torch.int32
                    and dL_dmeans2D.dtype == torch.float32 and dL_dmeans2D.shape[-1] == 3):
                raise RuntimeError("add_view: expects contiguous dL_dmeans2D [P,3] f32 and radii [P] int32")
            idx = dL_dmeans2D.device.index
            with R._guard(idx):
                N.check(N.lib().gs_densify_stats(R._ctx(idx), radii.numel(), radii.data_ptr(), dL_dmeans2D.data_ptr(),
                                                 self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(),
                                                 self.max_radii2D.data_ptr(), R._raw_stream(idx)))
            return
        vis = radii > 0
        self.max_radii2D[vis] = torch.max(self.max_radii2D[vis], radii[vis].to(self.max_radii2D.dtype))
        self.xyz_gradient_accum[vis] += torch.norm(dL_dmeans2D[vis, :2], dim=-1, keepdim=True)
        self.denom[vis] += 1

    def allreduce(self, group=None) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)


def allreduce_bucket(bucket: GradBucket, group=None, average: bool = False, async_op: bool = False):
    """Sum (or mean) of the per-rank gradient buckets: the only collective of a shared-model step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if average and async_op:
        raise ValueError("allreduce_bucket: average=True needs the result; divide after work.wait() when async_op=True")
    work = dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if average and not async_op:
        bucket.flat.div_(dist.get_world_size(group))
    return work


def view_step(params: dict, settings, cotangent: torch.Tensor, bucket: Optional[GradBucket] = None,
              means2D_grad: Optional[torch.Tensor] = None):
    """Forward + backward of ONE view straight through the C ABI (no autograd graph).

    params: dict(means3D, shs, opacities, scales, rotations) of CUDA tensors (activated values, like the
    reference's render() passes them).  settings: GaussianRasterizationSettings.  Gradients of the five
    parameter tensors are written into `bucket` (allocated if None).  Returns (color, depth, radii, bucket,
    dL_dmeans2D)."""
    from . import rasterizer as R

    rs = settings
    e = torch.empty(0)
    prep = R._prepare(rs.bg, params["means3D"], e, params["opacities"], params["scales"], params["rotations"],
                      rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                      rs.image_width, params["shs"], rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
    _nr, color, depth, radii, geom, binning, img, cap = R._forward_impl(prep)
    if bucket is None:
        bucket = GradBucket(prep.P, prep.M, prep.device)
    dm2 = means2D_grad if means2D_grad is not None else torch.empty((prep.P, 3), device=prep.device)
    if isinstance(bucket, SymmGradBucket):
        R._backward_impl(prep, radii, geom, binning, img, cap, cotangent, False, False, out=dict(dm2=dm2),
                         peers=bucket.peers)
    else:
        R._backward_impl(prep, radii, geom, binning, img, cap, cotangent, False, False,
                         out=dict(dm3=bucket.means3D, dm2=dm2, dop=bucket.opacities, dsh=bucket.shs, dsc=bucket.scales,
                                  drot=bucket.rotations))
    return color, depth, radii, bucket, dm2


def multi_view_step(params: dict, settings_list: Sequence, cotangents: Sequence, bucket: GradBucket, rank: int = 0,
                    world: int = 1, stats: Optional[DensifyStats] = None):
    """This rank's share of a shared-model optimisation step over `settings_list` (config 5): forward+backward of
    every local view, parameter gradients SUMMED over the views into `bucket`.  With a SymmGradBucket the sum also
    runs over all ranks (call bucket.begin_step() before and bucket.end_step() after); with a plain GradBucket call
    allreduce_bucket(bucket) afterwards.  Returns {view: (color, depth, radii)}."""
    out = {}
    views = shard_views(len(settings_list), rank, world)
    symm = isinstance(bucket, SymmGradBucket)
    tmp = None
    for n, v in enumerate(views):
        if symm or n == 0:
            target = bucket
        else:
            if tmp is None:
                tmp = GradBucket(bucket.P, bucket.M, bucket.flat.device)
            target = tmp
        color, depth, radii, _b, dm2 = view_step(params, settings_list[v], cotangents[v], bucket=target)
        if target is tmp:
            bucket.flat.add_(tmp.flat)
        if stats is not None:
            stats.add_view(dm2, radii)
        out[v] = (color, depth, radii)
    if not views and not symm:
        bucket.flat.zero_()
    return out


def render_views(params: dict, settings_list: Sequence, rank: int = 0, world: int = 1):
    """Forward-only batch render of this rank's share of `settings_list` (config 4).  Returns {view: (color, depth)}."""
    from . import rasterizer as R

    out = {}
    with torch.no_grad():
        for v in shard_views(len(settings_list), rank, world):
            rast = R.GaussianRasterizer(settings_list[v])
            color, _radii, depth = rast(params["means3D"], torch.empty(0), params["opacities"], shs=params["shs"],
                                        scales=params["scales"], rotations=params["rotations"])
            out[v] = (color, depth)
    return out
