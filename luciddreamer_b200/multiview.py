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

        bucket.begin_step()                 # switch to the pre-zeroed buffer; the other one is cleared in the background
        view_step(params, settings, cot, bucket=bucket)     # forward + backward, gradients land in all buckets
        bucket.end_step()                   # bucket.flat now holds the sum over all views of all ranks

    Nothing of the exchange sits on the step's critical path except the adds themselves:
      * two buffers alternate; the one used by the previous step is zeroed on a side stream while this step computes;
      * the two cross-rank barriers (before the first add: every rank has cleared; after the last add: every rank's adds
        have landed) are one-warp kernels of the library on peer-mapped signal words (GsGrads.peer_signals); the first one
        waits on the context's side stream beside the tile pass of the backward.  `device_sync=False` keeps host-issued
        barriers (needed when some rank has no view in a step)."""

    def __init__(self, P: int, M: int, device, group=None, use_multicast: Optional[bool] = None, device_sync: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.P, self.M = int(P), int(M)
        self.width = 3 + 3 * self.M + 1 + 3 + 4
        widths = (3, 3 * self.M, 1, 3, 4)
        seg = [(self.P * w + 63) // 64 * 64 for w in widths]
        self.group = group if group is not None else dist.group.WORLD
        n = sum(seg)
        self._both = symm_mem.empty(2 * n, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self._both, self.group)
        self._both.zero_()
        self._sig = symm_mem.empty(64, dtype=torch.int32, device=device)      # signal words of the folded barriers
        self._sig_handle = symm_mem.rendezvous(self._sig, self.group)
        self._sig.zero_()
        offs = [0]
        for x in seg[:-1]:
            offs.append(offs[-1] + x)
        self.seg_off = offs
        shapes = ((self.P, 3), (self.P, self.M, 3), (self.P, 1), (self.P, 3), (self.P, 4))
        self.rank, self.world = self.handle.rank, self.handle.world_size
        if use_multicast is None:
            # one multimem.red per element beats `world` peer atomics from 4 ranks on; with 2 ranks the two plain atomics
            # are faster (measured on B200: reduce kernel 37 vs 63 us at 2 ranks, 179 vs 128 us at 8)
            use_multicast = self.world > 2
        mc = 0
        if use_multicast and getattr(self.handle, "has_multicast_support", False):
            mc = int(self.handle.multicast_ptr or 0)
        self._bufs = []
        for b in range(2):
            flat = self._both[b * n:(b + 1) * n]
            views = [flat[o:o + self.P * w].view(*sh) for o, w, sh in zip(offs, widths, shapes)]
            peers = dict(world=self.world, ptrs=[int(p) + b * n * 4 for p in self.handle.buffer_ptrs],
                         mc=(mc + b * n * 4) if mc else 0, seg_off=offs, rank=self.rank,
                         signals=[int(p) for p in self._sig_handle.buffer_ptrs] if device_sync else None,
                         epoch_begin=0, epoch_end=0)
            self._bufs.append((flat, views, peers))
        self.device_sync = bool(device_sync)
        self._cur, self._epoch = 0, 0
        self._side = torch.cuda.Stream(device=device)
        self._zeroed = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in self._zeroed:
            ev.record()
        self._select(0)
        torch.cuda.synchronize(device)
        self.handle.barrier(channel=0)          # every rank's buffers and signal words are zero before the first step

    def _select(self, b):
        self._cur = b
        self.flat, views, self.peers = self._bufs[b]
        self.means3D, self.shs, self.opacities, self.scales, self.rotations = views

    def begin_step(self, first_and_last: bool = True):
        """Call before the first view of the step.  `first_and_last`: the step has exactly one view_step on this rank (it
        then carries both folded barriers); multi_view_step manages the flags itself."""
        main = torch.cuda.current_stream(self.flat.device)
        prev = self._cur
        self._select(prev ^ 1)
        main.wait_event(self._zeroed[self._cur])              # cleared during the previous step
        ev = torch.cuda.Event()
        ev.record(main)                                       # consumers of the previous result enqueued so far
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._bufs[prev][0].zero_()
            self._zeroed[prev].record(self._side)
        self._epoch += 1
        if self.device_sync:
            self.peers["epoch_begin"] = self._epoch if first_and_last else 0
            self.peers["epoch_end"] = self._epoch if first_and_last else 0
        else:
            self.handle.barrier(channel=0)

    def mark(self, first: bool, last: bool):
        """Which folded barriers the NEXT view_step of this step carries (several views per rank and step)."""
        if self.device_sync:
            self.peers["epoch_begin"] = self._epoch if first else 0
            self.peers["epoch_end"] = self._epoch if last else 0

    def end_step(self):
        if not self.device_sync:
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
    if symm and bucket.device_sync and len(settings_list) < world:
        raise RuntimeError("SymmGradBucket(device_sync=True) needs at least one view per rank and step")
    tmp = None
    for n, v in enumerate(views):
        if symm:
            bucket.mark(first=(n == 0), last=(n == len(views) - 1))
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


def render_views_batched(params: dict, settings_list: Sequence, rank: int = 0, world: int = 1, n_streams: int = 2,
                         pair_capacity: Optional[int] = None):
    """Forward-only batch render of this rank's share of `settings_list` (config 4) through the batched C entry point
    gs_forward_views: up to `n_streams` views IN FLIGHT on the context's internal streams (forked from / joined into
    the current stream), nothing blocks until every view is enqueued, then ONE host wait collects all counts.  The
    reference renders view by view and blocks the host once per view (rasterizer_impl.cu:282; luciddreamer.py:250-262).
    All views must share P and the image size.  Returns ({view: (color, depth)}, {view: counts dict})."""
    import ctypes as C

    from . import _native as N
    from . import rasterizer as R

    views = shard_views(len(settings_list), rank, world)
    if not views:
        return {}, {}
    dev = params["means3D"].device
    if not params["means3D"].is_cuda:
        raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    e = torch.empty(0)
    preps = []
    for v in views:
        rs = settings_list[v]
        preps.append(R._prepare(rs.bg, params["means3D"], e, params["opacities"], params["scales"], params["rotations"],
                                rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                                rs.image_width, params["shs"], rs.sh_degree, rs.campos, rs.prefiltered, rs.debug))
    n, P = len(views), preps[0].P
    H, W = preps[0].frame.H, preps[0].frame.W
    if any(p.frame.H != H or p.frame.W != W or p.P != P for p in preps):
        raise RuntimeError("render_views_batched: all views must share one image size and one Gaussian set")
    L = N.lib()
    out, counts = {}, {}
    with R._guard(idx), torch.no_grad():
        ctx, stream = R._ctx(idx), R._raw_stream(idx)
        hint = R._cap_hint.get(idx)
        if pair_capacity is None and hint is None:
            # nothing known about this scene yet: learn the pair count from one ordinary (synchronising) render
            rs = settings_list[views[0]]
            R.GaussianRasterizer(rs)(params["means3D"], e, params["opacities"], shs=params["shs"], scales=params["scales"],
                                     rotations=params["rotations"])
            hint = R._cap_hint.get(idx)
        cap = R._round_cap(pair_capacity if pair_capacity is not None else hint * 1.5 + 4096)
        color = torch.empty((n, 3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((n, 1, H, W), dtype=torch.float32, device=dev)
        for first in range(0, n, 64):                       # a context has 64 status slots
            m = min(64, n - first)
            ns = max(1, min(int(n_streams), m, 8))
            u8 = dict(dtype=torch.uint8, device=dev)
            scr_t = [(torch.empty((L.gs_geom_bytes(P),), **u8), torch.empty((L.gs_image_bytes(W, H),), **u8),
                      torch.empty((L.gs_binning_bytes(cap),), **u8), torch.empty((P,), dtype=torch.int32, device=dev))
                     for _ in range(ns)]
            scr = (N.GsViewScratch * ns)()
            for k, (g_, i_, b_, r_) in enumerate(scr_t):
                scr[k].geom_buffer, scr[k].image_buffer, scr[k].binning_buffer = g_.data_ptr(), i_.data_ptr(), b_.data_ptr()
                scr[k].pair_capacity, scr[k].radii = cap, r_.data_ptr()
            frames = (N.GsFrame * m)()
            res = (N.GsViewResult * m)()
            for k in range(m):
                C.memmove(C.byref(frames[k]), C.byref(preps[first + k].frame), C.sizeof(N.GsFrame))
                res[k].out_color, res[k].out_depth, res[k].radii = color[first + k].data_ptr(), depth[first + k].data_ptr(), None
            rc = L.gs_forward_views(ctx, frames, m, scr, ns, res, stream)
            if rc not in (0, -3):                           # GS_ECAPACITY (-3) is handled per view below
                N.check(rc)
            for k in range(m):
                v = views[first + k]
                c = res[k].counts
                counts[v] = dict(num_rendered=int(c.num_rendered), num_pairs=int(c.num_pairs), num_visible=int(c.num_visible))
                R._note_pairs(idx, c.num_pairs)
                if res[k].status != 0:                      # this view alone needs a larger buffer: ordinary path
                    rs = settings_list[v]
                    col, _r, dep = R.GaussianRasterizer(rs)(params["means3D"], e, params["opacities"], shs=params["shs"],
                                                            scales=params["scales"], rotations=params["rotations"])
                    color[first + k].copy_(col)
                    depth[first + k].copy_(dep)
                out[v] = (color[first + k], depth[first + k])
    return out, counts


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
