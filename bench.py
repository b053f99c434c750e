#!/usr/bin/env python
"""bench.py -- forward+backward Mrays/s of the 3DGS rasterizer at BASELINE.json config 3
(1 M Gaussians, 1920x1080, SH degree 3, synthetic shell scene of SURVEY.md 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scene shell|stress]

One "step" = one forward + one backward of one 1080p view through the public operator API
(GaussianRasterizer autograd.Function), parameters resident in HBM.  N > 1 (torchrun, one rank per GPU): every
rank renders its own view of the same replicated model per step (weak scaling over views, no data-path
collective); value = total rays of all ranks / max-over-ranks device time.  Prints ONE JSON line on rank 0.

--impl reference times the UNMODIFIED reference CUDA rasterizer (oracle/_ref/libref_rasterizer.so, compiled
from the reference's own sources) on the same GPU, same inputs, same protocol; if that library is absent it
falls back to the CPU oracle port.  Under torchrun (N > 1) rank 0 alone runs it (one GPU, "n_gpus": 1 in its line) and
the other ranks exit 0, as the bench contract asks.  The reference has no CPU implementation of this path (SURVEY.md
0.3), so `cpu_baseline` is always the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from luciddreamer_b200 import synthetic as syn

METRIC = "forward+backward Mrays/s @1080p/1M Gaussians"
CFG_ID = 3


# ------------------------------------------------------------------------------------------------ helpers

def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def view_pose(rank: int, step: int = 0):
    """Per-rank camera: identity for rank 0, other ranks yaw around the vertical axis on the
    rotate360 path (camera at the origin) so every GPU renders a different view of the same shell."""
    if rank == 0:
        return None
    return syn.rotate360_poses(64)[(rank * 8) % 64]


def alg_bytes(P, Pv, pairs, N, G, M):
    """Algorithmic (compulsory) HBM bytes per launch of each kernel -- DESIGN.md section 4 states the model."""
    dense = 2 * Pv > P                    # the device picks k_grad_dense over (k_grad_vis, k_grad_write) on this test
    return {
        "k_project": 12 * P + 4 * P + Pv * (12 + 16 + 4 + 32 + 4),
        "k_count_tiles": Pv * (4 + 32 + 4) + 4 * pairs + 12 * G,   # + the scan of the G tile counts by its last CTA
        "k_tile_scan": 0,                                # a launch of its own only for empty models / GS_SCAN_KERNEL=1
        "k_shade_emit": Pv * (4 + 32 + 12 + 12 * M + 4 + 32 + 48) + 12 * pairs,
        "k_tile_sort": 8 * pairs + 4 * pairs,
        "k_tile_sort_big": 0,
        "k_blend_fwd": 4 * pairs + 48 * Pv + 24 * N,
        "k_blend_bwd": 20 * N + 4 * pairs + 48 * Pv + 72 * Pv,
        "k_grad_vis": 0 if dense else Pv * (4 + 96 + 16 + 12 + 12 + 16 + 12 * M + 12 + 12 + 12 * M + 4 + 12 + 16),
        "k_grad_write": 0,                               # replaced by k_fill_zero + the scatter of k_grad_vis (sparse regime)
        "k_grad_dense": (4 * P + Pv * (96 + 4 + 12 + 12 + 16 + 12 * M) + (12 + 12 + 12 * M + 4 + 12 + 16) * P) if dense else 0,
        # zero-fill of the dense gradient tensors beside the tile pass (gs_backward_prefill); its time is measured while
        # k_blend_bwd runs next to it
        "k_fill_zero": 0 if dense else (12 + 12 + 12 * M + 4 + 12 + 16) * P,
    }


def make_workload(args, rank):
    c = syn.CONFIGS[CFG_ID]
    P, W, H, D = c["P"], c["W"], c["H"], c["sh_degree"]
    if args.P:
        P = args.P
    if args.W and args.H:
        W, H = args.W, args.H
    scale_mult = 4.0 if args.scene == "stress" else 1.0
    if args.scene == "frustum":          # LucidDreamer-shaped population (synthetic.make_frustum_scene), raster memory order
        scene = syn.make_frustum_scene(P, 1000 + CFG_ID, W, H, raster=not args.shuffled)
    else:
        scene = syn.make_scene(P, 1000 + CFG_ID, scale_mult=scale_mult)
    cam = syn.make_camera(W, H, c2w=view_pose(rank))
    cot = syn.make_cotangent(H, W, 1000 + CFG_ID)
    wl = dict(P=P, W=W, H=H, D=D, scale_mult=scale_mult, cams=None)
    if getattr(args, "random_camera", 0):
        # one of N jittered training views per step, drawn pseudo-randomly (luciddreamer.py:291-292)
        wl["cams"] = [syn.make_camera(W, H, c2w=m) for m in syn.jitter_poses(args.random_camera, 4711 + rank)]
    return scene, cam, cot, wl


# ------------------------------------------------------------------------------------------- implementations

class Ours:
    name = "ours"

    def __init__(self, scene, cam, dev, D):
        from luciddreamer_b200 import rasterizer as R
        self.R, self.dev, self.D = R, dev, D
        self.leaves = {k: scene[k].to(dev).requires_grad_(True)
                       for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        self.P = scene["means3D"].shape[0]
        self.m2 = torch.zeros(self.P, 3, device=dev, requires_grad=True)
        self.bg = torch.zeros(3, device=dev)
        self.vm, self.pm, self.cp = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
        self.cam = cam
        self.settings = R.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                                        self.bg, 1.0, self.vm, self.pm, D, self.cp, False, False)
        self.rast = R.GaussianRasterizer(self.settings)
        # our kernels only, no torch kernels (profiles/r02_launches_ours.csv + the fold of the tile scan):
        # k_clear_words, k_project, k_count_tiles (its last CTA scans the tile histogram), k_shade_emit, k_tile_sort,
        # k_tile_sort_mid, k_tile_sort_big, k_blend_fwd, k_fill_zero (side stream) | k_blend_bwd, k_grad_vis
        # (+ k_grad_dense unless the host can prove the sparse regime)
        self.kernels_per_step = 11
        self.last = None
        self.rasts, self.order, self.k = None, None, 0

    def set_cameras(self, cams, seed=99):
        """Random-camera mode: a rasterizer per training view (camera tensors resident), one drawn per step."""
        R, dev = self.R, self.dev
        self.rasts = [R.GaussianRasterizer(R.GaussianRasterizationSettings(
            c.image_height, c.image_width, c.tanfovx, c.tanfovy, self.bg, 1.0, c.viewmatrix.to(dev), c.projmatrix.to(dev),
            self.D, c.campos.to(dev), False, False)) for c in cams]
        self.order = np.random.RandomState(seed).randint(0, len(cams), size=4096)

    def forward(self):
        L = self.leaves
        self.last = None                       # drop the previous step's autograd graph before building the next one
        for t in L.values():
            t.grad = None
        self.m2.grad = None
        rast = self.rast
        if self.rasts is not None:
            rast = self.rasts[self.order[self.k & 4095]]
            self.k += 1
        color, radii, depth = rast(L["means3D"], self.m2, L["opacities"], shs=L["shs"], scales=L["scales"],
                                   rotations=L["rotations"])
        self.last = (color, radii)
        return color

    def backward(self, color, cot):
        torch.autograd.backward(color, grad_tensors=cot)

    def step(self, cot):
        color = self.forward()
        self.backward(color, cot)
        return color

    def bind_camera(self, d_cam):          # e2e: the operator reads the camera straight from the upload buffer
        R, cam = self.R, self.cam
        self.vm, self.pm, self.cp = d_cam[0:16].view(4, 4), d_cam[16:32].view(4, 4), d_cam[32:35]
        self.settings = R.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                                        self.bg, 1.0, self.vm, self.pm, self.D, self.cp, False, False)
        self.rast = R.GaussianRasterizer(self.settings)

    def stats(self):
        idx = self.dev.index or 0
        pairs = self.R._last_pairs.get(idx, 0)
        radii = self.last[1]
        return dict(P_vis=int((radii > 0).sum().item()), pairs=int(pairs))

    def profile(self, cot, n=5):
        import ctypes as C
        from luciddreamer_b200 import _native as N
        L = N.lib()
        ctx = self.R._ctx(self.dev.index or 0)
        nk = L.gs_profile_num_kernels()
        names = [L.gs_profile_kernel_name(i).decode() for i in range(nk)]
        acc = np.zeros(nk)
        L.gs_profile_enable(ctx, 1)
        for _ in range(n):
            self.step(cot)
            buf = (C.c_float * nk)()
            N.check(L.gs_profile_read(ctx, buf))
            acc += np.maximum(np.array(buf[:]), 0.0)
        L.gs_profile_enable(ctx, 0)
        return dict(zip(names, (acc / n).tolist()))


class RefCuda:
    """The unmodified reference CUDA rasterizer (oracle/_ref), driven like its torch binding drives it."""
    name = "reference"

    def __init__(self, scene, cam, dev, D):
        from oracle import ref_cuda
        self.rc, self.dev, self.D = ref_cuda, dev, D
        self.ctx = ref_cuda.RefContext()
        self.t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        self.bg = torch.zeros(3, device=dev)
        self.vm, self.pm, self.cp = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
        self.cam = cam
        self.kernels_per_step = 0
        self.cams, self.order, self.k = None, None, 0

    def set_cameras(self, cams, seed=99):
        dev = self.dev
        self.cams = [(c, c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev)) for c in cams]
        self.order = np.random.RandomState(seed).randint(0, len(cams), size=4096)

    def forward(self):
        t, cam, vm, pm, cp = self.t, self.cam, self.vm, self.pm, self.cp
        if self.cams is not None:
            cam, vm, pm, cp = self.cams[self.order[self.k & 4095]]
            self.k += 1
        R, color, depth, radii = self.rc.rasterize_gaussians(
            self.ctx, self.bg, t["means3D"], None, t["opacities"], t["scales"], t["rotations"], 1.0, None, vm,
            pm, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, t["shs"], self.D, cp)
        self.last = (color, radii, R)
        return color

    def backward(self, color, cot):
        self.rc.rasterize_gaussians_backward(self.ctx, self.last[1], cot)

    def step(self, cot):
        color = self.forward()
        self.backward(color, cot)
        return color

    def bind_camera(self, d_cam):
        self.vm, self.pm, self.cp = d_cam[0:16].view(4, 4), d_cam[16:32].view(4, 4), d_cam[32:35]

    def stats(self):
        return dict(P_vis=int((self.last[1] > 0).sum().item()), pairs=int(self.last[2]))


def time_steps(impl, cot, steps, warmup, world):
    for _ in range(warmup):
        impl.step(cot)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        impl.step(cot)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def time_e2e(impl, cam, target_u8_cpu, steps, warmup, world, fused_loss, graphed=False):
    """Same metric end to end, shaped like one training iteration of the reference (luciddreamer.py:291-304): every
    step uploads that step's inputs from pinned host memory -- the camera (viewmatrix, projmatrix, campos: 35 floats)
    and the uint8 target image [H,W,3] -- computes the L1 photometric loss and its gradient on the device, runs
    forward+backward through the operator API, and reads the scalar loss back.  Target uploads of step k+1 overlap the
    backward of step k on a copy stream (double buffered); everything is inside the timed region.
    fused_loss: our fused L1 kernel (luciddreamer_b200.losses); the reference arm uses plain torch ops.
    graphed: forward -> loss -> backward of a step is ONE CUDA graph (luciddreamer_b200.graphs.GraphedStep, the public
    API for it; one graph per target buffer).  The uploads and the loss read-back stay outside the graph, per step.  The
    reference cannot be captured: its forward blocks on a D2H copy (rasterizer_impl.cu:282)."""
    dev = impl.dev
    H, W = cam.image_height, cam.image_width
    h_tgt = target_u8_cpu.contiguous().pin_memory()
    h_cam = torch.cat([cam.viewmatrix.reshape(-1), cam.projmatrix.reshape(-1), cam.campos.reshape(-1)]).contiguous().pin_memory()
    d_tgt = [torch.empty_like(target_u8_cpu, device=dev) for _ in range(2)]
    d_cam = torch.empty(35, device=dev)
    h_loss = torch.zeros(steps + warmup, dtype=torch.float32).pin_memory()
    copy_s = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    h2d = h_tgt.numel() + 35 * 4
    d2h = 4
    if fused_loss:
        from luciddreamer_b200 import losses

    dbg = os.environ.get("GS_E2E_SKIP", "")          # diagnostics only: "u" skips the uploads, "c" the camera copy, "l" the loss copy

    def upload(k):
        b = k & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_free[b])
            d_tgt[b].copy_(h_tgt, non_blocking=True)
            ev_up[b].record(copy_s)

    graphs = None

    def run(n, base):
        if "u" not in dbg:
            upload(base)
        for k in range(base, base + n):
            b = k & 1
            if "c" not in dbg:
                d_cam.copy_(h_cam, non_blocking=True)             # 140 bytes, main stream; the operator reads it in place
            if "u" not in dbg:
                main.wait_event(ev_up[b])
            if graphs is not None:
                loss = graphs[b].replay()
                if "l" not in dbg:
                    h_loss[k].copy_(loss.reshape(()), non_blocking=True)
                if "u" not in dbg:
                    ev_free[b].record(main)
                    if k + 1 < base + n:
                        upload(k + 1)
                continue
            color = impl.forward()
            if fused_loss:
                loss, cot = losses.l1_loss_with_grad(color, d_tgt[b])
            else:
                diff = color.detach() - d_tgt[b].permute(2, 0, 1).float().div_(255.0)
                loss = diff.abs().mean()
                cot = torch.sign(diff).div_(diff.numel())
            impl.backward(color, cot)
            h_loss[k].copy_(loss.reshape(()), non_blocking=True)
            ev_free[b].record(main)
            if k + 1 < base + n:                                  # next target: in flight during this step's backward;
                upload(k + 1)                                     # issued last so the host reaches backward() early

    impl.bind_camera(d_cam)
    d_cam.copy_(h_cam)
    for b in range(2):
        d_tgt[b].copy_(h_tgt)
    if graphed:
        from luciddreamer_b200.graphs import GraphedStep
        # ONE host->device copy per step: the step's camera (35 floats) and target image travel together in one pinned
        # staging buffer [256 B header | H*W*3 B image] on the copy stream -- a separate 140-byte copy on the compute stream
        # would queue behind the next step's image upload in the same copy engine
        nimg = h_tgt.numel()
        h_in = torch.empty(256 + nimg, dtype=torch.uint8).pin_memory()
        h_in[:140].copy_(h_cam.view(torch.uint8))
        h_in[256:].copy_(h_tgt.reshape(-1))
        d_in = [torch.empty(256 + nimg, dtype=torch.uint8, device=dev) for _ in range(2)]
        for b in range(2):
            d_in[b].copy_(h_in)
        h2d = 256 + nimg

        def upload(k):                               # noqa: F811 (replaces the two-copy version above)
            b = k & 1
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_free[b])
                d_in[b].copy_(h_in, non_blocking=True)
                ev_up[b].record(copy_s)

        def make_fn(b):
            tgt = d_in[b][256:].view(H, W, 3)

            def fn():
                color = impl.forward()
                loss, cot = losses.l1_loss_with_grad(color, tgt)
                impl.backward(color, cot)
                return loss
            return fn
        graphs = []
        for b in range(2):
            impl.bind_camera(d_in[b][:140].view(torch.float32))   # graph b reads the camera of staging buffer b
            graphs.append(GraphedStep(make_fn(b), warmup=2))
        dbg = dbg + "c"                              # no separate camera copy in this mode
    for b in range(2):
        ev_free[b].record(main)
    run(warmup, 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(steps, warmup)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    ms = max(e0.elapsed_time(e1), wall_ms)      # the host-visible result is only there after the final sync
    if graphs is not None:
        for g in graphs:
            g.check()                           # raises if a replay overflowed the static pair capacity
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, h2d, d2h, float(h_loss[-1])


def cpu_baseline(scene, cam, cot, D, budget_s=30.0, max_steps=5):
    """Oracle port (C + OpenMP, fp32 faithful mode) on the host cores: full forward+backward steps of the bench workload,
    threads pinned (OMP_PROC_BIND / OMP_PLACES, set in main() before libgomp starts), MEDIAN of the steps reported."""
    from oracle import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    oracle.set_threads(threads)
    args = (torch.zeros(3), scene["means3D"], None, scene["opacities"], scene["scales"], scene["rotations"], 1.0, None,
            cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width, scene["shs"],
            D, cam.campos)
    cotn = cot.numpy()
    times = []
    t_all = time.perf_counter()
    while len(times) < max_steps and (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        f = oracle.rasterize_gaussians(*args)
        oracle.rasterize_gaussians_backward(f, cotn)
        times.append(time.perf_counter() - t0)
        f.free()
    best, med = min(times), float(np.median(times))
    rays = cam.image_height * cam.image_width
    return {"value": rays / med / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} full forward+backward steps of the bench workload (oracle/gs_oracle.c, OpenMP "
                      f"{threads} pinned threads, fp32): median {med * 1e3:.0f} ms/step, best {best * 1e3:.0f}, "
                      f"worst {max(times) * 1e3:.0f}"}


def loss_leg(H, W, dev, iters=20):
    """SURVEY 8f-1 ("next" row): the reference's training loss 0.8*L1 + 0.2*(1-SSIM) forward+backward at the bench
    resolution -- our fused kernels vs the same formula in torch ops (what utils/loss.py runs), same inputs."""
    import torch.nn.functional as F
    from luciddreamer_b200 import losses
    g1 = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
    g1 = (g1 / g1.sum()).unsqueeze(1)
    win = (g1 @ g1.t()).expand(3, 1, 11, 11).contiguous().to(dev)

    def torch_ssim(x, y):                      # the formula of utils/loss.py:38-69 in torch ops (baseline only)
        x4, y4 = x.unsqueeze(0), y.unsqueeze(0)
        conv = lambda t: F.conv2d(t, win, padding=5, groups=3)
        mu1, mu2 = conv(x4), conv(y4)
        s1, s2, s12 = conv(x4 * x4) - mu1 * mu1, conv(y4 * y4) - mu2 * mu2, conv(x4 * y4) - mu1 * mu2
        C1, C2 = 0.01 ** 2, 0.03 ** 2
        return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()

    g = torch.Generator().manual_seed(7)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    img = (gt + 0.1 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)

    def ours():
        return losses.photometric_loss_with_grad(img, gt, 0.2)

    def torch_ref():
        x = img.clone().requires_grad_(True)
        l = 0.8 * (x - gt).abs().mean() + 0.2 * (1.0 - torch_ssim(x, gt))
        l.backward()
        return l, x.grad

    out = {}
    for name, fn in (("fused_ms", ours), ("torch_ops_ms", torch_ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            r = fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters
    a, b = ours(), torch_ref()
    out["rel_err_grad_vs_torch"] = float(((a[1] - b[1]).norm() / b[1].norm()).item())
    out["what"] = "0.8*L1 + 0.2*(1-SSIM) loss + gradient at the bench resolution (luciddreamer.py:301-304)"
    return out


def adam_leg(scene, dev, iters=10):
    """SURVEY 8f-2 ("next" row): one optimiser step over all parameters of the bench model -- our fused
    activation-backward + Adam launch vs autograd through the activations + torch.optim.Adam (what the reference runs)."""
    from luciddreamer_b200 import multiview as MV
    from luciddreamer_b200 import optim
    P = scene["means3D"].shape[0]
    g = torch.Generator().manual_seed(11)
    raw = dict(xyz=scene["means3D"].clone(), f_dc=scene["shs"][:, :1].contiguous(), f_rest=scene["shs"][:, 1:].contiguous(),
               opacity=torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)), scaling=torch.log(scene["scales"]),
               rotation=scene["rotations"].clone())
    dv = {k: v.to(dev).contiguous() for k, v in raw.items()}
    bucket = MV.GradBucket(P, 16, dev)
    bucket.flat.copy_(torch.randn(bucket.flat.numel(), generator=g) * 1e-3)
    mine = optim.FusedGaussianAdam(dv["xyz"], dv["f_dc"], dv["f_rest"], dv["opacity"], dv["scaling"], dv["rotation"])
    tp = {k: torch.nn.Parameter(v.clone()) for k, v in dv.items()}
    opt = torch.optim.Adam([{"params": [tp[k]], "lr": optim.DEFAULT_LRS[k]} for k in tp], lr=0.0, eps=1e-15)

    def torch_step():
        acts = (tp["xyz"], torch.cat((tp["f_dc"], tp["f_rest"]), dim=1), torch.sigmoid(tp["opacity"]),
                torch.exp(tp["scaling"]), torch.nn.functional.normalize(tp["rotation"]))
        opt.zero_grad(set_to_none=True)
        torch.autograd.backward(acts, [bucket.means3D, bucket.shs, bucket.opacities, bucket.scales, bucket.rotations])
        opt.step()

    out = {}
    for name, fn in (("fused_ms", lambda: mine.step(bucket)), ("torch_adam_ms", torch_step)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters
    nbytes = 59 * P * 4 * 7
    out["fused_GBps"] = nbytes / (out["fused_ms"] * 1e-3) / 1e9
    out["what"] = "activation chain rule + Adam over all 59 parameters/Gaussian (28 B/parameter of compulsory traffic)"
    return out


SHAPED = [
    # BASELINE.json configs[1]: the reference's own resolution, 100 k Gaussians (shell scene, seed 1002)
    ("config2_100k_512", dict(scene="shell", P=100_000, W=512, H=512, seed=1002, cams=0)),
    # what the reference's optimisation loop renders (luciddreamer.py:283-327, arguments.py:42-46): 512x512, 1-2 M Gaussians
    # nearly all inside the frustum, about a pixel in size, a random training camera per iteration
    ("frustum_1M_512_randcam", dict(scene="frustum", P=1_000_000, W=512, H=512, seed=2001, cams=16)),
    ("frustum_2M_512_randcam", dict(scene="frustum", P=2_000_000, W=512, H=512, seed=2002, cams=16)),
    ("frustum_1M_1080p", dict(scene="frustum", P=1_000_000, W=1920, H=1080, seed=2003, cams=0)),
    # the same population in RANDOM memory order (the frustum rows above are in LucidDreamer's raster order)
    ("frustum_1M_512_shuffled", dict(scene="frustum", P=1_000_000, W=512, H=512, seed=2001, cams=0, shuffled=True)),
]


def graphed_steps(impl, cam, cot, ncams, seed, steps, warmup):
    """forward+backward of `impl` captured once (static pair capacity, camera read from a device buffer) and replayed; with
    ncams > 0 a jittered training camera is copied into that buffer before every replay (device-to-device, 35 floats)."""
    from luciddreamer_b200.graphs import GraphedStep
    dev = impl.dev
    cams = [cam] + ([syn.make_camera(cam.image_width, cam.image_height, c2w=m) for m in syn.jitter_poses(ncams, seed)] if ncams else [])
    table = torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos.reshape(-1)]) for c in cams]).to(dev)
    d_cam = table[0].clone()
    impl.rasts = None
    impl.bind_camera(d_cam)
    g = GraphedStep(lambda: impl.step(cot), warmup=2, headroom=2.5)
    order = np.random.RandomState(99).randint(0, len(cams), size=warmup + steps)

    def run(k0, n):
        for k in range(k0, k0 + n):
            if ncams:
                d_cam.copy_(table[order[k]], non_blocking=True)
            g.replay()
    run(0, warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(warmup, steps)
    e1.record(); torch.cuda.synchronize()
    g.check()
    return e0.elapsed_time(e1) / steps


def shaped_legs(impl_cls, dev, D=3, steps=10, warmup=3, only=None):
    """`next` rows of the measurement: forward+backward on the reference's real workload shape and on
    BASELINE config 2, same protocol as the headline (CUDA events, parameters resident in HBM).  The product arm and the
    reference arm (`--impl reference`) each report their own side."""
    out = {}
    for name, s in SHAPED:
        if only and name not in only:
            continue
        try:
            W, H = s["W"], s["H"]
            scene = (syn.make_frustum_scene(s["P"], s["seed"], W, H, raster=not s.get("shuffled", False))
                     if s["scene"] == "frustum" else syn.make_scene(s["P"], s["seed"]))
            cam = syn.make_camera(W, H)
            cot = syn.make_cotangent(H, W, s["seed"]).to(dev)
            impl = impl_cls(scene, cam, dev, D)
            if s["cams"]:
                impl.set_cameras([syn.make_camera(W, H, c2w=m) for m in syn.jitter_poses(s["cams"], s["seed"])])
            ms = time_steps(impl, cot, steps, warmup, 1) / steps
            st = impl.stats()
            out[name] = {"ms_per_step": ms, "value": W * H / ms / 1e3, "unit": "Mrays/s", "P": s["P"], "W": W, "H": H,
                         "P_vis": st["P_vis"], "pairs": st["pairs"], "random_cameras": s["cams"], "steps": steps}
            if hasattr(impl, "profile"):
                out[name]["kernels_ms"] = impl.profile(cot, n=3)
                try:                                 # the same step as ONE CUDA graph (graphs.GraphedStep), camera rewritten per step
                    out[name]["graph_ms_per_step"] = graphed_steps(impl, cam, cot, s["cams"], s["seed"], steps, warmup)
                    out[name]["graph_value"] = W * H / out[name]["graph_ms_per_step"] / 1e3
                except Exception as ex:
                    out[name]["graph_error"] = str(ex)[:200]
            del impl, scene, cot
            torch.cuda.empty_cache()
        except Exception as ex:                      # a `next` row must never take the headline line down
            out[name] = {"error": str(ex)[:200]}
    return out


def _tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def knn_leg(impl="ours"):
    """simple_knn.distCUDA2 over 1 M points (depth-map shaped cloud).  The `ours` arm times our kernels, the
    `reference` arm the reference's own simple_knn.cu compiled unmodified (oracle/_ref) -- each in its own bench run."""
    out = _tool("knn_bench").run(1_000_000, which=("depthmap",), impls=(impl,))["depthmap"]
    out["what"] = "mean squared distance to the 3 nearest neighbours, 1 M points (bit-identical outputs: tests/)"
    return out


def video_leg(impl="ours"):
    """forward-only clip at the bench resolution.  `ours`: device-side packing + one host copy;
    `reference`: the reference rasterizer inside the per-frame loop of luciddreamer.py:250-262."""
    out = _tool("video_bench").run(16, impls=(impl,))
    out["what"] = "rotate360 clip, uint8 frames + masked depth + dmin/dmax delivered to host memory"
    return out


def shared_model_leg(scene, cam, cot, dev, D, steps, warmup, world):
    """Config-5 style step on every rank: forward+backward of its view, then the ONE exchange of the path -- the sum
    of the per-Gaussian gradients over all views.  Two implementations are timed:
      fused  the final backward kernel adds the rows of the VISIBLE Gaussians straight into every rank's symmetric
             bucket (multimem.red through the NVSwitch multicast address, or peer stores), bracketed by two
             device-side barriers -- traffic ~ 59 floats x P_vis per rank
      nccl   dense local bucket + NCCL all-reduce(sum) of 59 floats x P (the baseline this replaces)"""
    from luciddreamer_b200 import multiview as MV
    from luciddreamer_b200 import rasterizer as R
    params = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    bg = torch.zeros(3, device=dev)
    rs = R.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, bg, 1.0,
                                         cam.viewmatrix.to(dev), cam.projmatrix.to(dev), D, cam.campos.to(dev), False, False)
    P, M = params["means3D"].shape[0], params["shs"].shape[1]
    dm2 = torch.empty(P, 3, device=dev)
    rays = cam.image_height * cam.image_width

    def timed(step):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    dense = MV.GradBucket(P, M, dev)

    def step_nccl():
        MV.view_step(params, rs, cot, bucket=dense, means2D_grad=dm2)
        MV.allreduce_bucket(dense)

    ms_nccl = timed(step_nccl)
    out = {"nccl_allreduce": {"ms_per_step": ms_nccl, "value": world * rays / ms_nccl / 1e3, "unit": "Mrays/s",
                              "bytes_reduced_per_rank": dense.nbytes()}}
    try:
        best = None
        for key, mc in (("fused_peer_reduce", True), ("fused_peer_stores", False)):
            symm = MV.SymmGradBucket(P, M, dev, use_multicast=mc)
            if not mc or symm.peers["mc"]:
                def step_fused():
                    symm.begin_step()
                    MV.view_step(params, rs, cot, bucket=symm, means2D_grad=dm2)
                    symm.end_step()

                ms_f = timed(step_fused)
                torch.cuda.synchronize()
                err = float(((symm.flat - dense.flat).norm() / dense.flat.norm().clamp_min(1e-30)).item())
                out[key] = {"ms_per_step": ms_f, "value": world * rays / ms_f / 1e3, "unit": "Mrays/s",
                            "multicast": bool(symm.peers["mc"]), "rel_err_vs_nccl": err}
                if best is None or ms_f < best:
                    best = ms_f
            del symm
            torch.cuda.empty_cache()
        out["ms_per_step"], out["value"], out["unit"] = best, world * rays / best / 1e3, "Mrays/s"
    except Exception as ex:                     # symmetric memory unavailable: the NCCL leg is the result
        out["fused_peer_reduce"] = {"unavailable": str(ex)[:200]}
        out["ms_per_step"], out["value"], out["unit"] = ms_nccl, world * rays / ms_nccl / 1e3, "Mrays/s"
    return out


def _maxrank(ms, world):
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def config4_leg(dev, rank, world, impl="ours", D=3, n_views=64, reps=3):
    """BASELINE config 4 at its stated size: 1 M Gaussians, 64 views of the rotate360 path at 1920x1080, forward only,
    views sharded round-robin over the ranks (strong scaling: 64 views in total whatever N is).  Ours: the batched C entry
    point gs_forward_views with 1 and with 2 views in flight per GPU; reference arm (N = 1): its per-view loop."""
    from luciddreamer_b200 import multiview as MV
    from luciddreamer_b200 import rasterizer as R
    W, H, P = 1920, 1080, 1_000_000
    scene = syn.make_scene(P, 1004)
    poses = syn.rotate360_poses(720)[::11][:n_views]          # every floor(720/64)-th frame
    cams = [syn.make_camera(W, H, c2w=m) for m in poses]
    out = {"P": P, "W": W, "H": H, "views": n_views, "views_per_rank": len(MV.shard_views(n_views, rank, world)), "reps": reps}

    def timed(fn):
        fn()                                                  # warm-up (learns the binning capacity)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return _maxrank(e0.elapsed_time(e1), world) / reps

    if impl == "reference":
        ref = RefCuda(scene, cams[0], dev, D)
        ref.set_cameras(cams)
        ref.order = np.arange(4096) % len(cams)

        def loop():
            ref.k = 0
            for _ in range(n_views):
                ref.forward()
        ms = timed(loop)
        out.update(ms_per_batch=ms, value=n_views * W * H / ms / 1e3, unit="Mrays/s", what="reference per-view loop")
        return out
    params = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    bg = torch.zeros(3, device=dev)
    sl = [R.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev), c.projmatrix.to(dev), D,
                                          c.campos.to(dev), False, False) for c in cams]
    for ns in (1, 2):
        ms = timed(lambda: MV.render_views_batched(params, sl, rank, world, n_streams=ns))
        out[f"in_flight_{ns}"] = {"ms_per_batch": ms, "value": n_views * W * H / ms / 1e3, "unit": "Mrays/s"}
    best = min((1, 2), key=lambda ns: out[f"in_flight_{ns}"]["ms_per_batch"])
    out.update(ms_per_batch=out[f"in_flight_{best}"]["ms_per_batch"], value=out[f"in_flight_{best}"]["value"], unit="Mrays/s",
               views_in_flight=best, scaling="strong")
    return out


def config5_leg(dev, rank, world, D=3, steps=5, warmup=2):
    """BASELINE config 5 at its stated size: 2 M Gaussians, 8 views of the llff path at 1920x1080, ONE shared-model
    optimisation step = forward+backward of every view (8 / N per rank) + the sum of the per-Gaussian gradients over
    all views of all ranks (472 MB bucket).  nccl: dense buckets + one all-reduce; fused: the final backward kernel of every
    view adds its visible rows straight into every rank's symmetric bucket (multimem.red / peer stores)."""
    from luciddreamer_b200 import multiview as MV
    from luciddreamer_b200 import rasterizer as R
    W, H, P, NV = 1920, 1080, 2_000_000, 8
    scene = syn.make_scene(P, 1005)
    cams = [syn.make_camera(W, H, c2w=m) for m in syn.llff_poses(400)[::50][:NV]]    # every 50th frame
    params = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    bg = torch.zeros(3, device=dev)
    sl = [R.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev), c.projmatrix.to(dev), D,
                                          c.campos.to(dev), False, False) for c in cams]
    mine = MV.shard_views(NV, rank, world)
    cots = [syn.make_cotangent(H, W, 1005 + v).to(dev) if v in mine else None for v in range(NV)]
    M = params["shs"].shape[1]
    out = {"P": P, "W": W, "H": H, "views": NV, "views_per_rank": len(mine), "bucket_bytes": P * (3 + 3 * M + 1 + 3 + 4) * 4}

    def timed(step):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return _maxrank(e0.elapsed_time(e1), world) / steps

    dense = MV.GradBucket(P, M, dev)

    def step_nccl():
        MV.multi_view_step(params, sl, cots, dense, rank, world)
        MV.allreduce_bucket(dense)
    ms = timed(step_nccl)
    out["nccl_allreduce" if world > 1 else "single_gpu"] = {"ms_per_step": ms, "value": NV * W * H / ms / 1e3, "unit": "Mrays/s"}
    out.update(ms_per_step=ms, value=NV * W * H / ms / 1e3, unit="Mrays/s", scaling="strong")
    if world > 1:
        try:
            symm = MV.SymmGradBucket(P, M, dev)

            def step_fused():
                symm.begin_step()
                MV.multi_view_step(params, sl, cots, symm, rank, world)
                symm.end_step()
            ms_f = timed(step_fused)
            torch.cuda.synchronize()
            err = float(((symm.flat - dense.flat).norm() / dense.flat.norm().clamp_min(1e-30)).item())
            out["fused_peer_reduce"] = {"ms_per_step": ms_f, "value": NV * W * H / ms_f / 1e3, "unit": "Mrays/s",
                                        "multicast": bool(symm.peers["mc"]), "rel_err_vs_nccl": err}
            if ms_f < ms:
                out.update(ms_per_step=ms_f, value=NV * W * H / ms_f / 1e3)
        except Exception as ex:
            out["fused_peer_reduce"] = {"unavailable": str(ex)[:200]}
    return out


# ----------------------------------------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="shell", choices=["shell", "stress", "frustum"])
    ap.add_argument("--random-camera", type=int, default=0, help="N jittered training views, one drawn per step")
    ap.add_argument("--no-next-rows", action="store_true")
    ap.add_argument("--P", type=int, default=0); ap.add_argument("--W", type=int, default=0); ap.add_argument("--H", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shared-model", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="e2e through the eager loop only (no CUDA-graph step)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 4 / config 5 legs")
    ap.add_argument("--shuffled", action="store_true", help="frustum scene: Gaussians in random memory order (default: "
                    "LucidDreamer's raster order)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    os.environ.setdefault("OMP_PROC_BIND", "close")      # cpu_baseline: pinned OpenMP threads (read when libgomp starts)
    os.environ.setdefault("OMP_PLACES", "cores")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0                                   # rank 0 alone runs the reference arm
    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a GPU (no CPU fallback for the product path)"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    else:
        world = 1 if args.impl == "reference" else world

    scene, cam, cot_cpu, wl = make_workload(args, rank)
    W, H, P, D = wl["W"], wl["H"], wl["P"], wl["D"]
    rays = W * H
    cot = cot_cpu.to(dev)
    peak, peak_src = load_peaks()

    use_ref_cuda = False
    if args.impl == "reference":
        from oracle import ref_cuda
        use_ref_cuda = ref_cuda.available()
        if not use_ref_cuda:                       # reference CUDA not compiled here: time the oracle port instead
            cb = cpu_baseline(scene, cam, cot_cpu, D, budget_s=60.0, max_steps=max(1, min(args.steps, 5)))
            line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Mrays/s", "n_gpus": 0,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": rays / cb["value"] / 1e3,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic", "config": config_dict(wl, args, 1), "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        impl = RefCuda(scene, cam, dev, D)
    else:
        impl = Ours(scene, cam, dev, D)

    if wl["cams"]:
        impl.set_cameras(wl["cams"])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = time_steps(impl, cot, args.steps, args.warmup, world)
    clocks = sampler.stop() if rank == 0 else None
    st = impl.stats()
    value = world * rays * args.steps / ms / 1e3           # Mrays/s, whole job
    g = torch.Generator().manual_seed(4242)
    target_u8 = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    e2e_ms, h2d, d2h, loss = time_e2e(impl, cam, target_u8, args.steps, args.warmup, world, fused_loss=(args.impl == "ours"))
    e2e_eager = None
    if args.impl == "ours" and not args.no_graph:
        # headline e2e: the step as ONE CUDA graph through the public API (graphs.GraphedStep); the eager loop is kept
        # beside it (`e2e.eager`) -- it is host-bound on boxes with slow cores, the graph is not
        try:
            g_ms, h2d, d2h, g_loss = time_e2e(impl, cam, target_u8, args.steps, args.warmup, world, fused_loss=True, graphed=True)
            e2e_eager = {"value": world * rays * args.steps / e2e_ms / 1e3, "ms_per_step": e2e_ms / args.steps,
                         "loss": loss}
            assert abs(g_loss - loss) <= 1e-6 * max(1.0, abs(loss)), (g_loss, loss)
            e2e_ms, loss = g_ms, g_loss
        except Exception as ex:
            e2e_eager = {"graph_error": str(ex)[:300]}
    e2e_val = world * rays * args.steps / e2e_ms / 1e3

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(wl, args, world),
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "loss": loss,
                    "mode": ("cuda_graph" if (e2e_eager and "value" in e2e_eager) else "eager"), "eager": e2e_eager,
                    "protocol": "per step: H2D of the step's camera (35 floats) + uint8 target image [H,W,3] from pinned memory "
                                "(copy stream, double buffered; cuda_graph mode: both in ONE staged copy), forward, on-device L1 "
                                "loss + gradient, backward, D2H scalar loss"},
            "clocks": clocks, "scene_stats": {"P_vis": st["P_vis"], "pairs": st["pairs"]}}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["reference_kind"] = "reference CUDA extension core (oracle/_ref) on the same GPU"
        line["gpu_launches"] = 0
    else:
        line["gpu_launches"] = impl.kernels_per_step * args.steps
        # ---- roofline of the dominant kernel (CUDA events around every launch, on the launching stream)
        kt = impl.profile(cot, n=5)
        N_, G_, M_ = W * H, ((W + 15) // 16) * ((H + 15) // 16), scene["shs"].shape[1]
        ab = alg_bytes(P, st["P_vis"], st["pairs"], N_, G_, M_)
        dom = max(kt, key=lambda k: kt[k])
        traffic, warp_inst = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import make_traffic
                if tj.get("sources_sha") != make_traffic.sources_sha():
                    raise ValueError("profiles/traffic.json was captured from other kernel sources")   # -> traffic: null
                traffic = tj.get(args.scene, {}).get(dom)
                warp_inst = tj.get(args.scene + "_warp_inst", {}).get(dom)
            except Exception:
                traffic, warp_inst = None, None
        ach = ab[dom] / (kt[dom] * 1e-3) / 1e9 if kt[dom] > 0 else 0.0
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes": ab[dom], "kernel_ms": kt[dom]}
        if warp_inst and kt[dom] > 0 and clocks and clocks.get("sm_mhz"):
            # the dominant kernel is FP32-issue bound, not HBM bound: its own ceiling is the warp
            # instruction issue rate = SMs x 4 schedulers x SM clock; instructions per launch are the ncu count of this
            # workload (profiles/traffic.json), the duration is measured live
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6
            ach_issue = warp_inst / (kt[dom] * 1e-3)
            line["roofline"]["issue"] = {"warp_instructions_per_launch": warp_inst, "achieved_ginst_s": ach_issue / 1e9,
                                         "peak_ginst_s": peak_issue / 1e9, "frac": ach_issue / peak_issue,
                                         "note": "supplementary: issue-slot ceiling of the FP32-bound blend kernel"}
        total_alg = sum(ab.values())
        line["roofline_step"] = {"algorithmic_bytes": total_alg, "achieved": total_alg / (ms / args.steps * 1e-3) / 1e9,
                                 "unit": "GB/s", "frac": total_alg / (ms / args.steps * 1e-3) / 1e9 / peak}
        line["kernels_ms"] = kt
        line["kernels_frac_of_hbm_peak"] = {k: (ab[k] / (kt[k] * 1e-3) / 1e9 / peak if kt[k] > 0 else None) for k in kt}
        if world > 1 and not args.no_shared_model:
            line["shared_model_step"] = shared_model_leg(scene, cam, cot, dev, D, args.steps, args.warmup, world)

    if not args.no_configs and args.scene == "shell" and not args.P:
        # BASELINE configs 4 and 5 at their stated sizes (every rank takes part; strong scaling over the view list)
        cfgs = {}
        del impl
        torch.cuda.empty_cache()
        for name, fn in (("config4_rotate360_64views_fwd", lambda: config4_leg(dev, rank, world, impl=args.impl)),
                         ("config5_llff_8views_shared_model", lambda: config5_leg(dev, rank, world))):
            if args.impl == "reference" and name.startswith("config5"):
                continue                                  # the reference has no multi-view / multi-GPU step
            try:
                cfgs[name] = fn()
            except Exception as ex:
                cfgs[name] = {"error": str(ex)[:300]}
            torch.cuda.empty_cache()
        line["baseline_configs"] = cfgs
    if rank == 0 and world == 1 and args.impl == "ours" and not args.no_next_rows:
        nr = {}
        for name, fn in (("photometric_loss", lambda: loss_leg(H, W, dev)), ("optimizer_step", lambda: adam_leg(scene, dev)),
                         ("dist_cuda2", knn_leg), ("video_render", video_leg),
                         ("luciddreamer_shaped", lambda: shaped_legs(Ours, dev))):     # reference side: --impl reference
            try:
                nr[name] = fn()
            except Exception as ex:                      # a "next" row must never take the headline line down
                nr[name] = {"error": str(ex)[:200]}
        line["next_rows"] = nr
    if rank == 0 and world == 1 and args.impl == "reference" and use_ref_cuda and not args.no_next_rows:
        nr = {}
        for name, fn in (("dist_cuda2", lambda: knn_leg("reference")), ("video_render", lambda: video_leg("reference")),
                         ("luciddreamer_shaped", lambda: shaped_legs(RefCuda, dev))):
            try:
                nr[name] = fn()
            except Exception as ex:
                nr[name] = {"error": str(ex)[:200]}
        line["next_rows"] = nr
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(scene, cam, cot_cpu, D)
    if rank == 0:
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


def config_dict(wl, args, world):
    return {"workload": f"BASELINE config {CFG_ID}: {wl['P']} Gaussians, {wl['W']}x{wl['H']}, SH degree {wl['D']}, "
                        f"forward+backward, one view per GPU per step",
            "scene": (f"{args.scene} (SURVEY.md 8d shell, seed {1000 + CFG_ID}, scale x{wl['scale_mult']})" if args.scene != "frustum"
                      else f"frustum (LucidDreamer-shaped, synthetic.make_frustum_scene, seed {1000 + CFG_ID}, "
                           f"{'random' if args.shuffled else 'raster'} memory order)"),
            "random_cameras": len(wl["cams"]) if wl.get("cams") else 0,
            "views_per_step": world, "parallelism": f"view-parallel x{world}",
            "l2_policy": "inputs larger than L2 (236 MB of Gaussian parameters + 25 MB cotangent per step)"}


if __name__ == "__main__":
    sys.exit(main())
