"""Ad-hoc GPU sanity run: product vs CPU oracle vs reference CUDA (oracle/_ref), plus rough timings.
Usage (on the GPU box): python tools/gpu_check.py [cfg ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from luciddreamer_b200 import synthetic as syn
from luciddreamer_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from oracle import oracle, ref_cuda


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run(P, W, H, D, seed, scale_mult=1.0, do_oracle=True, iters=0):
    dev = torch.device("cuda:0")
    sc = syn.make_scene(P, seed, scale_mult=scale_mult)
    cam = syn.make_camera(W, H)
    w = syn.make_cotangent(H, W, seed)
    d = {k: v.to(dev) for k, v in sc.items()}
    bg = torch.zeros(3, device=dev)
    rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix.to(dev),
                                       cam.projmatrix.to(dev), D, cam.campos.to(dev), False, False)
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
    rast = GaussianRasterizer(rs)
    color, radii, depth = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                               rotations=leaves["rotations"])
    torch.autograd.backward(color, grad_tensors=w.to(dev))
    torch.cuda.synchronize()
    mine = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy(), radii=radii.cpu().numpy(),
                m2=m2.grad.cpu().numpy(), op=leaves["opacities"].grad.cpu().numpy(), m3=leaves["means3D"].grad.cpu().numpy(),
                sh=leaves["shs"].grad.cpu().numpy(), sc=leaves["scales"].grad.cpu().numpy(), rot=leaves["rotations"].grad.cpu().numpy())
    print(f"== P={P} {W}x{H} D={D} scale_mult={scale_mult}: vis={(mine['radii']>0).sum()} color mean={mine['color'].mean():.5f}")
    names = ["m2", "col", "op", "m3", "cov", "sh", "sc", "rot"]
    if ref_cuda.available():
        rc = ref_cuda.RefContext()
        R, rcol, rdep, rrad = ref_cuda.rasterize_gaussians(rc, bg, d["means3D"], None, d["opacities"], d["scales"], d["rotations"], 1.0, None,
                                                           rs.viewmatrix, rs.projmatrix, cam.tanfovx, cam.tanfovy, H, W, d["shs"], D, rs.campos)
        rg = ref_cuda.rasterize_gaussians_backward(rc, rrad, w.to(dev))
        torch.cuda.synchronize()
        print(f"   [ref ] R={R} color maxabs={np.abs(mine['color']-rcol.cpu().numpy()).max():.3e} depth maxabs={np.abs(mine['depth']-rdep.cpu().numpy()).max():.3e} "
              f"radii mismatches={(mine['radii']!=rrad.cpu().numpy()).sum()}")
        for n, g in zip(names, rg):
            if n in mine:
                print(f"   [ref ] grad {n:4s} rel={rel(mine[n], g.cpu().numpy()):.3e}")
    if do_oracle:
        t0 = time.time()
        f = oracle.rasterize_gaussians(torch.zeros(3), sc["means3D"], None, sc["opacities"], sc["scales"], sc["rotations"], 1.0, None,
                                       cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, sc["shs"], D, cam.campos)
        og = oracle.rasterize_gaussians_backward(f, w.numpy())
        print(f"   [orac] R={f.num_rendered} ({time.time()-t0:.2f}s) color maxabs={np.abs(mine['color']-f.color).max():.3e} depth maxabs={np.abs(mine['depth']-f.depth).max():.3e} "
              f"radii mismatches={(mine['radii']!=f.radii).sum()}")
        for n, g in zip(names, og):
            if n in mine:
                print(f"   [orac] grad {n:4s} rel={rel(mine[n], g):.3e}")
    if iters:
        def step():
            for v in leaves.values(): v.grad = None
            m2.grad = None
            c, r, dd = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            torch.autograd.backward(c, grad_tensors=wd)
        wd = w.to(dev)
        for _ in range(5): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"   [time] mine fwd+bwd {ms:.3f} ms -> {W*H/ms/1e3:.1f} Mrays/s")
        if ref_cuda.available():
            def rstep():
                R, rcol, rdep, rrad = ref_cuda.rasterize_gaussians(rc, bg, d["means3D"], None, d["opacities"], d["scales"], d["rotations"], 1.0, None,
                                                                   rs.viewmatrix, rs.projmatrix, cam.tanfovx, cam.tanfovy, H, W, d["shs"], D, rs.campos)
                ref_cuda.rasterize_gaussians_backward(rc, rrad, wd)
            for _ in range(3): rstep()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters): rstep()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"   [time] ref  fwd+bwd {ms:.3f} ms -> {W*H/ms/1e3:.1f} Mrays/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "host cores", os.cpu_count())
    run(2000, 128, 96, 3, 1)
    run(10_000, 256, 256, 0, 1001)
    run(100_000, 512, 512, 3, 1002, iters=20)
    run(100_000, 512, 512, 3, 1002, scale_mult=4.0, iters=20)
    run(1_000_000, 1920, 1080, 3, 1003, do_oracle=False, iters=20)
    run(1_000_000, 1920, 1080, 3, 1003, scale_mult=4.0, do_oracle=False, iters=10)
