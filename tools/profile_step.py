"""One workload, a few fwd+bwd steps of the product (or the reference CUDA with --ref) -- the target for
`ncu --metrics gpu__time_duration.sum` launch lists and `ncu --set full` captures."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from luciddreamer_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1_000_000); ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--H", type=int, default=1080); ap.add_argument("--D", type=int, default=3)
ap.add_argument("--seed", type=int, default=1003); ap.add_argument("--scale-mult", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--ref", action="store_true")
ap.add_argument("--scene", default="shell", choices=["shell", "frustum"])
ap.add_argument("--shuffled", action="store_true", help="frustum scene in random memory order (default: raster order)")
a = ap.parse_args()
dev = torch.device("cuda:0")
_scene = (syn.make_frustum_scene(a.P, a.seed, a.W, a.H, raster=not a.shuffled) if a.scene == "frustum"
          else syn.make_scene(a.P, a.seed, scale_mult=a.scale_mult))
sc = {k: v.to(dev) for k, v in _scene.items()}
cam = syn.make_camera(a.W, a.H)
w = syn.make_cotangent(a.H, a.W, a.seed).to(dev)
bg = torch.zeros(3, device=dev)
vm, pm, cp = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
if a.ref:
    from oracle import ref_cuda
    rc = ref_cuda.RefContext()
    for _ in range(a.steps):
        R, c, d, r = ref_cuda.rasterize_gaussians(rc, bg, sc["means3D"], None, sc["opacities"], sc["scales"], sc["rotations"], 1.0, None,
                                                  vm, pm, cam.tanfovx, cam.tanfovy, a.H, a.W, sc["shs"], a.D, cp)
        ref_cuda.rasterize_gaussians_backward(rc, r, w)
else:
    from luciddreamer_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    rs = GaussianRasterizationSettings(a.H, a.W, cam.tanfovx, cam.tanfovy, bg, 1.0, vm, pm, a.D, cp, False, False)
    rast = GaussianRasterizer(rs)
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros(a.P, 3, device=dev, requires_grad=True)
    for _ in range(a.steps):
        c, r, d = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        torch.autograd.backward(c, grad_tensors=w)
torch.cuda.synchronize()
print("done")
