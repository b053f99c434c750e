"""Host-side cost of one forward+backward through the autograd API: a tiny scene (GPU work ~0) timed by wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from luciddreamer_b200 import synthetic as syn
from luciddreamer_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
dev = torch.device("cuda:0")
P, W, H = 2000, 64, 64
sc = {k: v.to(dev) for k, v in syn.make_scene(P, 1).items()}
cam = syn.make_camera(W, H)
w = syn.make_cotangent(H, W, 1).to(dev)
rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
                                   cam.projmatrix.to(dev), 3, cam.campos.to(dev), False, False)
rast = GaussianRasterizer(rs)
L = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
def step():
    for t in L.values(): t.grad = None
    m2.grad = None
    c, r, d = rast(L["means3D"], m2, L["opacities"], shs=L["shs"], scales=L["scales"], rotations=L["rotations"])
    t1 = time.perf_counter()
    torch.autograd.backward(c, grad_tensors=w)
    return t1
for _ in range(20): step()
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter(); tf = 0.0
for _ in range(n):
    ta = time.perf_counter(); t1 = step(); tf += t1 - ta
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / n
print(f"host: {tot*1e6:.1f} us per fwd+bwd step (forward call {tf/n*1e6:.1f} us, backward call {(tot - tf/n)*1e6:.1f} us) at P={P}, {W}x{H}")

import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:4500])

# ---- the e2e loop of bench.py on the same tiny scene: host cost of the extra per-step operations
import bench as B
class _I:  # minimal impl wrapper like bench.Ours
    pass
cam_small = cam
impl = B.Ours({k: sc[k].cpu() for k in sc}, cam_small, dev, 3)
tgt = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8)
for _ in range(2):
    ms, h2d, d2h, loss = B.time_e2e(impl, cam_small, tgt, 200, 20, 1, fused_loss=True)
print(f"host: e2e loop {ms/200*1e3:.1f} us per step at P={P}, {W}x{H}")
pr = cProfile.Profile(); pr.enable()
B.time_e2e(impl, cam_small, tgt, 200, 5, 1, fused_loss=True)
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(16); print(st.getvalue()[:3800])
