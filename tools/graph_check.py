"""GPU check of the sync-free forward and of the CUDA-graph step (luciddreamer_b200.graphs.GraphedStep):
results identical to the default path, overflow detected, and the host cost per step of each variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from luciddreamer_b200 import synthetic as syn, rasterizer as R, losses
from luciddreamer_b200.graphs import GraphedStep

dev = torch.device("cuda:0")
P, W, H, D = int(os.environ.get("GC_P", 1_000_000)), 1920, 1080, 3
sc = {k: v.to(dev) for k, v in syn.make_scene(P, 1003).items()}
cam = syn.make_camera(W, H)
d_cam = torch.cat([cam.viewmatrix.reshape(-1), cam.projmatrix.reshape(-1), cam.campos.reshape(-1)]).to(dev)
bg = torch.zeros(3, device=dev)
rs = R.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, d_cam[0:16].view(4, 4), d_cam[16:32].view(4, 4), D,
                                     d_cam[32:35], False, False)
rast = R.GaussianRasterizer(rs)
leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
tgt = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev)
out = {}


def step():
    out.clear()                                   # no reference to the previous step's autograd graph
    for t in leaves.values():
        t.grad = None
    m2.grad = None
    color, radii, depth = rast(leaves["means3D"], m2, leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                               rotations=leaves["rotations"])
    loss, cot = losses.l1_loss_with_grad(color, tgt)
    torch.autograd.backward(color, grad_tensors=cot)
    out["color"], out["loss"] = color, loss
    return color, loss


def grads():
    return [leaves[k].grad.clone() for k in leaves] + [m2.grad.clone()]


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th = time.perf_counter()
    for _ in range(n):
        fn()
    host = (time.perf_counter() - th) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3, host


step(); torch.cuda.synchronize()
ref_color, ref_grads = out["color"].detach().clone(), grads()   # detach: a clone with a grad_fn would keep the graph alive
pairs = R._last_pairs[0]
print("pairs", pairs)
print("eager sync   : gpu %.3f ms/step, wall %.3f, host-enqueue %.3f" % timeit(step))

with R.static_capacity(int(pairs * 1.5)):
    step(); torch.cuda.synchronize()
    assert torch.equal(out["color"], ref_color), "static-capacity forward differs"
    for a, b in zip(grads(), ref_grads):
        assert (a - b).norm() <= 1e-5 * b.norm().clamp_min(1e-20)
    print("eager static : gpu %.3f ms/step, wall %.3f, host-enqueue %.3f" % timeit(step))
print("check_static:", R.check_static())

# overflow is detected, loudly
try:
    with R.static_capacity(max(64, pairs // 2)):
        step()
    R.check_static()
    print("ERROR: overflow not detected"); sys.exit(1)
except RuntimeError as ex:
    print("overflow detected:", str(ex)[:90])
pass
step(); torch.cuda.synchronize()                  # the device survived the overflow (guards in every kernel)
assert torch.equal(out["color"], ref_color)

g = GraphedStep(step, warmup=3)
g.replay(); c = g.check()
print("graph counts", c, "capacity", g.pair_capacity)
assert torch.equal(g.outputs[0], ref_color), "graphed forward differs"
for a, b in zip([leaves[k].grad for k in leaves] + [m2.grad], ref_grads):
    assert (a - b).norm() <= 1e-5 * b.norm().clamp_min(1e-20)
print("graph replay : gpu %.3f ms/step, wall %.3f, host-enqueue %.3f" % timeit(g.replay))
g.check()
# a different camera through the same graph (the captured kernels read the camera from d_cam)
cam2 = syn.make_camera(W, H, c2w=syn.rotate360_poses(64)[5])
d_cam.copy_(torch.cat([cam2.viewmatrix.reshape(-1), cam2.projmatrix.reshape(-1), cam2.campos.reshape(-1)]).to(dev))
g.replay(); c2 = g.check()
gc = g.outputs[0].detach().clone()
step(); torch.cuda.synchronize()
assert torch.equal(gc, out["color"]), "graph replay with a new camera differs from eager"
print("new camera through the graph ok:", c2)
print("GRAPH_CHECK_OK")
