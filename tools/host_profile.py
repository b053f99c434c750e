"""cProfile of the host side of forward + fused L1 + backward through the operator API (GPU box).
Usage: python tools/host_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from luciddreamer_b200 import losses, synthetic as syn


class A:
    P = None; W = None; H = None; scene = "shell"


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
scene, cam, cot, meta = bench.make_workload(A, 0)
dev = torch.device("cuda:0")
impl = bench.Ours(scene, cam, dev, meta["D"])
tgt = (torch.rand(cam.image_height, cam.image_width, 3) * 255).to(torch.uint8).to(dev)


def step():
    color = impl.forward()
    loss, cotg = losses.l1_loss_with_grad(color, tgt)
    impl.backward(color, cotg)


for _ in range(20):
    step()
torch.cuda.synchronize()
# pure host time per step when the GPU is NOT the bottleneck is not observable directly (the counts wait blocks), so
# report wall per step and the profile of everything except the wait
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / steps * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
