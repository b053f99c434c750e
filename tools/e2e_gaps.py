"""Where does the e2e step lose time against the device-only step?  Same loop as bench.time_e2e, with CUDA events
around forward..backward of every step: busy = event(before forward) -> event(after backward); gap = the rest
(camera copy, upload wait, loss read-back, and any time the GPU waits for the host).  Usage: python tools/e2e_gaps.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from luciddreamer_b200 import losses


class A:
    P = None; W = None; H = None; scene = "shell"


steps, warm = 60, 10
scene, cam, cot, meta = bench.make_workload(A, 0)
dev = torch.device("cuda:0")
impl = bench.Ours(scene, cam, dev, meta["D"])
H, W = cam.image_height, cam.image_width
tgt = (torch.rand(H, W, 3) * 255).to(torch.uint8)
h_tgt = tgt.contiguous().pin_memory()
h_cam = torch.cat([cam.viewmatrix.reshape(-1), cam.projmatrix.reshape(-1), cam.campos.reshape(-1)]).contiguous().pin_memory()
d_tgt = [torch.empty_like(tgt, device=dev) for _ in range(2)]
d_cam = torch.empty(35, device=dev)
impl.bind_camera(d_cam)
N = steps + warm
h_loss = torch.zeros(N, dtype=torch.float32).pin_memory()
copy_s = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream(dev)
ev_up = [torch.cuda.Event() for _ in range(2)]
ev_free = [torch.cuda.Event() for _ in range(2)]
ea = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
ef = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
el = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
eb = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
host = {"fwd": 0.0, "loss": 0.0, "bwd": 0.0, "rest": 0.0}


def upload(k):
    b = k & 1
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(ev_free[b])
        d_tgt[b].copy_(h_tgt, non_blocking=True)
        ev_up[b].record(copy_s)


for b in range(2):
    ev_free[b].record(main)
upload(0)
t_prev = time.perf_counter()
for k in range(N):
    b = k & 1
    d_cam.copy_(h_cam, non_blocking=True)
    main.wait_event(ev_up[b])
    ea[k].record(main)
    t0 = time.perf_counter()
    color = impl.forward()
    ef[k].record(main)
    t1 = time.perf_counter()
    loss, cotg = losses.l1_loss_with_grad(color, d_tgt[b])
    el[k].record(main)
    t2 = time.perf_counter()
    impl.backward(color, cotg)
    eb[k].record(main)
    t3 = time.perf_counter()
    h_loss[k].copy_(loss.reshape(()), non_blocking=True)
    ev_free[b].record(main)
    if k + 1 < N:
        upload(k + 1)
    t4 = time.perf_counter()
    if k >= warm:
        host["fwd"] += t1 - t0; host["loss"] += t2 - t1; host["bwd"] += t3 - t2; host["rest"] += (t4 - t3) + (t0 - t_prev)
    t_prev = t4
torch.cuda.synchronize()
r = range(warm, N - 1)
fwd = sum(ea[k].elapsed_time(ef[k]) for k in r) / len(r)
los = sum(ef[k].elapsed_time(el[k]) for k in r) / len(r)
bwd = sum(el[k].elapsed_time(eb[k]) for k in r) / len(r)
gap = sum(eb[k].elapsed_time(ea[k + 1]) for k in r) / len(r)
print(f"GPU timeline per step (us): forward {fwd*1e3:.1f} | loss {los*1e3:.1f} | backward {bwd*1e3:.1f} | "
      f"between steps {gap*1e3:.1f} | total {(fwd+los+bwd+gap)*1e3:.1f}")
print("host per step (us, incl. the counts wait inside forward):", {k: round(v / steps * 1e6, 1) for k, v in host.items()})
