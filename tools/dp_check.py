"""Multi-GPU check of the fused gradient -> peer-reduce path (run under torchrun on >= 2 GPUs):
SymmGradBucket (peer stores / multicast) must equal the NCCL all-reduce of per-rank dense buckets."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from luciddreamer_b200 import synthetic as syn, multiview as MV
from luciddreamer_b200.rasterizer import GaussianRasterizationSettings
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
P, W, H, D = int(os.environ.get("DP_P", 200000)), int(os.environ.get("DP_W", 640)), int(os.environ.get("DP_H", 360)), 3
sc = {k: v.to(dev) for k, v in syn.make_scene(P, 7, scale_mult=2.0 if W < 1000 else 1.0).items()}
cam = syn.make_camera(W, H, c2w=syn.rotate360_poses(16)[rank * 2 % 16])
cot = syn.make_cotangent(H, W, 7).to(dev)
rs = GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
                                   cam.projmatrix.to(dev), D, cam.campos.to(dev), False, False)
ref = MV.GradBucket(P, 16, dev)
MV.view_step(sc, rs, cot, bucket=ref)
MV.allreduce_bucket(ref)
torch.cuda.synchronize()
for mc, ds in ((False, True), (True, True), (True, False)):
    b = MV.SymmGradBucket(P, 16, dev, use_multicast=mc, device_sync=ds)
    if mc and not b.peers["mc"]:
        if rank == 0: print("multicast not supported here; skipped")
        continue
    for it in range(5):        # both buffers of the double-buffered bucket get used, cleared and reused
        b.begin_step(); MV.view_step(sc, rs, cot, bucket=b); b.end_step()
        torch.cuda.synchronize()
        e_it = ((b.flat - ref.flat).norm() / ref.flat.norm()).item()
        assert e_it < 1e-5, (it, e_it)
    torch.cuda.synchronize()
    err = (b.flat - ref.flat).norm() / ref.flat.norm()
    nz = (ref.flat != 0).sum().item()
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); e0.record()
    for it in range(10):
        b.begin_step(); MV.view_step(sc, rs, cot, bucket=b); b.end_step()
    e1.record(); torch.cuda.synchronize()
    t_fused = e0.elapsed_time(e1) / 10
    dist.barrier(); e0.record()
    for it in range(10):
        MV.view_step(sc, rs, cot, bucket=ref); MV.allreduce_bucket(ref)
    e1.record(); torch.cuda.synchronize()
    t_nccl = e0.elapsed_time(e1) / 10
    if rank == 0:                                 # per-kernel times of the fused step (CUDA events around every launch)
        import ctypes as C
        from luciddreamer_b200 import _native as N, rasterizer as R
        L = N.lib(); ctx = R._ctx(lr); nk = L.gs_profile_num_kernels()
        L.gs_profile_enable(ctx, 1)
    acc_ms = None
    for it in range(4):
        b.begin_step(); MV.view_step(sc, rs, cot, bucket=b); b.end_step()
        if rank == 0:
            buf = (C.c_float * nk)(); N.check(L.gs_profile_read(ctx, buf))
            acc_ms = [max(x, 0.0) for x in buf] if acc_ms is None else [a + max(x, 0.0) for a, x in zip(acc_ms, buf)]
    if rank == 0:
        L.gs_profile_enable(ctx, 0)
        print("   kernels (us):", {L.gs_profile_kernel_name(i).decode(): round(acc_ms[i] / 4 * 1000, 1) for i in range(nk) if acc_ms[i] > 0})
    print(f"rank {rank}: multicast={bool(b.peers['mc'])} device_sync={ds} rel err vs NCCL all-reduce {err.item():.2e} (nonzeros {nz}); "
          f"step fused {t_fused:.3f} ms vs dense all-reduce {t_nccl:.3f} ms")
dist.destroy_process_group()
