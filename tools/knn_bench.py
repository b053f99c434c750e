"""Times simple_knn.distCUDA2: ours vs the reference's own kernel (oracle/_ref/libref_simpleknn.so) on the same points.
Usage (GPU box): python tools/knn_bench.py [P]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from luciddreamer_b200.simple_knn import distCUDA2



def run(P=1_000_000, which=("depthmap", "uniform"), impls=("ours", "reference")):
    ref_cuda = None
    if "reference" in impls:
        from oracle import ref_cuda          # the reference's own kernels (oracle/_ref), only when asked for
    side = int(round(P ** 0.5))
    rng = np.random.default_rng(11)
    u, v = np.meshgrid(np.linspace(-1, 1, side), np.linspace(-0.6, 0.6, side))
    z = 2.0 + 0.5 * np.sin(3 * u) * np.cos(2 * v) + 0.002 * rng.normal(size=u.shape)
    clouds = {"depthmap": np.stack([u * z, v * z, z], -1).reshape(-1, 3).astype(np.float32),
              "uniform": rng.uniform(-1, 1, size=(side * side, 3)).astype(np.float32)}
    out = {}
    for name in which:
        pts = clouds[name]
        t = torch.from_numpy(pts).cuda()
        res = {}
        for impl, fn in (("ours", distCUDA2 if "ours" in impls else None),
                         ("reference", ref_cuda.distCUDA2 if ref_cuda is not None and ref_cuda.knn_available() else None)):
            if fn is None:
                continue
            fn(t); torch.cuda.synchronize()
            n = 5 if impl == "ours" else 2
            t0 = time.perf_counter()
            for _ in range(n):
                r = fn(t)
            torch.cuda.synchronize()
            res[impl + "_ms"] = (time.perf_counter() - t0) / n * 1e3
            res[impl] = r
        if "reference" in res and "ours" in res:
            res["bit_identical"] = bool(torch.equal(res["ours"], res["reference"]))
        res.pop("ours", None); res.pop("reference", None)
        out[name] = dict(P=len(pts), **res)
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)))
