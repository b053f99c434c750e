"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA rasterizer (oracle/_ref, built by
oracle/Makefile from /root/reference) on a B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  Each fixture stores the reference's outputs (colour, depth,
radii, num_rendered, the 8 gradient tensors row-sparse) for the seeded inputs of tests/cases.py, plus the
reference's own run-to-run gradient jitter (float atomics, 5 repeats)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import cases
from oracle import ref_cuda


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    for case in cases.CASES:
        if not case.golden:
            continue
        inp = cases.build_inputs(case)
        args = cases.binding_args(inp, dev)
        ctx = ref_cuda.RefContext()
        R, color, depth, radii = ref_cuda.rasterize_gaussians(ctx, *args)
        cot = inp["cot"].to(dev)
        runs = []
        for _ in range(5):
            g = ref_cuda.rasterize_gaussians_backward(ctx, radii, cot)
            torch.cuda.synchronize()
            runs.append([t.cpu().numpy() for t in g])
        store = dict(num_rendered=np.int64(R), color=color.cpu().numpy(), depth=depth.cpu().numpy(),
                     radii=radii.cpu().numpy().astype(np.int32), device=torch.cuda.get_device_name(0))
        jit = {}
        for k, name in enumerate(cases.GRAD_NAMES):
            a = runs[0][k]
            idx, rows = cases.sparse_rows(a)
            store[name + "_idx"] = idx
            store[name + "_rows"] = rows
            store[name + "_shape"] = np.array(a.shape, np.int64)
            nrm = max(float(np.linalg.norm(a.astype(np.float64))), 1e-30)
            jit[name] = max(float(np.linalg.norm((r[k] - a).astype(np.float64))) / nrm for r in runs[1:])
            store[name + "_jitter"] = np.float64(jit[name])
        path = os.path.join(outdir, case.name + ".npz")
        np.savez_compressed(path, **store)
        print(f"{case.name}: R={R} vis={(store['radii'] > 0).sum()} jitter max={max(jit.values()):.2e} "
              f"-> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
