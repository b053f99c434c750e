"""Golden vectors for simple_knn.distCUDA2, produced by the reference's OWN kernels (oracle/_ref/libref_simpleknn.so =
submodules/simple-knn/simple_knn.cu compiled unmodified) on a B200.  Run on the GPU box:
    python tests/golden/make_knn_golden.py gpurun_out/knn_golden.npz     (then copy into tests/golden/)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import ref_cuda

rng = np.random.default_rng(2024)
clouds = {
    "normal_3000": rng.normal(size=(3000, 3)).astype(np.float32),
    "clusters_dups_2500": np.concatenate([rng.normal(size=(1000, 3)), 1e-3 * rng.normal(size=(1000, 3)) + 3.0,
                                          rng.uniform(-5, 5, size=(480, 3)) * [1, 1, 0], np.zeros((20, 3))]).astype(np.float32),
    "ragged_1025": rng.uniform(-2, 2, size=(1025, 3)).astype(np.float32),
    "line_5": np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [7, 0, 0], [15, 0, 0]], np.float32),
}
out = {}
for k, pts in clouds.items():
    out[k + "/points"] = pts
    out[k + "/dist2"] = ref_cuda.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
np.savez_compressed(sys.argv[1], **out)
print("wrote", sys.argv[1])
