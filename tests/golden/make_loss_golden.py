"""Golden vectors of the photometric loss from the reference's OWN functions (utils/loss.py), generated in the build
container where /root/reference is mounted:  python tests/golden/make_loss_golden.py  -> tests/golden/loss_golden.npz"""
import os
import sys

import numpy as np
import torch

# utils/loss.py calls .cuda() at import time (line 88, Canny filter), so the module cannot be imported without a GPU;
# execute only its loss-function part (everything before the Canny section) -- the reference's own code, run in
# place, not copied.
_src = open("/root/reference/utils/loss.py").read().split("import numpy as np")[0]
_ns = {}
exec(compile(_src, "/root/reference/utils/loss.py", "exec"), _ns)
l1_loss, ssim = _ns["l1_loss"], _ns["ssim"]

out = {}
for name, (H, W, seed) in {"a_37x53": (37, 53, 1), "b_16x16": (16, 16, 2), "c_70x100": (70, 100, 3), "d_9x40": (9, 40, 4)}.items():
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    loss = (1.0 - 0.2) * Ll1 + 0.2 * (1.0 - s)     # luciddreamer.py:302, lambda_dssim = 0.2 (arguments.py:29)
    loss.backward()
    out[name + "_img"] = img.detach().numpy(); out[name + "_gt"] = gt.numpy()
    out[name + "_loss"] = np.float64(loss.item()); out[name + "_l1"] = np.float64(Ll1.item()); out[name + "_ssim"] = np.float64(s.item())
    out[name + "_grad"] = img.grad.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_golden.npz"), **out)
print({k: float(v) for k, v in out.items() if k.endswith("_loss")})
