"""CPU oracle of the photometric training loss (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Restates, in plain torch on the CPU (float32 or float64), the reference's
    loss = (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))      luciddreamer.py:301-303
with l1_loss = |x - y|.mean() (utils/loss.py:18) and ssim = the windowed statistic of utils/loss.py:26-69 (11-tap
Gaussian, sigma 1.5, outer-product window, per-channel conv2d with zero padding 5, C1 = 0.01^2, C2 = 0.03^2, mean).
Pinned against outputs of the reference's own functions: tests/golden/loss_golden.npz (make_loss_golden.py)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _window(channel: int, dtype) -> torch.Tensor:
    g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    w2 = (g @ g.t()).to(dtype)
    return w2.expand(channel, 1, 11, 11).contiguous()


def l1_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return (x - y).abs().mean()


def ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    ch = x.shape[-3]
    w = _window(ch, x.dtype)
    x4, y4 = x.reshape(1, ch, *x.shape[-2:]), y.reshape(1, ch, *y.shape[-2:])
    conv = lambda t: F.conv2d(t, w, padding=5, groups=ch)
    mu1, mu2 = conv(x4), conv(y4)
    s1 = conv(x4 * x4) - mu1 * mu1
    s2 = conv(y4 * y4) - mu2 * mu2
    s12 = conv(x4 * y4) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return m.mean()


def photometric_loss_with_grad(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2, dtype=torch.float32):
    """Returns (loss, l1, ssim, dL/dimage) as CPU tensors of `dtype`."""
    x = image.detach().cpu().to(dtype).clone().requires_grad_(True)
    y = gt.detach().cpu().to(dtype)
    l1, s = l1_loss(x, y), ssim(x, y)
    loss = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - s)
    loss.backward()
    return loss.detach(), l1.detach(), s.detach(), x.grad.detach()
