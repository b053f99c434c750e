"""Fused photometric loss pieces (first slice): the L1 term of LucidDreamer's training loss
(luciddreamer.py:301-303, utils/loss.py:18) and its gradient in one CUDA pass over the image."""
from __future__ import annotations

import torch

from . import _native as N
from . import rasterizer as R


def l1_loss_with_grad(color: torch.Tensor, target_u8: torch.Tensor, weight: float = 1.0):
    """color [3,H,W] f32 (CUDA), target_u8 [H,W,3] uint8 (CUDA).  Returns (loss [1] f32 on device,
    dL_dcolor [3,H,W] f32) with loss = weight * mean|color - target/255| -- `l1_loss` of utils/loss.py:18."""
    if not (color.is_cuda and target_u8.is_cuda):
        raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
    _, H, W = color.shape
    if target_u8.dtype != torch.uint8 or tuple(target_u8.shape) != (H, W, 3):
        raise RuntimeError("target_u8 must be uint8 [H, W, 3]")
    color = color.detach().contiguous()
    target_u8 = target_u8.contiguous()
    dev = color.device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    grad = torch.empty_like(color)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(idx):
        N.check(N.lib().gs_l1_loss_backward(R._ctx(idx), color.data_ptr(), target_u8.data_ptr(), H, W, float(weight),
                                            grad.data_ptr(), loss.data_ptr(), torch.cuda.current_stream(idx).cuda_stream))
    return loss, grad
