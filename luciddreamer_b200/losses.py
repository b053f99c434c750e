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
    with R._guard(idx):
        N.check(N.lib().gs_l1_loss_backward(R._ctx(idx), color.data_ptr(), target_u8.data_ptr(), H, W, float(weight),
                                            grad.data_ptr(), loss.data_ptr(), R._raw_stream(idx)))
    return loss, grad


def photometric_loss_with_grad(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2, need_grad: bool = True):
    """LucidDreamer's training loss (luciddreamer.py:301-303) and its gradient, fused:
        loss = (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
    image, gt: [3,H,W] f32 CUDA.  Returns (loss3, grad): loss3 = [loss, l1, ssim] on the device (f32 [3]),
    grad = dloss/dimage ([3,H,W]) or None."""
    if not (image.is_cuda and gt.is_cuda):
        raise RuntimeError("luciddreamer_b200: tensors must live on a CUDA device (no CPU fallback)")
    if image.dim() != 3 or image.shape[0] != 3 or image.shape != gt.shape:
        raise RuntimeError("image and gt must both be [3, H, W]")
    image = image.detach().to(torch.float32).contiguous()
    gt = gt.detach().to(torch.float32).contiguous()
    _, H, W = image.shape
    dev = image.device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    L = N.lib()
    scratch = torch.empty(L.gs_photometric_scratch_bytes(H, W), dtype=torch.uint8, device=dev)
    grad = torch.empty_like(image) if need_grad else None
    loss3 = torch.empty(3, dtype=torch.float32, device=dev)
    with R._guard(idx):
        N.check(L.gs_photometric_loss_backward(R._ctx(idx), image.data_ptr(), gt.data_ptr(), H, W, float(lambda_dssim),
                                               scratch.data_ptr(), grad.data_ptr() if need_grad else None,
                                               loss3.data_ptr(), R._raw_stream(idx)))
    return loss3, grad


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        loss3, grad = photometric_loss_with_grad(image, gt, lambda_dssim, need_grad=image.requires_grad)
        ctx.save_for_backward(grad)
        return loss3[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """Differentiable drop-in for the loss line of the reference's training loop (luciddreamer.py:302):
    `loss = photometric_loss(image, gt_image, opt.lambda_dssim); loss.backward()`."""
    return _PhotometricLoss.apply(image, gt, float(lambda_dssim))
