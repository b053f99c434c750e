"""Fused optimiser step for the Gaussian parameters.

`FusedGaussianAdam` holds the six raw parameter tensors of scene/gaussian_model.py (xyz, features_dc, features_rest,
opacity logit, log scaling, raw rotation) like `GaussianModel.training_setup` registers them with torch.optim.Adam
(gaussian_model.py:151-169), and consumes the rasterizer's gradients w.r.t. the ACTIVATED values -- a
`multiview.GradBucket` (or any object with .means3D/.shs/.opacities/.scales/.rotations) -- in ONE kernel launch:
activation chain rule (sigmoid, exp, F.normalize, cat) + Adam, in place.  No autograd graph, no per-group Python loop."""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import _native as N
from . import rasterizer as R

# arguments.py:19-27 (GSParams) defaults; xyz is scheduled by the caller via set_lr (gaussian_model.py:171-177)
DEFAULT_LRS = {"xyz": 0.00016, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0, "opacity": 0.05, "scaling": 0.005,
               "rotation": 0.001}


def expon_lr(step: int, lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
             max_steps: int = 1000000) -> float:
    """The xyz learning-rate schedule (utils/general.py:31-64): log-linear interpolation lr_init -> lr_final over
    max_steps, optionally eased in by a sine ramp from lr_delay_mult over lr_delay_steps; 0 disables the group."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    delay = 1.0
    if lr_delay_steps > 0:
        delay = lr_delay_mult + (1.0 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay * math.exp(math.log(lr_init) * (1.0 - t) + math.log(lr_final) * t)


class FusedGaussianAdam:
    def __init__(self, xyz, features_dc, features_rest, opacity, scaling, rotation, lrs: Dict[str, float] = None,
                 betas=(0.9, 0.999), eps: float = 1e-15):
        self.params = {"xyz": xyz, "f_dc": features_dc, "f_rest": features_rest, "opacity": opacity,
                       "scaling": scaling, "rotation": rotation}
        for k, t in self.params.items():
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f"{k}: parameters must be contiguous float32 CUDA tensors (no CPU fallback)")
        self.lrs = dict(DEFAULT_LRS)
        if lrs:
            self.lrs.update(lrs)
        self.betas, self.eps, self.step_count = betas, eps, 0
        self._sched = (DEFAULT_LRS["xyz"], 0.0000016, 0, 0.01, 2990)
        self.exp_avg = {k: torch.zeros_like(t) for k, t in self.params.items()}
        self.exp_avg_sq = {k: torch.zeros_like(t) for k, t in self.params.items()}
        self.P = xyz.shape[0]
        self.M = 1 + features_rest.shape[1]

    def set_lr(self, name: str, lr: float) -> None:
        self.lrs[name] = float(lr)

    def set_xyz_schedule(self, lr_init: float, lr_final: float, lr_delay_mult: float = 0.01, max_steps: int = 2990,
                         spatial_lr_scale: float = 1.0) -> None:
        """GaussianModel.training_setup's xyz scheduler (gaussian_model.py:166-169; defaults arguments.py:20-23)."""
        self._sched = (lr_init * spatial_lr_scale, lr_final * spatial_lr_scale, 0, lr_delay_mult, max_steps)

    def update_learning_rate(self, iteration: int) -> float:
        """GaussianModel.update_learning_rate (gaussian_model.py:171-177): host arithmetic, no launch."""
        lr = expon_lr(iteration, *self._sched)
        self.lrs["xyz"] = lr
        return lr

    @torch.no_grad()
    def step(self, grads) -> None:
        """grads: gradients w.r.t. the activated values (means3D [P,3], shs [P,M,3], opacities [P,1], scales [P,3],
        rotations [P,4]) -- e.g. the GradBucket the backward wrote."""
        P, M = self.P, self.M
        groups = (N.GsAdamGroup * 6)()
        spec = [  # name, grad tensor, rows, row_width, grad_row_width, grad_offset, activation
            ("xyz", grads.means3D, P, 3, 3, 0, 0),
            ("f_dc", grads.shs, P, 3, 3 * M, 0, 0),
            ("f_rest", grads.shs, P, 3 * (M - 1), 3 * M, 3, 0),
            ("opacity", grads.opacities, P, 1, 1, 0, 1),
            ("scaling", grads.scales, P, 3, 3, 0, 2),
            ("rotation", grads.rotations, P, 4, 4, 0, 3),
        ]
        dev = self.params["xyz"].device
        spec = [s for s in spec if s[2] * s[3] > 0]        # an SH-degree-0 model (M = 1) has an empty f_rest group
        for k, (name, gt, rows, rw, grw, goff, act) in enumerate(spec):
            if not (gt.is_cuda and gt.dtype == torch.float32 and gt.is_contiguous() and gt.device == dev):
                raise RuntimeError(f"gradient for {name} must be a contiguous float32 tensor on {dev}")
            gr = groups[k]
            gr.param, gr.grad = self.params[name].data_ptr(), gt.data_ptr()
            gr.exp_avg, gr.exp_avg_sq = self.exp_avg[name].data_ptr(), self.exp_avg_sq[name].data_ptr()
            gr.rows, gr.row_width, gr.grad_row_width, gr.grad_offset, gr.activation = rows, rw, grw, goff, act
            gr.lr = float(self.lrs[name])
        self.step_count += 1
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        with R._guard(idx):
            N.check(N.lib().gs_gaussian_adam_step(R._ctx(idx), groups, len(spec), float(self.betas[0]), float(self.betas[1]),
                                                  float(self.eps), self.step_count,
                                                  R._raw_stream(idx)))
