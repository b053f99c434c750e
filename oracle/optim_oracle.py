"""CPU oracle of the fused optimiser step (TEST INFRASTRUCTURE ONLY).

Runs what the reference runs: the parameter activations of scene/gaussian_model.py:97-117 (sigmoid, exp, F.normalize,
cat) under autograd and torch.optim.Adam(lr=0, eps=1e-15) with the six groups of gaussian_model.py:156-165, driven
by given gradients w.r.t. the ACTIVATED values (the rasterizer's outputs)."""
from __future__ import annotations

import torch


class ReferenceStepper:
    def __init__(self, xyz, f_dc, f_rest, opacity, scaling, rotation, lrs, dtype=torch.float32):
        mk = lambda t: torch.nn.Parameter(t.detach().cpu().to(dtype).clone())
        self.p = {"xyz": mk(xyz), "f_dc": mk(f_dc), "f_rest": mk(f_rest), "opacity": mk(opacity), "scaling": mk(scaling),
                  "rotation": mk(rotation)}
        self.opt = torch.optim.Adam([{"params": [self.p[k]], "lr": lrs[k], "name": k} for k in
                                     ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")], lr=0.0, eps=1e-15)
        self.dtype = dtype

    def step(self, g_means3D, g_shs, g_opac, g_scales, g_rots):
        p = self.p
        acts = (p["xyz"], torch.cat((p["f_dc"], p["f_rest"]), dim=1), torch.sigmoid(p["opacity"]), torch.exp(p["scaling"]),
                torch.nn.functional.normalize(p["rotation"]))
        gs = [t.detach().cpu().to(self.dtype) for t in (g_means3D, g_shs, g_opac, g_scales, g_rots)]
        self.opt.zero_grad(set_to_none=True)
        torch.autograd.backward(acts, gs)
        self.opt.step()

    def state(self, name):
        st = self.opt.state[self.p[name]]
        return self.p[name].detach(), st["exp_avg"], st["exp_avg_sq"]
