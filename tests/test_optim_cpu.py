"""Host logic of the optimiser row (no GPU): the xyz learning-rate schedule against golden values produced by the
reference's own function, and the torch oracle of the fused Adam step against hand-derived activation gradients."""
import json
import os

import numpy as np
import pytest
import torch

from luciddreamer_b200 import optim
from oracle import optim_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lr_golden.json")


def test_expon_lr_matches_reference_golden():
    for case in json.load(open(GOLD)):
        a = case["args"]
        for step, want in zip(case["steps"], case["lr"]):
            got = optim.expon_lr(step, a["lr_init"], a["lr_final"], a.get("lr_delay_steps", 0), a.get("lr_delay_mult", 1.0),
                                 a["max_steps"])
            assert got == pytest.approx(want, rel=1e-12, abs=0.0), (a, step)


def test_oracle_first_step_is_sign_of_activation_gradient():
    """After one Adam step from zero moments the update is -lr * sign(dL/draw) (bias-corrected m / sqrt(v) = +-1), with
    dL/draw the chain rule through sigmoid / exp / normalize: pins the oracle's use of the reference's activations."""
    P = 50
    g = torch.Generator().manual_seed(1)
    raw = dict(xyz=torch.randn(P, 3, generator=g), f_dc=torch.randn(P, 1, 3, generator=g), f_rest=torch.randn(P, 15, 3, generator=g),
               opacity=torch.randn(P, 1, generator=g), scaling=torch.randn(P, 3, generator=g), rotation=torch.randn(P, 4, generator=g))
    lrs = dict(optim.DEFAULT_LRS)
    st = optim_oracle.ReferenceStepper(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"], lrs,
                                       dtype=torch.float64)
    grads = [torch.randn(P, 3, generator=g), torch.randn(P, 16, 3, generator=g), torch.randn(P, 1, generator=g),
             torch.randn(P, 3, generator=g), torch.randn(P, 4, generator=g)]
    st.step(*grads)
    s = torch.sigmoid(raw["opacity"].double())
    q = raw["rotation"].double(); n = q.norm(dim=1, keepdim=True); h = q / n
    gq = grads[4].double()
    want = dict(xyz=grads[0].double(), f_dc=grads[1][:, :1].double(), f_rest=grads[1][:, 1:].double(),
                opacity=grads[2].double() * s * (1 - s), scaling=grads[3].double() * torch.exp(raw["scaling"].double()),
                rotation=(gq - h * (h * gq).sum(1, keepdim=True)) / n)
    for k in want:
        p, m, v = st.state(k)
        np.testing.assert_allclose(m.numpy(), 0.1 * want[k].numpy(), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(p.numpy(), (raw[k].double() - lrs[k] * torch.sign(want[k])).numpy(), rtol=1e-8, atol=1e-10)
