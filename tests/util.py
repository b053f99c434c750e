"""Shared helpers of the parity tests."""
import os

import numpy as np

import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances: forward colour/depth within 1e-4 abs, gradients within 1e-3 rel of the reference
FWD_ABS_TOL = 1e-4
GRAD_REL_TOL = 1e-3


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    out = dict(num_rendered=int(z["num_rendered"]), color=z["color"], depth=z["depth"], radii=z["radii"])
    grads = {}
    for g in cases.GRAD_NAMES:
        shape = tuple(int(x) for x in z[g + "_shape"])
        grads[g] = cases.dense_rows(z[g + "_idx"], z[g + "_rows"], shape)
        out[g + "_jitter"] = float(z[g + "_jitter"])
    out["grads"] = grads
    return out


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def assert_forward_close(color, depth, radii, gold, what="", audit=None, flip_div=20000):
    """Colour/depth within 1e-4 abs.  Discontinuities (alpha<1/255, T<1e-4, the
    acc>0.5 depth gate) may flip for isolated pixels when two implementations differ in the last ulp; those
    are counted and bounded instead of hidden: at most 1 pixel in 20 000 may exceed the tolerance, and -- when the
    oracle's forward state is passed as `audit` -- every one of them must be ATTRIBUTED to a branch operand sitting on
    its threshold (tests/flip_audit.py); an unexplained outlier fails."""
    color = np.asarray(color); depth = np.asarray(depth)
    dc = np.abs(color - gold["color"]).max(axis=0)
    dd = np.abs(depth - gold["depth"]).reshape(dc.shape)
    n = dc.size
    bad_c, bad_d = int((dc > FWD_ABS_TOL).sum()), int((dd > FWD_ABS_TOL).sum())
    assert bad_c <= n // flip_div, f"{what}: {bad_c}/{n} pixels off by > {FWD_ABS_TOL} in colour (max {dc.max():.3e})"
    assert bad_d <= n // flip_div, f"{what}: {bad_d}/{n} pixels off by > {FWD_ABS_TOL} in depth (max {dd.max():.3e})"
    if audit is not None and (bad_c or bad_d):
        import flip_audit
        explained, unexplained = flip_audit.explain_outliers(audit, color, depth, gold["color"], gold["depth"], FWD_ABS_TOL)
        print(f"[flip audit] {what}: {len(explained)} outlier pixel(s) attributed to branch flips: {explained}")
        assert not unexplained, f"{what}: outliers with no branch operand near a threshold: {unexplained}"
    mism = int((np.asarray(radii) != gold["radii"]).sum())
    assert mism <= max(0, len(gold["radii"]) // 50000), f"{what}: {mism} radii differ"
    return dict(max_color=float(dc.max()), max_depth=float(dd.max()), flips_color=bad_c, flips_depth=bad_d,
                radii_mismatch=mism)


def assert_grads_close(grads, gold_grads, names=None, tol=GRAD_REL_TOL, what=""):
    out = {}
    for k, g in grads.items():
        if names is not None and k not in names:
            continue
        ref = gold_grads[k]
        if ref.size == 0:
            continue
        e = rel_err(np.asarray(g).reshape(ref.shape), ref)
        out[k] = e
        assert e <= tol, f"{what}: {k} rel err {e:.3e} > {tol}"
    return out
