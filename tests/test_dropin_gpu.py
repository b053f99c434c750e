"""Drop-in by EXECUTION: the reference's own Python -- `gaussian_renderer.render()` (gaussian_renderer/__init__.py:18-104),
`scene.GaussianModel` (scene/gaussian_model.py:27-407) and its operator wrapper (RAST/depth_diff_gaussian_rasterization_min/
__init__.py:21-221) -- run UNCHANGED on top of this repository's packages.

The reference checkout does not exist on the GPU box, and its sources may not be copied into the repo; what travels is
BYTECODE compiled by oracle/build_pyref.py from the sources where they lie (`make -C oracle pyref`, part of
`__graft_entry__.build()`), into the git-ignored oracle/_ref/pyref/.  The modules are imported sourceless here:

    gaussian_renderer, scene.gaussian_model, utils.*, arguments      <- the reference's
    depth_diff_gaussian_rasterization_min, simple_knn                <- THIS repo's packages (repo root)
    refrast                                                          <- the reference's operator wrapper, whose
                                                                        `from . import _C` gets this repo's `_C`
`plyfile` (absent in this image, only used by save_ply / load_ply) is stubbed.  Results are checked against the CPU oracle.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

import cases
import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
needs_pyref = pytest.mark.skipif(not os.path.exists(os.path.join(PYREF, "gaussian_renderer", "__init__.pyc")),
                                 reason="oracle/_ref/pyref not built (make -C oracle pyref needs /root/reference)")


@pytest.fixture(scope="module")
def refmods():
    """Imports the reference's bytecode with this repo's packages underneath."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, PYREF)
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = object
    sys.modules["plyfile"] = ply
    for name in [m for m in sys.modules if m.split(".")[0] in ("utils", "scene", "gaussian_renderer", "arguments", "refrast")]:
        del sys.modules[name]
    import depth_diff_gaussian_rasterization_min as ours_rast           # this repo (alias package at the repo root)
    import simple_knn._C as ours_knn                                    # this repo
    assert os.path.dirname(os.path.abspath(ours_rast.__file__)).startswith(ROOT)
    assert os.path.abspath(ours_knn.__file__).startswith(ROOT)
    shim = types.ModuleType("refrast._C")                               # what RAST/ext.cpp:15-19 exports, from OUR `_C`
    for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        setattr(shim, n, getattr(ours_rast._C, n))
    sys.modules["refrast._C"] = shim
    import arguments
    import gaussian_renderer
    import refrast
    from scene.gaussian_model import GaussianModel
    from utils.graphics import BasicPointCloud
    for mod in (gaussian_renderer, refrast, sys.modules["scene.gaussian_model"]):
        assert mod.__file__.startswith(PYREF) and mod.__file__.endswith(".pyc")      # the reference's bytecode, not ours
    assert gaussian_renderer.GaussianRasterizer is ours_rast.GaussianRasterizer          # ... calling OUR operator
    yield types.SimpleNamespace(render=gaussian_renderer.render, GaussianModel=GaussianModel, GSParams=arguments.GSParams,
                                BasicPointCloud=BasicPointCloud, refrast=refrast, ours=ours_rast)
    sys.path[:] = saved_path
    # drop only what this fixture put there: torch imports modules lazily (optimizer.step() pulls in torch._dynamo /
    # torch._inductor), and deleting those would make a later import re-run their TORCH_LIBRARY registrations
    ours = ("utils", "scene", "gaussian_renderer", "arguments", "refrast", "plyfile")
    for name in [m for m in sys.modules if m not in saved_mods and m.split(".")[0] in ours]:
        del sys.modules[name]


def _camera(case, inp, dev):
    cam = inp["cam"]
    import math
    return types.SimpleNamespace(FoVx=2.0 * math.atan(cam.tanfovx), FoVy=2.0 * math.atan(cam.tanfovy),
                                 image_height=cam.image_height, image_width=cam.image_width,
                                 world_view_transform=cam.viewmatrix.to(dev), full_proj_transform=cam.projmatrix.to(dev),
                                 camera_center=cam.campos.to(dev))


@needs_pyref
def test_reference_render_and_gaussian_model_run_unchanged(refmods):
    """One iteration of the reference's optimisation loop (luciddreamer.py:283-327) with the reference's OWN
    GaussianModel and render(): create_from_pcd (-> simple_knn.distCUDA2, ours), training_setup, render (-> our
    rasterizer), loss.backward, add_densification_stats, optimizer.step -- checked against the CPU oracle."""
    from oracle import oracle, knn_oracle
    dev = torch.device("cuda:0")
    case = cases.BY_NAME["micro_1k_64"]
    inp = cases.build_inputs(case)
    g = torch.Generator().manual_seed(5)
    colors = torch.rand(case.P, 3, generator=g)
    pcd = refmods.BasicPointCloud(points=inp["means3D"].numpy(), colors=colors.numpy(), normals=np.zeros((case.P, 3)))
    gm = refmods.GaussianModel(3)
    gm.create_from_pcd(pcd, 1.0)                      # reference code; distCUDA2 is this repo's kernel
    # initial scales = sqrt(mean squared distance to the 3 nearest neighbours): bit-identical to the brute-force oracle
    d2 = knn_oracle.dist2(inp["means3D"].numpy())
    assert np.allclose(gm.get_scaling[:, 0].detach().cpu().numpy(), np.sqrt(np.maximum(d2, 1e-7)), rtol=2e-6, atol=0)
    opt = refmods.GSParams()
    gm.training_setup(opt)
    gm.active_sh_degree = 2                           # luciddreamer.py:287-288 ramps it during training
    with torch.no_grad():                             # make the scene interesting: the case's scales / opacities / SH
        gm._scaling.copy_(torch.log(inp["scales"].to(dev)))
        gm._rotation.copy_(inp["rotations"].to(dev) * 1.7)       # un-normalised on purpose: get_rotation normalises
        gm._opacity.copy_(torch.logit(inp["opacities"].to(dev).clamp(1e-4, 1 - 1e-4)))
        gm._features_rest.copy_(inp["shs"][:, 1:].to(dev))
    cam = _camera(case, inp, dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    pkg = refmods.render(cam, gm, opt, bg)            # the reference's render(), unchanged
    image, depth, vsp, vis, radii = pkg["render"], pkg["depth"], pkg["viewspace_points"], pkg["visibility_filter"], pkg["radii"]
    cot = inp["cot"].to(dev)
    (image * cot).sum().backward()
    torch.cuda.synchronize()

    # oracle on the same post-activation tensors
    with torch.no_grad():
        act = dict(means3D=gm.get_xyz.cpu(), opac=gm.get_opacity.cpu(), scales=gm.get_scaling.cpu(),
                   rots=gm.get_rotation.cpu(), shs=gm.get_features.cpu())
    c = inp["cam"]
    f = oracle.rasterize_gaussians(bg.cpu(), act["means3D"], None, act["opac"], act["scales"], act["rots"], 1.0, None,
                                   c.viewmatrix, c.projmatrix, c.tanfovx, c.tanfovy, c.image_height, c.image_width,
                                   act["shs"], 2, c.campos)
    og = oracle.rasterize_gaussians_backward(f, inp["cot"].numpy())
    gold = dict(color=f.color, depth=f.depth, radii=f.radii)
    util.assert_forward_close(image.detach().cpu().numpy(), depth.detach().cpu().numpy(), radii.cpu().numpy(), gold,
                              what="reference render() on our rasterizer", audit=f)
    assert torch.equal(vis.cpu(), torch.from_numpy(f.radii > 0))
    assert util.rel_err(vsp.grad.cpu().numpy(), og[0]) <= util.GRAD_REL_TOL          # dL_dmeans2D: the densification signal
    assert util.rel_err(gm._xyz.grad.cpu().numpy(), og[3]) <= util.GRAD_REL_TOL       # identity activation -> dL_dmeans3D
    # chain rule through the reference's activations, against the oracle's post-activation gradients
    s = act["opac"].double().numpy()
    assert util.rel_err(gm._opacity.grad.cpu().numpy(), og[2] * s * (1 - s)) <= util.GRAD_REL_TOL
    assert util.rel_err(gm._scaling.grad.cpu().numpy(), og[6] * act["scales"].double().numpy()) <= util.GRAD_REL_TOL
    # the rest of the iteration (luciddreamer.py:306-327)
    before = gm._xyz.detach().clone()
    gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])
    gm.add_densification_stats(vsp, vis)
    gm.optimizer.step()
    gm.optimizer.zero_grad(set_to_none=True)
    assert float(gm.denom.sum()) == float(vis.sum()) and float((gm._xyz.detach() - before).abs().max()) > 0
    # forward only, as render_video does (luciddreamer.py:250-262)
    out = refmods.render(cam, gm, opt, bg, render_only=True)
    assert set(out) == {"render", "depth"} and out["render"].shape == (3, case.H, case.W)


@needs_pyref
@pytest.mark.parametrize("name", ["micro_1k_64", "precomp_2k_80", "odd_3k_100x70_bg"])
def test_reference_operator_wrapper_on_our_C(refmods, name):
    """The reference's own `__init__.py` (GaussianRasterizer / _RasterizeGaussians / rasterize_gaussians) with its `_C`
    bound to this repo's `_C` surface: forward + backward against the golden vectors of the reference CUDA extension."""
    R = refmods.refrast
    dev = torch.device("cuda:0")
    case = cases.BY_NAME[name]
    inp = cases.build_inputs(case)
    gold = util.load_golden(case.name)
    cam = inp["cam"]
    rs = R.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, inp["bg"].to(dev),
                                         case.scale_modifier, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), case.D,
                                         cam.campos.to(dev), False, False)
    leaves = {k: inp[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations",
                                                               "colors_precomp", "cov3D_precomp") if inp[k] is not None}
    m2 = torch.zeros(case.P, 3, device=dev, requires_grad=True)
    color, radii, depth = R.GaussianRasterizer(rs)(
        leaves["means3D"], m2, leaves["opacities"], shs=leaves.get("shs"), colors_precomp=leaves.get("colors_precomp"),
        scales=leaves.get("scales"), rotations=leaves.get("rotations"), cov3D_precomp=leaves.get("cov3D_precomp"))
    torch.autograd.backward(color, grad_tensors=inp["cot"].to(dev))
    torch.cuda.synchronize()
    util.assert_forward_close(color.detach().cpu().numpy(), depth.detach().cpu().numpy(), radii.cpu().numpy(), gold,
                              what="reference wrapper on our _C: " + name)
    got = {"dL_dmeans3D": leaves["means3D"].grad, "dL_dmeans2D": m2.grad, "dL_dopacity": leaves["opacities"].grad}
    if "shs" in leaves:
        got.update(dL_dsh=leaves["shs"].grad, dL_dscales=leaves["scales"].grad, dL_drotations=leaves["rotations"].grad)
    else:
        got.update(dL_dcolors=leaves["colors_precomp"].grad, dL_dcov3D=leaves["cov3D_precomp"].grad)
    util.assert_grads_close({k: v.cpu().numpy() for k, v in got.items()}, gold["grads"], what=name)
    # markVisible through the reference wrapper
    present = R.GaussianRasterizer(rs).markVisible(leaves["means3D"].detach())
    from oracle import oracle
    assert np.array_equal(present.cpu().numpy(), oracle.mark_visible(inp["means3D"], cam.viewmatrix))
