"""Compiles the reference's own Python modules of the hot path's CALLERS to bytecode, from the sources where they
lie under /root/reference, into oracle/_ref/pyref/ (git-ignored like every other oracle/_ref output; it travels to
the GPU box with the repo snapshot).  No reference source text enters the repository.

Test infrastructure only (tests/test_dropin_gpu.py): the GPU tests import these modules *sourceless* and run them
UNCHANGED on top of this repo's `depth_diff_gaussian_rasterization_min` / `simple_knn` packages -- the proof by
execution that `gaussian_renderer.render()`, `scene.GaussianModel` and the reference's own operator wrapper drop in.

    package in pyref/            compiled from (under /root/reference)
    gaussian_renderer/__init__   gaussian_renderer/__init__.py            (render(), :18-104)
    scene/gaussian_model         scene/gaussian_model.py                  (GaussianModel)
    utils/{__init__,general,system,graphics,sh}                           (their pure-torch helpers)
    arguments                    arguments.py                             (GSParams defaults)
    refrast/__init__             submodules/depth-diff-gaussian-rasterization-min/
                                 depth_diff_gaussian_rasterization_min/__init__.py   (the reference's operator API;
                                 its `from . import _C` is satisfied by the test with THIS repo's `_C` surface)
"""
import os
import py_compile
import sys

REF = os.environ.get("GS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pyref")

MODULES = [
    ("gaussian_renderer/__init__.py", "gaussian_renderer/__init__.pyc"),
    ("scene/gaussian_model.py", "scene/gaussian_model.pyc"),
    ("utils/__init__.py", "utils/__init__.pyc"),
    ("utils/general.py", "utils/general.pyc"),
    ("utils/system.py", "utils/system.pyc"),
    ("utils/graphics.py", "utils/graphics.pyc"),
    ("utils/sh.py", "utils/sh.pyc"),
    ("arguments.py", "arguments.pyc"),
    ("submodules/depth-diff-gaussian-rasterization-min/depth_diff_gaussian_rasterization_min/__init__.py",
     "refrast/__init__.pyc"),
]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"reference not mounted: {REF}")
        return 1
    for src, dst in MODULES:
        s, d = os.path.join(REF, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        # dfile: the traceback path points back at the reference checkout, not into this repo
        py_compile.compile(s, cfile=d, dfile=s, doraise=True)
    print(f"compiled {len(MODULES)} reference modules -> {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
