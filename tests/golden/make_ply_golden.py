"""Golden bytes of the on-disk scene format, produced by EXECUTING the reference's own GaussianModel.save_ply and
GaussianModel.load_ply (scene/gaussian_model.py:179-208, 215-256) imported from /root/reference in the build container.

The reference writes through the third-party `plyfile` package (requirements.txt: plyfile==0.8.1), which is not in this
image.  The stand-in below restates the two calls the reference makes on it, after plyfile 0.8.1's documented behaviour:
  * PlyElement.describe(structured_array, name) + PlyData([el]).write(path): header lines "ply",
    "format binary_little_endian 1.0" (native byte order of the array on a little-endian host), "element <name> <count>",
    one "property <type> <field>" per dtype field with numpy 'f4' spelled "float", "end_header"; then the records, packed,
    in the file's byte order;
  * PlyData.read(path): .elements[0][field] -> column, .elements[0].properties -> objects with .name, header order.
So the pin is: the reference's field order / transposes / dtypes by execution, the container format by restatement.
simple_knn._C (CUDA, import-time only) is stubbed; load_ply builds its tensors with device="cuda", which this GPU-less
container cannot do, so torch.tensor is wrapped to drop the device argument while the reference function runs.

Run: python tests/golden/make_ply_golden.py   (writes ref_save_ply_d3.ply, ref_save_ply_d1.ply, ply_golden.npz)"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- stand-in for plyfile 0.8.1 (only what gaussian_model.py touches) ------------------------------------------------
_TYPE_NAME = {"f4": "float", "f8": "double", "i4": "int", "u4": "uint", "i2": "short", "u2": "ushort", "i1": "char", "u1": "uchar"}
_TYPE_CODE = {v: k for k, v in _TYPE_NAME.items()}


class _Prop:
    def __init__(self, name, code):
        self.name, self.code = name, code


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = [_Prop(n, data.dtype.fields[n][0].str[1:]) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        assert isinstance(data, np.ndarray) and data.ndim == 1 and data.dtype.names
        return PlyElement(name, data)

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        with open(path, "wb") as fh:
            lines = ["ply", "format binary_little_endian 1.0"]
            for el in self.elements:
                lines.append(f"element {el.name} {len(el.data)}")
                lines += [f"property {_TYPE_NAME[p.code]} {p.name}" for p in el.properties]
            lines.append("end_header")
            fh.write(("\n".join(lines) + "\n").encode("ascii"))
            for el in self.elements:
                el.data.astype(el.data.dtype.newbyteorder("<"), copy=False).tofile(fh)

    @staticmethod
    def read(path):
        with open(path, "rb") as fh:
            assert fh.readline() == b"ply\n"
            assert fh.readline() == b"format binary_little_endian 1.0\n"
            name, count, fields = None, 0, []
            while True:
                tok = fh.readline().decode("ascii").split()
                if tok[0] == "element":
                    assert name is None, "one element is all the reference writes"
                    name, count = tok[1], int(tok[2])
                elif tok[0] == "property":
                    fields.append((tok[2], "<" + _TYPE_CODE[tok[1]]))
                elif tok[0] == "end_header":
                    break
            data = np.fromfile(fh, dtype=np.dtype(fields), count=count)
        return PlyData([PlyElement(name, data)])


mod = types.ModuleType("plyfile")
mod.PlyData, mod.PlyElement = PlyData, PlyElement
sys.modules["plyfile"] = mod
knn = types.ModuleType("simple_knn")
knn._C = types.ModuleType("simple_knn._C")
knn._C.distCUDA2 = None
sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn._C
sys.path.insert(0, "/root/reference")
pkg = types.ModuleType("scene")                 # scene/__init__.py pulls in imageio etc.; only gaussian_model.py is wanted
pkg.__path__ = ["/root/reference/scene"]
sys.modules["scene"] = pkg
from scene.gaussian_model import GaussianModel  # noqa: E402


def model(P, deg, seed):
    g = torch.Generator().manual_seed(seed)
    M = (deg + 1) ** 2
    m = types.SimpleNamespace()
    m._xyz = torch.randn(P, 3, generator=g)
    m._features_dc = torch.randn(P, 1, 3, generator=g)
    m._features_rest = torch.randn(P, M - 1, 3, generator=g)
    m._opacity = torch.randn(P, 1, generator=g)
    m._scaling = torch.randn(P, 3, generator=g)
    m._rotation = torch.randn(P, 4, generator=g)
    m.max_sh_degree = deg
    m.construct_list_of_attributes = lambda: GaussianModel.construct_list_of_attributes(m)
    return m


FIELDS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")
out = {}
for tag, P, deg, seed in (("d3", 6, 3, 11), ("d1", 3, 1, 12)):
    m = model(P, deg, seed)
    path = os.path.join(HERE, f"ref_save_ply_{tag}.ply")
    GaussianModel.save_ply(m, path)                                   # the reference's writer
    back = types.SimpleNamespace(max_sh_degree=deg)
    real_tensor = torch.tensor
    torch.tensor = lambda *a, device=None, **k: real_tensor(*a, **k)   # no GPU here; values do not depend on the device
    try:
        GaussianModel.load_ply(back, path)                            # the reference's reader on the reference's file
    finally:
        torch.tensor = real_tensor
    for f in FIELDS:
        out[f"{tag}_in{f}"] = getattr(m, f).numpy()
        out[f"{tag}_loaded{f}"] = getattr(back, f).detach().numpy()
        assert np.array_equal(out[f"{tag}_in{f}"], out[f"{tag}_loaded{f}"]), f   # the reference round-trips itself
    out[f"{tag}_active_sh_degree"] = np.int32(back.active_sh_degree)
    print(tag, os.path.getsize(path), "bytes")
np.savez(os.path.join(HERE, "ply_golden.npz"), **out)
