"""The on-disk scene format: binary little-endian PLY exactly as GaussianModel.save_ply writes it and
load_ply reads it (scene/gaussian_model.py:179-208, 215-256).

One `vertex` element of float32 properties in this order (construct_list_of_attributes, :179-191):
    x y z | nx ny nz (zeros) | f_dc_0..2 | f_rest_0..(3*(M-1)-1) | opacity | scale_0..2 | rot_0..3
with the SH features stored CHANNEL-major (features.transpose(1, 2).flatten(1): f_rest_k = rest[:, k % (M-1), k // (M-1)]).
All values are the RAW parameters (logit opacity, log scale, un-normalised quaternion).  The reference goes through the
`plyfile` package (plyfile==0.8.1, not in this image); the file layout it produces for such an element is the plain PLY
one restated here: an ASCII header, then P packed little-endian records -- read and written with one numpy
structured-array call.  Pinned by tests/golden/ref_save_ply_*.ply + ply_golden.npz: files written and read back by the
reference's own save_ply / load_ply (executed from /root/reference over a restated plyfile container writer,
tests/golden/make_ply_golden.py); save_ply here must reproduce those bytes, load_ply those tensors."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4):
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation) -> None:
    """Arguments are the raw parameter tensors in the reference's layouts: xyz [P,3], features_dc [P,1,3],
    features_rest [P,M-1,3], opacity [P,1], scaling [P,3], rotation [P,4] (torch, any device, or numpy)."""
    def host(t):
        return (t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)).astype(np.float32, copy=False)
    xyz, fdc, fr, op, sc, rot = map(host, (xyz, features_dc, features_rest, opacity, scaling, rotation))
    P = xyz.shape[0]
    fdc = np.transpose(fdc, (0, 2, 1)).reshape(P, fdc.shape[1] * fdc.shape[2])
    fr = np.transpose(fr, (0, 2, 1)).reshape(P, fr.shape[1] * fr.shape[2])
    table = np.concatenate([xyz, np.zeros_like(xyz), fdc, fr, op.reshape(P, 1), sc, rot], axis=1).astype("<f4")
    names = attribute_names(fdc.shape[1], fr.shape[1], sc.shape[1], rot.shape[1])
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n" + \
             "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(np.ascontiguousarray(table).tobytes())


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1",
              "int8": "i1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4",
              "uint": "u4", "uint32": "u4"}


def read_vertex_table(path: str):
    """Parses the header and returns the first element as a numpy structured array (binary LE/BE or ascii)."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_first, seen_elements = None, None, [], False, 0
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                seen_elements += 1
                in_first = seen_elements == 1
                if in_first:
                    count = int(tok[2])
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties in the vertex element are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or count is None:
            raise ValueError(f"{path}: malformed PLY header")
        if fmt == "ascii":
            rows = np.loadtxt(fh, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=[(n, "<" + t) for n, t in props])
            for k, (n, _t) in enumerate(props):
                out[n] = rows[:, k]
            return out
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, order + t) for n, t in props])
        buf = fh.read(dt.itemsize * count)
        if len(buf) != dt.itemsize * count:
            raise ValueError(f"{path}: truncated PLY body")
        return np.frombuffer(buf, dtype=dt, count=count)


def load_ply(path: str, max_sh_degree: int = 3, device="cuda") -> Dict[str, torch.Tensor]:
    """Raw parameter tensors in the reference's layouts (scene/gaussian_model.py:215-256); raises like the reference's
    assert when the number of f_rest_* properties does not match max_sh_degree."""
    v = read_vertex_table(path)
    names = v.dtype.names
    P = v.shape[0]
    M = (max_sh_degree + 1) ** 2

    def cols(prefix):
        sel = sorted([n for n in names if n.startswith(prefix)], key=lambda s: int(s.split("_")[-1]))
        return np.stack([np.asarray(v[n], dtype=np.float32) for n in sel], 1) if sel else np.zeros((P, 0), np.float32)

    xyz = np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32)
    fdc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], 1).astype(np.float32).reshape(P, 3, 1)
    rest = cols("f_rest_")
    if rest.shape[1] != 3 * M - 3:
        raise AssertionError(f"{path}: {rest.shape[1]} f_rest properties, expected {3 * M - 3} for SH degree {max_sh_degree}")
    rest = rest.reshape(P, 3, M - 1)
    out = dict(xyz=xyz, features_dc=np.ascontiguousarray(fdc.transpose(0, 2, 1)),
               features_rest=np.ascontiguousarray(rest.transpose(0, 2, 1)),
               opacity=np.asarray(v["opacity"], dtype=np.float32)[:, None], scaling=cols("scale_"), rotation=cols("rot"))
    # np.array(...) copies: views of the file buffer (np.frombuffer) are read-only, torch wants writable memory
    return {k: torch.from_numpy(np.array(a, dtype=np.float32, order="C")).to(device) for k, a in out.items()}
