"""Times the video-render path (BASELINE config 4 shape: forward-only view batch at 1080p, 1 M Gaussians):
  ours      = luciddreamer_b200.video.render_video_frames (device-side packing, one host copy per clip)
  reference = the reference rasterizer (oracle/_ref) driven by the per-frame loop of luciddreamer.py:250-262
              (two blocking float D2H copies + numpy per frame).
Usage (GPU box): python tools/video_bench.py [frames]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from luciddreamer_b200 import GaussianRasterizationSettings, synthetic as syn, video


def run(F=48, impls=("ours", "reference")):
    ref_cuda = None
    if "reference" in impls:
        from oracle import ref_cuda          # the reference rasterizer (oracle/_ref), only when asked for
    d = torch.device("cuda:0")
    c = syn.CONFIGS[3]
    P, W, H = c["P"], c["W"], c["H"]
    sc = {k: v.to(d) for k, v in syn.make_scene(P, 1003).items()}
    poses = syn.rotate360_poses(F)
    cams = [syn.make_camera(W, H, c2w=poses[k]) for k in range(F)]
    settings = [GaussianRasterizationSettings(H, W, cm.tanfovx, cm.tanfovy, torch.zeros(3, device=d), 1.0, cm.viewmatrix.to(d),
                                              cm.projmatrix.to(d), 3, cm.campos.to(d), False, False) for cm in cams]


    def ours():
        return video.render_video_frames(sc, settings)


    def reference():
        rc = ref_cuda.RefContext()
        framelist, depthlist, dmin, dmax = [], [], 1e8, -1e8
        for cm, rs in zip(cams, settings):
            _R, color, depth, _rad = ref_cuda.rasterize_gaussians(rc, rs.bg, sc["means3D"], None, sc["opacities"], sc["scales"],
                                                                  sc["rotations"], 1.0, None, rs.viewmatrix, rs.projmatrix,
                                                                  cm.tanfovx, cm.tanfovy, H, W, sc["shs"], 3, rs.campos)
            framelist.append(np.round(color.permute(1, 2, 0).detach().cpu().numpy().clip(0, 1) * 255.).astype(np.uint8))
            dd = -(depth * (depth > 0)).detach().cpu().numpy()
            dmin, dmax = min(dmin, dd.min().item()), max(dmax, dd.max().item())
            depthlist.append(dd)
        return framelist, depthlist, dmin, dmax


    out = {"frames": F, "H": H, "W": W, "P": P}
    res = {}
    for name, fn in (("ours", ours if "ours" in impls else None),
                     ("reference", reference if ref_cuda is not None and ref_cuda.available() else None)):
        if fn is None:
            continue
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = r
        out[name + "_ms_per_frame"] = dt / F * 1e3
        out[name + "_fps"] = F / dt
    if "reference" in res and "ours" in res:
        a, b = res["ours"], res["reference"]
        diff = np.abs(a[0].astype(np.int16) - np.stack(b[0]).astype(np.int16))
        out["max_uint8_diff_vs_reference"] = int(diff.max())
        out["pixels_differing"] = float((diff.max(-1) > 0).mean())
        out["dmin_dmax"] = [a[2], a[3], b[2], b[3]]
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 48)))
