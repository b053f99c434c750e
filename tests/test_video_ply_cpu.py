"""Host side of the video / on-disk row (no GPU): the PLY wire format (header restated from the format spec, field
order from scene/gaussian_model.py:179-191, channel-major SH), the preset-camera loader against the reference's own
camera files (fixture), and the colorize restatement's basic properties."""
import json
import os

import numpy as np
import pytest
import torch

from luciddreamer_b200 import ply, synthetic as syn, video

HERE = os.path.dirname(__file__)


def _model(P, M=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(xyz=torch.randn(P, 3, generator=g), features_dc=torch.randn(P, 1, 3, generator=g),
                features_rest=torch.randn(P, M - 1, 3, generator=g), opacity=torch.randn(P, 1, generator=g),
                scaling=torch.randn(P, 3, generator=g), rotation=torch.randn(P, 4, generator=g))


def test_ply_header_and_record_layout(tmp_path):
    m = _model(7)
    path = str(tmp_path / "point_cloud.ply")
    ply.save_ply(path, **m)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 7"]
    names = [l.split()[2] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    assert names == (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)] +
                     ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])
    rec = np.frombuffer(body, "<f4").reshape(7, 62)
    np.testing.assert_array_equal(rec[:, 0:3], m["xyz"].numpy())
    np.testing.assert_array_equal(rec[:, 3:6], 0)
    np.testing.assert_array_equal(rec[:, 6:9], m["features_dc"][:, 0].numpy())
    # channel-major SH: f_rest_k = rest[:, k % 15, k // 15]   (transpose(1,2).flatten(1), gaussian_model.py:197)
    for k in (0, 1, 14, 15, 31, 44):
        np.testing.assert_array_equal(rec[:, 9 + k], m["features_rest"][:, k % 15, k // 15].numpy())
    np.testing.assert_array_equal(rec[:, 54], m["opacity"][:, 0].numpy())
    np.testing.assert_array_equal(rec[:, 55:58], m["scaling"].numpy())
    np.testing.assert_array_equal(rec[:, 58:62], m["rotation"].numpy())


@pytest.mark.parametrize("tag,deg", [("d3", 3), ("d1", 1)])
def test_ply_bytes_match_the_reference_writer_and_reader(tmp_path, tag, deg):
    """Fixtures made by running the reference's own save_ply / load_ply (tests/golden/make_ply_golden.py): our writer
    produces the same BYTES for the same model, our reader returns what the reference's reader returned for that file."""
    g = np.load(os.path.join(HERE, "golden", "ply_golden.npz"))
    ref_file = os.path.join(HERE, "golden", f"ref_save_ply_{tag}.ply")
    key = dict(xyz="_xyz", features_dc="_features_dc", features_rest="_features_rest", opacity="_opacity",
               scaling="_scaling", rotation="_rotation")
    m = {k: torch.from_numpy(g[f"{tag}_in{v}"]) for k, v in key.items()}
    path = str(tmp_path / "ours.ply")
    ply.save_ply(path, **m)
    assert open(path, "rb").read() == open(ref_file, "rb").read()
    back = ply.load_ply(ref_file, max_sh_degree=deg, device="cpu")
    assert int(g[f"{tag}_active_sh_degree"]) == deg
    for k, v in key.items():
        want = g[f"{tag}_loaded{v}"]
        assert tuple(back[k].shape) == want.shape and back[k].dtype == torch.float32
        np.testing.assert_array_equal(back[k].numpy(), want)


@pytest.mark.parametrize("P", [0, 1, 1000])
def test_ply_round_trip(tmp_path, P):
    m = _model(P, seed=P)
    path = str(tmp_path / "m.ply")
    ply.save_ply(path, **m)
    back = ply.load_ply(path, max_sh_degree=3, device="cpu")
    for k in m:
        assert back[k].shape == m[k].shape and back[k].dtype == torch.float32
        assert torch.equal(back[k], m[k]), k


def test_ply_load_rejects_wrong_sh_degree_and_reads_ascii_and_shuffled_properties(tmp_path):
    m = _model(5, M=4)
    path = str(tmp_path / "d1.ply")
    ply.save_ply(path, **m)
    with pytest.raises(AssertionError):
        ply.load_ply(path, max_sh_degree=3, device="cpu")
    assert torch.equal(ply.load_ply(path, max_sh_degree=1, device="cpu")["features_rest"], m["features_rest"])
    # ascii file with properties in a different order + a comment: looked up by NAME like plyfile does
    names = ply.attribute_names(3, 9)
    t = np.arange(2 * len(names), dtype=np.float32).reshape(2, -1)
    order = list(reversed(range(len(names))))
    with open(tmp_path / "a.ply", "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\n")
        fh.write("".join(f"property float {names[i]}\n" for i in order) + "end_header\n")
        for r in t:
            fh.write(" ".join(str(float(r[i])) for i in order) + "\n")
    b = ply.load_ply(str(tmp_path / "a.ply"), max_sh_degree=1, device="cpu")
    np.testing.assert_array_equal(b["xyz"].numpy(), t[:, 0:3])
    np.testing.assert_array_equal(b["rotation"].numpy(), t[:, -4:])
    np.testing.assert_array_equal(b["features_rest"][:, :, 0].numpy(), t[:, 9:12])


def test_load_json_matches_reference_camera_files(tmp_path):
    """Frames of the reference's cameras/rotate360.json (fixture): load_json -> same matrices as make_camera on the
    converted pose, FoV swap quirk included (utils/camera.py:48)."""
    fx = json.load(open(os.path.join(HERE, "golden", "cameras_fixture.json")))
    frames = [dict(transform_matrix=fx["rotate360"][k]) for k in sorted(fx["rotate360"], key=int)]
    p = tmp_path / "preset.json"
    json.dump(dict(camera_angle_x=fx["camera_angle_x"], frames=frames), open(p, "w"))
    cams = video.load_json(str(p), 72, 128)
    assert len(cams) == len(frames)
    for cam, fr in zip(cams, frames):
        c2w = np.array(fr["transform_matrix"]); c2w[:3, 1:3] *= -1
        if c2w.shape[0] == 3:
            c2w = np.concatenate([c2w, [[0, 0, 0, 1.0]]], 0)
        w2c = np.linalg.inv(c2w)
        np.testing.assert_allclose(cam.viewmatrix.numpy(), np.float32(w2c).T, atol=1e-6)
        assert cam.image_height == 72 and cam.image_width == 128
        fovy = 2 * np.arctan(np.tan(fx["camera_angle_x"] / 2) * 72 / 128)
        assert cam.tanfovx == pytest.approx(np.tan(fovy / 2)) and cam.tanfovy == pytest.approx(np.tan(fx["camera_angle_x"] / 2))


def test_colorize_shape_background_and_monotone_hue():
    d = -np.linspace(0.5, 5.0, 64 * 48, dtype=np.float32).reshape(48, 64)
    d[0, :5] = -99
    img = video.colorize(d)
    assert img.shape == (48, 64, 4) and img.dtype == np.uint8
    assert (img[0, :5] == (128, 128, 128, 255)).all()
    lut = video.jet_lut()
    assert lut.shape == (256, 4) and tuple(lut[0]) == (0, 0, 127, 255) and tuple(lut[-1]) == (127, 0, 0, 255)
    assert lut[96:160, 1].min() == 255                         # green plateau of jet in the middle


def test_jet_lut_known_answers():
    """matplotlib is not installed in this image, so video.colorize cannot be compared with the reference's
    `matplotlib.cm.get_cmap('jet')(x, bytes=True)` (utils/depth.py:45-50) by execution.  Anchors instead: the
    known values of matplotlib's 256-entry jet table -- jet(0) = (0, 0, 0.5), jet(255) = (0.5, 0, 0),
    jet(128) = (0.4901960784313725, 1.0, 0.4775458570524984) -- and the properties of its construction (piecewise
    linear in i / 255 between the published segment knots; bytes=True truncates lut * 255 to uint8)."""
    from luciddreamer_b200 import video
    lut = video.jet_lut()
    assert lut.shape == (256, 4) and lut.dtype == np.uint8 and (lut[:, 3] == 255).all()
    assert tuple(lut[0, :3]) == (0, 0, 127) and tuple(lut[255, :3]) == (127, 0, 0)
    assert tuple(lut[128, :3]) == (int(0.4901960784313725 * 255), 255, int(0.4775458570524984 * 255))
    # knots of the segment data: red rises on [0.35, 0.66], green on [0.125, 0.375], blue falls on [0.34, 0.65]
    f = lut[:, :3].astype(int)
    i = np.arange(256) / 255.0
    assert (f[i <= 0.35, 0] == 0).all() and (f[(i >= 0.66) & (i <= 0.89), 0] == 255).all()
    assert (f[i <= 0.125, 1] == 0).all() and (f[(i >= 0.375) & (i <= 0.64), 1] == 255).all() and (f[i >= 0.91, 1] == 0).all()
    assert (f[(i >= 0.11) & (i <= 0.34), 2] == 255).all() and (f[i >= 0.65, 2] == 0).all()
    for ch, lo, hi in ((0, 0.35, 0.66), (1, 0.125, 0.375)):               # monotone ramps between the knots
        seg = f[(i >= lo) & (i <= hi), ch]
        assert (np.diff(seg) >= 0).all()
    # colorize(): percentile normalisation + invalid handling of utils/depth.py:25-52
    d = np.linspace(1.0, 5.0, 64 * 48, dtype=np.float32).reshape(48, 64)
    d[0, 0] = -99
    img = video.colorize(d)
    assert img.shape == (48, 64, 4) and tuple(img[0, 0]) == (128, 128, 128, 255)
    assert tuple(img[-1, -1, :3]) == (127, 0, 0) and tuple(img[0, 1, :3]) == (0, 0, 127)     # beyond the 98th / 2nd percentile
