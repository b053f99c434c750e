"""Host-side logic of the view-parallel path, on CPU with the gloo backend (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luciddreamer_b200 import multiview as MV
from luciddreamer_b200 import synthetic as syn


def test_shard_views_partitions():
    for n, w in ((64, 8), (8, 8), (7, 3), (1, 4), (0, 2)):
        parts = [MV.shard_views(n, r, w) for r in range(w)]
        flat = sorted(v for p in parts for v in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        MV.shard_views(4, 4, 4)


def test_bucket_layout_is_one_flat_buffer():
    b = MV.GradBucket(64, 16, "cpu")
    assert b.width == 59 and b.flat.numel() == 64 * 59 and b.nbytes() == 64 * 59 * 4
    b.means3D.fill_(1); b.shs.fill_(2); b.opacities.fill_(3); b.scales.fill_(4); b.rotations.fill_(5)
    f = b.flat.numpy()
    assert (f[:192] == 1).all() and (f[192:192 + 3072] == 2).all() and (f[3264:3328] == 3).all()
    assert (f[3328:3520] == 4).all() and (f[3520:] == 5).all()
    odd = MV.GradBucket(10, 16, "cpu")             # ragged P: segments are padded to 256-byte boundaries
    for bb in (b, odd):
        for v in (bb.means3D, bb.shs, bb.opacities, bb.scales, bb.rotations):
            assert v.is_contiguous() and (v.data_ptr() - bb.flat.data_ptr()) % 256 == 0
    assert odd.means3D.shape == (10, 3) and odd.shs.shape == (10, 16, 3) and odd.rotations.shape == (10, 4)


def test_camera_conventions():
    """world_view_transform / full_proj_transform / camera_center as scene/cameras.py builds them."""
    c2w = syn.rotate360_poses(8)[3]
    c2w[:3, 3] = [0.3, -0.2, 0.5]
    cam = syn.make_camera(64, 48, c2w=c2w)
    V = cam.viewmatrix.T.double().numpy()                       # W2C
    assert np.allclose(V @ c2w, np.eye(4), atol=1e-6)
    assert np.allclose(cam.campos.numpy(), c2w[:3, 3], atol=1e-6)
    Pm = syn.projection_matrix(0.01, 100.0, syn.REF_FOVX, 2 * np.arctan(np.tan(syn.REF_FOVX / 2) * 48 / 64)).double().numpy()
    assert np.allclose(cam.projmatrix.T.double().numpy(), Pm @ V, atol=1e-5)
    p = np.array([0.1, 0.2, 3.0, 1.0]) @ np.linalg.inv(c2w).T    # a point 3 units in front
    hom = cam.projmatrix.T.double().numpy() @ (c2w @ np.array([0.1, 0.2, 3.0, 1.0]))
    assert abs(hom[3] - 3.0) < 1e-5                              # w == z_view (A.1)
    sw = syn.make_camera(64, 48, swap_fov_like_load_json=True)
    assert abs(sw.tanfovx - cam.tanfovy) < 1e-12 and abs(sw.tanfovy - cam.tanfovx) < 1e-12


def test_rotate360_matches_reference_fixture():
    """Our rotate360 generator reproduces the frames of the reference preset (fixture extracted from
    cameras/rotate360.json by tests/golden/make_camera_fixture.py)."""
    import json
    p = os.path.join(os.path.dirname(__file__), "golden", "cameras_fixture.json")
    fx = json.load(open(p))
    frames = fx["rotate360"]
    n = fx["rotate360_total"]
    ours = syn.rotate360_poses(n)
    for idx, m in frames.items():
        ref = syn.nerf_c2w_to_colmap(np.array(m))
        assert np.allclose(ours[int(idx)], ref, atol=1e-6) or np.allclose(ours[(n - int(idx)) % n], ref, atol=1e-6)
    assert abs(fx["camera_angle_x"] - syn.REF_FOVX) < 1e-9


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, M = 50, 16
    b = MV.GradBucket(P, M, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    b.flat.copy_(torch.randn(b.flat.numel(), generator=g))
    mine = b.flat.clone()
    MV.allreduce_bucket(b)
    st = MV.DensifyStats(P, "cpu")
    radii = torch.zeros(P, dtype=torch.int32); radii[rank::3] = 5 + rank
    dm2 = torch.randn(P, 3, generator=g)
    st.add_view(dm2, radii)
    local = (st.xyz_gradient_accum.clone(), st.denom.clone(), st.max_radii2D.clone())
    st.allreduce()
    q.put((rank, mine.numpy(), b.flat.numpy(), [t.numpy() for t in local],
           [st.xyz_gradient_accum.numpy(), st.denom.numpy(), st.max_radii2D.numpy()], MV.shard_views(8, rank, world)))
    dist.destroy_process_group()


def test_gloo_world2_allreduce_and_densify_stats():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for r in res:
        assert np.allclose(r[2], total, atol=1e-6)             # every rank holds the summed bucket
        assert np.allclose(r[4][0], res[0][3][0] + res[1][3][0], atol=1e-6)   # norms summed, not norm of sums
        assert np.allclose(r[4][1], res[0][3][1] + res[1][3][1])
        assert np.allclose(r[4][2], np.maximum(res[0][3][2], res[1][3][2]))
    assert sorted(res[0][5] + res[1][5]) == list(range(8))


def test_pair_capacity_hint_is_a_decayed_maximum():
    """a random camera per iteration (luciddreamer.py:291-292) must not re-render whenever a view has
    more pairs than the PREVIOUS one -- the hint keeps the largest recent count and forgets it slowly."""
    from luciddreamer_b200 import rasterizer as R
    idx = 63                                        # a device index nothing else uses
    R._cap_hint.pop(idx, None)
    seq = [400_000, 100_000, 390_000, 120_000, 410_000, 90_000]
    hints = []
    for n in seq:
        R._note_pairs(idx, n)
        hints.append(R._cap_hint[idx])
    assert hints[0] == 400_000 and hints[4] == 410_000
    assert all(h >= n for h, n in zip(hints, seq))                      # never below the count just seen
    assert hints[1] >= 0.97 * 400_000 - 1 and hints[5] >= 0.97 * 410_000 - 1      # a small view does not erase a large one
    for _ in range(200):
        R._note_pairs(idx, 100_000)
    assert R._cap_hint[idx] == 100_000               # ... but the memory fades: 0.97^200 of 410 k is below 100 k
    R._cap_hint.pop(idx, None); R._last_pairs.pop(idx, None)
