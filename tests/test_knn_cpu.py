"""distCUDA2 oracle (no GPU): the brute-force C restatement against hand-computable cases, against an independent
float64 k-d tree, and its edge cases (duplicates, P < 4)."""
import os

import numpy as np

from oracle import knn_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "knn_golden.npz")


def test_oracle_bit_exact_vs_reference_golden():
    """Pins the oracle (incl. the FMA contraction of the distance) to outputs of the reference's own simple_knn kernels
    (tests/golden/make_knn_golden.py, run on a B200)."""
    g = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 4
    for n in names:
        np.testing.assert_array_equal(knn_oracle.dist2(g[n + "/points"]), g[n + "/dist2"], err_msg=n)


def test_lattice_known_answer():
    """Unit cubic lattice: every interior point has 6 neighbours at distance 1 -> mean of 3 nearest squared = 1;
    a corner has exactly 3 at distance 1; scaled lattice scales quadratically."""
    g = np.stack(np.meshgrid(np.arange(5), np.arange(5), np.arange(5), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    np.testing.assert_array_equal(knn_oracle.dist2(g), np.ones(len(g), np.float32))
    np.testing.assert_array_equal(knn_oracle.dist2(g * 0.5), np.full(len(g), 0.25, np.float32))


def test_hand_case_and_duplicates():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [10, 0, 0], [0, 0, 0]], np.float32)
    out = knn_oracle.dist2(pts)
    # point 0: duplicate (0), then 1, then 4 -> 5/3 ; point 4: 81, 100, 100 -> 281/3
    assert out[0] == np.float32(np.float32(0 + 1 + 4) / np.float32(3))
    assert out[5] == out[0]
    assert out[4] == np.float32(np.float32(81 + 100 + 100) / np.float32(3))


def test_fewer_than_four_points_overflow_like_reference():
    """Missing neighbours stay FLT_MAX (simple_knn.cu:154): two of them overflow the float sum to inf, one gives
    ~FLT_MAX / 3."""
    for P in (1, 2):
        out = knn_oracle.dist2(np.random.default_rng(P).normal(size=(P, 3)).astype(np.float32))
        assert np.all(np.isinf(out))
    out = knn_oracle.dist2(np.random.default_rng(3).normal(size=(3, 3)).astype(np.float32))
    assert np.all(np.isfinite(out)) and np.all(out > 1e38)


def test_matches_independent_kdtree():
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.normal(size=(4000, 3)), 0.01 * rng.normal(size=(3000, 3)) + 2.0,
                          rng.uniform(-5, 5, size=(3000, 3)) * [1, 1, 0]]).astype(np.float32)     # clusters + a plane
    np.testing.assert_allclose(knn_oracle.dist2(pts), knn_oracle.dist2_kdtree(pts), rtol=2e-5, atol=1e-12)
