"""The oracle's KD-tree search, voxel grid and feature matching against brute-force NumPy (tests/np_ref.py)."""
import numpy as np
import pytest

import np_ref


def cloud(n, seed, lattice=False):
    rng = np.random.default_rng(seed)
    if lattice:  # exact ties in distance: exercises the (d2, index) tie rule
        return rng.integers(0, 6, (n, 3)).astype(np.float32) * np.float32(0.25)
    return rng.uniform(-1, 1, (n, 3)).astype(np.float32)


@pytest.mark.parametrize("n,lattice,radius,k", [(300, False, 0.4, 10), (500, True, 0.6, 30), (64, False, 5.0, 100),
                                                (200, True, 0.26, 7), (5, False, 1.0, 30), (1, False, 1.0, 3)])
def test_knn_hybrid_matches_bruteforce(orc, n, lattice, radius, k):
    pts = cloud(n, n + 7, lattice)
    q = np.concatenate([pts[: min(n, 50)], cloud(20, 99)]).astype(np.float32)
    idx, d2, cnt = orc.knn_hybrid(pts, q, radius, k)
    ridx, rd2, rcnt = np_ref.knn_hybrid(pts, q, radius, k)
    assert np.array_equal(cnt, rcnt)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(d2, rd2)


def test_nn1_radius_is_strict_and_empty_inputs(orc):
    tgt = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    q = np.array([[0.5, 0, 0], [3, 0, 0], [0.25, 0, 0]], np.float32)
    idx, d2 = orc.nn1(tgt, q, 0.5)  # d2 == r2 is NOT a neighbour (A.3: strict <)
    assert list(idx) == [-1, -1, 0]
    idx, d2 = orc.nn1(tgt, q, 0.5000001)
    assert list(idx) == [0, -1, 0]  # tie at 0.5 -> lowest index
    idx, _ = orc.nn1(np.zeros((0, 3), np.float32), q, 1.0)
    assert list(idx) == [-1, -1, -1]
    idx, _ = orc.nn1(tgt, np.zeros((0, 3), np.float32), 1.0)
    assert len(idx) == 0


@pytest.mark.parametrize("n,voxel", [(2000, 0.1), (500, 0.5), (100, 3.0), (1, 0.3)])
def test_voxel_downsample(orc, n, voxel):
    pts = cloud(n, 3)
    out = orc.voxel_downsample(pts, voxel)
    ref = np_ref.voxel_downsample(pts, voxel)
    assert out.shape == ref.shape
    # fixed-point sums (rule D5) agree with fp64 means far below fp32 resolution
    assert np.allclose(out, ref, rtol=0, atol=1.2e-7)
    with pytest.raises(ValueError):
        orc.voxel_downsample(pts, 0.0)
    assert orc.voxel_downsample(np.zeros((0, 3), np.float32), 0.3).shape == (0, 3)


def test_voxel_downsample_on_grids_far_beyond_a_dense_table(orc):
    """The oracle sorts 64-bit voxel keys, so the extent / voxel ratio is unbounded (Open3D hashes the index): two dense
    clusters 10^5 units apart plus isolated points, 10^15 voxel ids.  Checked against the NumPy restatement (np.unique on
    the same keys, fp64 means) — this is the case the device's sorted-key path is compared with on the GPU."""
    rng = np.random.default_rng(11)
    a = cloud(1500, 7) * 2.0
    b = a[:700] + np.array([1e5, -3e4, 2e4], np.float32)
    lone = rng.uniform(-5e4, 5e4, (300, 3)).astype(np.float32)
    pts = np.concatenate([a, b, lone]).astype(np.float32)
    voxel = 0.3
    ext = pts.max(0).astype(np.float64) - pts.min(0)
    assert np.prod(np.floor(ext / voxel) + 1) > 1e15
    out = orc.voxel_downsample(pts, voxel)
    ref = np_ref.voxel_downsample(pts, voxel)
    assert out.shape == ref.shape and len(out) < len(pts)
    # means of fp32 coordinates of magnitude 1e5: the fixed-point sums agree with fp64 means to half an fp32 ulp there
    assert np.allclose(out, ref, rtol=0, atol=4e-3) and np.allclose(out[np.abs(ref).max(1) < 10], ref[np.abs(ref).max(1) < 10], rtol=0, atol=1e-6)
    with pytest.raises(ValueError):  # a dimension beyond int32
        orc.voxel_downsample(pts, 1e-6)


def test_voxel_is_order_independent(orc):
    pts = cloud(3000, 5)
    a = orc.voxel_downsample(pts, 0.2)
    perm = np.random.default_rng(0).permutation(len(pts))
    b = orc.voxel_downsample(pts[perm], 0.2)
    assert np.array_equal(a, b)  # bit-exact: int64 fixed-point sums


def test_feature_matching(orc):
    rng = np.random.default_rng(11)
    fs = rng.uniform(0, 200, (120, 33)).astype(np.float32)
    ft = rng.uniform(0, 200, (150, 33)).astype(np.float32)
    ft[10] = fs[3]
    ft[20] = fs[3]          # exact tie -> lowest index (10)
    fs[50] = 0.0
    ft[60] = 0.0
    ft[70] = 0.0            # all-zero descriptors tie exactly
    nn = orc.nn_features(fs, ft)
    assert np.array_equal(nn, np_ref.nn_features(fs, ft))
    assert nn[3] == 10 and nn[50] == 60
    one = orc.match_features(fs, ft, False)
    assert np.array_equal(one[:, 0], np.arange(120)) and np.array_equal(one[:, 1], nn)
    nn_t = np_ref.nn_features(ft, fs)
    mutual = np.array([(i, j) for i, j in enumerate(nn) if nn_t[j] == i], np.int32).reshape(-1, 2)
    got = orc.match_features(fs, ft, True, mutual_ratio=0.0)
    assert np.array_equal(got, mutual)
    # fewer than ratio*ms mutual pairs -> fall back to the one-directional set (A.5)
    assert np.array_equal(orc.match_features(fs, ft, True, mutual_ratio=0.99), one)
    assert orc.match_features(fs[:0], ft, True).shape == (0, 2)


def test_transform_spec(orc):
    rng = np.random.default_rng(4)
    pts = cloud(1000, 8)
    T = np.eye(4)
    T[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    T[:3, 3] = rng.normal(size=3)
    assert np.array_equal(orc.transform_points(T, pts), np_ref.transform_f32(T, pts))
