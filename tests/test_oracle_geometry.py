"""Known-answer tests for the oracle's normals, FPFH, Kabsch, RANSAC and ICP restatements."""
import numpy as np
import pytest

from pcr_b200 import synth


def test_fast_eigen3x3_against_eigh(orc):
    rng = np.random.default_rng(0)
    for k in range(300):
        a = rng.normal(size=(3, 3)) * rng.uniform(0.01, 3)
        cov = a @ a.T
        if k % 5 == 0:  # near-planar neighbourhoods: one tiny eigenvalue
            u = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            cov = u @ np.diag([1.0, 0.5, 1e-9]) @ u.T
        c6 = [cov[0, 0], cov[0, 1], cov[0, 2], cov[1, 1], cov[1, 2], cov[2, 2]]
        v = orc.fast_eigen3x3(c6)
        w, vecs = np.linalg.eigh(cov)
        ref = vecs[:, 0]
        assert abs(abs(v @ ref) - 1.0) < 1e-6, (k, v, ref)
    assert np.array_equal(orc.fast_eigen3x3([1, 0, 0, 1, 0, 1]), [0, 0, 1])  # identity covariance -> +z (A.2)
    assert np.array_equal(orc.fast_eigen3x3([0.5, 0, 0, 2, 0, 1]), [1, 0, 0])
    assert np.array_equal(orc.fast_eigen3x3([0, 0, 0, 0, 0, 0]), [0, 0, 0])


def test_normals_plane_and_sphere(orc):
    rng = np.random.default_rng(1)
    n = 4000
    uv = rng.uniform(-1, 1, (n, 2))
    nrm = np.array([1.0, 2.0, -0.5]); nrm /= np.linalg.norm(nrm)
    e1 = np.cross(nrm, [0, 0, 1.0]); e1 /= np.linalg.norm(e1)
    e2 = np.cross(nrm, e1)
    plane = (uv[:, :1] * e1 + uv[:, 1:] * e2 + 0.3 * nrm).astype(np.float32)
    got = orc.estimate_normals(plane, 0.15, 30).astype(np.float64)
    assert np.all(np.abs(np.abs(got @ nrm) - 1.0) < 1e-4)
    u = rng.normal(size=(n * 4, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    sph = u.astype(np.float32)
    got = orc.estimate_normals(sph, 0.12, 30).astype(np.float64)
    assert np.median(np.abs(np.sum(got * u, axis=1))) > 0.999
    # fewer than 3 neighbours -> covariance = I -> (0,0,1) (A.2)
    lonely = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0]], np.float32)
    assert np.array_equal(orc.estimate_normals(lonely, 0.5, 30), [[0, 0, 1]] * 3)


def test_pair_features_and_fpfh_of_a_plane(orc):
    # two points on a plane with parallel normals: f1 = f2 = 0, f0 = atan2(0, 1) = 0
    f = orc.pair_features([0, 0, 0], [0, 0, 1], [1, 0, 0], [0, 0, 1])
    assert f[3] == 1.0 and abs(f[0]) < 1e-15 and abs(f[1]) < 1e-15 and abs(f[2]) < 1e-15
    assert np.array_equal(orc.pair_features([0, 0, 0], [0, 0, 1], [0, 0, 0], [0, 0, 1]), [0, 0, 0, 0])
    # a flat grid with exact +z normals: all mass falls into bin 5 of every 11-bin block
    g = np.stack(np.meshgrid(np.arange(20), np.arange(20)), -1).reshape(-1, 2) * 0.1
    pts = np.concatenate([g, np.zeros((len(g), 1))], 1).astype(np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (len(pts), 1))
    F = orc.fpfh(pts, nrm, 0.35, 100).astype(np.float64)
    for b in range(3):
        blk = F[:, 11 * b: 11 * b + 11]
        assert np.allclose(blk[:, 5], 200.0, atol=1e-3)       # 100 (weighted neighbours) + 100 (own SPFH)
        assert np.allclose(np.delete(blk, 5, axis=1), 0.0)
    # isolated points keep an all-zero descriptor (A.4)
    lonely = np.array([[0, 0, 0], [50, 0, 0]], np.float32)
    assert np.all(orc.fpfh(lonely, nrm[:2], 1.0, 100) == 0)


def test_fpfh_block_sums_and_rigid_invariance(orc):
    v = 0.05
    pts = synth.surface(3000, v, 77, spacing_ratio=1.0).astype(np.float32)
    nrm = orc.estimate_normals(pts, 2 * v, 30)
    F = orc.fpfh(pts, nrm, 5 * v, 100).astype(np.float64)
    sums = F.reshape(len(F), 3, 11).sum(2)
    ok = sums[:, 0] > 0
    assert ok.mean() > 0.99
    assert np.allclose(sums[ok], 200.0, atol=1e-2)
    # rigid motion of the cloud (with rotated normals) leaves the descriptors nearly unchanged
    R = synth.euler_zyx(0.3, -0.2, 0.9)
    p2 = (pts.astype(np.float64) @ R.T + [0.5, -0.25, 1.0]).astype(np.float32)
    n2 = (nrm.astype(np.float64) @ R.T).astype(np.float32)
    F2 = orc.fpfh(p2, n2, 5 * v, 100).astype(np.float64)
    assert np.median(np.abs(F - F2).sum(1)) < 8.0  # out of 600: only neighbour-set/bin-edge flips differ


def load_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ransac_numpy_golden.npz"))


def test_kabsch_and_inlier_ratio_against_reference_numpy(orc):
    """Pinned against the reference's own functions (tests/golden/make_ransac_numpy_golden.py)."""
    g = load_golden()
    for k in range(int(g["n_cases"])):
        src, tgt, corr, idx, T = g[f"src_{k}"], g[f"tgt_{k}"], g[f"corr_{k}"], g[f"idx_{k}"], g[f"T_{k}"]
        s3 = src[corr[idx, 0]].astype(np.float64)
        t3 = tgt[corr[idx, 1]].astype(np.float64)
        mine = orc.kabsch3(s3, t3)
        # rotation to 1e-5, translation to 1e-5 * extent (the north-star tolerance); typical agreement ~1e-12
        sv = np.linalg.svd((s3 - s3.mean(0)).T @ (t3 - t3.mean(0)), compute_uv=False)
        if sv[1] > 1e-3 * sv[0]:  # rank-2 cross-covariance: the optimal rotation is unique
            assert np.abs(mine[:3, :3] - T[:3, :3]).max() < 1e-8, k
            assert np.abs(mine[:3, 3] - T[:3, 3]).max() < 1e-7, k
        # rigid and proper in every case
        assert np.allclose(mine[:3, :3] @ mine[:3, :3].T, np.eye(3), atol=1e-9)
        assert abs(np.linalg.det(mine[:3, :3]) - 1) < 1e-9
        cost = lambda M: np.sum((s3 @ M[:3, :3].T + M[:3, 3] - t3) ** 2)  # noqa: E731
        assert cost(mine) <= cost(T) * (1 + 1e-9) + 1e-18, k  # equally optimal even when not unique
        voxel = float(g[f"voxel_{k}"])
        c = orc.inlier_count(src, tgt, corr, T, voxel * 1.5)
        assert c / len(corr) == float(g[f"ratio_{k}"]), k
        c2 = orc.inlier_count(src, tgt, corr, T, (voxel * 1.5) ** 2, squared=True)
        assert abs(c2 / len(corr) - float(g[f"ratio_fast_{k}"])) <= 1.0 / len(corr), k


def test_soft_failures_match_reference(orc):
    g = load_golden()
    col = np.array([[0, 0, i] for i in range(10)], np.float64)
    dup = np.array([[1, 1, 1]] * 10, np.float64)
    assert np.allclose(g["T_collinear"], np.eye(4)) and np.allclose(g["T_duplicate"], np.eye(4))
    assert np.array_equal(orc.kabsch3(col[[1, 4, 7]], col[[1, 4, 7]]), np.eye(4))
    assert np.array_equal(orc.kabsch3(dup[:3], dup[:3]), np.eye(4))
    ident = np.stack([np.arange(10), np.arange(10)], 1).astype(np.int32)
    T, _ = orc.ransac_step(col.astype(np.float32), col.astype(np.float32), ident[:2], 0, 0)
    assert np.array_equal(T, g["T_two_corr"]) and np.array_equal(T, np.eye(4))  # < 3 pairs -> identity
    assert float(g["ratio_empty"]) == 0.0 and float(g["ratio_fast_empty"]) == 0.0
    big = np.eye(4) * 1000.0; big[3, 3] = 1
    pts = g["huge_pts"]
    cc = np.stack([np.arange(50), np.arange(50)], 1).astype(np.int32)
    assert orc.inlier_count(pts, pts, cc, big, 0.075) / 50 == float(g["ratio_huge"])
    # 1000 random draws never produce NaN/Inf (test_ransac_crash.py:227-271)
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 1, (30, 3)).astype(np.float32)
    c = np.stack([rng.integers(0, 30, 90), rng.integers(0, 30, 90)], 1).astype(np.int32)
    for h in range(1000):
        T, smp = orc.ransac_step(p, p, c, 1, h)
        assert np.isfinite(T).all() and len(set(smp.tolist())) == 3


@pytest.fixture(scope="module")
def pair20k(orc):
    v = 0.005
    src, tgt, T = synth.make_pair(20000, v, 20241)
    return v, src, tgt, T, orc.preprocess(src, v), orc.preprocess(tgt, v)


def test_ransac_recovers_known_transform_and_is_thread_invariant(orc, pair20k):
    v, src, tgt, T, S, G = pair20k
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    assert len(corr) >= 0.1 * len(S.pcd_down)
    r = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * v, 100000, 0.999, seed=1)
    assert r.fitness > 0.9 and r.best_hyp >= 0 and r.hyp_evaluated <= 100000
    assert np.abs(r.transformation[:3, :3] - T[:3, :3]).max() < 0.05
    nt = orc.num_threads()
    orc.set_num_threads(1)
    r1 = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * v, 100000, 0.999, seed=1)
    orc.set_num_threads(nt)
    assert r1.best_hyp == r.best_hyp and r1.sum_d2_fixed == r.sum_d2_fixed and r1.hyp_evaluated == r.hyp_evaluated
    assert np.array_equal(r1.transformation, r.transformation)
    # confidence 1.0 consumes every iteration (A.6)
    rfull = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * v, 3000, 1.0, seed=1)
    assert rfull.hyp_evaluated == 3000 and rfull.est_k == 3000
    # degenerate inputs return the default result
    d = orc.ransac(S.pcd_down, G.pcd_down, corr[:2], 1.5 * v, 100, 0.999)
    assert d.fitness == 0 and d.best_hyp == -1 and np.array_equal(d.transformation, np.eye(4))


def test_icp_converges_to_truth(orc, pair20k):
    v, src, tgt, T, S, G = pair20k
    pert = np.eye(4)
    pert[:3, :3] = synth.euler_zyx(0.002, -0.001, 0.0015)
    pert[:3, 3] = [2e-4, -3e-4, 1e-4]
    r = orc.refine_registration(S, G, pert @ T, v)
    assert r.fitness > 0.95 and r.iterations >= 1
    err0 = np.abs((pert @ T) - T).max()
    assert np.abs(r.transformation - T).max() < 0.3 * err0
    # fixed work: relative criteria 0 -> exactly max_iter updates
    r50 = orc.icp_point_to_plane(src, tgt, G.normals, 0.4 * v, pert @ T, 7, 0.0, 0.0)
    assert r50.iterations == 7 and not r50.converged
    # max_iter 0 -> evaluation only
    r0 = orc.icp_point_to_plane(src, tgt, G.normals, 0.4 * v, T, 0)
    assert r0.iterations == 0 and np.array_equal(r0.transformation, T)
    idx, d2 = orc.nn1(tgt, orc.transform_points(T, src), 0.4 * v)
    assert np.array_equal(r0.correspondence, idx) and r0.inlier_count == int((idx >= 0).sum())
    with pytest.raises(ValueError):
        orc.icp_point_to_plane(src, tgt, G.normals, 0.0, T, 1)
    nt = orc.num_threads()
    orc.set_num_threads(1)
    r1 = orc.refine_registration(S, G, pert @ T, v)
    orc.set_num_threads(nt)
    assert np.array_equal(r1.transformation, r.transformation) and r1.sum_d2_fixed == r.sum_d2_fixed
