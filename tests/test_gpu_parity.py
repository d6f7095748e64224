"""GPU parity tests proper: the CUDA path, called through the C ABI (pcr_b200.engine -> libpcr_b200.so), against
the CPU oracle on the same seeded inputs.  Bit-exact for every integer output (voxel assignment, neighbour
indices, correspondences, inlier counts, winning hypothesis) and — because the arithmetic specification is
shared — for the floating-point outputs as well; the north-star tolerance (1e-5 rotation,
1e-5 * extent translation) is asserted separately where a looser bound is the contract."""
import numpy as np
import pytest
import torch

from pcr_b200 import synth

pytestmark = pytest.mark.gpu


def xyz(t):
    return t[:, :3].cpu().numpy()


@pytest.fixture(scope="module")
def pair(orc, eng):
    v = 0.005
    src, tgt, T = synth.make_pair(20000, v, 20241)
    d = {"v": v, "src": src, "tgt": tgt, "T": T, "ds": eng.pack(src), "dt": eng.pack(tgt)}
    d["osd"], d["otd"] = orc.voxel_downsample(src, v), orc.voxel_downsample(tgt, v)
    d["osn"], d["otn"] = orc.estimate_normals(d["osd"], 2 * v, 30), orc.estimate_normals(d["otd"], 2 * v, 30)
    d["osf"], d["otf"] = orc.fpfh(d["osd"], d["osn"], 5 * v, 100), orc.fpfh(d["otd"], d["otn"], 5 * v, 100)
    d["ocorr"] = orc.match_features(d["osf"], d["otf"], True)
    d["sd"], d["td"] = eng.voxel_downsample(d["ds"], v).contiguous(), eng.voxel_downsample(d["dt"], v).contiguous()
    d["sn"], d["tn"] = eng.estimate_normals(d["sd"], 2 * v, 30), eng.estimate_normals(d["td"], 2 * v, 30)
    d["sf"], d["tf"] = eng.compute_fpfh(d["sd"], d["sn"], 5 * v, 100), eng.compute_fpfh(d["td"], d["tn"], 5 * v, 100)
    d["corr"] = eng.match_features(d["sf"], d["tf"], True).contiguous()
    return d


def test_pack_quantises_fp64(eng):
    a = np.random.default_rng(0).normal(size=(1000, 3))
    t = eng.pack(a)
    assert np.array_equal(xyz(t), a.astype(np.float32)) and torch.all(t[:, 3] == 0)
    assert eng.pack(np.zeros((0, 3))).shape == (0, 4)


def test_voxel_downsample(pair):
    assert np.array_equal(xyz(pair["sd"]), pair["osd"]) and np.array_equal(xyz(pair["td"]), pair["otd"])


@pytest.mark.parametrize("voxel", [0.002, 0.01, 0.05, 1.0])
def test_voxel_sizes(orc, eng, pair, voxel):
    assert np.array_equal(xyz(eng.voxel_downsample(pair["ds"], voxel)), orc.voxel_downsample(pair["src"], voxel))


def test_voxel_edge_cases(orc, eng):
    with pytest.raises(ValueError):
        eng.voxel_downsample(eng.pack(np.zeros((5, 3))), 0.0)
    assert eng.voxel_downsample(eng.pack(np.zeros((0, 3))), 0.3).shape[0] == 0
    dup = np.ones((10, 3), np.float32)
    assert np.array_equal(xyz(eng.voxel_downsample(eng.pack(dup), 0.3)), orc.voxel_downsample(dup, 0.3))
    one = np.array([[1.5, -2.0, 3.25]], np.float32)
    assert np.array_equal(xyz(eng.voxel_downsample(eng.pack(one), 0.3)), one)
    # 10^18 voxel ids: far beyond the dense table, down-sampled through the sorted-key path like any other cloud
    far = np.array([[0, 0, 0], [1e3, 1e3, 1e3], [1e3, 1e3, 1e3]], np.float32)
    assert np.array_equal(xyz(eng.voxel_downsample(eng.pack(far), 1e-3)), orc.voxel_downsample(far, 1e-3))
    with pytest.raises(ValueError):  # a dimension beyond int32 (PCR_ERR_TOO_LARGE), as the oracle refuses it
        eng.voxel_downsample(eng.pack(far), 1e-7)


@pytest.mark.parametrize("voxel", [0.3, 0.011])
def test_voxel_grid_beyond_the_dense_budget(orc, eng, pair, voxel):
    """a 200-unit scene at the reference's default voxel 0.3, or a lidar sweep at 5 cm, has more voxel ids than the
    dense table may hold (2^27).  Two dense clusters 10^5 units apart and a sprinkle of isolated points: the sorted-key
    path must give the oracle's points bit for bit, in ascending voxel id — voxels with hundreds of points and with one."""
    rng = np.random.default_rng(5)
    big = voxel > 0.1
    a = pair["src"].astype(np.float32) * (60.0 if big else 1.0)
    b = a[: len(a) // 2] + np.array([1e5, -3e4, 2e4] if big else [2e3, -5e2, 3e2], np.float32)
    lone = rng.uniform(-5e4, 5e4, (3000, 3)).astype(np.float32) * (1.0 if big else 0.02)
    pts = np.concatenate([a, b, lone]).astype(np.float32)
    rng.shuffle(pts)
    ext = pts.max(0).astype(np.float64) - pts.min(0)
    assert 2.0 ** 27 < np.prod(np.floor(ext / voxel) + 1) < 9.0e18
    got = xyz(eng.voxel_downsample(eng.pack(pts), voxel))
    want = orc.voxel_downsample(pts, voxel)
    assert got.shape == want.shape and len(want) < len(pts) - 1000
    assert np.array_equal(got, want)
    # and the rest of Ply._preprocess runs on it (the search grids enlarge their cells to fit): normals + FPFH = the oracle's
    d = eng.voxel_downsample(eng.pack(pts), voxel)
    n = eng.estimate_normals(d, 2 * voxel, 30)
    assert np.array_equal(xyz(n), orc.estimate_normals(want, 2 * voxel, 30))


@pytest.mark.parametrize("radius_v,k", [(2, 30), (5, 100), (1.2, 4), (8, 256)])
def test_knn_hybrid(orc, eng, pair, radius_v, k):
    v = pair["v"]
    q = pair["ds"][:3000].contiguous()
    idx, d2, cnt = eng.knn_hybrid(pair["ds"], q, radius_v * v, k)
    oi, od, oc = orc.knn_hybrid(pair["src"], pair["src"][:3000], radius_v * v, k)
    assert np.array_equal(cnt.cpu().numpy(), oc)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)


def test_knn_ties_duplicates_and_outside_queries(orc, eng):
    rng = np.random.default_rng(3)
    lat = (rng.integers(0, 5, (800, 3)) * 0.25).astype(np.float32)  # many exact ties and duplicate points
    q = np.concatenate([lat[:100], rng.uniform(-3, 4, (100, 3)).astype(np.float32)])
    idx, d2, cnt = eng.knn_hybrid(eng.pack(lat), eng.pack(q), 0.6, 40)
    oi, od, oc = orc.knn_hybrid(lat, q, 0.6, 40)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od) and np.array_equal(cnt.cpu().numpy(), oc)
    # buffer overflow path: > 1024 in-radius candidates per query
    dense = rng.normal(0, 0.05, (5000, 3)).astype(np.float32)
    idx, d2, cnt = eng.knn_hybrid(eng.pack(dense), eng.pack(dense[:200]), 0.2, 50)
    oi, od, oc = orc.knn_hybrid(dense, dense[:200], 0.2, 50)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
    # empty index
    idx, d2, cnt = eng.knn_hybrid(eng.pack(np.zeros((0, 3))), eng.pack(q), 0.5, 5)
    assert torch.all(idx == -1) and torch.all(cnt == 0)


def test_nn1(orc, eng, pair):
    v = pair["v"]
    moved = orc.transform_points(pair["T"], pair["src"])
    idx, d2 = eng.nn1(pair["dt"], eng.pack(moved), 0.4 * v)
    oi, od = orc.nn1(pair["tgt"], moved, 0.4 * v)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
    tgt = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    q = np.array([[0.5, 0, 0], [3, 0, 0], [0.25, 0, 0]], np.float32)
    assert eng.nn1(eng.pack(tgt), eng.pack(q), 0.5)[0].tolist() == [-1, -1, 0]       # strict radius
    assert eng.nn1(eng.pack(tgt), eng.pack(q), 0.5000001)[0].tolist() == [0, -1, 0]  # tie -> lowest index
    with pytest.raises(ValueError):
        eng.nn1(eng.pack(tgt), eng.pack(q), 0.0)


def test_transform_points(orc, eng, pair):
    assert np.array_equal(xyz(eng.transform_points(pair["ds"], pair["T"])), orc.transform_points(pair["T"], pair["src"]))


def test_normals(orc, eng, pair):
    assert np.array_equal(xyz(pair["sn"]), pair["osn"]) and np.array_equal(xyz(pair["tn"]), pair["otn"])
    full = eng.estimate_normals(pair["dt"], 2 * pair["v"], 30)
    assert np.array_equal(xyz(full), orc.estimate_normals(pair["tgt"], 2 * pair["v"], 30))
    lonely = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0]], np.float32)
    assert np.array_equal(xyz(eng.estimate_normals(eng.pack(lonely), 0.5, 30)), [[0, 0, 1]] * 3)
    plane = np.concatenate([np.random.default_rng(0).uniform(-1, 1, (2000, 2)), np.zeros((2000, 1))], 1).astype(np.float32)
    assert np.array_equal(xyz(eng.estimate_normals(eng.pack(plane), 0.2, 30)), orc.estimate_normals(plane, 0.2, 30))


def test_fpfh(orc, eng, pair):
    assert np.array_equal(pair["sf"].cpu().numpy(), pair["osf"]) and np.array_equal(pair["tf"].cpu().numpy(), pair["otf"])
    lonely = np.array([[0, 0, 0], [50, 0, 0]], np.float32)
    nrm = np.array([[0, 0, 1], [0, 0, 1]], np.float32)
    assert torch.all(eng.compute_fpfh(eng.pack(lonely), eng.pack(nrm), 1.0, 100) == 0)


def test_feature_matching(orc, eng, pair):
    assert np.array_equal(pair["corr"].cpu().numpy(), pair["ocorr"])
    one = eng.match_features(pair["sf"], pair["tf"], False)
    assert np.array_equal(one.cpu().numpy(), orc.match_features(pair["osf"], pair["otf"], False))
    assert np.array_equal(eng.nn_features(pair["tf"], pair["sf"]).cpu().numpy(), orc.nn_features(pair["otf"], pair["osf"]))
    # fall-back when too few mutual pairs; exact ties and all-zero descriptors
    rng = np.random.default_rng(11)
    fs = rng.uniform(0, 200, (300, 33)).astype(np.float32)
    ft = rng.uniform(0, 200, (260, 33)).astype(np.float32)
    ft[10] = fs[3]; ft[20] = fs[3]; fs[50] = 0; ft[60] = 0; ft[70] = 0
    dfs, dft = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
    for mutual, ratio in ((False, 0.1), (True, 0.0), (True, 0.1), (True, 0.99)):
        assert np.array_equal(eng.match_features(dfs, dft, mutual, ratio).cpu().numpy(), orc.match_features(fs, ft, mutual, ratio))
    assert eng.match_features(dfs[:0].contiguous(), dft, True).shape[0] == 0


@pytest.mark.parametrize("conf,iters,seed", [(0.999, 100000, 1), (0.999, 100000, 2), (1.0, 6000, 3), (0.9, 50, 4)])
def test_ransac(orc, eng, pair, conf, iters, seed):
    v = pair["v"]
    r = eng.ransac(pair["sd"], pair["td"], pair["corr"], 1.5 * v, iters, conf, seed)
    o = orc.ransac(pair["osd"], pair["otd"], pair["ocorr"], 1.5 * v, iters, conf, seed)
    assert (r.best_hyp, r.inlier_count, r.sum_d2_fixed, r.est_k, r.hyp_evaluated) == \
           (o.best_hyp, o.inlier_count, o.sum_d2_fixed, o.est_k, o.hyp_evaluated)
    assert np.array_equal(r.transformation, o.transformation)
    assert r.fitness == o.fitness and r.inlier_rmse == o.inlier_rmse
    if conf == 1.0:
        assert r.survivors == o.survivors


def test_ransac_waves_match_single_call(orc, eng, pair):
    """pcr_ransac_wave + pcr_ransac_scan (the multi-GPU building blocks) on one GPU, emulating 3 ranks."""
    from pcr_b200.dist import ransac_distributed, records_to_array
    import ctypes as C
    v = pair["v"]
    k_d = int(eng.lib.pcr_ransac_k_d(C.c_double(1.5 * v), C.c_int(pair["sd"].shape[0])))

    def wave_fn(lo, hi, bc, bs):
        parts, ns = [], 0
        for r in range(3):  # three emulated ranks, concatenated in rank order
            a = lo + (hi - lo) * r // 3
            b = lo + (hi - lo) * (r + 1) // 3
            if b > a:
                recs, n, s = eng.ransac_wave(pair["sd"], pair["td"], pair["corr"], 1.5 * v, a, b, 5, 0.9, 4096, bc, bs)
                parts.append(records_to_array(recs, n)); ns += s
        return (np.concatenate(parts) if parts else np.zeros((0, 16), np.int64)), ns
    st, _ = ransac_distributed(wave_fn, pair["corr"].shape[0], pair["sd"].shape[0], k_d, 100000, 0.999, lib=eng.lib, first_wave=1000)
    o = orc.ransac(pair["osd"], pair["otd"], pair["ocorr"], 1.5 * v, 100000, 0.999, 5)
    assert (st.best_hyp, st.inlier_count, st.sum_d2_fixed, st.est_k, st.hyp_evaluated) == \
           (o.best_hyp, o.inlier_count, o.sum_d2_fixed, o.est_k, o.hyp_evaluated)
    assert np.array_equal(np.array(st.transformation).reshape(4, 4), o.transformation)


def test_ransac_session_and_multi_gpu_driver(orc, eng, pair):
    """ransac_multi_gpu (world 1: the same wave loop the ranks run, inside a pcr_ransac_session) equals the oracle's
    sequential loop; waves inside and outside a session return the same records; a wave on other buffers while a
    session is open prepares its own work."""
    from pcr_b200.dist import prefix_maxima, ransac_multi_gpu, records_to_array
    v = pair["v"]

    def chain(w):  # a wave returns a SUPERSET of its prefix maxima (which extras it holds depends on timing); the chain is exact
        return prefix_maxima(records_to_array(w[0], w[1]), 0, 0)
    for conf, iters, seed in ((0.999, 60000, 5), (1.0, 9000, 11)):
        r, stats = ransac_multi_gpu(eng, pair["sd"], pair["td"], pair["corr"], 1.5 * v, iters, conf, seed, first_wave=512)
        o = orc.ransac(pair["osd"], pair["otd"], pair["ocorr"], 1.5 * v, iters, conf, seed)
        assert (r.best_hyp, r.inlier_count, r.sum_d2_fixed, r.hyp_evaluated) == (o.best_hyp, o.inlier_count, o.sum_d2_fixed, o.hyp_evaluated)
        assert np.array_equal(r.transformation, o.transformation)
        assert stats["waves"] >= (3 if conf == 1.0 else 1)  # confidence 1.0 never exits early: 512, 1024, 2048, ...
    plain = eng.ransac_wave(pair["sd"], pair["td"], pair["corr"], 1.5 * v, 0, 3000, 5)
    eng.ransac_session_begin(pair["sd"], pair["td"], 1.5 * v)
    try:
        inside = eng.ransac_wave(pair["sd"], pair["td"], pair["corr"], 1.5 * v, 0, 3000, 5)
        other = eng.ransac_wave(pair["td"], pair["sd"], pair["corr"][:, [1, 0]].contiguous(), 1.5 * v, 0, 3000, 5)  # not the session's clouds
    finally:
        eng.ransac_session_end()
    assert plain[2] == inside[2] and np.array_equal(chain(plain), chain(inside))
    after = eng.ransac_wave(pair["td"], pair["sd"], pair["corr"][:, [1, 0]].contiguous(), 1.5 * v, 0, 3000, 5)
    assert other[2] == after[2] and np.array_equal(chain(other), chain(after)) and len(chain(plain)) >= 1


def test_ransac_degenerate(eng, pair):
    v = pair["v"]
    r = eng.ransac(pair["sd"], pair["td"], pair["corr"][:2].contiguous(), 1.5 * v, 100)
    assert r.best_hyp == -1 and r.fitness == 0 and np.array_equal(r.transformation, np.eye(4))
    r = eng.ransac(pair["sd"], pair["td"], pair["corr"], 0.0, 100)
    assert r.best_hyp == -1


def test_manual_step_twins(orc, eng, pair):
    v = pair["v"]
    Ts = eng.ransac_step(pair["sd"], pair["td"], pair["corr"], 9, 100, 512)
    cnt = eng.inlier_count(pair["sd"], pair["td"], pair["corr"], Ts, 1.5 * v).cpu().numpy()
    cnt2 = eng.inlier_count(pair["sd"], pair["td"], pair["corr"], Ts, (1.5 * v) ** 2, squared=True).cpu().numpy()
    Tn = Ts.cpu().numpy()
    for i in range(0, 512, 37):
        oT, _ = orc.ransac_step(pair["osd"], pair["otd"], pair["ocorr"], 9, 100 + i)
        assert np.array_equal(Tn[i], oT)
        assert cnt[i] == orc.inlier_count(pair["osd"], pair["otd"], pair["ocorr"], oT, 1.5 * v)
        assert cnt2[i] == orc.inlier_count(pair["osd"], pair["otd"], pair["ocorr"], oT, (1.5 * v) ** 2, squared=True)


def test_device_kabsch_against_reference_golden(eng):
    """Three correspondences -> the sample is all of them: compare with the reference's NumPy Kabsch output."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ransac_numpy_golden.npz"))
    checked = 0
    for k in range(int(g["n_cases"])):
        src, tgt, corr, idx, T = g[f"src_{k}"], g[f"tgt_{k}"], g[f"corr_{k}"], g[f"idx_{k}"], g[f"T_{k}"]
        c3 = np.ascontiguousarray(corr[idx])
        s3, t3 = src[c3[:, 0]].astype(np.float64), tgt[c3[:, 1]].astype(np.float64)
        sv = np.linalg.svd((s3 - s3.mean(0)).T @ (t3 - t3.mean(0)), compute_uv=False)
        if sv[1] <= 1e-3 * sv[0]:
            continue
        mine = eng.ransac_step(eng.pack(src), eng.pack(tgt), torch.from_numpy(c3).cuda(), 0, 0, 1)[0].cpu().numpy()
        assert np.abs(mine[:3, :3] - T[:3, :3]).max() < 1e-5          # north-star rotation tolerance
        assert np.abs(mine[:3, 3] - T[:3, 3]).max() < 1e-5 * 4.0      # 1e-5 * extent
        assert np.abs(mine - T).max() < 1e-7                           # what is actually achieved
        checked += 1
    assert checked >= 12


@pytest.mark.parametrize("iters,rf,rr", [(30, 1e-6, 1e-6), (12, 0.0, 0.0), (0, 1e-6, 1e-6)])
def test_icp(orc, eng, pair, iters, rf, rr):
    v = pair["v"]
    pert = np.eye(4); pert[:3, :3] = synth.euler_zyx(0.002, -0.001, 0.0015); pert[:3, 3] = [2e-4, -3e-4, 1e-4]
    init = pert @ pair["T"]
    otn = orc.estimate_normals(pair["tgt"], 2 * v, 30)
    g, corr = eng.icp_point_to_plane(pair["ds"], pair["dt"], eng.pack(otn), 0.4 * v, init, iters, rf, rr)
    o = orc.icp_point_to_plane(pair["src"], pair["tgt"], otn, 0.4 * v, init, iters, rf, rr)
    assert np.array_equal(corr.cpu().numpy(), o.correspondence)
    assert (g.inlier_count, g.sum_d2_fixed, g.iterations, g.converged) == (o.inlier_count, o.sum_d2_fixed, o.iterations, o.converged)
    assert np.array_equal(g.transformation, o.transformation)
    assert g.fitness == o.fitness and g.inlier_rmse == o.inlier_rmse


def test_icp_certified_passes_ties_and_drift(orc, eng, pair):
    """Many passes (the certified streaming path runs from pass 3 on), a start far enough that correspondences appear,
    change and drop out while the cloud moves, and a target with exact duplicates (ties decided by the lowest index):
    correspondences, sums and the transform must still equal a search in every pass (the oracle) bit for bit."""
    v = pair["v"]
    tgt = np.concatenate([pair["tgt"], pair["tgt"][:700], pair["tgt"][100:300]])
    otn = orc.estimate_normals(tgt, 2 * v, 30)
    pert = np.eye(4); pert[:3, :3] = synth.euler_zyx(0.012, -0.009, 0.007); pert[:3, 3] = [1.5e-3, -1.1e-3, 0.9e-3]
    init = pert @ pair["T"]
    for iters in (3, 41):
        g, corr = eng.icp_point_to_plane(pair["ds"], eng.pack(tgt), eng.pack(otn), 0.4 * v, init, iters, 0.0, 0.0)
        o = orc.icp_point_to_plane(pair["src"], tgt, otn, 0.4 * v, init, iters, 0.0, 0.0)
        assert np.array_equal(corr.cpu().numpy(), o.correspondence)
        assert (g.inlier_count, g.sum_d2_fixed, g.iterations) == (o.inlier_count, o.sum_d2_fixed, o.iterations)
        assert np.array_equal(g.transformation, o.transformation)
    assert 0.2 < g.fitness < 1.0


@pytest.mark.parametrize("cert_pass", ["0", "1", "5"])
def test_icp_certificate_start_pass_cannot_change_a_result(orc, eng, pair, cert_pass, monkeypatch):
    """The certificates (squared-form test, pair tier, "still none") only decide between reuse and search.  Starting them
    at pass 0 (weak certificates of a cloud that still moves: most points take the pair tier or search again), at pass 1
    or only at pass 5 must reproduce the oracle — which searches in every pass — bit for bit (the switch is read per call)."""
    monkeypatch.setenv("PCR_ICP_CERT_PASS", cert_pass)
    v = pair["v"]
    otn = orc.estimate_normals(pair["tgt"], 2 * v, 30)
    pert = np.eye(4); pert[:3, :3] = synth.euler_zyx(0.006, -0.004, 0.003); pert[:3, 3] = [6e-4, -5e-4, 4e-4]
    init = pert @ pair["T"]
    g, corr = eng.icp_point_to_plane(pair["ds"], pair["dt"], eng.pack(otn), 0.4 * v, init, 25, 0.0, 0.0)
    o = orc.icp_point_to_plane(pair["src"], pair["tgt"], otn, 0.4 * v, init, 25, 0.0, 0.0)
    assert np.array_equal(corr.cpu().numpy(), o.correspondence)
    assert (g.inlier_count, g.sum_d2_fixed, g.iterations) == (o.inlier_count, o.sum_d2_fixed, o.iterations)
    assert np.array_equal(g.transformation, o.transformation)


@pytest.mark.parametrize("ctas", ["8", "3", "1"])
def test_icp_many_chunks_per_cta(orc, eng, pair, ctas, monkeypatch):
    """A small grid (PCR_ICP_MAX_CTAS) gives a 20k-point cloud the many-chunks-per-CTA shape of a 1M-point one (double-buffered
    row staging, per-group barriers, a thread owning several points): every correspondence, sum and bit of the transform
    must equal the oracle's, with a partial last chunk too."""
    monkeypatch.setenv("PCR_ICP_MAX_CTAS", ctas)
    v = pair["v"]
    otn = orc.estimate_normals(pair["tgt"], 2 * v, 30)
    pert = np.eye(4); pert[:3, :3] = synth.euler_zyx(0.004, -0.003, 0.002); pert[:3, 3] = [4e-4, -3e-4, 2e-4]
    init = pert @ pair["T"]
    for cut in (0, 77):
        src = pair["src"][: len(pair["src"]) - cut]
        g, corr = eng.icp_point_to_plane(eng.pack(src), pair["dt"], eng.pack(otn), 0.4 * v, init, 20, 0.0, 0.0)
        o = orc.icp_point_to_plane(src, pair["tgt"], otn, 0.4 * v, init, 20, 0.0, 0.0)
        assert np.array_equal(corr.cpu().numpy(), o.correspondence)
        assert (g.inlier_count, g.sum_d2_fixed, g.iterations) == (o.inlier_count, o.sum_d2_fixed, o.iterations)
        assert np.array_equal(g.transformation, o.transformation)


@pytest.mark.parametrize("ns", [1, 2, 31, 129, 257, 1000])
def test_icp_small_and_ragged_sizes(orc, eng, pair, ns):
    """Source sizes around the kernel's granularities (128-row groups, 256-thread CTAs), down to a single point."""
    v = pair["v"]
    rng = np.random.default_rng(ns)
    sub = np.ascontiguousarray(pair["src"][rng.choice(len(pair["src"]), ns, replace=False)])
    otn = orc.estimate_normals(pair["tgt"], 2 * v, 30)
    pert = np.eye(4); pert[:3, :3] = synth.euler_zyx(0.003, 0.002, -0.001); pert[:3, 3] = [3e-4, 1e-4, -2e-4]
    init = pert @ pair["T"]
    g, corr = eng.icp_point_to_plane(eng.pack(sub), pair["dt"], eng.pack(otn), 0.4 * v, init, 12, 0.0, 0.0)
    o = orc.icp_point_to_plane(sub, pair["tgt"], otn, 0.4 * v, init, 12, 0.0, 0.0)
    assert np.array_equal(corr.cpu().numpy(), o.correspondence)
    assert (g.inlier_count, g.sum_d2_fixed, g.iterations) == (o.inlier_count, o.sum_d2_fixed, o.iterations)
    assert np.array_equal(g.transformation, o.transformation)


@pytest.mark.parametrize("ms", [3, 40, 300, 513])
def test_ransac_small_and_ragged_source(orc, eng, pair, ms):
    """RANSAC validation with source clouds around the chunk sizes of the validation kernel (256 / 512 / 1024)."""
    v = pair["v"]
    keep = np.sort(np.random.default_rng(ms).choice(len(pair["osd"]), ms, replace=False))
    remap = -np.ones(len(pair["osd"]), np.int64); remap[keep] = np.arange(ms)
    corr = pair["ocorr"][np.isin(pair["ocorr"][:, 0], keep)].copy()
    corr[:, 0] = remap[corr[:, 0]]
    if len(corr) < 3:
        pytest.skip("too few correspondences survive the subsampling")
    sd = np.ascontiguousarray(pair["osd"][keep])
    r = eng.ransac(eng.pack(sd), pair["td"], torch.as_tensor(corr.astype(np.int32)).to(eng.tdev).contiguous(), 1.5 * v, 20000, 1.0, 9)
    o = orc.ransac(sd, pair["otd"], corr.astype(np.int32), 1.5 * v, 20000, 1.0, 9)
    assert (r.best_hyp, r.inlier_count, r.sum_d2_fixed, r.survivors) == (o.best_hyp, o.inlier_count, o.sum_d2_fixed, o.survivors)
    assert np.array_equal(r.transformation, o.transformation)


def test_icp_edge_cases(eng, pair):
    v = pair["v"]
    n = eng.pack(np.tile([[0, 0, 1.0]], (pair["dt"].shape[0], 1)))
    with pytest.raises(ValueError):
        eng.icp_point_to_plane(pair["ds"], pair["dt"], n, 0.0)
    far = np.eye(4); far[:3, 3] = 100.0  # no correspondences at all: identity updates, fitness 0
    g, corr = eng.icp_point_to_plane(pair["ds"], pair["dt"], n, 0.4 * v, far, 5)
    assert g.fitness == 0 and g.inlier_count == 0 and torch.all(corr == -1) and np.array_equal(g.transformation, far)
    empty = eng.pack(np.zeros((0, 3)))
    g, corr = eng.icp_point_to_plane(empty, pair["dt"], n, 0.4 * v)
    assert g.fitness == 0 and corr.shape[0] == 0


def test_align_end_to_end(orc, eng, pair):
    from pcr_b200 import align
    v = pair["v"]
    T, fit, rmse, info = align(pair["src"], pair["tgt"], v, ransac_iteration=100000, seed=1, return_info=True)
    S, G = orc.preprocess(pair["src"], v), orc.preprocess(pair["tgt"], v)
    ro = orc.global_registration(S, G, v, 100000, 0.999, 1)
    io = orc.refine_registration(S, G, ro.transformation, v)
    assert np.array_equal(T, io.transformation) and fit == io.fitness and rmse == io.inlier_rmse
    assert info.ransac.best_hyp == ro.best_hyp and info.n_corr == len(pair["ocorr"])
    # recovers the known SE(3): 1e-5-level agreement is with the oracle; vs ground truth it is the sampling noise
    assert np.abs(T[:3, :3] - pair["T"][:3, :3]).max() < 2e-3 and fit > 0.95
    # device-resident inputs give the same bits; the call is deterministic run to run
    T2, fit2, rmse2 = align(pair["ds"], pair["dt"], v, ransac_iteration=100000, seed=1)
    assert np.array_equal(T, T2) and fit == fit2 and rmse == rmse2
    with pytest.raises(ValueError):
        align(np.zeros((0, 3)), pair["tgt"], v)


def test_tensor_core_matching_is_exact(orc, eng):
    """Sizes that take the tcgen05 path (pcr_match_tc.cu): the certificate + fallback must reproduce the exact
    fp64 arg-min bit for bit, including exact ties, duplicated rows and all-zero descriptors."""
    rng = np.random.default_rng(21)
    for nq, nb in ((700, 1500), (1300, 2600)):
        fs = rng.uniform(0, 200, (nq, 33)).astype(np.float32)
        ft = rng.uniform(0, 200, (nb, 33)).astype(np.float32)
        ft[100:140] = ft[50]            # 41 identical base rows: top-4 candidates cannot certify -> fallback
        fs[7] = ft[50]                  # exact tie over those rows -> lowest index 50
        fs[300:320] = 0.0
        ft[600:700] = 0.0               # all-zero descriptors tie exactly
        ft[900] = fs[11]
        ft[901] = fs[11] + np.float32(1e-3)   # near-tie well inside the bf16 error bound
        # smooth, highly correlated descriptors (what real FPFH looks like): small nearest-neighbour gaps
        base = rng.uniform(0, 200, 33).astype(np.float32)
        fs[400:600] = base + rng.normal(0, 0.05, (200, 33)).astype(np.float32)
        ft[1000:1400] = base + rng.normal(0, 0.05, (400, 33)).astype(np.float32)
        dfs, dft = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
        nn = eng.nn_features(dfs, dft).cpu().numpy()
        ref = orc.nn_features(fs, ft)
        assert np.array_equal(nn, ref)
        assert nn[7] == 50 and nn[300] == 600
        assert np.array_equal(eng.match_features(dfs, dft, True, 0.0).cpu().numpy(), orc.match_features(fs, ft, True, 0.0))


def test_tensor_core_threshold_under_attack(orc, eng):
    """Adversarial inputs for the error bound eps of the tensor-core filter (pcr_match_tc.cu header):
    large-norm descriptors (the split-bf16 and fp32-accumulation errors scale with ||a|| ||b||) whose true nearest
    neighbours differ from dozens of decoys only in the last bits, at several magnitudes, plus rows where the decoys
    outnumber the list capacity (exact fallback).  If eps were too small the true minimiser would miss the candidate
    list and the result would differ from the exact fp64 arg-min."""
    rng = np.random.default_rng(5)
    for scale in (1.0, 200.0, 3.0e4, 1.0e7):
        nq, nb = 640, 2304
        ft = (rng.uniform(0.25, 1.0, (nb, 33)) * scale).astype(np.float32)
        fs = (rng.uniform(0.25, 1.0, (nq, 33)) * scale).astype(np.float32)
        # rows 0..199: the query equals a base row up to 1-2 ulp in a few coordinates; 8 decoys of that base row differ
        # from it by 1 ulp in ONE coordinate each (so exact distances differ in their last digits)
        for i in range(200):
            j = 10 * i
            ft[j + 1:j + 9] = ft[j]
            for d in range(8):
                ft[j + 1 + d, d] = np.nextafter(ft[j, d], np.float32(np.inf if d % 2 else -np.inf))
            fs[i] = ft[j]
            fs[i, 20] = np.nextafter(np.nextafter(fs[i, 20], np.float32(np.inf)), np.float32(np.inf))
        # rows 200..219: 40 last-bit decoys each — more than the pass-2 list holds -> exact scan
        for i in range(200, 220):
            j = 2010 + 13 * (i - 200) // 1
            if j + 41 > nb:
                break
            ft[j:j + 40] = ft[j]
            for d in range(40):
                ft[j + d, d % 33] = np.nextafter(ft[j, d % 33], np.float32(np.inf if d % 3 else -np.inf))
            fs[i] = ft[j + 17]
        dfs, dft = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
        assert np.array_equal(eng.nn_features(dfs, dft).cpu().numpy(), orc.nn_features(fs, ft)), scale
        assert np.array_equal(eng.nn_features(dft, dfs).cpu().numpy(), orc.nn_features(ft, fs)), scale


def _degenerate_cases():
    rng = np.random.default_rng(0)
    v = 0.005
    src, tgt, _ = synth.make_pair(3000, v, 9)
    line = lambda n, off: (np.stack([np.linspace(0, 1, n)] * 3, 1) + off).astype(np.float32)  # noqa: E731
    plane = lambda n: np.concatenate([rng.random((n, 2)), np.zeros((n, 1))], 1).astype(np.float32)  # noqa: E731
    return [
        ("three_points", rng.random((3, 3)).astype(np.float32), rng.random((3, 3)).astype(np.float32), 0.1),
        ("single_point", np.ones((1, 3), np.float32), 2 * np.ones((1, 3), np.float32), 0.1),
        ("all_identical", np.ones((500, 3), np.float32), np.ones((400, 3), np.float32), 0.05),
        ("one_voxel", src, tgt, 10.0),                      # both clouds collapse to one point: no correspondences
        ("disjoint", src, (tgt + 100.0).astype(np.float32), v),  # RANSAC finds a transform, ICP has no pairs
        ("collinear", line(800, 0.0), line(700, 0.001), 0.01),  # singular normal equations
        ("coplanar", plane(2000), plane(2100), 0.02),
        ("tiny_vs_big", src[:7], tgt, v),
    ]


@pytest.mark.parametrize("case", range(8))
def test_align_degenerate_inputs_follow_the_oracle(orc, eng, case):
    """Whole-pipeline soft failures (SURVEY §8b error conventions; the shapes of test_ransac_crash.py:27-79 pushed through
    align()): nothing raises, fewer than three correspondences give the identity with fitness 0 from RANSAC, singular
    or empty ICP systems leave the transform finite — and every number equals the oracle's."""
    from pcr_b200 import align
    name, s, t, v = _degenerate_cases()[case]
    T, fit, rmse, info = align(s, t, v, ransac_iteration=2000, seed=0, return_info=True)
    S, G = orc.preprocess(s, v), orc.preprocess(t, v)
    ro = orc.global_registration(S, G, v, 2000, 0.999, 0)
    io = orc.refine_registration(S, G, ro.transformation, v)
    assert np.isfinite(T).all(), name
    assert (info.n_src_down, info.n_tgt_down) == (len(S.pcd_down), len(G.pcd_down)), name
    assert info.ransac.best_hyp == ro.best_hyp and info.ransac.fitness == ro.fitness, name
    assert np.array_equal(np.array(info.ransac.transformation).reshape(4, 4), ro.transformation), name
    assert np.array_equal(T, io.transformation) and fit == io.fitness and rmse == io.inlier_rmse, name
    assert info.icp.iterations == io.iterations, name


def test_pipeline_matches_the_committed_golden(eng):
    """The CUDA path against a COMMITTED fixture (tests/golden/pipeline_golden.npz: the oracle's outputs for the pair that
    __graft_entry__.smoke() aligns), not only against a live oracle run."""
    import os
    from pcr_b200 import align
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_golden.npz"))
    v = 0.005
    src, tgt, T_true = synth.make_pair(8000, v, 123)
    assert np.array_equal(T_true, g["T_true"])
    T, fit, rmse, info = align(src, tgt, v, ransac_iteration=20000, seed=5, return_info=True)
    assert np.array_equal(T, g["icp_T"]) and fit == float(g["icp_fitness"]) and rmse == float(g["icp_inlier_rmse"])
    assert info.ransac.best_hyp == int(g["ransac_best_hyp"]) and info.icp.inlier_count == int(g["icp_inlier_count"])
    assert np.array_equal(np.array(info.ransac.transformation).reshape(4, 4), g["ransac_T"])
    assert (info.n_src_down, info.n_tgt_down, info.n_corr) == (len(g["src_down"]), len(g["tgt_down"]), len(g["corr"]))
    assert info.icp.iterations == int(g["icp_iterations"])


def test_compact_grid_equals_dense(orc, eng, pair, monkeypatch):
    """The two-level (compact) search grid that ICP and nn1 use on fine grids returns what the dense table returns:
    radius-limited nearest neighbours (index and fp32 d2 bits, also against the oracle) and a whole ICP run."""
    v = pair["v"]
    moved = orc.transform_points(pair["T"], pair["src"])
    dm = eng.pack(moved)
    n = eng.estimate_normals(pair["dt"], 2 * v, 30)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PCR_GRID_COMPACT", mode)
        idx, d2 = eng.nn1(pair["dt"], dm, 0.4 * v)          # (extent / 0.4 v)^3 cells: far above the compact threshold
        g, corr = eng.icp_point_to_plane(pair["ds"], pair["dt"], n, 0.4 * v, pair["T"], 8, 0.0, 0.0)
        out[mode] = (idx.cpu().numpy(), d2.cpu().numpy(), corr.cpu().numpy(), np.asarray(g.transformation), g.sum_d2_fixed)
    for a, b in zip(out["1"], out["0"]):
        assert np.array_equal(a, b)
    oi, od = orc.nn1(pair["tgt"], moved, 0.4 * v)
    assert np.array_equal(out["1"][0], oi) and np.array_equal(out["1"][1], od)
    # a cloud with many points per cell and whole empty blocks: two clusters far apart, fine radius
    rng = np.random.default_rng(5)
    tgt = np.concatenate([rng.normal(0, 0.01, (3000, 3)), rng.normal(0, 0.01, (3000, 3)) + 4.0]).astype(np.float32)
    q = np.concatenate([tgt[::3] + rng.normal(0, 0.002, (2000, 3)).astype(np.float32), rng.uniform(-1, 5, (500, 3)).astype(np.float32)])
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PCR_GRID_COMPACT", mode)
        idx, d2 = eng.nn1(eng.pack(tgt), eng.pack(q), 0.01)
        res[mode] = (idx.cpu().numpy(), d2.cpu().numpy())
    oi, od = orc.nn1(tgt, q, 0.01)
    assert np.array_equal(res["1"][0], res["0"][0]) and np.array_equal(res["1"][1], res["0"][1])
    assert np.array_equal(res["1"][0], oi) and np.array_equal(res["1"][1], od)


def test_cluster_grid_build_equals_multi_kernel_build(eng, pair, monkeypatch):
    """Small clouds build their search grids in one thread-block-cluster launch; the neighbour lists, normals and a whole
    RANSAC run must not depend on which build produced the grid."""
    v = pair["v"]
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PCR_GRID_CLUSTER", mode)
        idx, d2, cnt = eng.knn_hybrid(pair["sd"], pair["sd"], 5 * v, 100)
        nrm = eng.estimate_normals(pair["td"], 2 * v, 30)
        r = eng.ransac(pair["sd"], pair["td"], pair["corr"], 1.5 * v, 20000, 1.0, 7)
        out[mode] = (idx.cpu().numpy(), d2.cpu().numpy(), cnt.cpu().numpy(), nrm.cpu().numpy(), np.asarray(r.transformation),
                     np.array([r.best_hyp, r.inlier_count, r.sum_d2_fixed, r.survivors]))
    for a, b in zip(out["1"], out["0"]):
        assert np.array_equal(a, b)
