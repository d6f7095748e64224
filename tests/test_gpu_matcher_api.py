"""The reference-facing Python surface (matcher.ransac / matcher.icp / ply.Ply) on the GPU: same names, positional
order and soft-failure behaviour as src/matcher/*.py and src/ply/ply.py — the cases of the reference's own
test_ransac_crash.py, with assertions."""
import numpy as np
import pytest

from pcr_b200 import synth

pytestmark = pytest.mark.gpu


class Cloud:
    def __init__(self, pts):
        self.points = np.asarray(pts, np.float64)


class MockPly:  # test_ransac_crash.py:92-96
    def __init__(self, pts):
        self.pcd = Cloud(pts)
        self.pcd_down = Cloud(pts)
        self.pcd_fpfh = None


def test_minimal_and_degenerate_correspondences(eng):
    from matcher.ransac import compute_step_transformation, evaluate_inlier_ratio, evaluate_inlier_ratio_fast
    rng = np.random.default_rng(0)
    a, b = MockPly(rng.random((3, 3))), MockPly(rng.random((3, 3)))
    r = compute_step_transformation(a, b, np.array([[0, 0], [1, 1], [2, 2]]))
    assert r and r.fitness == 0.0 and np.isfinite(r.transformation).all()
    col = np.array([[0, 0, i] for i in range(10)], float)
    ident = np.stack([np.arange(10), np.arange(10)], 1)
    assert np.array_equal(compute_step_transformation(MockPly(col), MockPly(col), ident).transformation, np.eye(4))
    dup = np.array([[1, 1, 1]] * 10, float)
    assert np.array_equal(compute_step_transformation(MockPly(dup), MockPly(dup), ident).transformation, np.eye(4))
    cop = rng.random((10, 3)); cop[:, 2] = 0
    assert np.isfinite(compute_step_transformation(MockPly(cop), MockPly(cop), ident).transformation).all()
    r = compute_step_transformation(MockPly(col), MockPly(col), ident[:2])        # < 3 pairs (ransac.py:138-140)
    assert np.array_equal(r.transformation, np.eye(4)) and r.fitness == 0.0
    assert evaluate_inlier_ratio(a, b, np.zeros((0, 2), np.int32), np.eye(4), 0.05) == 0.0   # ransac.py:220-221
    assert evaluate_inlier_ratio_fast(np.zeros((0, 3)), np.zeros((0, 3)), np.eye(4), 0.01) == 0.0
    big = np.eye(4) * 1000.0; big[3, 3] = 1
    pts = rng.random((50, 3))
    cc = np.stack([np.arange(50), np.arange(50)], 1)
    assert evaluate_inlier_ratio(MockPly(pts), MockPly(pts), cc, big, 0.05) < 0.1        # huge transform scores ~0
    assert evaluate_inlier_ratio(MockPly(pts), MockPly(pts), cc, np.eye(4), 0.05) == 1.0
    assert evaluate_inlier_ratio_fast(pts, pts, np.eye(4), 1e-6) == 1.0
    for _ in range(200):  # stability: never NaN/Inf (test_ransac_crash.py:227-271)
        assert np.isfinite(compute_step_transformation(MockPly(pts), MockPly(pts), cc).transformation).all()


def test_ply_pipeline_and_main_style_calls(tmp_path, orc, eng):
    from matcher.icp import refine_registration
    from matcher.ransac import (compute_feature_correspondences, compute_step_transformations, evaluate_inlier_ratios,
                                global_registration)
    from pcr_b200.plyio import write_ply
    from ply import Ply
    v = 0.005
    src, tgt, T = synth.make_pair(20000, v, 31)
    write_ply(tmp_path / "sample.ply", src, binary=False)   # the reference's converter emits ASCII PLY
    write_ply(tmp_path / "target.ply", tgt, binary=True)
    with pytest.raises(FileNotFoundError):
        Ply(tmp_path / "nope.ply")
    (tmp_path / "x.txt").write_text("x")
    with pytest.raises(TypeError):
        Ply(tmp_path / "x.txt")
    write_ply(tmp_path / "empty.ply", src[:0])
    with pytest.raises(ValueError):
        Ply(tmp_path / "empty.ply")
    s = Ply(tmp_path / "sample.ply", v, noise_sigma=0.0)
    t = Ply(tmp_path / "target.ply", v, noise_sigma=0.0)
    assert s.voxel_size == v and len(s.pcd.points) == 20000 and s.pcd_fpfh.data.shape[0] == 33
    So, To = orc.preprocess(src, v), orc.preprocess(tgt, v)
    assert np.array_equal(s.pcd_down.points.astype(np.float32), So.pcd_down)
    assert np.array_equal(s.pcd_fpfh.data.T.astype(np.float32), So.pcd_fpfh)
    # global_registration(src, tgt, voxel_size, iteration) — reference positional order (ransac.py:20-25)
    res = global_registration(s, t, v, 100000)
    want = orc.global_registration(So, To, v, 100000, 0.999, 0)
    assert res and np.array_equal(res.transformation, want.transformation)
    assert res.fitness == want.fitness and res.inlier_rmse == want.inlier_rmse
    assert len(res.correspondence_set) == want.inlier_count
    # src/main.py:34,38 call without voxel_size -> falls back to the Ply's own
    assert np.array_equal(global_registration(s, t, iteration=100000).transformation, res.transformation)
    icp = refine_registration(s, t, res.transformation, v)
    wicp = orc.refine_registration(So, To, want.transformation, v)
    assert np.array_equal(icp.transformation, wicp.transformation) and icp.fitness == wicp.fitness
    assert np.array_equal(icp.correspondence_set[:, 1], wicp.correspondence[wicp.correspondence >= 0])
    assert np.array_equal(refine_registration(s, t, res).transformation, icp.transformation)  # main.py:38 arity
    icp50 = refine_registration(s, t, res.transformation, v, max_iteration=50, relative_fitness=0.0, relative_rmse=0.0)
    assert icp50.info["iterations"] == 50
    # correspondences with injected noise: (1 + ratio) * C rows (ransac.py:89-99)
    c0 = compute_feature_correspondences(s, t)
    assert np.array_equal(c0, orc.match_features(So.pcd_fpfh, To.pcd_fpfh, False))
    for ratio in (1, 5):
        c = compute_feature_correspondences(s, t, noise_ratio=ratio, seed=1)
        assert len(c) == (1 + ratio) * len(c0) and c.dtype == np.int32
    Ts = compute_step_transformations(s, t, c0, 256, seed=2)
    ratios = evaluate_inlier_ratios(s, t, c0, Ts, v)
    assert ratios.shape == (256,) and ratios.max() <= 1.0
    # the noisy Ply default (sigma 0.05, ply.py:61-62) only moves pcd_down, not the descriptors
    sn = Ply(tmp_path / "sample.ply", v, seed=3)
    assert np.array_equal(sn.pcd_fpfh.data, s.pcd_fpfh.data) and not np.array_equal(sn.pcd_down.points, s.pcd_down.points)
    # point-to-plane without target normals is an error, as in Open3D
    with pytest.raises(RuntimeError):
        refine_registration(MockPly(src), MockPly(tgt), np.eye(4), v)


def test_full_size_properties_100k(orc, eng):
    """BASELINE config 2 sizes: size-independent properties + oracle parity on the cheap stages."""
    from pcr_b200 import align
    v = 0.005
    src, tgt, T = synth.make_pair(100000, v, 20242)
    Tg, fit, rmse, info = align(src, tgt, v, ransac_iteration=100000, icp_max_iteration=50, seed=7, return_info=True)
    assert fit > 0.95 and np.abs(Tg[:3, :3] - T[:3, :3]).max() < 1e-3
    R = Tg[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-9) and abs(np.linalg.det(R) - 1) < 1e-9
    # rigid-motion equivariance: moving the source by M changes the answer to T M^-1 (up to RANSAC/ICP noise)
    M = np.eye(4); M[:3, :3] = synth.euler_zyx(0.4, 0.1, -0.3); M[:3, 3] = [0.01, -0.02, 0.03]
    src2 = (src.astype(np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    T2, fit2, _ = align(src2, tgt, v, ransac_iteration=100000, icp_max_iteration=50, seed=7)
    assert fit2 > 0.95 and np.abs(T2 @ M - Tg).max() < 1e-3
    # oracle parity at full size for the bandwidth-bound stages
    ds, dt = eng.pack(src), eng.pack(tgt)
    assert np.array_equal(eng.voxel_downsample(ds, v)[:, :3].cpu().numpy(), orc.voxel_downsample(src, v))
    otn = orc.estimate_normals(tgt, 2 * v, 30)
    assert np.array_equal(eng.estimate_normals(dt, 2 * v, 30)[:, :3].cpu().numpy(), otn)
    g, corr = eng.icp_point_to_plane(ds, dt, eng.pack(otn), 0.4 * v, Tg, 10, 0.0, 0.0)
    o = orc.icp_point_to_plane(src, tgt, otn, 0.4 * v, Tg, 10, 0.0, 0.0)
    assert np.array_equal(g.transformation, o.transformation) and np.array_equal(corr.cpu().numpy(), o.correspondence)


def test_icp_1m_properties(eng):
    """BASELINE config 3 size (1M points): every reported correspondence is a true radius-limited nearest neighbour."""
    import torch
    v = 0.005
    src, tgt, T = synth.make_icp_pair(1000000, v, 20243)
    ds, dt = eng.pack(src), eng.pack(tgt)
    n = eng.estimate_normals(dt, 2 * v, 30)
    g, corr = eng.icp_point_to_plane(ds, dt, n, 0.4 * v, np.eye(4), 50, 0.0, 0.0)
    assert g.iterations == 50 and g.fitness > 0.9
    assert g.inlier_count == int((corr >= 0).sum())
    moved = eng.transform_points(ds, g.transformation)
    idx, d2 = eng.nn1(dt, moved, 0.4 * v)
    assert torch.equal(idx, corr)
    sel = torch.nonzero(corr >= 0)[:200000, 0]
    d = (moved[sel, :3] - dt[corr[sel].long(), :3]).double().norm(dim=1)
    assert float(d.max()) < 0.4 * v * (1 + 1e-6)
    assert abs(g.inlier_rmse - float(torch.sqrt((d2[corr >= 0].double()).mean()))) < 1e-9
    assert np.abs(g.transformation - T).max() < 5e-5


def test_benchmark_cli_report_and_export(tmp_path, eng):
    """The benchmark_ransac.py counterpart: same phases and report table as the reference's harness
    (benchmark_ransac.py:223-280, src/utils/profiler.py:151-215), plus the headless export."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("b200_benchmark_ransac", os.path.join(root, "3d-matching_b200", "benchmark_ransac.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from pcr_b200.plyio import read_ply, write_ply
    v = 0.005
    src, tgt, T_true = synth.make_pair(8000, v, 4242)
    write_ply(tmp_path / "sample.ply", src)
    write_ply(tmp_path / "target.ply", tgt)
    prof, res = mod.run_comprehensive_benchmark(tmp_path / "sample.ply", tmp_path / "target.ply", v, 0.5, 20, 20000, icp=True,
                                                export_dir=tmp_path / "out", report_path=tmp_path / "benchmark_results.txt",
                                                quiet=True)
    report = (tmp_path / "benchmark_results.txt").read_text()
    assert "PROFILING REPORT" in report and "Median (ms)" in report and "TOTAL" in report
    for name in ("ply_loading", "correspondence_computation", "ransac_iteration", "compute_transformation",
                 "evaluate_inliers", "full_ransac", "icp_refinement"):
        assert name in report
    st = prof.stats()
    assert len(st["ransac_iteration"]) == 20 and len(st["full_ransac"]) == 1
    out = json.loads((tmp_path / "out" / "registration.json").read_text())
    T = np.array(out["icp"]["transformation"])
    assert np.abs(T[:3, :3] - T_true[:3, :3]).max() < 5e-3 and out["icp"]["fitness"] > 0.9
    aligned, _ = read_ply(tmp_path / "out" / "source_aligned.ply")
    assert aligned.shape == src.shape
    # the exported cloud is the source moved by the exported transform
    assert np.abs(aligned - (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3])).max() < 1e-5


def test_align_batch_workers_match_sequential(eng):
    """Batch of independent pairs: concurrent worker threads (own context + stream each) give the same
    table as the sequential loop, and each row equals a single align of that pair."""
    from pcr_b200.dist import align_batch
    v = 0.005
    clouds = [synth.make_pair(6000, v, 500 + i) for i in range(5)]
    pairs = [(eng.pack(s), eng.pack(t)) for s, t, _ in clouds]
    p = eng.default_params(v)
    p.ransac_max_iter = 20000
    p.seed = 3
    seq = align_batch(eng, pairs, p, workers=1)
    par = align_batch(eng, pairs, p, workers=3)
    assert seq.shape == (5, 18) and np.array_equal(seq, par)
    one = eng.align_device(pairs[2][0], pairs[2][1], p)
    assert np.array_equal(seq[2, :16], np.array(one.icp.transformation)) and seq[2, 16] == one.icp.fitness
    for i, (_, _, T) in enumerate(clouds):
        assert np.abs(seq[i, :16].reshape(4, 4)[:3, :3] - T[:3, :3]).max() < 5e-3 and seq[i, 16] > 0.9


def test_run_ransac_manual_matches_the_gui_loop(eng):
    """run_ransac_manual = the reference GUI worker's loop (_visualize_matcher.py:343-470): strictly-greater best,
    early-stop formula, update callbacks — checked against a plain NumPy replay over the same hypothesis stream."""
    from matcher.ransac import (compute_feature_correspondences, compute_step_transformations, required_iterations,
                                run_ransac_manual)
    from ply import Ply
    v = 0.005
    s, t, _ = synth.make_pair(12000, v, 909)
    src, tgt = Ply.from_points(s, v), Ply.from_points(t, v)
    corres = compute_feature_correspondences(src, tgt, noise_ratio=2.0, seed=4)
    assert len(corres) == 3 * (len(corres) // 3) or len(corres) > 0
    ps, pt = src.pcd_down.points[corres[:, 0]], tgt.pcd_down.points[corres[:, 1]]
    thr2 = (1.5 * v) ** 2

    def replay(max_iter, early, interval):
        Ts = compute_step_transformations(src, tgt, corres, max_iter, seed=4, start=0).cpu().numpy()
        best, bw, it, events = None, -1.0, 0, []
        while it < max_iter:
            T = Ts[it]
            it += 1
            d2 = (((ps @ T[:3, :3].T + T[:3, 3]) - pt) ** 2).sum(1)
            w = float((d2 < thr2).sum()) / len(corres)
            new = best is None or w > bw
            if new:
                best, bw = T, w
            if early and bw > 0.5 and it >= required_iterations(bw, 0.99, 3, max_iter):
                events.append(("stop", it))
                return best, bw, it, True, events
            if it % interval == 0 or new:
                events.append(("upd", it))
        return best, bw, it, False, events

    for max_iter, early, interval in ((300, True, 10), (300, False, 7), (5000, True, 10)):
        seen = []
        r = run_ransac_manual(src, tgt, v, max_iter, correspondences=corres, early_stop_enabled=early, update_interval=interval,
                              callback=lambda res, it, w, bw: seen.append(it), seed=4, batch=128)
        T, bw, it, stopped, events = replay(max_iter, early, interval)
        # the device counts inliers with the same strict squared threshold; ratios are integer counts / C, so equal
        assert r.info["iterations"] == it and r.info["stopped_early"] == stopped
        assert abs(r.fitness - bw) < 1e-12 and np.allclose(r.transformation, T, atol=1e-12)
        assert seen == [e[1] for e in events]
    # cooperative stop: polled before every iteration
    calls = {"n": 0}

    def stop():
        calls["n"] += 1
        return calls["n"] > 25
    r = run_ransac_manual(src, tgt, v, 1000, correspondences=corres, should_stop=stop, early_stop_enabled=False, seed=4)
    assert r.info["iterations"] == 25 and r.info["stopped_early"]
    # fewer than three correspondences: identity, fitness 0 (ransac.py:133-140)
    r = run_ransac_manual(src, tgt, v, 100, correspondences=corres[:2])
    assert np.array_equal(r.transformation, np.eye(4)) and r.fitness == 0.0


def test_align_deterministic_under_overlap_and_threads(eng):
    """pcr_align overlaps stages on a helper context and may be called from several host threads (one context each):
    every run must return the same bits (tools/gpu_stress_align.py is the long version)."""
    import threading
    import torch
    from pcr_b200.engine import Engine
    v = 0.005
    src, tgt, _ = synth.make_pair(15000, v, 4711)

    def run(e, out, reps):
        ds, dt = e.pack(src), e.pack(tgt)
        p = e.default_params(v)
        p.ransac_max_iter = 20000
        p.seed = 5
        for _ in range(reps):
            r = e.align_device(ds, dt, p)
            out.append((tuple(r.icp.transformation), r.icp.fitness, r.icp.inlier_rmse, r.ransac.best_hyp, r.icp.iterations))
    base = []
    run(eng, base, 10)
    assert all(x == base[0] for x in base)
    outs = [[], []]
    ths = [threading.Thread(target=lambda k=k: (torch.cuda.set_device(eng.tdev), run(Engine(eng.tdev.index), outs[k], 10))) for k in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert all(len(o) == 10 and all(x == base[0] for x in o) for o in outs)


def test_main_pipeline_and_headless_draw(tmp_path, eng, capsys):
    """src/main.py:24-39 on the engine: Ply x2 -> global_registration -> draw -> refine_registration -> draw, from PLY
    files through the native reader; the 'viewer' writes a coloured PLY + JSON per call."""
    import importlib.util
    import json
    import os
    from pcr_b200.plyio import probe_ply, read_ply, write_ply
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("b200_main", os.path.join(root, "3d-matching_b200", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    v = 0.005
    src, tgt, T_true = synth.make_pair(12000, v, 77)
    write_ply(tmp_path / "sample.ply", src, binary=False)
    write_ply(tmp_path / "target.ply", tgt)
    out = tmp_path / "viz"
    rc = mod.main(["--source", str(tmp_path / "sample.ply"), "--target", str(tmp_path / "target.ply"), "--voxel-size", str(v),
                   "--ransac-iterations", "50000", "--noise-sigma", "0", "--export-dir", str(out)])
    assert rc == 0
    text = capsys.readouterr().out
    assert "transformation (ICP):" in text and "fitness" in text
    files = sorted(p.name for p in out.iterdir())
    assert len(files) == 4 and files[0].endswith(".json") and files[1].endswith(".ply")
    plys = sorted(out.glob("*.ply"))
    meta = json.loads(plys[-1].with_suffix(".json").read_text())
    T = np.array(meta["transformation"])
    assert np.abs(T[:3, :3] - T_true[:3, :3]).max() < 5e-3 and meta["fitness"] > 0.9
    info = probe_ply(plys[-1])
    assert info.has_colors == 1 and info.n_vertex == meta["n_source"] + meta["n_target"]
    # the first n_source vertices are the down-sampled source moved by T; the rest is the down-sampled target, unmoved
    from ply import Ply
    s, t = Ply(tmp_path / "sample.ply", v, noise_sigma=0.0), Ply(tmp_path / "target.ply", v, noise_sigma=0.0)
    pts, _ = read_ply(plys[-1])
    ns = meta["n_source"]
    assert ns == len(s.pcd_down.points) and meta["n_target"] == len(t.pcd_down.points)
    assert np.array_equal(pts[ns:], t.pcd_down.points)
    assert np.abs(pts[:ns] - (s.pcd_down.points @ T[:3, :3].T + T[:3, 3])).max() < 1e-6
    # missing input: exit code 1 and a message, as Ply raises FileNotFoundError (src/ply/ply.py:46-47)
    assert mod.main(["--source", str(tmp_path / "nope.ply"), "--target", str(tmp_path / "target.ply")]) == 1
    # align() takes the same files through the native reader
    from pcr_b200 import align
    Ta, fit, rmse = align(tmp_path / "sample.ply", tmp_path / "target.ply", v, ransac_iteration=50000)
    Tb, fitb, rmseb = align(src, tgt, v, ransac_iteration=50000)
    assert np.array_equal(Ta, Tb) and fit == fitb and rmse == rmseb
    # pcr_align_files error paths: a truncated file is a ValueError naming the file, an empty cloud the reference's
    # "Point cloud is empty" (src/ply/ply.py:81-84); the context stays usable afterwards
    bad = tmp_path / "bad.ply"
    bad.write_bytes((tmp_path / "target.ply").read_bytes()[:-40])
    with pytest.raises(ValueError, match="truncated"):
        align(tmp_path / "sample.ply", bad, v)
    write_ply(tmp_path / "empty.ply", src[:0])
    with pytest.raises(ValueError, match="empty"):
        align(tmp_path / "empty.ply", tmp_path / "target.ply", v)
    with pytest.raises(TypeError):
        align(tmp_path / "sample.ply", __file__, v)
    Tc, fitc, _ = align(tmp_path / "sample.ply", tmp_path / "target.ply", v, ransac_iteration=50000)
    assert np.array_equal(Tc, Ta) and fitc == fit
    # one path + one array still works (the file goes through read_ply_xyzw)
    Td, fitd, _ = align(tmp_path / "sample.ply", tgt, v, ransac_iteration=50000)
    assert np.array_equal(Td, Ta) and fitd == fit
