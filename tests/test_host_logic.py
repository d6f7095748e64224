"""Host-side logic that needs no GPU: PLY I/O, the RANSAC replay (pcr_ransac_scan), wave slicing, containers."""
import ctypes as C

import numpy as np
import pytest

from pcr_b200 import _capi, synth
from pcr_b200.dist import array_to_records, prefix_maxima, records_to_array, slice_bounds
from pcr_b200.plyio import read_ply, write_ply


@pytest.mark.parametrize("binary", [True, False])
def test_ply_roundtrip(tmp_path, binary):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(257, 3)).astype(np.float32)
    nrm = rng.normal(size=(257, 3)).astype(np.float32)
    p = tmp_path / "a.ply"
    write_ply(p, pts, nrm, binary=binary)
    rp, rn = read_ply(p)
    assert np.allclose(rp, pts, atol=1e-6) and np.allclose(rn, nrm, atol=1e-6)
    write_ply(p, pts, None, binary=binary)
    rp, rn = read_ply(p)
    assert rn is None and rp.shape == (257, 3)
    write_ply(p, pts[:0], None, binary=binary)
    assert read_ply(p)[0].shape == (0, 3)
    (tmp_path / "bad.ply").write_text("nope\n")
    with pytest.raises(ValueError):
        read_ply(tmp_path / "bad.ply")


def test_slices_cover_the_wave():
    for world in (1, 2, 3, 8):
        for b, e in ((0, 4096), (100, 101), (7, 7), (5, 1000003)):
            parts = [slice_bounds(b, e, r, world) for r in range(world)]
            assert parts[0][0] == b and parts[-1][1] == e
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def mk(hyp, cnt, sumq, cin):
    r = np.zeros(16, np.int64)
    r[0], r[1], r[2] = hyp, cnt, sumq
    r[3] = cin  # low 32 bits = corr_inliers
    r[4:] = np.arange(12, dtype=np.float64).view(np.int64) + hyp
    return r


def run_scan(recs, begin, end, c, ms, conf, state=None, max_iter=10 ** 6):
    lib = _capi.load()
    st = state or _capi.RegResult()
    if state is None:
        st.best_hyp = -1
        st.est_k = max_iter
    arr = np.array(recs, np.int64).reshape(-1, 16)
    stop = C.c_int(0)
    lib.pcr_ransac_scan(array_to_records(arr), C.c_int(len(arr)), C.c_int64(begin), C.c_int64(end), C.c_int(c), C.c_int(ms),
                        C.c_double(conf), C.c_int32(40), C.byref(st), C.byref(stop))
    return st, stop.value


def sequential(recs, c, conf, max_iter):
    """The loop of SURVEY A.6, single thread, over explicit survivor records."""
    by_h = {int(r[0]): r for r in recs}
    est_k, bc, bs, best, evaluated = max_iter, 0, 0, -1, 0
    for itr in range(max_iter):
        if itr >= est_k:
            break
        evaluated += 1
        r = by_h.get(itr)
        if r is None:
            continue
        cnt, sq, cin = int(r[1]), int(r[2]), int(r[3])
        if cnt > bc or (cnt == bc and bc > 0 and sq < bs):
            bc, bs, best = cnt, sq, itr
            ratio = cin / c
            with np.errstate(divide="ignore", invalid="ignore"):
                est = np.log(1 - conf) / np.log(1 - ratio ** 3)
            if est >= 0 and est < est_k:
                est_k = int(np.ceil(est))
    return best, bc, bs, est_k, evaluated


def test_scan_equals_sequential_loop_for_any_wave_split():
    rng = np.random.default_rng(5)
    c, ms, conf, max_iter = 500, 1000, 0.999, 5000
    for trial in range(30):
        hyps = np.sort(rng.choice(max_iter, 200, replace=False))
        recs = [mk(h, int(rng.integers(1, 900)), int(rng.integers(1, 10 ** 9)), int(rng.integers(3, 200 + trial * 10))) for h in hyps]
        want = sequential(recs, c, conf, max_iter)
        for wave in (64, 1000, 4096, 10 ** 6):
            st = None
            b = 0
            while b < max_iter:
                e = min(max_iter, b + wave)
                chunk = [r for r in recs if b <= r[0] < e]
                if st is not None:
                    chunk = list(prefix_maxima(np.array(chunk, np.int64).reshape(-1, 16), st.inlier_count, st.sum_d2_fixed))
                st, stop = run_scan(chunk, b, e, c, ms, conf, st, max_iter)
                b = e
                if stop:
                    break
            assert (st.best_hyp, st.inlier_count, st.sum_d2_fixed, st.est_k, st.hyp_evaluated) == want, (trial, wave)
            assert st.fitness == st.inlier_count / ms


def test_scan_edge_cases():
    st, stop = run_scan([], 0, 100, 10, 10, 0.999, max_iter=1000)
    assert stop == 0 and st.best_hyp == -1 and st.hyp_evaluated == 100 and st.fitness == 0.0
    st, stop = run_scan([], 0, 100, 10, 10, 0.999, max_iter=100)   # range exhausted: est_k == max_iter == end
    assert stop == 1 and st.hyp_evaluated == 100
    # confidence 1.0 never shortens the loop
    st, stop = run_scan([mk(3, 5, 10, 10)], 0, 100, 10, 10, 1.0, max_iter=100)
    assert st.est_k == 100 and st.best_hyp == 3
    # zero fitness never beats the default result (IsBetterRANSACThan)
    st, stop = run_scan([mk(3, 0, 0, 3)], 0, 100, 10, 10, 0.999, max_iter=100)
    assert st.best_hyp == -1
    # equal count, smaller sum wins; equal both keeps the first
    st, _ = run_scan([mk(1, 5, 10, 3), mk(2, 5, 9, 3), mk(3, 5, 9, 3)], 0, 100, 1000, 10, 0.5, max_iter=100)
    assert st.best_hyp == 2


def test_record_roundtrip_and_prefix_maxima():
    arr = np.array([mk(5, 3, 7, 3), mk(9, 3, 6, 4), mk(11, 2, 1, 5), mk(12, 4, 100, 6)], np.int64)
    back = records_to_array(array_to_records(arr), 4)
    assert np.array_equal(arr, back)
    assert [int(r[0]) for r in prefix_maxima(arr, 0, 0)] == [5, 9, 12]
    assert [int(r[0]) for r in prefix_maxima(arr, 3, 6)] == [12]
    assert prefix_maxima(arr[:0], 0, 0).shape == (0, 16)


def test_synth_is_deterministic_and_consistent():
    a = synth.make_pair(2000, 0.005, 5)
    b = synth.make_pair(2000, 0.005, 5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    src, tgt, T = a
    assert src.dtype == np.float32 and tgt.dtype == np.float32
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
    moved = src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    assert abs(np.linalg.norm(moved.mean(0) - tgt.mean(0))) < 0.1 * np.ptp(tgt)


def test_required_iterations_formula():
    """The GUI worker's early-stop bound (src/visualize_matcher/_visualize_matcher.py:356-370):
    int(log(1 - confidence) / log(1 - ratio^3)), max_iter below a 1 % inlier ratio."""
    import math
    from matcher.ransac import required_iterations
    assert required_iterations(0.005, 0.99, 3, 1234) == 1234
    for ratio, conf in ((0.5, 0.99), (0.9, 0.99), (0.51, 0.999), (0.2, 0.9)):
        assert required_iterations(ratio, conf, 3, 10) == int(math.log(1 - conf) / math.log(1 - ratio ** 3))
    assert required_iterations(0.5, 0.99) == 34
    assert required_iterations(1.0, 0.99) == 0  # log(0) = -inf: the loop stops at once


def test_benchmark_cli_arguments_match_the_reference():
    """Same flags and defaults as the reference's benchmark_ransac.py:283-343 (the run itself needs a GPU)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "3d-matching_b200", "benchmark_ransac.py")).read()
    for flag, default in (("--source", '"sample.ply"'), ("--target", '"target.ply"'), ("--voxel-size", "0.3"),
                          ("--noise-ratio", "0.0"), ("--test-iterations", "100"), ("--ransac-iterations", "30")):
        line = next(ln for ln in src.splitlines() if f'"{flag}"' in ln)
        assert f"default={default}" in line, line
    assert importlib.util.find_spec("torch") is not None


def test_main_py_imports_resolve_without_the_reference(tmp_path):
    """src/main.py:13-17 imports matcher, ply, utils.setup_logging and visualization as top-level packages; with
    3d-matching_b200/ on the path they all resolve to this engine."""
    import importlib
    import logging
    for name in ("matcher.icp", "matcher.ransac", "ply", "utils.setup_logging", "visualization.draw_registration_result"):
        m = importlib.import_module(name)
        assert "3d-matching_b200" in m.__file__, (name, m.__file__)
    from utils.setup_logging import setup_logging
    log = setup_logging("pcr.test.logger")
    assert log.level == logging.INFO and setup_logging("pcr.test.logger") is log and len(log.handlers) <= 1
    from visualization.draw_registration_result import SOURCE_COLOR, TARGET_COLOR, _as_matrix
    assert SOURCE_COLOR == (1.0, 0.706, 0.0) and TARGET_COLOR == (0.0, 0.651, 0.929)  # draw_registration_result.py:37-38
    from pcr_b200.containers import RegistrationResult
    T = np.arange(16.0).reshape(4, 4)
    assert np.array_equal(_as_matrix(RegistrationResult(T)), T) and np.array_equal(_as_matrix(T.tolist()), T)
    with pytest.raises(ValueError):
        _as_matrix(np.eye(3))
