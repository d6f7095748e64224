"""Boundary hardening (round 2): correspondence validation, NaN descriptors, in-place mutated inputs, lazy normals,
PCR_ERR_BUSY under concurrent use of ONE context, and worker contexts created from a cold process."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

from pcr_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Cloud:
    def __init__(self, pts):
        self.points = np.asarray(pts, np.float64)


class MockPly:  # test_ransac_crash.py:92-96
    def __init__(self, pts):
        self.pcd = Cloud(pts)
        self.pcd_down = Cloud(pts)
        self.pcd_fpfh = None


def test_out_of_range_correspondences_raise_index_error(eng):
    """The reference indexes the clouds with the pairs (src/matcher/ransac.py:147-148, :223-224): IndexError."""
    from matcher.ransac import compute_step_transformation, evaluate_inlier_ratio, run_ransac_manual
    rng = np.random.default_rng(0)
    a, b = MockPly(rng.random((50, 3))), MockPly(rng.random((40, 3)))
    good = np.stack([np.arange(40), np.arange(40)], 1)
    assert np.isfinite(compute_step_transformation(a, b, good).transformation).all()
    for bad in ([[0, 0], [1, 1], [2, 40]], [[0, 0], [50, 1], [2, 2]], [[0, 0], [1, -1], [2, 2]], [[-7, 0], [1, 1], [2, 2]]):
        bad = np.array(bad)
        with pytest.raises(IndexError):
            compute_step_transformation(a, b, bad)
        with pytest.raises(IndexError):
            evaluate_inlier_ratio(a, b, bad, np.eye(4), 0.05)
        with pytest.raises(IndexError):
            run_ransac_manual(a, b, 0.05, 10, correspondences=bad)
    # the C ABI itself: PCR_ERR_INVALID -> ValueError, and the context stays usable (no sticky CUDA error)
    s, t = eng.pack(a.pcd.points), eng.pack(b.pcd.points)
    c = torch.tensor([[0, 0], [1, 1], [2, 99]], dtype=torch.int32, device=eng.tdev)
    with pytest.raises(ValueError, match="indexes outside"):
        eng.ransac(s, t, c, 0.1, 100, 0.999, 1)
    with pytest.raises(ValueError, match="indexes outside"):
        eng.ransac_wave(s, t, c, 0.1, 0, 64)
    eng.ransac_session_begin(s, t, 0.1)
    with pytest.raises(ValueError, match="indexes outside"):
        eng.ransac_wave(s, t, c, 0.1, 0, 64)
    eng.ransac_session_end()
    ok = torch.tensor(good, dtype=torch.int32, device=eng.tdev)
    assert eng.ransac(s, t, ok, 0.1, 100, 0.999, 1).hyp_evaluated > 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [300, 3000])  # exact CUDA-core kernel / tcgen05 path
def test_nan_descriptor_yields_no_pair(orc, eng, n):
    rng = np.random.default_rng(4)
    fs = (rng.random((n, 33)) * 100).astype(np.float32)
    ft = fs[rng.permutation(n)] + rng.normal(0, 0.01, (n, 33)).astype(np.float32)
    fs[7, 3] = np.nan
    fs[n - 1, :] = np.nan
    for mutual in (False, True):
        got = eng.match_features(torch.from_numpy(fs).to(eng.tdev), torch.from_numpy(ft).to(eng.tdev), mutual).cpu().numpy()
        want = orc.match_features(fs, ft, mutual)
        assert np.array_equal(got, want)
        assert 7 not in got[:, 0] and n - 1 not in got[:, 0] and got.min() >= 0


def test_in_place_mutation_is_seen(eng):
    """a cached device copy keyed by id()/pointer went stale under `pts += ...`; there is no cache now."""
    from matcher.ransac import evaluate_inlier_ratio, evaluate_inlier_ratio_fast
    rng = np.random.default_rng(1)
    p = rng.random((200, 3))
    q = p.copy()
    assert evaluate_inlier_ratio_fast(p, q, np.eye(4), 1e-6) == 1.0
    p += 0.5  # same object, same shape, same data pointer
    assert evaluate_inlier_ratio_fast(p, q, np.eye(4), 1e-6) == 0.0
    a, b = MockPly(q.copy()), MockPly(q.copy())
    cc = np.stack([np.arange(200), np.arange(200)], 1)
    assert evaluate_inlier_ratio(a, b, cc, np.eye(4), 0.05) == 1.0
    a.pcd_down.points[:] += 1.0
    assert evaluate_inlier_ratio(a, b, cc, np.eye(4), 0.05) == 0.0


def test_lazy_normals_follow_open3d_semantics(orc, eng):
    """The reference estimates full-resolution normals eagerly (src/ply/ply.py:65) and Open3D's transform rotates them:
    a lazy estimate must be taken BEFORE the points move, then rotate with them."""
    from ply import Ply
    v = 0.005
    src, _, _ = synth.make_pair(6000, v, 77)
    want = orc.estimate_normals(src, 2 * v, 30)
    ang = 0.7
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [0.3, -0.2, 0.1]
    a = Ply.from_points(src, v)
    a.pcd.transform(T)                       # before the first normals access
    got = a.pcd.normals
    assert np.abs(got - want.astype(np.float64) @ T[:3, :3].T).max() < 1e-6
    b = Ply.from_points(src, v)
    n0 = b.pcd.normals.copy()
    assert np.array_equal(n0.astype(np.float32), want)
    b.pcd.points = b.pcd.points + 1.0        # setter: normals stay (Open3D keeps them), estimated on the original cloud
    assert np.array_equal(b.pcd.normals, n0)
    c = Ply.from_points(src, v)
    c.pcd.points = c.pcd.points + 1.0        # setter before the first access: still the ORIGINAL cloud's normals
    assert np.array_equal(c.pcd.normals, n0)


def test_one_context_two_threads_reports_busy(eng):
    """include/pcr.h: one context may be used by one thread at a time, PCR_ERR_BUSY otherwise.  Two threads call the
    raw C ABI on ONE context (bypassing the Python engine's lock): every call either succeeds or returns PCR_ERR_BUSY,
    some of them do collide, and the context works afterwards."""
    from pcr_b200 import _capi
    lib = _capi.load()
    ctx = C.c_void_p()
    assert lib.pcr_create(C.c_int(0), C.byref(ctx)) == 0
    pts = eng.pack(synth.make_pair(20000, 0.005, 3)[0])
    torch.cuda.synchronize()
    n = pts.shape[0]
    outs = [torch.empty_like(pts) for _ in range(2)]
    codes = [[], []]
    start = threading.Barrier(2)

    def work(k):
        m = C.c_int(0)
        start.wait()
        for _ in range(300):
            codes[k].append(lib.pcr_voxel_downsample(ctx, C.c_void_p(pts.data_ptr()), C.c_int(n), C.c_double(0.005),
                                                     C.c_void_p(outs[k].data_ptr()), C.byref(m)))
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    allc = codes[0] + codes[1]
    assert set(allc) <= {0, _capi.PCR_ERR_BUSY}, set(allc)
    assert allc.count(0) > 0
    assert allc.count(_capi.PCR_ERR_BUSY) > 0, "600 overlapping calls never collided"
    m = C.c_int(0)
    assert lib.pcr_voxel_downsample(ctx, C.c_void_p(pts.data_ptr()), C.c_int(n), C.c_double(0.005),
                                    C.c_void_p(outs[0].data_ptr()), C.byref(m)) == 0 and m.value > 0
    lib.pcr_destroy(ctx)


def test_align_batch_with_three_workers_from_a_cold_process():
    """Worker contexts are created from several host threads in a process that has not touched the library yet
    (per-context occupancy / function-attribute caches)."""
    code = f"""
import sys
sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, '3d-matching_b200')!r}]
import numpy as np
from pcr_b200 import synth
from pcr_b200.engine import get_engine
from pcr_b200.dist import align_batch
eng = get_engine(0)
v = 0.005
pairs = []
for i in range(6):
    s, t, _ = synth.make_pair(12000, v, 900 + i)
    pairs.append((eng.pack(s), eng.pack(t)))
p = eng.default_params(v)
p.ransac_max_iter = 20000
p.seed = 3
a = align_batch(eng, pairs, p, workers=3)
b = align_batch(eng, pairs, p, workers=1)
assert np.array_equal(a, b), np.abs(a - b).max()
assert a[:, 16].min() > 0.9
print("cold ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cold ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_c_side_sharding_entry_points_on_one_gpu(orc, eng):
    """pcr_ransac_multi / pcr_align_batch without a communicator (world 1) are the single-GPU paths: same result as
    pcr_ransac / as aligning the pairs one by one, for any wave schedule and any number of native workers."""
    v = 0.005
    src, tgt, _ = synth.make_pair(20000, v, 20241)
    S, G = orc.preprocess(src, v, full_normals=False), orc.preprocess(tgt, v, full_normals=False)
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    sd, td = eng.pack(S.pcd_down), eng.pack(G.pcd_down)
    dc = torch.from_numpy(np.ascontiguousarray(corr, np.int32)).to(eng.tdev)
    eng.comm_init()
    for conf, iters in ((0.999, 100000), (1.0, 30000)):
        want = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * v, iters, conf, seed=3)
        for first, growth in ((0, 0), (512, 2), (100000, 8)):
            got, waves = eng.ransac_multi(sd, td, dc, 1.5 * v, iters, conf, 3, first_wave=first, growth=growth)
            assert (got.best_hyp, got.inlier_count, got.sum_d2_fixed, got.hyp_evaluated) == \
                   (want.best_hyp, want.inlier_count, want.sum_d2_fixed, want.hyp_evaluated), (conf, first, growth)
            if conf == 1.0:  # with an early exit the engine has scored (and counts) whole waves, the oracle 103 hypotheses
                assert got.survivors == want.survivors
            assert np.array_equal(got.transformation, want.transformation) and waves >= 1
    with pytest.raises(ValueError, match="indexes outside"):
        eng.ransac_multi(sd, td, torch.tensor([[0, 0], [1, 1], [2, 10 ** 6]], dtype=torch.int32, device=eng.tdev), 1.5 * v, 100)
    pairs = []
    for i in range(5):
        s, t, _ = synth.make_pair(12000, v, 700 + i)
        pairs.append((eng.pack(s), eng.pack(t)))
    p = eng.default_params(v)
    p.ransac_max_iter = 20000
    p.seed = 3
    one = np.stack([np.concatenate([np.asarray(r.icp.transformation), [r.icp.fitness, r.icp.inlier_rmse]])
                    for r in (eng.align_device(s, t, p) for s, t in pairs)])
    for workers in (1, 3, 8):
        assert np.array_equal(eng.align_batch(pairs, p, len(pairs), workers=workers), one), workers
    assert eng.align_batch([], p, 0).shape == (0, 18)
    with pytest.raises(ValueError):
        eng.align_batch(pairs[:2], p, 5)  # this rank must hold 5 of 5 pairs
