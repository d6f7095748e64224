"""world_size-2 gloo test of the hypothesis-sharded RANSAC driver (pcr_b200.dist.ransac_distributed).

The per-rank scorer is the CPU oracle (allowed in tests): what is under test is the host logic of the N>1 path —
slicing, prefix-maxima chains, the all-gather and the replay — which must reproduce the single-process result
for any world size."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
    from oracle import pcr_oracle as orc
    from pcr_b200 import synth
    v = 0.02
    src, tgt, T = synth.make_pair(3000, v, 4242)
    S, G = orc.preprocess(src, v, full_normals=False), orc.preprocess(tgt, v, full_normals=False)
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    return orc, v, S.pcd_down, G.pcd_down, corr


def _wave_fn_factory(orc, sd, td, corr, max_dist, seed):
    def wave_fn(lo, hi, bc, bs):
        rows, nsurv = [], 0
        for h in range(lo, hi):
            e = orc.ransac_eval_one(sd, td, corr, max_dist, seed, h)
            if e is None:
                continue
            nsurv += 1
            T, cnt, sq, cin = e
            if cnt > bc or (cnt == bc and bc > 0 and sq < bs):
                r = np.zeros(16, np.int64)
                r[0], r[1], r[2], r[3] = h, cnt, sq, cin
                r[4:] = T[:3].reshape(-1).view(np.int64)
                rows.append(r)
        return np.array(rows, np.int64).reshape(-1, 16), nsurv
    return wave_fn


def _worker(rank, world, port, conf, max_iter, q, fixed_cap=None):
    try:
        _worker_body(rank, world, port, conf, max_iter, q, fixed_cap)
    except Exception as e:  # surface the failure instead of letting the parent wait for its timeout
        q.put(("error", repr(e)))
        raise


def _worker_body(rank, world, port, conf, max_iter, q, fixed_cap=None):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc, v, sd, td, corr = _problem()
    orc.set_num_threads(1)
    from pcr_b200 import _capi
    from pcr_b200 import dist as pdist
    from pcr_b200.dist import ransac_distributed
    if fixed_cap is not None:
        pdist.FIXED_CAP = fixed_cap  # force the second all-gather (a chain longer than the per-wave record budget)
    lib = _capi.load()
    import ctypes as C
    k_d = int(lib.pcr_ransac_k_d(C.c_double(1.5 * v), C.c_int(len(sd))))
    st, stats = ransac_distributed(_wave_fn_factory(orc, sd, td, corr, 1.5 * v, 3), len(corr), len(sd), k_d, max_iter, conf,
                                   device="cpu", first_wave=64, lib=lib)
    q.put((rank, st.best_hyp, st.inlier_count, st.sum_d2_fixed, st.est_k, st.hyp_evaluated, list(st.transformation),
           st.fitness, st.inlier_rmse))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("conf,max_iter,fixed_cap", [(0.999, 3000, None), (1.0, 600, None), (1.0, 400, 1)])
def test_two_ranks_reproduce_the_sequential_result(conf, max_iter, fixed_cap):
    orc, v, sd, td, corr = _problem()
    want = orc.ransac(sd, td, corr, 1.5 * v, max_iter, conf, seed=3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, conf, max_iter, q, fixed_cap)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    assert all(g[0] != "error" for g in got), got
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g in got:
        _, best, cnt, sq, est_k, ev, T, fit, rmse = g
        assert (best, cnt, sq, est_k, ev) == (want.best_hyp, want.inlier_count, want.sum_d2_fixed, want.est_k, want.hyp_evaluated)
        assert np.array_equal(np.array(T).reshape(4, 4), want.transformation)
        assert fit == want.fitness and rmse == want.inlier_rmse


class _OracleEngine:
    """Duck-typed stand-in for pcr_b200.engine.Engine in align_batch: `tdev` and `align_device(src, tgt, params)`.
    The alignment itself is the CPU oracle's (allowed in tests); what is under test is the pair -> rank mapping, the
    ragged last round (5 pairs on 2 ranks) and the final all-gather."""

    def __init__(self, orc, v):
        import torch
        self.orc, self.v, self.tdev = orc, v, torch.device("cpu")
        self.calls = []

    def align_device(self, src, tgt, params):
        from types import SimpleNamespace
        S, G = self.orc.preprocess(src, self.v), self.orc.preprocess(tgt, self.v)
        ro = self.orc.global_registration(S, G, self.v, 2000, 0.999, 7)
        io = self.orc.refine_registration(S, G, ro.transformation, self.v)
        self.calls.append(len(src))
        return SimpleNamespace(icp=SimpleNamespace(transformation=io.transformation.reshape(-1), fitness=io.fitness,
                                                   inlier_rmse=io.inlier_rmse))


def _batch_pairs():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
    from pcr_b200 import synth
    return [synth.make_pair(1500 + 100 * i, 0.02, 900 + i)[:2] for i in range(5)]


def _batch_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        pairs = _batch_pairs()
        from oracle import pcr_oracle as orc
        orc.build()
        orc.set_num_threads(1)
        from pcr_b200.dist import align_batch
        eng = _OracleEngine(orc, 0.02)
        # a rank only materialises its own pairs (the others are None, as bench.py's batch leg does)
        mine = [p if i % world == rank else None for i, p in enumerate(pairs)]
        tab = align_batch(eng, mine, None)
        q.put((rank, tab.tolist(), eng.calls))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        q.put(("error", repr(e), None))
        raise


def test_align_batch_shards_pairs_and_gathers_on_two_ranks():
    pairs = _batch_pairs()
    from oracle import pcr_oracle as orc
    orc.build()
    eng = _OracleEngine(orc, 0.02)
    sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
    from pcr_b200.dist import align_batch
    want = align_batch(eng, pairs, None)  # no process group: everything on this process
    assert want.shape == (5, 18) and (want[:, 16] > 0.5).all()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    assert all(g[0] != "error" for g in got), got
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tab, calls in got:
        assert np.array_equal(np.array(tab), want)                      # every rank holds the whole table, pair order kept
        assert calls == [len(pairs[i][0]) for i in range(rank, 5, 2)]   # and aligned exactly its own pairs, in order
