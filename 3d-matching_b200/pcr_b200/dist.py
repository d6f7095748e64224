"""Multi-GPU drivers (one process per GPU, torch.distributed over NCCL/NVLink; gloo on CPU for tests).

Two shardings, both without any data-path collective:
  * RANSAC hypotheses: each wave [begin, end) of the global hypothesis stream is split into contiguous slices,
    one per rank; every rank scores its slice (pcr_ransac_wave), keeps the chain of prefix maxima of its slice,
    and ONE small all-gather per wave (a count + <= cap fixed-size records per rank) lets every rank replay the
    sequential loop (pcr_ransac_scan) and reach the identical winner.  Philox is keyed by the global index, so
    the result does not depend on the number of GPUs.
  * batches of independent pairs: pair i -> rank i mod world; results are all-gathered at the end.
Single-pair ICP stays on one GPU ("replicas only").
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _capi

REC_WORDS = 16  # pcr_hyp_record = 128 bytes = 16 x int64


def records_to_array(recs, n: int) -> np.ndarray:
    """ctypes pcr_hyp_record[n] -> (n,16) int64 (bit copy)."""
    if n == 0:
        return np.zeros((0, REC_WORDS), np.int64)
    buf = (C.c_char * (C.sizeof(_capi.HypRecord) * n)).from_address(C.addressof(recs))
    return np.frombuffer(buf, dtype=np.int64).reshape(n, REC_WORDS).copy()


def array_to_records(arr: np.ndarray):
    n = arr.shape[0]
    recs = (_capi.HypRecord * max(n, 1))()
    if n:
        C.memmove(C.addressof(recs), np.ascontiguousarray(arr, np.int64).ctypes.data, n * C.sizeof(_capi.HypRecord))
    return recs


def prefix_maxima(arr: np.ndarray, best_count: int, best_sum: int) -> np.ndarray:
    """Rows of `arr` (sorted by hypothesis index) that improve on the running best — the only survivors that can
    change the state of the sequential loop.  Column 1 = inlier_count, column 2 = sum_d2_fixed."""
    keep = []
    bc, bs = best_count, best_sum
    for i in range(arr.shape[0]):
        c, s = int(arr[i, 1]), int(arr[i, 2])
        if c > bc or (c == bc and bc > 0 and s < bs):
            keep.append(i)
            bc, bs = c, s
    return arr[keep] if keep else arr[:0]


def slice_bounds(begin: int, end: int, rank: int, world: int):
    n = end - begin
    return begin + n * rank // world, begin + n * (rank + 1) // world


FIXED_CAP = 24    # records per rank carried by the per-wave all-gather
WAVE_GROWTH = int(os.environ.get("PCR_DIST_GROWTH", "4"))  # measured at 10M hypotheses: x2 / x4 / x8 = 160 / 162 / 162 M hyp/s on 1 GPU, 295 / 309 / 308 on 2


def ransac_distributed(wave_fn, n_corr: int, n_src: int, k_d: int, max_iter: int, confidence: float, *,
                       group=None, device=None, first_wave: int = 4096, max_wave: int = 1 << 22, lib=None):
    """Generic driver.  wave_fn(lo, hi, best_count, best_sum) -> (n,16) int64 records sorted by hypothesis index
    (a superset of the slice's prefix maxima) and the number of survivors.  Returns (_capi.RegResult, stats)."""
    lib = lib or _capi.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    state = _capi.RegResult()
    for i in (0, 5, 10, 15):
        state.transformation[i] = 1.0
    state.best_hyp = -1
    state.est_k = max_iter
    begin = 0
    wave = first_wave * world
    survivors = 0
    waves = 0
    trace = os.environ.get("PCR_DIST_TRACE") and rank == 0
    while begin < max_iter and begin < state.est_k:
        end = min(max_iter, begin + wave)
        lo, hi = slice_bounds(begin, end, rank, world)
        t0 = time.perf_counter()
        arr, nsurv = wave_fn(lo, hi, int(state.inlier_count), int(state.sum_d2_fixed))
        t1 = time.perf_counter()
        chain = prefix_maxima(arr, int(state.inlier_count), int(state.sum_d2_fixed))
        if world > 1:
            # ONE all-gather per wave in the common case: [count, survivors, FIXED_CAP records] per rank (a chain of
            # prefix maxima is ~ln(n) long); a second one only if some rank's chain does not fit
            buf = np.zeros((2 + FIXED_CAP * REC_WORDS,), np.int64)
            buf[0], buf[1] = chain.shape[0], nsurv
            k = min(chain.shape[0], FIXED_CAP)
            buf[2:2 + k * REC_WORDS] = chain[:k].reshape(-1)
            mine = torch.from_numpy(buf).to(device)
            allb = torch.empty((world * buf.shape[0],), dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(allb, mine, group=group)
            allb_h = allb.cpu().numpy().reshape(world, buf.shape[0])
            counts = allb_h[:, 0].astype(np.int64)
            survivors += int(allb_h[:, 1].sum())
            cap = int(counts.max())
            if cap <= FIXED_CAP:
                merged = np.concatenate([allb_h[r, 2:2 + int(counts[r]) * REC_WORDS].reshape(-1, REC_WORDS)
                                         for r in range(world)], axis=0)
            else:
                pad = np.zeros((cap, REC_WORDS), np.int64)
                pad[: chain.shape[0]] = chain
                mine = torch.from_numpy(pad.reshape(-1)).to(device)
                allr = torch.empty((world * cap * REC_WORDS,), dtype=torch.int64, device=device)
                dist.all_gather_into_tensor(allr, mine, group=group)
                allr_h = allr.cpu().numpy().reshape(world, cap, REC_WORDS)
                merged = np.concatenate([allr_h[r, : int(counts[r])] for r in range(world)], axis=0)
        else:
            merged = chain
            survivors += nsurv
        if trace:
            print(f"[pcr dist] wave [{begin}, {end}): wave_fn {(t1 - t0) * 1e3:.2f} ms, exchange {(time.perf_counter() - t1) * 1e3:.2f} ms",
                  file=sys.stderr)
        recs = array_to_records(merged)
        stop = C.c_int(0)
        lib.pcr_ransac_scan(recs, C.c_int(merged.shape[0]), C.c_int64(begin), C.c_int64(end), C.c_int(n_corr),
                            C.c_int(n_src), C.c_double(confidence), C.c_int32(k_d), C.byref(state), C.byref(stop))
        waves += 1
        begin = end
        if stop.value:
            break
        if wave < max_wave * world:
            wave *= WAVE_GROWTH
    state.survivors = survivors
    if state.hyp_evaluated > max_iter:
        state.hyp_evaluated = max_iter
    return state, {"waves": waves, "world": world}


def ransac_multi_gpu(eng, src, tgt, corr, max_dist: float, max_iter: int, confidence: float = 0.999, seed: int = 0,
                     edge_sim: float = 0.9, group=None, first_wave: int = 4096, max_wave: int = 1 << 22):
    """RANSAC over all ranks of `group` (NCCL).  Every rank holds the same clouds/correspondences and returns the
    same result (engine.DeviceRegResult)."""
    from .engine import DeviceRegResult

    k_d = int(eng.lib.pcr_ransac_k_d(C.c_double(max_dist), C.c_int(src.shape[0])))

    def wave_fn(lo, hi, bc, bs):
        if hi <= lo:
            return np.zeros((0, REC_WORDS), np.int64), 0
        cap = 4096
        while True:
            try:
                recs, n, nsurv = eng.ransac_wave(src, tgt, corr, max_dist, lo, hi, seed, edge_sim, cap, bc, bs)
                return records_to_array(recs, n), nsurv
            except ValueError:
                if cap >= (1 << 22):
                    raise
                cap *= 16

    if corr.shape[0] < 3 or not (max_dist > 0.0) or src.shape[0] == 0 or tgt.shape[0] == 0 or max_iter <= 0:
        st = _capi.RegResult()
        for i in (0, 5, 10, 15):
            st.transformation[i] = 1.0
        st.best_hyp = -1
        st.est_k = max_iter
        return DeviceRegResult.from_c(st), {"waves": 0}
    eng.ransac_session_begin(src, tgt, max_dist)  # grid + sorted source once, not once per wave
    try:
        state, stats = ransac_distributed(wave_fn, int(corr.shape[0]), int(src.shape[0]), k_d, int(max_iter), confidence,
                                          group=group, device=eng.tdev, first_wave=first_wave, max_wave=max_wave, lib=eng.lib)
    finally:
        eng.ransac_session_end()
    return DeviceRegResult.from_c(state), stats


def align_batch(eng, pairs, params, group=None, workers: int = 1):
    """Batch of independent pairs sharded pair i -> rank i mod world.  `pairs` is a sequence of (src, tgt) packed
    CUDA tensors or a callable i -> (src, tgt) plus its length as (fn, n).  Returns an (n, 18) float64 array on
    every rank: 16 transform entries, fitness, inlier RMSE.

    workers > 1: the rank's pairs are aligned by that many host threads, each with its own engine context and CUDA
    stream, so that the host synchronisations of one alignment (voxel counts, RANSAC waves) overlap with the kernels
    of another; the result of a pair does not depend on it."""
    if isinstance(pairs, tuple) and callable(pairs[0]):
        fn, n = pairs
    else:
        fn, n = (lambda i: pairs[i]), len(pairs)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per = (n + world - 1) // world
    out = torch.zeros((per, 18), dtype=torch.float64)
    mine_idx = list(range(rank, n, world))

    def run(e, k, i):
        s, t = fn(i)
        r = e.align_device(s, t, params)
        out[k, :16] = torch.tensor(list(r.icp.transformation), dtype=torch.float64)
        out[k, 16] = r.icp.fitness
        out[k, 17] = r.icp.inlier_rmse

    if workers <= 1 or len(mine_idx) <= 1:
        for k, i in enumerate(mine_idx):
            run(eng, k, i)
    else:
        import threading
        from .engine import Engine
        pool = _worker_engines(eng, workers, Engine)
        errors = []

        def work(w):
            try:
                torch.cuda.set_device(eng.tdev)
                with torch.cuda.stream(pool[w][1]):
                    for k in range(w, len(mine_idx), workers):
                        run(pool[w][0], k, mine_idx[k])
                    pool[w][1].synchronize()
            except Exception as exc:  # surfaced on the calling thread
                errors.append(exc)
        torch.cuda.current_stream(eng.tdev).synchronize()  # inputs produced on the caller's stream are complete
        threads = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    if world > 1:
        mine = out.reshape(-1).to(eng.tdev)
        allr = torch.empty((world * per * 18,), dtype=torch.float64, device=eng.tdev)
        dist.all_gather_into_tensor(allr, mine, group=group)
        allr = allr.cpu().reshape(world, per, 18)
        res = torch.zeros((n, 18), dtype=torch.float64)
        for r in range(world):
            idx = list(range(r, n, world))
            res[idx] = allr[r, : len(idx)]
        return res.numpy()
    return out[:n].numpy()


_worker_pool: dict = {}


def _worker_engines(eng, workers: int, Engine):
    """(engine, stream) per worker thread, created once per device: worker 0 reuses the caller's engine."""
    key = (eng.tdev.index, workers)
    if key not in _worker_pool:
        pool = [(eng, torch.cuda.Stream(device=eng.tdev))]
        for _ in range(workers - 1):
            pool.append((Engine(eng.tdev.index), torch.cuda.Stream(device=eng.tdev)))
        _worker_pool[key] = pool
    return _worker_pool[key]
