"""Mirror of the reference's src/matcher/ransac.py on the B200 engine — same function names, positional
order and soft-failure behaviour; every Open3D / NumPy hot loop runs in libpcr_b200.so.

    global_registration              ransac.py:20-59    -> pcr_match_features + pcr_ransac
    compute_feature_correspondences  ransac.py:62-101   -> pcr_match_features (+ host noise pairs, :89-99)
    compute_step_transformation      ransac.py:104-192  -> pcr_ransac_step
    evaluate_inlier_ratio            ransac.py:195-236  -> pcr_inlier_count
    evaluate_inlier_ratio_fast       ransac.py:239-277  -> pcr_inlier_count (squared)
    run_ransac_manual                src/visualize_matcher/_visualize_matcher.py:343-470 (the GUI's step-by-step loop:
                                     best = strictly greater inlier ratio, early stop, update callbacks) -> batched
                                     pcr_ransac_step + pcr_inlier_count, the loop itself replayed on the host

Keyword-only extras (never disturb positional use): confidence, seed.  Randomness is Philox keyed by
(seed, hypothesis index) instead of the reference's unseeded global generators.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from pcr_b200.containers import RegistrationResult
from pcr_b200.engine import get_engine

from ._common import device_cloud, device_corr, device_feature, index_errors, voxel_of

_step_counter = itertools.count()


def _inlier_correspondences(eng, src_xyzw, tgt_xyzw, T, max_dist):
    """Open3D's RegistrationResult.correspondence_set: (i, nn(i)) for transformed source points within max_dist."""
    def make():
        moved = eng.transform_points(src_xyzw, T)
        idx, _ = eng.nn1(tgt_xyzw, moved, max_dist)
        idx = idx.cpu().numpy()
        keep = np.nonzero(idx >= 0)[0]
        return np.stack([keep.astype(np.int32), idx[keep]], axis=1)
    return make


def global_registration(src, tgt, voxel_size: float | None = None, iteration: int = 30, *, confidence: float = 0.999,
                        seed: int = 0) -> RegistrationResult:
    voxel_size = voxel_of(src, voxel_size)
    eng = get_engine()
    dist_thresh = voxel_size * 1.5
    s, t = device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng)
    fs, ft = device_feature(src.pcd_fpfh, eng), device_feature(tgt.pcd_fpfh, eng)
    corr = eng.match_features(fs, ft, mutual=True)  # mutual_filter=True, ransac.py:47
    r = eng.ransac(s, t, corr.contiguous(), dist_thresh, int(iteration), confidence, seed, edge_sim=0.9)
    res = RegistrationResult(r.transformation, r.fitness, r.inlier_rmse,
                             _inlier_correspondences(eng, s, t, r.transformation, dist_thresh) if r.best_hyp >= 0 else None)
    res.info = {"best_hyp": r.best_hyp, "hyp_evaluated": r.hyp_evaluated, "survivors": r.survivors, "est_k": r.est_k,
                "inlier_count": r.inlier_count, "n_corr": int(corr.shape[0])}
    return res


def compute_feature_correspondences(src, tgt, mutual_filter: bool = False, noise_ratio: float = 0.0, *,
                                    seed: int | None = None) -> np.ndarray:
    eng = get_engine()
    fs, ft = device_feature(src.pcd_fpfh, eng), device_feature(tgt.pcd_fpfh, eng)
    corres_np = eng.match_features(fs, ft, mutual=bool(mutual_filter)).cpu().numpy()
    if noise_ratio > 0:
        n_original = len(corres_np)
        n_noise = int(n_original * noise_ratio)
        if n_noise > 0:
            rng = np.random if seed is None else np.random.RandomState(seed)
            src_indices = rng.randint(0, len(src.pcd_down.points), n_noise)
            tgt_indices = rng.randint(0, len(tgt.pcd_down.points), n_noise)
            noise_corres = np.stack((src_indices, tgt_indices), axis=1).astype(np.int32)
            corres_np = np.vstack((corres_np, noise_corres))
            rng.shuffle(corres_np)
    return np.ascontiguousarray(corres_np, dtype=np.int32)


def compute_step_transformation(src, tgt, correspondences, *, seed: int = 0, index: int | None = None) -> RegistrationResult:
    res = RegistrationResult()  # identity, fitness 0.0 (ransac.py:134-136)
    corr = np.asarray(correspondences)
    if len(corr) < 3:
        return res
    eng = get_engine()
    h = next(_step_counter) if index is None else int(index)
    with index_errors():
        T = eng.ransac_step(device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng), device_corr(corr, eng), seed, h, 1)
    res.transformation = T[0].cpu().numpy()
    return res


def compute_step_transformations(src, tgt, correspondences, count: int, *, seed: int = 0, start: int = 0) -> torch.Tensor:
    """Batched twin: `count` hypotheses in one launch; returns a (count,4,4) fp64 CUDA tensor."""
    eng = get_engine()
    with index_errors():
        return eng.ransac_step(device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng),
                               device_corr(correspondences, eng), seed, start, count)


def evaluate_inlier_ratio(src, tgt, correspondences, transform, voxel_size: float) -> float:
    dist_thresh = voxel_size * 1.5
    corr = np.asarray(correspondences)
    if len(corr) == 0:
        return 0.0
    eng = get_engine()
    T = torch.as_tensor(np.asarray(transform, np.float64).reshape(1, 4, 4))
    with index_errors():
        cnt = eng.inlier_count(device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng), device_corr(corr, eng), T,
                               dist_thresh, squared=False)
    return float(cnt[0].item()) / len(corr)


def evaluate_inlier_ratios(src, tgt, correspondences, transforms, voxel_size: float) -> np.ndarray:
    """Batched twin of evaluate_inlier_ratio for a (k,4,4) stack of transforms."""
    corr = np.asarray(correspondences) if not isinstance(correspondences, torch.Tensor) else correspondences
    if len(corr) == 0:
        return np.zeros(len(transforms))
    eng = get_engine()
    with index_errors():
        cnt = eng.inlier_count(device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng), device_corr(corr, eng),
                               torch.as_tensor(transforms), voxel_size * 1.5, squared=False)
    return cnt.cpu().numpy() / float(len(corr))


def evaluate_inlier_ratio_fast(p_src, p_tgt, transform, dist_thresh_sq: float) -> float:
    if len(p_src) == 0:
        return 0.0
    eng = get_engine()
    n = len(p_src)
    ident = torch.arange(n, dtype=torch.int32, device=eng.tdev)
    corr = torch.stack([ident, ident], dim=1).contiguous()
    T = torch.as_tensor(np.asarray(transform, np.float64).reshape(1, 4, 4))
    cnt = eng.inlier_count(device_cloud(p_src, eng), device_cloud(p_tgt, eng), corr, T, dist_thresh_sq, squared=True)
    return float(cnt[0].item()) / n


def required_iterations(inlier_ratio: float, confidence: float = 0.99, sample_size: int = 3, max_iter: int = 0) -> int:
    """RANSAC iteration count for a given inlier ratio (the GUI's compute_required_iterations,
    _visualize_matcher.py:356-370): N = int(log(1 - confidence) / log(1 - ratio^sample_size)); ratio < 0.01 -> max_iter."""
    if inlier_ratio < 0.01:
        return int(max_iter)
    with np.errstate(divide="ignore"):
        return int(np.log(1 - confidence) / np.log(1 - inlier_ratio ** sample_size))


def run_ransac_manual(src, tgt, voxel_size: float | None = None, max_iter: int = 1000, *, correspondences=None,
                      noise_ratio: float = 2.0, early_stop_enabled: bool = True, early_stop_threshold: float = 0.5,
                      early_stop_confidence: float = 0.99, update_interval: int = 10, callback=None, should_stop=None,
                      seed: int = 0, batch: int = 4096) -> RegistrationResult:
    """The step-by-step RANSAC of the reference's GUI worker (_visualize_matcher.py:343-470) with the same loop
    semantics, minus the window: per iteration one 3-point Kabsch hypothesis (compute_step_transformation) scored by
    evaluate_inlier_ratio_fast on the cached correspondence pairs with the squared threshold (1.5 v)^2; the best is
    replaced only by a STRICTLY greater ratio (the first best wins ties, :426-429); early stop when
    best > early_stop_threshold and iter >= required_iterations(best, confidence) (:432-450); `callback(result, iter,
    ratio, best_ratio)` fires every `update_interval` iterations and on every new best (:453-466); `should_stop()` is
    polled before each iteration (:396-409).  Defaults are MatcherSettings' (:151-173).

    Hypotheses are generated and scored on the GPU `batch` at a time (two launches per batch) and the sequential loop
    is replayed over them on the host, so the result is what the one-at-a-time loop returns for the same hypothesis
    stream (Philox, seed) — iterations past an early stop are simply discarded.  Returns the best RegistrationResult
    (fitness = inlier ratio; `.info` holds iterations, stopped_early, n_correspondences), or an identity result when
    there are fewer than three correspondences (ransac.py:133-140)."""
    voxel_size = voxel_of(src, voxel_size)
    corres = compute_feature_correspondences(src, tgt, noise_ratio=noise_ratio, seed=seed) if correspondences is None \
        else np.ascontiguousarray(np.asarray(correspondences), dtype=np.int32)
    best = RegistrationResult()
    best.info = {"iterations": 0, "stopped_early": False, "n_correspondences": int(len(corres))}
    if len(corres) < 3 or max_iter <= 0:
        return best
    eng = get_engine()
    s, t, c = device_cloud(src.pcd_down, eng), device_cloud(tgt.pcd_down, eng), device_corr(corres, eng)
    thresh_sq = (voxel_size * 1.5) ** 2
    best_fitness, have_best, it = -1.0, False, 0
    while it < max_iter:
        n = min(batch, max_iter - it)
        with index_errors():
            Ts = eng.ransac_step(s, t, c, seed, it, n)
            w = (eng.inlier_count(s, t, c, Ts, thresh_sq, squared=True).cpu().numpy() / float(len(corres)))
        Ts_h = None
        for k in range(n):
            if should_stop is not None and should_stop():
                best.info.update(iterations=it, stopped_early=True)
                return best
            it += 1
            w_cur = float(w[k])
            is_new_best = (not have_best) or w_cur > best_fitness
            if is_new_best or (callback is not None and it % update_interval == 0):
                if Ts_h is None:
                    Ts_h = Ts.cpu().numpy()
                cur = RegistrationResult(Ts_h[k].copy(), w_cur, 0.0, None)
                if is_new_best:
                    info = best.info
                    best, best_fitness, have_best = cur, w_cur, True
                    best.info = info
            if early_stop_enabled and best_fitness > early_stop_threshold and \
                    it >= required_iterations(best_fitness, early_stop_confidence, 3, max_iter):
                if callback is not None:
                    callback(best, it, best_fitness, best_fitness)
                best.info.update(iterations=it, stopped_early=True)
                return best
            if callback is not None and (it % update_interval == 0 or is_new_best):
                callback(cur, it, w_cur, best_fitness)
    best.info.update(iterations=it, stopped_early=False)
    return best
