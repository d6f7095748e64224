"""The C oracle against INDEPENDENT NumPy restatements of the published Open3D 0.19.0 algorithms.

Open3D itself cannot be installed here (parity with the real wheel stays unpinned), so the next best
anchor is a second implementation written from the algorithm descriptions alone: brute-force neighbours
(tests/np_ref.py), libm / numpy.linalg instead of the oracle's fixed polynomials, eigen-solver and LDL^T, no shared
code with oracle/pcr_oracle.c.  Agreement is to tolerance, not bit level, because the two sides round differently —
what it pins is the SEMANTICS (which neighbours, which bins, which weights, which update), for the stages whose
arithmetic lives in the absent wheel:
    normals   src/ply/ply.py:110-112   EstimateNormals(Hybrid(2v, 30))          A.2, A.3
    FPFH      src/ply/ply.py:117-120   ComputeFPFHFeature(Hybrid(5v, 100))      A.4
    ICP step  src/matcher/icp.py:42-48 RegistrationICP + PointToPlane           A.7
"""
import math

import numpy as np
import pytest

import np_ref
from pcr_b200 import synth


def np_normals(pts, radius, max_nn):
    """A.2: covariance of the hybrid neighbourhood (query included) -> eigenvector of the smallest eigenvalue;
    fewer than 3 neighbours -> (0, 0, 1).  Unoriented."""
    idx, _, cnt = np_ref.knn_hybrid(pts, pts, radius, max_nn)
    out = np.zeros((len(pts), 3))
    gap = np.zeros(len(pts))
    P = np.asarray(pts, np.float64)
    for i in range(len(pts)):
        if cnt[i] < 3:
            out[i] = [0, 0, 1]
            gap[i] = 1.0
            continue
        q = P[idx[i, : cnt[i]]]
        cov = (q.T @ q) / cnt[i] - np.outer(q.mean(0), q.mean(0))
        w, v = np.linalg.eigh(cov)
        out[i] = v[:, 0]
        gap[i] = (w[1] - w[0]) / max(w[2], 1e-300)
    return out, gap


def np_pair_feature(p1, n1, p2, n2):
    """A.4 (Rusu's Darboux-frame angles as Open3D's ComputePairFeatures orders them)."""
    d = p2 - p1
    dist = math.sqrt(float(d @ d))
    if dist == 0.0:
        return np.zeros(4)
    a1, a2 = float(n1 @ d) / dist, float(n2 @ d) / dist
    if math.acos(min(1.0, abs(a1))) > math.acos(min(1.0, abs(a2))):
        n1, n2, d, f2 = n2, n1, -d, -a2
    else:
        f2 = a1
    v = np.cross(d, n1)
    vn = math.sqrt(float(v @ v))
    if vn == 0.0:
        return np.zeros(4)
    v = v / vn
    w = np.cross(n1, v)
    return np.array([math.atan2(float(w @ n2), float(n1 @ n2)), float(v @ n2), f2, dist])


def np_fpfh(pts, nrm, radius, max_nn):
    """A.4: SPFH histograms (3 x 11 bins, 100/(k-1) per neighbour), then the 1/d^2-weighted neighbour sum, each
    11-bin block rescaled to 100, plus the point's own SPFH."""
    P, N = np.asarray(pts, np.float64), np.asarray(nrm, np.float64)
    idx, d2, cnt = np_ref.knn_hybrid(pts, pts, radius, max_nn)
    n = len(P)
    spfh = np.zeros((n, 33))
    margins = np.full(n, np.inf)  # distance of the closest pair feature to a bin edge, in bins
    for i in range(n):
        k = cnt[i]
        if k <= 1:
            continue
        inc = 100.0 / (k - 1)
        for j in idx[i, 1:k]:
            f = np_pair_feature(P[i], N[i], P[j], N[j])
            pos = [11 * (f[0] + math.pi) / (2 * math.pi), 11 * (f[1] + 1.0) * 0.5, 11 * (f[2] + 1.0) * 0.5]
            for b, x in enumerate(pos):
                h = min(10, max(0, int(math.floor(x))))
                spfh[i, 11 * b + h] += inc
                margins[i] = min(margins[i], abs(x - round(x)))
    F = np.zeros((n, 33))
    for i in range(n):
        k = cnt[i]
        if k <= 1:
            continue
        acc, tot = np.zeros(33), np.zeros(3)
        for j, dd in zip(idx[i, 1:k], d2[i, 1:k].astype(np.float64)):
            if dd == 0.0:
                continue
            val = spfh[j] / dd
            acc += val
            tot += val.reshape(3, 11).sum(1)
        scale = np.where(tot != 0.0, 100.0 / np.where(tot != 0.0, tot, 1.0), 0.0)
        F[i] = acc * np.repeat(scale, 11) + spfh[i]
    return F, margins, idx, cnt


def test_normals_match_numpy_eigh(orc):
    v = 0.05
    pts = synth.surface(1200, v, 5, spacing_ratio=1.0).astype(np.float32)
    mine = orc.estimate_normals(pts, 2 * v, 30).astype(np.float64)
    ref, gap = np_normals(pts, 2 * v, 30)
    ok = gap > 1e-3  # the smallest eigenvector is well defined
    assert ok.mean() > 0.95
    sine = np.linalg.norm(np.cross(mine, ref), axis=1)  # unoriented: only the line matters
    assert sine[ok].max() < 1e-6, sine[ok].max()          # normals are stored fp32 (6e-8 per component)
    assert np.allclose(np.linalg.norm(mine, axis=1), 1.0, atol=1e-6)


def test_fpfh_matches_numpy_restatement(orc):
    v = 0.05
    pts = synth.surface(700, v, 11, spacing_ratio=1.0).astype(np.float32)
    nrm = orc.estimate_normals(pts, 2 * v, 30)
    got = orc.fpfh(pts, nrm, 5 * v, 100).astype(np.float64)
    want, margins, idx, cnt = np_fpfh(pts, nrm, 5 * v, 100)
    assert (cnt > 1).mean() > 0.99 and cnt.max() > 30
    # A pair feature closer than 1e-9 bins to a bin edge could legitimately land on either side (the oracle evaluates
    # atan2/acos with its own <= 4 ulp polynomials); a row is comparable when neither it nor any of its neighbours has
    # such a feature.  On this cloud that is every row.
    risky = margins < 1e-9
    tainted = risky.copy()
    for i in range(len(pts)):
        if risky[idx[i, : cnt[i]]].any():
            tainted[i] = True
    assert tainted.mean() < 0.01
    err = np.abs(got - want).max(1)
    assert err[~tainted].max() < 2e-4, err[~tainted].max()  # descriptors are stored fp32: 200 * 2^-24 = 1.2e-5 per bin
    assert np.array_equal(got[cnt <= 1], np.zeros(((cnt <= 1).sum(), 33)))


def test_icp_step_matches_numpy_gauss_newton(orc):
    """One RegistrationICP iteration: NN pass, J = [s x n ; n], r = (s - t).n, x = solve(JtJ, -Jtr),
    update = Rz(x2) Ry(x1) Rx(x0) | (x3, x4, x5), T <- update T; then the evaluation pass (fitness, inlier RMSE)."""
    v = 0.05
    tgt = synth.surface(1500, v, 21, spacing_ratio=0.5).astype(np.float32)
    tn = orc.estimate_normals(tgt, 2 * v, 30)
    T0 = np.eye(4)
    T0[:3, :3] = synth.euler_zyx(0.01, -0.008, 0.012)
    T0[:3, 3] = [0.004, -0.003, 0.002]
    src = synth.surface(1300, v, 22, spacing_ratio=0.5).astype(np.float32)
    max_dist = 0.4 * v
    for it in (1, 2):
        T = T0.copy()
        for _ in range(it):
            s = np_ref.transform_f32(T, src)
            j, _ = np_ref.nn1(tgt, s, max_dist)
            m = j >= 0
            assert m.sum() > 300
            S, Q, Nn = s[m].astype(np.float64), tgt[j[m]].astype(np.float64), tn[j[m]].astype(np.float64)
            r = ((S - Q) * Nn).sum(1)
            J = np.concatenate([np.cross(S, Nn), Nn], axis=1)
            x = np.linalg.solve(J.T @ J, -(J.T @ r))
            U = np.eye(4)
            U[:3, :3] = synth.euler_zyx(x[0], x[1], x[2])
            U[:3, 3] = x[3:]
            T = U @ T
        got = orc.icp_point_to_plane(src, tgt, tn, max_dist, T0, it, 0.0, 0.0)
        assert got.iterations == it
        assert np.abs(got.transformation - T).max() < 1e-8, np.abs(got.transformation - T).max()
        s = np_ref.transform_f32(T, src)
        j, d2 = np_ref.nn1(tgt, s, max_dist)
        m = j >= 0
        # the last-bit difference between the two transforms can move a point across the radius: allow one
        assert abs(got.inlier_count - int(m.sum())) <= 1
        assert abs(got.fitness - m.mean()) <= 1.0 / len(src)
        assert abs(got.inlier_rmse - math.sqrt(d2[m].astype(np.float64).sum() / m.sum())) < 1e-6 * max_dist


def py_philox4x32_10(counter: int, key: int):
    """Philox4x32-10 (Salmon et al., SC'11) on Python integers: 128-bit counter, 64-bit key -> four 32-bit words."""
    M = 0xFFFFFFFF
    c = [(counter >> (32 * i)) & M for i in range(4)]
    k0, k1 = key & M, (key >> 32) & M
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & M, (p0 >> 32) ^ c[3] ^ k1, p0 & M]
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c


def np_umeyama(X, Y):
    """Eigen::umeyama(X, Y, with_scaling=false) as published: Sigma = Yc^T Xc / n, U S V^T = svd(Sigma),
    S = diag(1, 1, sign(det U det V)), R = U S V^T, t = mu_y - R mu_x."""
    mx, my = X.mean(0), Y.mean(0)
    sigma = (Y - my).T @ (X - mx) / len(X)
    U, _, Vt = np.linalg.svd(sigma)
    S = np.diag([1.0, 1.0, 1.0 if np.linalg.det(U) * np.linalg.det(Vt) > 0 else -1.0])
    T = np.eye(4)
    T[:3, :3] = U @ S @ Vt
    T[:3, 3] = my - T[:3, :3] @ mx
    return T


@pytest.mark.parametrize("conf,max_iter", [(0.999, 4000), (1.0, 1500)])
def test_ransac_loop_matches_sequential_numpy_restatement(orc, conf, max_iter):
    """A.6, single-threaded: Philox-keyed 3-samples with replacement, edge-length(0.9) and distance checkers, Umeyama,
    full NN validation of the survivors, best = (more inliers, then smaller RMSE), est_k = ln(1-conf)/ln(1-ratio^3)."""
    v = 0.05
    src_full, tgt_full, T_true = synth.make_pair(4000, v, 61)
    S, G = orc.preprocess(src_full, v, full_normals=False), orc.preprocess(tgt_full, v, full_normals=False)
    src, tgt = S.pcd_down, G.pcd_down
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    c, max_dist, seed = len(corr), 1.5 * v, 5
    assert 30 < c and len(src) < 1500
    s64, t64 = src.astype(np.float64), tgt.astype(np.float64)
    est_k, best, evaluated, survivors = max_iter, None, 0, 0
    h = 0
    while h < max_iter and h < est_k:
        r = py_philox4x32_10(h, seed)
        ids = [(r[k] * c) >> 32 for k in range(3)]
        X, Y = s64[corr[ids, 0]], t64[corr[ids, 1]]
        h += 1
        evaluated += 1
        ok = True
        for i in range(3):
            for j in range(i + 1, 3):
                ds, dt = np.linalg.norm(X[i] - X[j]), np.linalg.norm(Y[i] - Y[j])
                ok &= not (ds < 0.9 * dt or dt < 0.9 * ds)
        if not ok:
            continue
        T = np_umeyama(X, Y)
        if not np.isfinite(T).all() or (np.linalg.norm(X @ T[:3, :3].T + T[:3, 3] - Y, axis=1) > max_dist).any():
            continue
        survivors += 1
        j, d2 = np_ref.nn1(tgt, np_ref.transform_f32(T, src), max_dist)
        m = j >= 0
        cnt, sumd2 = int(m.sum()), float(d2[m].astype(np.float64).sum())
        if best is None or cnt > best[0] or (cnt == best[0] and cnt > 0 and sumd2 < best[1]):
            best = (cnt, sumd2, h - 1, T)
            p = src[corr[:, 0]].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
            ratio = float((np.linalg.norm(p - t64[corr[:, 1]], axis=1) < max_dist).mean())
            if 0.0 < ratio < 1.0 and conf < 1.0:  # confidence 1.0 consumes every iteration
                est = math.log(1.0 - conf) / math.log(1.0 - ratio ** 3)
                if 0.0 <= est < est_k:
                    est_k = int(math.ceil(est))
            elif ratio >= 1.0 and conf < 1.0:
                est_k = 0
    got = orc.ransac(src, tgt, corr, max_dist, max_iter, conf, seed)
    assert best is not None and survivors >= (3 if conf < 1.0 else 100)
    assert (got.best_hyp, got.hyp_evaluated, got.survivors, got.est_k) == (best[2], evaluated, survivors, est_k)
    assert got.inlier_count == best[0]
    assert np.abs(got.transformation - best[3]).max() < 1e-9
    assert abs(got.inlier_rmse - math.sqrt(best[1] / best[0])) < 1e-9
    assert got.fitness == best[0] / len(src)
