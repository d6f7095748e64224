"""Brute-force NumPy restatements used to cross-check the C oracle on small inputs (independent code path:
no KD-tree, no C).  Arithmetic follows the same determinism rules (fp32 distances in the order
(dx*dx + dy*dy) + dz*dz; ties by index)."""
import numpy as np


def dist2_f32(q, pts):
    q = np.asarray(q, np.float32)
    pts = np.asarray(pts, np.float32)
    d = q[None, :] - pts
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def knn_hybrid(pts, queries, radius, max_nn):
    r2 = np.float32(radius * radius)
    nq = len(queries)
    idx = -np.ones((nq, max_nn), np.int32)
    d2o = np.zeros((nq, max_nn), np.float32)
    cnt = np.zeros(nq, np.int32)
    for i, q in enumerate(queries):
        d2 = dist2_f32(q, pts)
        order = np.lexsort((np.arange(len(pts)), d2))  # by d2, ties by index
        order = order[d2[order] < r2][:max_nn]
        cnt[i] = len(order)
        idx[i, : len(order)] = order
        d2o[i, : len(order)] = d2[order]
    return idx, d2o, cnt


def nn1(tgt, queries, radius):
    idx, d2, cnt = knn_hybrid(tgt, queries, radius, 1)
    return idx[:, 0], d2[:, 0]


def voxel_downsample(pts, voxel):
    pts = np.asarray(pts, np.float32)
    mn, mx = pts.min(0), pts.max(0)
    org = mn.astype(np.float64) - voxel * 0.5
    dims = np.floor((mx.astype(np.float64) - org) / voxel).astype(np.int64) + 1
    c = np.floor((pts.astype(np.float64) - org) / voxel).astype(np.int64)
    key = (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]
    out = []
    for k in np.unique(key):
        out.append(pts[key == k].astype(np.float64).mean(0))
    return np.array(out)


def nn_features(fq, fb):
    fq = np.asarray(fq, np.float32).astype(np.float64)
    fb = np.asarray(fb, np.float32).astype(np.float64)
    out = np.zeros(len(fq), np.int32)
    for i in range(len(fq)):
        acc = np.zeros(len(fb))
        for k in range(33):
            df = fq[i, k] - fb[:, k]
            acc = acc + df * df
        out[i] = int(np.argmin(acc))  # first minimum = lowest index
    return out


def transform_f32(T, pts):
    pts = np.asarray(pts, np.float32).astype(np.float64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = np.empty((len(pts), 3), np.float32)
    for r in range(3):
        out[:, r] = (((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]).astype(np.float32)
    return out
