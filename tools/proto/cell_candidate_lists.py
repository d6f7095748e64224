"""Prototype (NumPy, CPU) of a next-round idea for RANSAC validation: per-fine-cell candidate lists with an exact
certificate, so that a radius-limited nearest-neighbour query tests 3-6 target points instead of walking 27 grid cells
(~20 candidates, ~300 instructions per query in k_ransac_validate today).

For a fine cell C (cube of side c = v / DIV) the candidate set is
    K(C) = { t : dmin(t, C) <= min(min_t' dmax(t', C), r) * (1 + SLACK) }
(dmin / dmax = smallest / largest distance from t to the cube).  For every query q in C the nearest target point lies
in K(C): its distance is at most dist(q, t') <= dmax(t', C) for the minimiser t', and dmin(NN, C) <= dist(q, NN).  The
final choice among K(C) is made with the SAME fp32 arithmetic and (d2, index) tie-break as the full search (rules D1,
D2), so the result is bit-identical; SLACK = 1e-5 covers the fp32 rounding of the distances (<= 2.4e-7 relative) and
of the cell mapping.  This script checks that claim against the CPU oracle's KD-tree search on the bench pair for
near-correct and for bad hypotheses, and prints the list statistics that size the kernel.

    python tools/proto/cell_candidate_lists.py [DIV]
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "3d-matching_b200"), ROOT]
from oracle import pcr_oracle as orc  # noqa: E402  (a prototype's checker, not product code)
from pcr_b200 import synth  # noqa: E402

SLACK = 1e-5


def d2_f32(q, P):
    """Rule D1: fp32 (dx*dx + dy*dy) + dz*dz on fp32 operands."""
    d = q[None, :].astype(np.float32) - P.astype(np.float32)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


class CellLists:
    def __init__(self, tgt32, r, c):
        self.t32 = np.ascontiguousarray(tgt32, np.float32)
        self.t64 = self.t32.astype(np.float64)
        self.tree = cKDTree(self.t64)
        self.r, self.c = float(r), float(c)
        self.r2 = np.float32(r * r)
        self.lo = self.t64.min(0) - r - c
        self.cache = {}

    def cell_of(self, q32):
        return tuple(np.floor((q32.astype(np.float64) - self.lo) / self.c).astype(np.int64))

    def candidates(self, cell):
        """K(C), built lazily here; on the GPU a build kernel fills a table for every cell within r of the cloud."""
        k = self.cache.get(cell)
        if k is not None:
            return k
        half = 0.5 * self.c * (1.0 + 2.0 ** -20)  # the fp64 cell mapping may put q an ulp outside the nominal cube
        ctr = self.lo + (np.array(cell) + 0.5) * self.c
        diag = np.sqrt(3.0) * half
        dn, _ = self.tree.query(ctr)
        idx = np.array(self.tree.query_ball_point(ctr, min(dn + 2 * diag, self.r + diag) * (1 + SLACK) + 1e-300), np.int64)
        if idx.size == 0:
            k = idx
        else:
            a = np.abs(self.t64[idx] - ctr)
            dmin = np.sqrt((np.maximum(a - half, 0.0) ** 2).sum(1))
            dmax = np.sqrt(((a + half) ** 2).sum(1))
            bound = min(dmax.min(), self.r) * (1 + SLACK)
            k = np.sort(idx[dmin <= bound])
        self.cache[cell] = k
        return k

    def nn1(self, q32):
        """(index or -1, d2) exactly as the full radius-limited search returns them."""
        k = self.candidates(self.cell_of(q32))
        if k.size == 0:
            return -1, np.float32(0)
        d2 = d2_f32(q32, self.t32[k])
        j = int(np.lexsort((k, d2))[0])  # ascending (d2, index) — rule D2
        return (int(k[j]), d2[j]) if d2[j] < self.r2 else (-1, np.float32(0))


def main():
    div = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    orc.build()
    v = 0.005
    r = 1.5 * v
    src, tgt, T = synth.make_pair(100000, v, 20241)
    sd, td = orc.voxel_downsample(src, v), orc.voxel_downsample(tgt, v)
    cl = CellLists(td, r, v / div)
    rng = np.random.default_rng(3)
    total = mism = 0
    sizes = []
    for trial in range(6):
        Th = T.copy()
        if trial:  # hypotheses of decreasing quality: small to large perturbations of the true transform
            ang = rng.normal(0, [0.0, 0.002, 0.01, 0.03, 0.1, 0.5][trial], 3)
            Th[:3, :3] = synth.euler_zyx(*ang) @ T[:3, :3]
            Th[:3, 3] += rng.normal(0, [0.0, 0.2, 0.5, 1.0, 3.0, 10.0][trial] * v, 3)
        q = orc.transform_points(Th, sd)
        want_i, want_d = orc.nn1(td, q, r)
        sub = rng.choice(len(q), 3000, replace=False)
        for i in sub:
            gi, gd = cl.nn1(q[i])
            total += 1
            if gi != want_i[i] or (gi >= 0 and np.float32(gd).tobytes() != np.float32(want_d[i]).tobytes()):
                mism += 1
            sizes.append(len(cl.candidates(cl.cell_of(q[i]))))
        print(f"trial {trial}: inliers {np.mean(want_i >= 0):.3f}  mismatches so far {mism} of {total}")
    sizes = np.array(sizes)
    print(f"c = v/{div}: candidates per query: mean {sizes.mean():.2f}, median {np.median(sizes):.0f}, p95 {np.percentile(sizes, 95):.0f}, "
          f"max {sizes.max()}, empty cells {np.mean(sizes == 0):.3f}; distinct cells touched {len(cl.cache)}")
    assert mism == 0, "candidate lists changed a nearest neighbour"
    print("exact: every index and every fp32 d2 equals the oracle's full search")


if __name__ == "__main__":
    main()
