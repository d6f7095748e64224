"""Generate tests/golden/ransac_numpy_golden.npz by running the REFERENCE's own NumPy functions.

Runs only in the development container (needs /root/reference).  open3d is absent there, so a ~20-line
stub supplies the three container types the reference touches (Vector2iVector / Vector3dVector as ndarray
casts, RegistrationResult / Feature as empty classes) — no arithmetic is stubbed: compute_step_transformation,
evaluate_inlier_ratio and evaluate_inlier_ratio_fast (src/matcher/ransac.py:104-277) run unmodified.

    python tests/golden/make_ransac_numpy_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"


def install_stub():
    o3d = types.ModuleType("open3d")
    util = types.ModuleType("open3d.utility")
    util.Vector2iVector = lambda a: np.asarray(a, dtype=np.int32)
    util.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64)
    pipelines = types.ModuleType("open3d.pipelines")
    reg = types.ModuleType("open3d.pipelines.registration")

    class RegistrationResult:
        def __init__(self):
            self.transformation = np.eye(4)
            self.fitness = 0.0
            self.inlier_rmse = 0.0
            self.correspondence_set = np.zeros((0, 2), np.int32)

    class Feature:
        pass

    reg.RegistrationResult = RegistrationResult
    reg.Feature = Feature
    pipelines.registration = reg
    o3d.utility = util
    o3d.pipelines = pipelines
    o3d.geometry = types.ModuleType("open3d.geometry")
    o3d.io = types.ModuleType("open3d.io")
    for name, mod in (("open3d", o3d), ("open3d.utility", util), ("open3d.pipelines", pipelines),
                      ("open3d.pipelines.registration", reg), ("open3d.geometry", o3d.geometry), ("open3d.io", o3d.io)):
        sys.modules[name] = mod
    # `from ply import Ply` in the reference imports open3d-dependent code; give it a trivial module
    ply = types.ModuleType("ply")

    class Ply:
        pass

    ply.Ply = Ply
    sys.modules["ply"] = ply


class Cloud:
    def __init__(self, pts):
        self.points = pts


class MockPly:  # the reference's own duck type (test_ransac_crash.py:92-96)
    def __init__(self, pts):
        self.pcd = Cloud(pts)
        self.pcd_down = Cloud(pts)
        self.pcd_fpfh = None


def euler(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
    return rz @ ry @ rx


def main():
    install_stub()
    sys.path.insert(0, REF)
    from matcher.ransac import compute_step_transformation, evaluate_inlier_ratio, evaluate_inlier_ratio_fast

    rng = np.random.default_rng(12345)
    out = {}
    n_cases = 24
    for k in range(n_cases):
        n = int(rng.integers(8, 200))
        src = rng.uniform(-1, 1, (n, 3)).astype(np.float32).astype(np.float64)  # fp32-representable (rule D1)
        R = euler(*rng.uniform(-np.pi, np.pi, 3))
        t = rng.uniform(-2, 2, 3)
        noise = rng.normal(0, 0.02 if k % 3 else 0.0, (n, 3))
        tgt = ((src @ R.T + t) + noise).astype(np.float32).astype(np.float64)
        c = int(rng.integers(3, 3 * n))
        corr = np.stack([rng.integers(0, n, c), rng.integers(0, n, c)], axis=1).astype(np.int32)
        if k % 2 == 0:  # half of the cases: mostly true correspondences
            corr[:, 1] = corr[:, 0]
        seed = 1000 + k
        np.random.seed(seed)
        idxs = np.random.choice(c, 3, replace=False)  # what src/matcher/ransac.py:143 will draw
        np.random.seed(seed)
        res = compute_step_transformation(MockPly(src), MockPly(tgt), corr)
        T = np.asarray(res.transformation, np.float64)
        voxel = 0.05
        ratio = evaluate_inlier_ratio(MockPly(src), MockPly(tgt), corr, T, voxel)
        thr = voxel * 1.5
        ratio_fast = evaluate_inlier_ratio_fast(src[corr[:, 0]], tgt[corr[:, 1]], T, thr * thr)
        out[f"src_{k}"] = src.astype(np.float32)
        out[f"tgt_{k}"] = tgt.astype(np.float32)
        out[f"corr_{k}"] = corr
        out[f"idx_{k}"] = idxs.astype(np.int32)
        out[f"T_{k}"] = T
        out[f"ratio_{k}"] = np.float64(ratio)
        out[f"ratio_fast_{k}"] = np.float64(ratio_fast)
        out[f"voxel_{k}"] = np.float64(voxel)
    out["n_cases"] = np.int32(n_cases)

    # soft-failure behaviours pinned by test_ransac_crash.py
    col = np.array([[0, 0, i] for i in range(10)], dtype=np.float64)
    dup = np.array([[1, 1, 1]] * 10, dtype=np.float64)
    ident_corr = np.stack([np.arange(10), np.arange(10)], 1).astype(np.int32)
    np.random.seed(5)
    out["T_collinear"] = np.asarray(compute_step_transformation(MockPly(col), MockPly(col), ident_corr).transformation)
    np.random.seed(5)
    out["T_duplicate"] = np.asarray(compute_step_transformation(MockPly(dup), MockPly(dup), ident_corr).transformation)
    two = ident_corr[:2]
    out["T_two_corr"] = np.asarray(compute_step_transformation(MockPly(col), MockPly(col), two).transformation)
    out["ratio_empty"] = np.float64(evaluate_inlier_ratio(MockPly(col), MockPly(col), np.zeros((0, 2), np.int32), np.eye(4), 0.05))
    out["ratio_fast_empty"] = np.float64(evaluate_inlier_ratio_fast(np.zeros((0, 3)), np.zeros((0, 3)), np.eye(4), 0.01))
    big = np.eye(4) * 1000.0
    big[3, 3] = 1.0
    rs = rng.uniform(0, 1, (50, 3)).astype(np.float32).astype(np.float64)
    cc = np.stack([np.arange(50), np.arange(50)], 1).astype(np.int32)
    out["huge_pts"] = rs.astype(np.float32)
    out["ratio_huge"] = np.float64(evaluate_inlier_ratio(MockPly(rs), MockPly(rs), cc, big, 0.05))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ransac_numpy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("collinear ->", np.allclose(out["T_collinear"], np.eye(4)), " duplicate ->", np.allclose(out["T_duplicate"], np.eye(4)))


if __name__ == "__main__":
    main()
