"""ctypes front-end of the CPU oracle (oracle/pcr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (3d-matching_b200/) never does.

All point arrays are (n,3) float32 C-contiguous; features (n,33) float32; correspondences (c,2) int32;
transforms (4,4) float64.  Function-by-function citations of the reference are in pcr_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpcr_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the path of the shared library."""
    src = os.path.join(_HERE, "pcr_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "pcr_detmath.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


class _RansacResult(C.Structure):
    _fields_ = [
        ("T", C.c_double * 16),
        ("fitness", C.c_double),
        ("inlier_rmse", C.c_double),
        ("best_hyp", C.c_longlong),
        ("inlier_count", C.c_longlong),
        ("sum_d2_fixed", C.c_longlong),
        ("k_d", C.c_int),
        ("hyp_evaluated", C.c_longlong),
        ("survivors", C.c_longlong),
        ("est_k", C.c_longlong),
    ]


class _IcpResult(C.Structure):
    _fields_ = [
        ("T", C.c_double * 16),
        ("fitness", C.c_double),
        ("inlier_rmse", C.c_double),
        ("inlier_count", C.c_longlong),
        ("sum_d2_fixed", C.c_longlong),
        ("k_d", C.c_int),
        ("iterations", C.c_int),
        ("converged", C.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_kabsch3.restype = None
        _lib.orc_philox.restype = None
        _lib.orc_detmath.restype = None
        _lib.orc_fast_eigen3x3.restype = None
        _lib.orc_pair_features.restype = None
        _lib.orc_transform_points.restype = None
        _lib.orc_set_num_threads.restype = None
    return _lib


def _f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected (n,{cols}) array, got {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def voxel_downsample(pts, voxel: float) -> np.ndarray:
    pts = _f32(pts, 3)
    out = np.empty_like(pts)
    m = C.c_int(0)
    rc = lib().orc_voxel_downsample(_p(pts), C.c_int(len(pts)), C.c_double(voxel), _p(out), C.byref(m))
    if rc != 0:
        raise ValueError("voxel_size <= 0 or voxel grid too large")
    return out[: m.value].copy()


def knn_hybrid(pts, queries, radius: float, max_nn: int):
    pts, queries = _f32(pts, 3), _f32(queries, 3)
    nq = len(queries)
    idx = np.empty((nq, max_nn), np.int32)
    d2 = np.empty((nq, max_nn), np.float32)
    cnt = np.empty(nq, np.int32)
    lib().orc_knn_hybrid(_p(pts), C.c_int(len(pts)), _p(queries), C.c_int(nq), C.c_double(radius), C.c_int(max_nn),
                         _p(idx), _p(d2), _p(cnt))
    return idx, d2, cnt


def nn1(tgt, queries, radius: float):
    tgt, queries = _f32(tgt, 3), _f32(queries, 3)
    idx = np.empty(len(queries), np.int32)
    d2 = np.empty(len(queries), np.float32)
    lib().orc_nn1(_p(tgt), C.c_int(len(tgt)), _p(queries), C.c_int(len(queries)), C.c_double(radius), _p(idx), _p(d2))
    return idx, d2


def estimate_normals(pts, radius: float, max_nn: int) -> np.ndarray:
    pts = _f32(pts, 3)
    out = np.empty_like(pts)
    lib().orc_estimate_normals(_p(pts), C.c_int(len(pts)), C.c_double(radius), C.c_int(max_nn), _p(out))
    return out


def fpfh(pts, normals, radius: float, max_nn: int) -> np.ndarray:
    pts, normals = _f32(pts, 3), _f32(normals, 3)
    out = np.empty((len(pts), 33), np.float32)
    lib().orc_fpfh(_p(pts), _p(normals), C.c_int(len(pts)), C.c_double(radius), C.c_int(max_nn), _p(out))
    return out


def nn_features(fq, fb) -> np.ndarray:
    fq, fb = _f32(fq, 33), _f32(fb, 33)
    nn = np.empty(len(fq), np.int32)
    lib().orc_nn_features(_p(fq), C.c_int(len(fq)), _p(fb), C.c_int(len(fb)), _p(nn))
    return nn


def match_features(fs, ft, mutual: bool = False, mutual_ratio: float = 0.1) -> np.ndarray:
    fs, ft = _f32(fs, 33), _f32(ft, 33)
    corr = np.empty((max(len(fs), 1), 2), np.int32)
    c = C.c_int(0)
    lib().orc_match_features(_p(fs), C.c_int(len(fs)), _p(ft), C.c_int(len(ft)), C.c_int(int(mutual)),
                             C.c_double(mutual_ratio), _p(corr), C.byref(c))
    return corr[: c.value].copy()


def philox(ctr_lo: int, ctr_hi: int, key: int) -> np.ndarray:
    out = np.empty(4, np.uint32)
    lib().orc_philox(C.c_uint64(ctr_lo), C.c_uint64(ctr_hi), C.c_uint64(key), _p(out))
    return out


def kabsch3(src3, tgt3) -> np.ndarray:
    s = np.ascontiguousarray(src3, np.float64).reshape(3, 3)
    t = np.ascontiguousarray(tgt3, np.float64).reshape(3, 3)
    T = np.empty((4, 4), np.float64)
    lib().orc_kabsch3(_p(s), _p(t), _p(T))
    return T


def ransac_step(src, tgt, corr, seed: int, h: int):
    src, tgt = _f32(src, 3), _f32(tgt, 3)
    corr = np.ascontiguousarray(corr, np.int32).reshape(-1, 2)
    T = np.empty((4, 4), np.float64)
    smp = np.zeros(3, np.int32)
    lib().orc_ransac_step(_p(src), _p(tgt), _p(corr), C.c_int(len(corr)), C.c_uint64(seed), C.c_uint64(h), _p(T), _p(smp))
    return T, smp


def inlier_count(src, tgt, corr, T, thresh: float, squared: bool = False) -> int:
    src, tgt = _f32(src, 3), _f32(tgt, 3)
    corr = np.ascontiguousarray(corr, np.int32).reshape(-1, 2)
    T = np.ascontiguousarray(T, np.float64)
    fn = lib().orc_inlier_count_sq if squared else lib().orc_inlier_count
    return int(fn(_p(src), _p(tgt), _p(corr), C.c_int(len(corr)), _p(T), C.c_double(thresh)))


@dataclass
class RansacResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    best_hyp: int
    inlier_count: int
    sum_d2_fixed: int
    k_d: int
    hyp_evaluated: int
    survivors: int
    est_k: int


def ransac(src, tgt, corr, max_dist: float, max_iter: int, confidence: float = 0.999, seed: int = 0,
           edge_sim: float = 0.9) -> RansacResult:
    src, tgt = _f32(src, 3), _f32(tgt, 3)
    corr = np.ascontiguousarray(corr, np.int32).reshape(-1, 2)
    r = _RansacResult()
    lib().orc_ransac(_p(src), C.c_int(len(src)), _p(tgt), C.c_int(len(tgt)), _p(corr), C.c_int(len(corr)),
                     C.c_double(max_dist), C.c_double(edge_sim), C.c_longlong(max_iter), C.c_double(confidence),
                     C.c_uint64(seed), C.byref(r))
    return RansacResult(np.array(r.T, np.float64).reshape(4, 4), r.fitness, r.inlier_rmse, r.best_hyp, r.inlier_count,
                        r.sum_d2_fixed, r.k_d, r.hyp_evaluated, r.survivors, r.est_k)


def ransac_eval_one(src, tgt, corr, max_dist: float, seed: int, h: int, edge_sim: float = 0.9):
    """(T, count, sum_d2_fixed, corr_inliers) of hypothesis h, or None when a checker rejects it."""
    src, tgt = _f32(src, 3), _f32(tgt, 3)
    corr = np.ascontiguousarray(corr, np.int32).reshape(-1, 2)
    T = np.empty((4, 4), np.float64)
    cnt, sq, ci = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
    ok = lib().orc_ransac_eval_one(_p(src), C.c_int(len(src)), _p(tgt), C.c_int(len(tgt)), _p(corr), C.c_int(len(corr)),
                                   C.c_double(max_dist), C.c_double(edge_sim), C.c_uint64(seed), C.c_longlong(h), _p(T),
                                   C.byref(cnt), C.byref(sq), C.byref(ci))
    return (T, cnt.value, sq.value, ci.value) if ok else None


@dataclass
class IcpResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    inlier_count: int
    sum_d2_fixed: int
    k_d: int
    iterations: int
    converged: bool
    correspondence: np.ndarray = field(repr=False, default=None)


def icp_point_to_plane(src, tgt, tgt_normals, max_dist: float, init=None, max_iter: int = 30,
                       rel_fitness: float = 1e-6, rel_rmse: float = 1e-6) -> IcpResult:
    src, tgt, tn = _f32(src, 3), _f32(tgt, 3), _f32(tgt_normals, 3)
    T0 = np.ascontiguousarray(np.eye(4) if init is None else init, np.float64)
    r = _IcpResult()
    corr = np.empty(len(src), np.int32)
    rc = lib().orc_icp_point_to_plane(_p(src), C.c_int(len(src)), _p(tgt), _p(tn), C.c_int(len(tgt)), C.c_double(max_dist),
                                      _p(T0), C.c_int(max_iter), C.c_double(rel_fitness), C.c_double(rel_rmse),
                                      C.byref(r), _p(corr))
    if rc != 0:
        raise ValueError("max_correspondence_distance must be > 0")
    return IcpResult(np.array(r.T, np.float64).reshape(4, 4), r.fitness, r.inlier_rmse, r.inlier_count, r.sum_d2_fixed,
                     r.k_d, r.iterations, bool(r.converged), corr)


def detmath(x: float, y: float = 0.0) -> np.ndarray:
    """[sin(x), cos(x), atan2(y,x), acos(x)] from include/pcr_detmath.h."""
    out = np.empty(4, np.float64)
    lib().orc_detmath(C.c_double(x), C.c_double(y), _p(out))
    return out


def fast_eigen3x3(c6) -> np.ndarray:
    c6 = np.ascontiguousarray(c6, np.float64)
    out = np.empty(3, np.float64)
    lib().orc_fast_eigen3x3(_p(c6), _p(out))
    return out


def pair_features(p1, n1, p2, n2) -> np.ndarray:
    a = [np.ascontiguousarray(v, np.float32) for v in (p1, n1, p2, n2)]
    out = np.empty(4, np.float64)
    lib().orc_pair_features(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(out))
    return out


def transform_points(T, pts) -> np.ndarray:
    pts = _f32(pts, 3)
    T = np.ascontiguousarray(T, np.float64)
    out = np.empty_like(pts)
    lib().orc_transform_points(_p(T), _p(pts), C.c_int(len(pts)), _p(out))
    return out


# ---- whole-pipeline restatement of the reference's call sequence (used as the CPU baseline) ------

@dataclass
class OraclePly:
    """Mirror of the reference's Ply container (src/ply/ply.py:20-66) over oracle arrays."""
    pcd: np.ndarray
    normals: np.ndarray
    pcd_down: np.ndarray
    down_normals: np.ndarray
    pcd_fpfh: np.ndarray
    voxel_size: float


def preprocess(pts, voxel: float, noise_sigma: float = 0.0, rng=None, full_normals: bool = True) -> OraclePly:
    """Ply.__init__ (src/ply/ply.py:53-65): voxel -> normals(2v,30) -> FPFH(5v,100) -> noise -> full normals."""
    pts = _f32(pts, 3)
    down = voxel_downsample(pts, voxel)
    dn = estimate_normals(down, 2.0 * voxel, 30)
    f = fpfh(down, dn, 5.0 * voxel, 100)
    if noise_sigma > 0.0:
        rng = rng or np.random.default_rng(0)
        down = (down.astype(np.float64) + noise_sigma * rng.standard_normal(down.shape)).astype(np.float32)
    fn = estimate_normals(pts, 2.0 * voxel, 30) if full_normals else None
    return OraclePly(pts, fn, down, dn, f, voxel)


def global_registration(src: OraclePly, tgt: OraclePly, voxel: float, iteration: int = 30, confidence: float = 0.999,
                        seed: int = 0) -> RansacResult:
    """src/matcher/ransac.py:20-59."""
    corr = match_features(src.pcd_fpfh, tgt.pcd_fpfh, mutual=True)
    return ransac(src.pcd_down, tgt.pcd_down, corr, 1.5 * voxel, iteration, confidence, seed)


def refine_registration(src: OraclePly, tgt: OraclePly, init, voxel: float, max_iter: int = 30,
                        rel_fitness: float = 1e-6, rel_rmse: float = 1e-6) -> IcpResult:
    """src/matcher/icp.py:17-48."""
    return icp_point_to_plane(src.pcd, tgt.pcd, tgt.normals, 0.4 * voxel, init, max_iter, rel_fitness, rel_rmse)
