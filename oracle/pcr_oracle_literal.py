"""ctypes front-end of oracle/pcr_oracle_literal.c — the Appendix-A-LITERAL fp64 restatement (libm, double distances,
incremental ICP transform, running FPFH normaliser, SVD Umeyama, pivoted LDL^T).  TEST INFRASTRUCTURE ONLY: imported by
tests/test_oracle_literal.py to bound what the determinism rules D1-D9 of the shared specification changed.

All arrays are float64: points (n,3), features (n,33), transforms (4,4); correspondences (c,2) int32."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpcr_oracle_literal.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "pcr_oracle_literal.c")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        _lib.lit_umeyama.restype = None
    return _lib


def _f64(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected (n,{cols}) array, got {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def voxel_downsample(pts, voxel: float) -> np.ndarray:
    pts = _f64(pts, 3)
    out = np.empty_like(pts)
    m = C.c_int(0)
    if lib().lit_voxel_downsample(_p(pts), C.c_int(len(pts)), C.c_double(voxel), _p(out), C.byref(m)) != 0:
        raise ValueError("voxel_size must be > 0")
    return out[: m.value].copy()


def estimate_normals(pts, radius: float, max_nn: int, want_neighbours: bool = False):
    pts = _f64(pts, 3)
    out = np.empty_like(pts)
    nb = np.empty((len(pts), max_nn), np.int32) if want_neighbours else None
    lib().lit_estimate_normals(_p(pts), C.c_int(len(pts)), C.c_double(radius), C.c_int(max_nn), _p(out),
                               _p(nb) if want_neighbours else None)
    return (out, nb) if want_neighbours else out


def fpfh(pts, normals, radius: float, max_nn: int) -> np.ndarray:
    pts, normals = _f64(pts, 3), _f64(normals, 3)
    out = np.empty((len(pts), 33), np.float64)
    lib().lit_fpfh(_p(pts), _p(normals), C.c_int(len(pts)), C.c_double(radius), C.c_int(max_nn), _p(out))
    return out


def nn_features(fq, fb) -> np.ndarray:
    fq, fb = _f64(fq, 33), _f64(fb, 33)
    nn = np.empty(len(fq), np.int32)
    lib().lit_nn_features(_p(fq), C.c_int(len(fq)), _p(fb), C.c_int(len(fb)), _p(nn))
    return nn


def match_features(fs, ft, mutual: bool = False, ratio: float = 0.1) -> np.ndarray:
    fs, ft = _f64(fs, 33), _f64(ft, 33)
    corr = np.empty((len(fs), 2), np.int32)
    c = C.c_int(0)
    lib().lit_match_features(_p(fs), C.c_int(len(fs)), _p(ft), C.c_int(len(ft)), C.c_int(int(mutual)), C.c_double(ratio),
                             _p(corr), C.byref(c))
    return corr[: c.value].copy()


def umeyama(src3, tgt3) -> np.ndarray:
    s, t = _f64(src3, 3), _f64(tgt3, 3)
    T = np.empty((4, 4), np.float64)
    lib().lit_umeyama(_p(s), _p(t), _p(T))
    return T


class _Ransac(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double), ("best_hyp", C.c_longlong),
                ("inlier_count", C.c_longlong), ("hyp_evaluated", C.c_longlong), ("survivors", C.c_longlong),
                ("est_k", C.c_longlong)]


@dataclass
class RansacResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    best_hyp: int
    inlier_count: int
    hyp_evaluated: int
    survivors: int
    est_k: int


def ransac(src, tgt, corr, max_dist: float, max_iter: int, confidence: float = 0.999, seed: int = 0,
           edge_sim: float = 0.9) -> RansacResult:
    src, tgt = _f64(src, 3), _f64(tgt, 3)
    corr = np.ascontiguousarray(corr, np.int32).reshape(-1, 2)
    r = _Ransac()
    lib().lit_ransac(_p(src), C.c_int(len(src)), _p(tgt), C.c_int(len(tgt)), _p(corr), C.c_int(len(corr)),
                     C.c_double(max_dist), C.c_double(edge_sim), C.c_longlong(max_iter), C.c_double(confidence),
                     C.c_uint64(seed), C.byref(r))
    return RansacResult(np.array(r.T, np.float64).reshape(4, 4), r.fitness, r.inlier_rmse, r.best_hyp, r.inlier_count,
                        r.hyp_evaluated, r.survivors, r.est_k)


class _Icp(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double), ("inlier_count", C.c_longlong),
                ("iterations", C.c_int), ("converged", C.c_int)]


@dataclass
class IcpResult:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    inlier_count: int
    iterations: int
    converged: bool
    correspondence: np.ndarray = field(repr=False, default=None)


def icp_point_to_plane(src, tgt, tgt_normals, max_dist: float, init=None, max_iter: int = 30, rel_fitness: float = 1e-6,
                       rel_rmse: float = 1e-6) -> IcpResult:
    src, tgt, tn = _f64(src, 3), _f64(tgt, 3), _f64(tgt_normals, 3)
    T0 = np.ascontiguousarray(np.eye(4) if init is None else init, np.float64)
    r = _Icp()
    corr = np.empty(len(src), np.int32)
    if lib().lit_icp_point_to_plane(_p(src), C.c_int(len(src)), _p(tgt), _p(tn), C.c_int(len(tgt)), C.c_double(max_dist),
                                    _p(T0), C.c_int(max_iter), C.c_double(rel_fitness), C.c_double(rel_rmse), C.byref(r),
                                    _p(corr)) != 0:
        raise ValueError("max_correspondence_distance must be > 0")
    return IcpResult(np.array(r.T, np.float64).reshape(4, 4), r.fitness, r.inlier_rmse, r.inlier_count, r.iterations,
                     bool(r.converged), corr)


def nn1(tgt, queries, radius: float) -> np.ndarray:
    tgt, q = _f64(tgt, 3), _f64(queries, 3)
    idx = np.empty(len(q), np.int32)
    lib().lit_nn1(_p(tgt), C.c_int(len(tgt)), _p(q), C.c_int(len(q)), C.c_double(radius), _p(idx))
    return idx
