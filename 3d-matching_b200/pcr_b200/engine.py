"""Engine: thin Python host layer over the C ABI (include/pcr.h).

PyTorch supplies device memory (torch tensors) and the CUDA stream; every computation happens in
libpcr_b200.so.  Clouds handed to the engine are packed float4 CUDA tensors of shape (n, 4) fp32 (x, y, z, 0);
`Engine.pack` converts (n, 3) fp32/fp64 host or device arrays.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@dataclass
class DeviceRegResult:
    """Plain-data view of pcr_reg_result."""
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    inlier_count: int
    sum_d2_fixed: int
    k_d: int
    iterations: int
    converged: bool
    best_hyp: int
    hyp_evaluated: int
    survivors: int
    est_k: int

    @staticmethod
    def from_c(r: _capi.RegResult) -> "DeviceRegResult":
        return DeviceRegResult(np.array(r.transformation, np.float64).reshape(4, 4), r.fitness, r.inlier_rmse,
                               r.inlier_count, r.sum_d2_fixed, r.k_d, r.iterations, bool(r.converged), r.best_hyp,
                               r.hyp_evaluated, r.survivors, r.est_k)


class Engine:
    """One context per (process, device).  Not re-entrant: calls on one Engine are serialised by a lock
    (the reference's GUI calls the matcher from worker threads, _visualize_matcher.py:264,275,292)."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("pcr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _capi.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.ctx = C.c_void_p(0)
        rc = self.lib.pcr_create(C.c_int(self.device), C.byref(self.ctx))
        if rc != 0:
            raise RuntimeError(f"pcr_create failed with {rc} (needs an sm_100 device)")
        self._lock = threading.Lock()
        self.tdev = torch.device("cuda", self.device)

    def close(self):
        if self.ctx:
            self.lib.pcr_destroy(self.ctx)
            self.ctx = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------------------------
    def _bind_stream(self):
        s = torch.cuda.current_stream(self.tdev).cuda_stream
        self.lib.pcr_set_stream(self.ctx, C.c_void_p(s))

    def _check(self, rc):
        _capi.check(self.lib, self.ctx, rc)

    def launch_count(self) -> int:
        return int(self.lib.pcr_launch_count(self.ctx))

    def set_profiling(self, enabled: bool) -> None:
        self.lib.pcr_set_profiling(self.ctx, C.c_int(int(enabled)))

    def kernel_stats(self, reset: bool = True) -> dict:
        """{class name: {ms, launches, bytes, flops}} accumulated since the last reset (synchronises the stream)."""
        n = int(self.lib.pcr_kernel_class_count())
        arr = (_capi.KernelStat * n)()
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_kernel_stats(self.ctx, arr, C.c_int(n), C.c_int(int(reset))))
        return {self.lib.pcr_kernel_class_name(i).decode(): {"ms": arr[i].total_ms, "overlapped_ms": arr[i].overlapped_ms, "launches": arr[i].launches,
                                                              "bytes": arr[i].bytes, "flops": arr[i].flops}
                for i in range(n) if arr[i].launches}

    def pack(self, xyz) -> torch.Tensor:
        """(n,3) or (n,4) fp32/fp64 numpy / torch (host or device) -> (n,4) fp32 CUDA tensor (quantised, D1)."""
        if isinstance(xyz, torch.Tensor):
            t = xyz
        else:
            a = np.asarray(xyz)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(a))
        if t.ndim != 2 or t.shape[1] not in (3, 4):
            raise ValueError(f"expected an (n,3) point array, got {tuple(t.shape)}")
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        t = t.to(self.tdev, non_blocking=True).contiguous()
        if t.shape[1] == 4 and t.dtype == torch.float32:
            return t
        if t.shape[1] == 4:
            t = t[:, :3].contiguous()
        n = t.shape[0]
        out = torch.empty((n, 4), dtype=torch.float32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            fn = self.lib.pcr_pack_xyz_f32 if t.dtype == torch.float32 else self.lib.pcr_pack_xyz_f64
            self._check(fn(self.ctx, _ptr(t), C.c_int(n), _ptr(out)))
        return out

    def transform_points(self, xyzw: torch.Tensor, T) -> torch.Tensor:
        """out = fp32(R p + t) in the specified fp64 operation order (rule D7)."""
        self._need_xyzw(xyzw, "points")
        T = np.ascontiguousarray(np.asarray(T, np.float64).reshape(4, 4))
        out = torch.empty_like(xyzw)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_transform_points(self.ctx, _ptr(xyzw), C.c_int(xyzw.shape[0]),
                                                      T.ctypes.data_as(C.c_void_p), _ptr(out)))
        return out

    @staticmethod
    def _need_xyzw(t: torch.Tensor, name: str):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.ndim == 2
                and t.shape[1] == 4 and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous (n,4) fp32 CUDA tensor (use Engine.pack)")

    # ---- preprocessing ---------------------------------------------------------------------------------
    def voxel_downsample(self, xyzw: torch.Tensor, voxel: float) -> torch.Tensor:
        self._need_xyzw(xyzw, "points")
        n = xyzw.shape[0]
        out = torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.tdev)
        m = C.c_int(0)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_voxel_downsample(self.ctx, _ptr(xyzw), C.c_int(n), C.c_double(voxel), _ptr(out),
                                                      C.byref(m)))
        return out[: m.value]

    def estimate_normals(self, xyzw: torch.Tensor, radius: float, max_nn: int) -> torch.Tensor:
        self._need_xyzw(xyzw, "points")
        n = xyzw.shape[0]
        out = torch.empty((n, 4), dtype=torch.float32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_estimate_normals(self.ctx, _ptr(xyzw), C.c_int(n), C.c_double(radius),
                                                      C.c_int(max_nn), _ptr(out)))
        return out

    def compute_fpfh(self, xyzw: torch.Tensor, normals: torch.Tensor, radius: float, max_nn: int) -> torch.Tensor:
        self._need_xyzw(xyzw, "points")
        self._need_xyzw(normals, "normals")
        n = xyzw.shape[0]
        out = torch.empty((n, 33), dtype=torch.float32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_compute_fpfh(self.ctx, _ptr(xyzw), _ptr(normals), C.c_int(n), C.c_double(radius),
                                                  C.c_int(max_nn), _ptr(out)))
        return out

    def knn_hybrid(self, xyzw: torch.Tensor, queries: torch.Tensor, radius: float, max_nn: int):
        self._need_xyzw(xyzw, "points")
        self._need_xyzw(queries, "queries")
        nq = queries.shape[0]
        idx = torch.empty((nq, max_nn), dtype=torch.int32, device=self.tdev)
        d2 = torch.empty((nq, max_nn), dtype=torch.float32, device=self.tdev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_knn_hybrid(self.ctx, _ptr(xyzw), C.c_int(xyzw.shape[0]), _ptr(queries), C.c_int(nq),
                                                C.c_double(radius), C.c_int(max_nn), _ptr(idx), _ptr(d2), _ptr(cnt)))
        return idx, d2, cnt

    def nn1(self, tgt: torch.Tensor, queries: torch.Tensor, radius: float):
        self._need_xyzw(tgt, "target")
        self._need_xyzw(queries, "queries")
        nq = queries.shape[0]
        idx = torch.empty((nq,), dtype=torch.int32, device=self.tdev)
        d2 = torch.empty((nq,), dtype=torch.float32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_nn1(self.ctx, _ptr(tgt), C.c_int(tgt.shape[0]), _ptr(queries), C.c_int(nq),
                                         C.c_double(radius), _ptr(idx), _ptr(d2)))
        return idx, d2

    # ---- feature matching ------------------------------------------------------------------------------
    @staticmethod
    def _need_feat(t, name):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.ndim == 2
                and t.shape[1] == 33 and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous (n,33) fp32 CUDA tensor")

    def nn_features(self, fq: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
        self._need_feat(fq, "query features")
        self._need_feat(fb, "base features")
        nn = torch.empty((fq.shape[0],), dtype=torch.int32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_nn_features(self.ctx, _ptr(fq), C.c_int(fq.shape[0]), _ptr(fb),
                                                 C.c_int(fb.shape[0]), _ptr(nn)))
        return nn

    def match_features(self, fs: torch.Tensor, ft: torch.Tensor, mutual: bool = False,
                       mutual_ratio: float = 0.1) -> torch.Tensor:
        self._need_feat(fs, "source features")
        self._need_feat(ft, "target features")
        ms = fs.shape[0]
        corr = torch.empty((max(ms, 1), 2), dtype=torch.int32, device=self.tdev)
        c = C.c_int(0)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_match_features(self.ctx, _ptr(fs), C.c_int(ms), _ptr(ft), C.c_int(ft.shape[0]),
                                                    C.c_int(int(bool(mutual))), C.c_double(mutual_ratio), _ptr(corr),
                                                    C.byref(c)))
        return corr[: c.value]

    # ---- RANSAC ----------------------------------------------------------------------------------------
    @staticmethod
    def _need_corr(corr):
        if not (isinstance(corr, torch.Tensor) and corr.is_cuda and corr.dtype == torch.int32 and corr.ndim == 2
                and corr.shape[1] == 2 and corr.is_contiguous()):
            raise ValueError("correspondences must be a contiguous (c,2) int32 CUDA tensor")

    def ransac(self, src: torch.Tensor, tgt: torch.Tensor, corr: torch.Tensor, max_dist: float, max_iter: int,
               confidence: float = 0.999, seed: int = 0, edge_sim: float = 0.9) -> DeviceRegResult:
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_corr(corr)
        r = _capi.RegResult()
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_ransac(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt), C.c_int(tgt.shape[0]),
                                            _ptr(corr), C.c_int(corr.shape[0]), C.c_double(max_dist),
                                            C.c_double(edge_sim), C.c_int64(max_iter), C.c_double(confidence),
                                            C.c_uint64(seed), C.byref(r)))
        return DeviceRegResult.from_c(r)

    def ransac_wave(self, src, tgt, corr, max_dist: float, hyp_begin: int, hyp_end: int, seed: int = 0,
                    edge_sim: float = 0.9, cap: int = 4096, best_count: int = 0, best_sum: int = 0):
        """Score hypotheses [hyp_begin, hyp_end); returns (records ctypes array, n_records, n_survivors)."""
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_corr(corr)
        recs = (_capi.HypRecord * cap)()
        n = C.c_int(0)
        ns = C.c_int64(0)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_ransac_wave(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt),
                                                 C.c_int(tgt.shape[0]), _ptr(corr), C.c_int(corr.shape[0]),
                                                 C.c_double(max_dist), C.c_double(edge_sim), C.c_int64(hyp_begin),
                                                 C.c_int64(hyp_end), C.c_uint64(seed), C.c_int64(best_count),
                                                 C.c_int64(best_sum), recs, C.c_int(cap), C.byref(n), C.byref(ns)))
        return recs, n.value, ns.value

    def ransac_session_begin(self, src, tgt, max_dist: float) -> None:
        """Build the target grid / sorted source once for a series of ransac_wave calls on these clouds."""
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_ransac_session_begin(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt),
                                                          C.c_int(tgt.shape[0]), C.c_double(max_dist)))

    def ransac_session_end(self) -> None:
        with self._lock:
            self._check(self.lib.pcr_ransac_session_end(self.ctx))

    def ransac_step(self, src, tgt, corr, seed: int, h_begin: int, count: int) -> torch.Tensor:
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_corr(corr)
        T = torch.empty((count, 4, 4), dtype=torch.float64, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_ransac_step(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt), C.c_int(tgt.shape[0]),
                                                 _ptr(corr), C.c_int(corr.shape[0]),
                                                 C.c_uint64(seed), C.c_int64(h_begin), C.c_int(count), _ptr(T)))
        return T

    def inlier_count(self, src, tgt, corr, T: torch.Tensor, thresh: float, squared: bool = False) -> torch.Tensor:
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_corr(corr)
        T = T.to(self.tdev, torch.float64).reshape(-1, 4, 4).contiguous()
        out = torch.empty((T.shape[0],), dtype=torch.int32, device=self.tdev)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_inlier_count(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt), C.c_int(tgt.shape[0]),
                                                  _ptr(corr), C.c_int(corr.shape[0]),
                                                  _ptr(T), C.c_int(T.shape[0]), C.c_double(thresh),
                                                  C.c_int(int(squared)), _ptr(out)))
        return out

    # ---- multi-GPU on the C side (pcr_dist.cu): NCCL communicator owned by the context ------------------------------
    def comm_init(self, group=None) -> int:
        """Create the context's NCCL communicator over the ranks of the (already initialised) torch.distributed group:
        rank 0 draws the id, torch.distributed only carries its 128 bytes.  Returns the world size (1: nothing to do)."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            with self._lock:
                self._check(self.lib.pcr_comm_init(self.ctx, None, C.c_int(0), C.c_int(0), C.c_int(1)))
            return 1
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        idb = (C.c_ubyte * 128)()
        if rank == 0:
            rc = self.lib.pcr_comm_unique_id(idb, C.c_int(128))
            if rc != 0:
                raise RuntimeError(f"pcr_comm_unique_id failed ({rc}): is libnccl.so.2 loadable?")
        if dist.get_backend(group) == "nccl":
            t = torch.tensor(list(idb), dtype=torch.uint8, device=self.tdev)
        else:
            t = torch.tensor(list(idb), dtype=torch.uint8)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().tolist())
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        with self._lock:
            self._check(self.lib.pcr_comm_init(self.ctx, buf, C.c_int(128), C.c_int(rank), C.c_int(world)))
        return world

    def comm_destroy(self) -> None:
        with self._lock:
            self._check(self.lib.pcr_comm_destroy(self.ctx))

    def ransac_multi(self, src, tgt, corr, max_dist: float, max_iter: int, confidence: float = 0.999, seed: int = 0,
                     edge_sim: float = 0.9, first_wave: int = 0, growth: int = 0):
        """RANSAC sharded over the ranks of the context's communicator (pcr_ransac_multi; the whole wave loop, the
        exchange and the replay run in C).  Every rank passes the same inputs and gets the same result.
        Returns (DeviceRegResult, waves)."""
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_corr(corr)
        r = _capi.RegResult()
        nw = C.c_int(0)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_ransac_multi(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt), C.c_int(tgt.shape[0]),
                                                  _ptr(corr), C.c_int(corr.shape[0]), C.c_double(max_dist), C.c_double(edge_sim),
                                                  C.c_int64(max_iter), C.c_double(confidence), C.c_uint64(seed),
                                                  C.c_int64(first_wave), C.c_int(growth), C.byref(r), C.byref(nw)))
        return DeviceRegResult.from_c(r), nw.value

    def align_batch(self, local_pairs, params: _capi.AlignParams, n_total: int, workers: int = 1) -> np.ndarray:
        """pcr_align_batch: this rank's pairs (packed CUDA tensors, global pair index rank, rank + world, ...) aligned by
        `workers` native threads; returns the (n_total, 18) float64 table of ALL pairs (identical on every rank)."""
        n = len(local_pairs)
        for s, t in local_pairs:
            self._need_xyzw(s, "source")
            self._need_xyzw(t, "target")
        sp = (C.c_void_p * max(n, 1))(*[s.data_ptr() for s, _ in local_pairs])
        tp = (C.c_void_p * max(n, 1))(*[t.data_ptr() for _, t in local_pairs])
        ns = (C.c_int * max(n, 1))(*[int(s.shape[0]) for s, _ in local_pairs])
        nt = (C.c_int * max(n, 1))(*[int(t.shape[0]) for _, t in local_pairs])
        out = np.zeros((max(n_total, 1), 18), np.float64)
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_align_batch(self.ctx, C.c_int(n), sp, ns, tp, nt, C.byref(params), C.c_int(workers),
                                                 C.c_int(n_total), out.ctypes.data_as(C.c_void_p)))
        return out[:n_total]

    # ---- ICP -------------------------------------------------------------------------------------------
    def icp_point_to_plane(self, src, tgt, tgt_normals, max_dist: float, init=None, max_iter: int = 30,
                           rel_fitness: float = 1e-6, rel_rmse: float = 1e-6, want_corr: bool = True):
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        self._need_xyzw(tgt_normals, "target normals")
        if tgt_normals.shape[0] != tgt.shape[0]:
            raise RuntimeError("TransformationEstimationPointToPlane requires target normals")
        T0 = np.ascontiguousarray(np.eye(4) if init is None else np.asarray(init, np.float64).reshape(4, 4))
        r = _capi.RegResult()
        corr = torch.empty((src.shape[0],), dtype=torch.int32, device=self.tdev) if want_corr else None
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_icp_point_to_plane(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt),
                                                        _ptr(tgt_normals), C.c_int(tgt.shape[0]), C.c_double(max_dist),
                                                        T0.ctypes.data_as(C.c_void_p), C.c_int(max_iter),
                                                        C.c_double(rel_fitness), C.c_double(rel_rmse), C.byref(r),
                                                        _ptr(corr)))
        return DeviceRegResult.from_c(r), corr

    # ---- end to end --------------------------------------------------------------------------------------
    def default_params(self, voxel_size: float) -> _capi.AlignParams:
        p = _capi.AlignParams()
        self.lib.pcr_align_default_params(C.byref(p))
        p.voxel_size = voxel_size
        return p

    def align_device(self, src: torch.Tensor, tgt: torch.Tensor, params: _capi.AlignParams) -> _capi.AlignResult:
        self._need_xyzw(src, "source")
        self._need_xyzw(tgt, "target")
        res = _capi.AlignResult()
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_align(self.ctx, _ptr(src), C.c_int(src.shape[0]), _ptr(tgt), C.c_int(tgt.shape[0]),
                                           C.byref(params), C.byref(res)))
        return res

    def align_host(self, src_xyz: np.ndarray, tgt_xyz: np.ndarray, params: _capi.AlignParams) -> _capi.AlignResult:
        s = np.ascontiguousarray(src_xyz, np.float32)
        t = np.ascontiguousarray(tgt_xyz, np.float32)
        if s.ndim != 2 or s.shape[1] != 3 or t.ndim != 2 or t.shape[1] != 3:
            raise ValueError("align_host expects (n,3) arrays")
        res = _capi.AlignResult()
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_align_host(self.ctx, s.ctypes.data_as(C.c_void_p), C.c_int(len(s)),
                                                t.ctypes.data_as(C.c_void_p), C.c_int(len(t)), C.byref(params),
                                                C.byref(res)))
        return res

    def align_files(self, src_path, tgt_path, params: _capi.AlignParams) -> _capi.AlignResult:
        """PLY paths in, result out — one C call (pcr_align_files: concurrent native decode into pinned staging)."""
        import os
        res = _capi.AlignResult()
        with self._lock:
            self._bind_stream()
            self._check(self.lib.pcr_align_files(self.ctx, os.fsencode(src_path), os.fsencode(tgt_path), C.byref(params),
                                                 C.byref(res)))
        return res


_default_engines: dict[int, Engine] = {}


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine for a device (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("pcr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    d = torch.cuda.current_device() if device is None else int(device)
    if d not in _default_engines:
        _default_engines[d] = Engine(d)
    return _default_engines[d]
