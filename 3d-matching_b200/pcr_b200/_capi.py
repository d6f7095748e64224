"""ctypes binding of libpcr_b200.so (the C ABI in include/pcr.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is not available the import of
the engine fails loudly.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcr_b200.so")

PCR_OK = 0
PCR_ERR_INVALID = -1
PCR_ERR_CUDA = -2
PCR_ERR_OOM = -3
PCR_ERR_BUSY = -4
PCR_ERR_TOO_LARGE = -5
PCR_ERR_IO = -6


class RegResult(C.Structure):
    """pcr_reg_result"""
    _fields_ = [
        ("transformation", C.c_double * 16),
        ("fitness", C.c_double),
        ("inlier_rmse", C.c_double),
        ("inlier_count", C.c_int64),
        ("sum_d2_fixed", C.c_int64),
        ("k_d", C.c_int32),
        ("iterations", C.c_int32),
        ("converged", C.c_int32),
        ("reserved", C.c_int32),
        ("best_hyp", C.c_int64),
        ("hyp_evaluated", C.c_int64),
        ("survivors", C.c_int64),
        ("est_k", C.c_int64),
    ]


class HypRecord(C.Structure):
    """pcr_hyp_record"""
    _fields_ = [
        ("hyp", C.c_int64),
        ("inlier_count", C.c_int64),
        ("sum_d2_fixed", C.c_int64),
        ("corr_inliers", C.c_int32),
        ("reserved", C.c_int32),
        ("transformation", C.c_double * 12),
    ]


class KernelStat(C.Structure):
    """pcr_kernel_stat"""
    _fields_ = [("total_ms", C.c_double), ("launches", C.c_int64), ("bytes", C.c_double), ("flops", C.c_double),
                ("overlapped_ms", C.c_double)]


class AlignParams(C.Structure):
    """pcr_align_params"""
    _fields_ = [
        ("voxel_size", C.c_double),
        ("ransac_max_iter", C.c_int64),
        ("ransac_confidence", C.c_double),
        ("seed", C.c_uint64),
        ("icp_max_iter", C.c_int32),
        ("icp_rel_fitness", C.c_double),
        ("icp_rel_rmse", C.c_double),
        ("source_normals", C.c_int32),
        ("reserved", C.c_int32),
    ]


class AlignResult(C.Structure):
    """pcr_align_result"""
    _fields_ = [
        ("ransac", RegResult),
        ("icp", RegResult),
        ("n_src_down", C.c_int32),
        ("n_tgt_down", C.c_int32),
        ("n_corr", C.c_int32),
        ("reserved", C.c_int32),
        ("stage_ms", C.c_float * 8),
    ]


class PlyInfo(C.Structure):
    """pcr_ply_info"""
    _fields_ = [("n_vertex", C.c_int64), ("data_offset", C.c_int64), ("format", C.c_int32), ("has_normals", C.c_int32),
                ("has_colors", C.c_int32), ("n_props", C.c_int32), ("vertex_stride", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/pcr.h declares (tests/test_capi_exports.py checks the list against the header)
EXPORTS = [
    "pcr_create", "pcr_destroy", "pcr_last_error", "pcr_set_stream", "pcr_version", "pcr_launch_count",
    "pcr_set_profiling", "pcr_kernel_class_count", "pcr_kernel_class_name", "pcr_kernel_stats",
    "pcr_pack_xyz_f32", "pcr_pack_xyz_f64", "pcr_unpack_xyz_f32", "pcr_transform_points",
    "pcr_voxel_downsample", "pcr_estimate_normals", "pcr_compute_fpfh", "pcr_knn_hybrid", "pcr_nn1",
    "pcr_match_features", "pcr_nn_features",
    "pcr_ransac", "pcr_ransac_wave", "pcr_ransac_session_begin", "pcr_ransac_session_end", "pcr_ransac_scan", "pcr_ransac_k_d",
    "pcr_ransac_step", "pcr_inlier_count",
    "pcr_icp_point_to_plane",
    "pcr_align_default_params", "pcr_align", "pcr_align_host", "pcr_align_files",
    "pcr_comm_unique_id", "pcr_comm_init", "pcr_comm_destroy", "pcr_ransac_multi", "pcr_align_batch",
    "pcr_ply_probe", "pcr_ply_read", "pcr_ply_write",
]

_lib = None


def load():
    """Load the shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C 3d-matching_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU fallback for this engine."
        )
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} does not export {missing}: rebuild it (make -C 3d-matching_b200/csrc)")
    lib.pcr_last_error.restype = C.c_char_p
    lib.pcr_launch_count.restype = C.c_int64
    lib.pcr_kernel_class_name.restype = C.c_char_p
    lib.pcr_align_default_params.restype = None
    _lib = lib
    return lib


class PcrError(RuntimeError):
    pass


def check(lib, ctx, rc: int) -> None:
    """Map pcr_status onto the exceptions the reference raises (src/ply/ply.py:46-51,81-84; Open3D RuntimeError)."""
    if rc == PCR_OK:
        return
    if rc == PCR_ERR_BUSY:  # the message buffer belongs to the call in progress on that context
        raise PcrError("libpcr_b200: context is in use by another call (PCR_ERR_BUSY)")
    msg = lib.pcr_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc in (PCR_ERR_INVALID, PCR_ERR_TOO_LARGE):
        raise ValueError(msg or "invalid argument")
    if rc == PCR_ERR_OOM:
        raise MemoryError(msg or "device out of memory")
    if rc == PCR_ERR_IO:
        raise OSError(msg or "file I/O error")
    raise PcrError(f"libpcr_b200 error {rc}: {msg}")
