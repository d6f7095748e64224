"""Shared helpers of the matcher mirror: duck-typed access to Ply-shaped arguments."""
from __future__ import annotations

import numpy as np
import torch

from pcr_b200.containers import Feature, PointCloud
from pcr_b200.engine import get_engine


def device_cloud(pcd, eng=None) -> torch.Tensor:
    """Packed float4 device tensor of a PointCloud or of any object with an array-like `.points`
    (the reference's tests pass Open3D clouds inside a MockPly, test_ransac_crash.py:92-96)."""
    eng = eng or get_engine()
    if isinstance(pcd, PointCloud):
        return pcd.xyzw
    if isinstance(pcd, torch.Tensor):
        return pcd if (pcd.is_cuda and pcd.ndim == 2 and pcd.shape[1] == 4 and pcd.dtype == torch.float32) else eng.pack(pcd)
    pts = pcd.points if hasattr(pcd, "points") else pcd
    arr = np.asarray(pts)
    if arr.size == 0:
        return torch.zeros((0, 4), dtype=torch.float32, device=eng.tdev)
    # No cache: the reference re-reads np.asarray(pcd.points) on every call (src/matcher/ransac.py:147-148, :223-224), and
    # callers do change clouds in place between calls (pcd.transform, `pts += ...`); a 10k-point pack is ~20 us.
    return eng.pack(arr)


def device_feature(f, eng=None) -> torch.Tensor:
    eng = eng or get_engine()
    if isinstance(f, Feature):
        return f.dev
    data = np.asarray(f.data if hasattr(f, "data") else f, dtype=np.float64)  # (33, n) as in Open3D
    if data.ndim != 2 or data.shape[0] != 33:
        raise ValueError("FPFH feature must be (33, n)")
    return torch.from_numpy(np.ascontiguousarray(data.T.astype(np.float32))).to(eng.tdev)


def device_corr(corr, eng=None) -> torch.Tensor:
    eng = eng or get_engine()
    if isinstance(corr, torch.Tensor):
        return corr.to(eng.tdev, torch.int32).reshape(-1, 2).contiguous()
    a = np.ascontiguousarray(np.asarray(corr, dtype=np.int32).reshape(-1, 2))
    return torch.from_numpy(a).to(eng.tdev)


class index_errors:
    """The reference raises IndexError for a correspondence that points outside a cloud (NumPy fancy indexing,
    src/matcher/ransac.py:147-148, :223-224); the C ABI reports it as PCR_ERR_INVALID -> ValueError."""

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is ValueError and "indexes outside the clouds" in str(ev):
            raise IndexError(str(ev)) from None
        return False


def voxel_of(obj, voxel_size):
    """voxel_size falls back to the Ply's own (tolerates the arity of src/main.py:34,38)."""
    if voxel_size is not None:
        return float(voxel_size)
    v = getattr(obj, "voxel_size", None)
    return float(v) if v is not None else 0.3  # Ply default, src/ply/ply.py:32
