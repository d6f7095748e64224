"""align(source, target, voxel_size) -> (T 4x4, fitness, inlier_rmse): the north-star call.

Runs the reference's whole pipeline — Ply preprocessing (src/ply/ply.py:87-135), global_registration
(src/matcher/ransac.py:20-59), refine_registration (src/matcher/icp.py:17-48) — as ONE call into the C ABI
(pcr_align_files for two PLY paths, pcr_align_host for host arrays, pcr_align for device tensors).
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch

from .engine import get_engine
from .plyio import read_ply_xyzw


def _is_path(x) -> bool:
    return isinstance(x, (str, os.PathLike, Path))


def _check_path(x) -> Path:
    """The reference's input validation (src/ply/ply.py:46-51)."""
    p = Path(x)
    if not p.exists():
        raise FileNotFoundError(f"Ply file not found: {p}")
    if p.suffix.lower() != ".ply":
        raise TypeError(f"File is not a ply file: {p}")
    return p


def _load(x):
    if _is_path(x):
        p = _check_path(x)
        xyzw, _ = read_ply_xyzw(p)  # native reader: file -> pinned packed float4 (one H2D copy, no pack kernel)
        if len(xyzw) == 0:
            raise ValueError(f"Point cloud is empty: {p}")
        return xyzw
    return x


def align(source, target, voxel_size: float, *, ransac_iteration: int = 100000, confidence: float = 0.999,
          seed: int = 0, icp_max_iteration: int = 30, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
          source_normals: bool = True, device: int | None = None, return_info: bool = False):
    """source/target: (n,3) numpy arrays, torch tensors (host or CUDA), or PLY paths."""
    eng = get_engine(device)
    p = eng.default_params(float(voxel_size))
    p.ransac_max_iter = int(ransac_iteration)
    p.ransac_confidence = float(confidence)
    p.seed = int(seed)
    p.icp_max_iter = int(icp_max_iteration)
    p.icp_rel_fitness = float(relative_fitness)
    p.icp_rel_rmse = float(relative_rmse)
    p.source_normals = int(bool(source_normals))
    if _is_path(source) and _is_path(target):
        # both clouds come from files: one C call decodes them concurrently into pinned staging and aligns
        res = eng.align_files(_check_path(source), _check_path(target), p)
        T = np.array(res.icp.transformation, np.float64).reshape(4, 4)
        return (T, res.icp.fitness, res.icp.inlier_rmse, res) if return_info else (T, res.icp.fitness, res.icp.inlier_rmse)
    s, t = _load(source), _load(target)
    packed = lambda a: isinstance(a, torch.Tensor) and (a.is_cuda or a.shape[-1] == 4)
    if packed(s) or packed(t):
        res = eng.align_device(eng.pack(s), eng.pack(t), p)
    else:
        s = s.cpu().numpy() if isinstance(s, torch.Tensor) else np.asarray(s)
        t = t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        res = eng.align_host(s, t, p)
    T = np.array(res.icp.transformation, np.float64).reshape(4, 4)
    if return_info:
        return T, res.icp.fitness, res.icp.inlier_rmse, res
    return T, res.icp.fitness, res.icp.inlier_rmse
