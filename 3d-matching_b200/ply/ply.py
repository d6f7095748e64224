"""Mirror of the reference's src/ply/ply.py on the B200 engine.

`Ply(path, voxel_size=0.3)` keeps the reference's attribute protocol — path, voxel_size, pcd, pcd_down, pcd_fpfh
(src/ply/ply.py:20-66) — and its error behaviour (FileNotFoundError / TypeError / ValueError, :46-51, :81-84).
Every Open3D call of `_preprocess` runs on the GPU through the C ABI:
    voxel_down_sample(v)                 ply.py:106      -> pcr_voxel_downsample
    estimate_normals(Hybrid(2v, 30))     ply.py:110-112  -> pcr_estimate_normals
    compute_fpfh_feature(Hybrid(5v,100)) ply.py:117-120  -> pcr_compute_fpfh
    pcd_down.points += 0.05*randn        ply.py:61-62    -> host RNG (seedable), re-quantised to fp32
    estimate_normals on the full cloud   ply.py:65       -> pcr_estimate_normals (evaluated on first use)

Differences, all documented in DESIGN.md: `noise_sigma` and `seed` keyword arguments (the reference hard-codes
sigma = 0.05 with an unseeded generator); `Ply.from_points` builds the same object from an array.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from pcr_b200.containers import Feature, PointCloud
from pcr_b200.engine import get_engine
from pcr_b200.plyio import read_ply_xyzw


class Ply:
    def __init__(self, path, voxel_size: float = 0.3, *, noise_sigma: float = 0.05, seed: int | None = None,
                 device: int | None = None, _points=None) -> None:
        self.path = Path(path) if path is not None else None
        self.voxel_size = voxel_size
        if _points is None:
            if not self.path.exists():
                raise FileNotFoundError(f"Ply file not found: {self.path}")
            if self.path.suffix.lower() != ".ply":
                raise TypeError(f"File is not a ply file: {self.path}")
            pts, _ = read_ply_xyzw(self.path)  # pcr_ply_read: pinned packed float4, replaces ply.py:80
            if len(pts) == 0:
                raise ValueError(f"Point cloud is empty: {self.path}")
        else:
            pts = _points
            if len(pts) == 0:
                raise ValueError("Point cloud is empty")
        eng = get_engine(device)
        self._eng = eng
        xyzw = eng.pack(pts)
        v = float(voxel_size)
        # full-resolution normals are only needed by point-to-plane ICP on the target: computed on first access
        self.pcd = PointCloud(xyzw, normals_fn=lambda cur: eng.estimate_normals(cur.contiguous(), 2.0 * v, 30))
        self.pcd_down, self.pcd_fpfh = self._preprocess(self.pcd, v)
        if noise_sigma and noise_sigma > 0.0:
            rng = np.random.default_rng(seed)
            noisy = self.pcd_down.points + noise_sigma * rng.standard_normal(self.pcd_down.points.shape)
            self.pcd_down.points = noisy  # descriptors are NOT recomputed (reference behaviour, ply.py:61-62)

    @classmethod
    def from_points(cls, points, voxel_size: float = 0.3, *, noise_sigma: float = 0.0, seed: int | None = None,
                    device: int | None = None) -> "Ply":
        return cls(None, voxel_size, noise_sigma=noise_sigma, seed=seed, device=device, _points=points)

    def _preprocess(self, pcd: PointCloud, voxel_size: float):
        eng = self._eng
        down = eng.voxel_downsample(pcd.xyzw, voxel_size).contiguous()
        normals = eng.estimate_normals(down, voxel_size * 2.0, 30)
        fpfh = eng.compute_fpfh(down, normals, voxel_size * 5.0, 100)
        return PointCloud(down, normals), Feature(fpfh)
