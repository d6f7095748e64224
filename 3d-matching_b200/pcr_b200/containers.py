"""Host-side containers mirroring the Open3D types the reference's matcher touches.

PointCloud  ~ open3d.geometry.PointCloud  (only what src/matcher and src/ply use: .points, .normals, has_points,
              transform, len(points))
Feature     ~ open3d.pipelines.registration.Feature (.data is (33, n) float64, column = point)
RegistrationResult ~ open3d.pipelines.registration.RegistrationResult: plain, mutable, truthy
              (constructed at src/matcher/ransac.py:134-136; written at _visualize_matcher.py:423)

Device data are packed float4 torch tensors; numpy views are produced lazily and cached.
"""
from __future__ import annotations

import numpy as np
import torch


class PointCloud:
    def __init__(self, xyzw: torch.Tensor, normals_xyzw: torch.Tensor | None = None, normals_fn=None):
        self._xyzw = xyzw
        self._normals = normals_xyzw
        self._normals_fn = normals_fn  # lazy estimator: (xyzw of the cloud at evaluation time) -> (n,4) tensor
        self._points_np = None

    # ---- device side --------------------------------------------------------------------------------
    @property
    def xyzw(self) -> torch.Tensor:
        return self._xyzw

    @property
    def normals_xyzw(self) -> torch.Tensor | None:
        self._materialise_normals()
        return self._normals

    def _materialise_normals(self) -> None:
        """Run the lazy estimator (on the CURRENT points: it receives the live tensor) exactly once."""
        if self._normals is None and self._normals_fn is not None:
            fn, self._normals_fn = self._normals_fn, None
            self._normals = fn(self._xyzw)

    # ---- Open3D-like host side ------------------------------------------------------------------------
    @property
    def points(self) -> np.ndarray:
        """(n,3) float64 host copy (what np.asarray(pcd.points) yields in the reference)."""
        if self._points_np is None:
            self._points_np = self._xyzw[:, :3].to(torch.float64).cpu().numpy()
        return self._points_np

    @points.setter
    def points(self, value):
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float64))
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError("points must be (n,3)")
        # The reference estimates normals EAGERLY (src/ply/ply.py:65, :110) and Open3D keeps them when the points are
        # replaced: a lazy estimate must therefore be taken on the cloud as it was, before the points change.
        self._materialise_normals()
        dev = self._xyzw.device
        t = torch.zeros((a.shape[0], 4), dtype=torch.float32, device=dev)
        t[:, :3] = torch.from_numpy(a).to(dev).to(torch.float32)  # quantise to fp32 (rule D1)
        self._xyzw = t
        self._points_np = None

    @property
    def normals(self) -> np.ndarray:
        n = self.normals_xyzw
        return np.zeros((0, 3)) if n is None else n[:, :3].to(torch.float64).cpu().numpy()

    def has_points(self) -> bool:
        return self._xyzw.shape[0] > 0

    def has_normals(self) -> bool:
        return self._normals is not None or self._normals_fn is not None

    def __len__(self) -> int:
        return int(self._xyzw.shape[0])

    def transform(self, T) -> "PointCloud":
        """In-place rigid transform (points and normals), like open3d's pcd.transform."""
        from .engine import get_engine
        self._materialise_normals()  # Open3D's transform rotates the (eagerly estimated) normals with the points
        self._xyzw = get_engine(self._xyzw.device.index).transform_points(self._xyzw.contiguous(), T)
        T = torch.as_tensor(np.asarray(T, np.float64), device=self._xyzw.device)
        self._points_np = None
        if self._normals is not None:
            n = self._normals[:, :3].to(torch.float64) @ T[:3, :3].T
            self._normals = torch.cat([n.to(torch.float32), torch.zeros_like(self._normals[:, 3:])], dim=1).contiguous()
        return self


class Feature:
    def __init__(self, dev: torch.Tensor):
        self.dev = dev  # (n, 33) fp32
        self._data = None

    @property
    def data(self) -> np.ndarray:
        if self._data is None:
            self._data = self.dev.to(torch.float64).cpu().numpy().T.copy()
        return self._data

    def dimension(self) -> int:
        return 33

    def num(self) -> int:
        return int(self.dev.shape[0])


class RegistrationResult:
    def __init__(self, transformation=None, fitness: float = 0.0, inlier_rmse: float = 0.0, correspondence_set=None):
        self.transformation = np.eye(4) if transformation is None else np.array(transformation, np.float64).reshape(4, 4)
        self.fitness = float(fitness)
        self.inlier_rmse = float(inlier_rmse)
        self._corr = correspondence_set
        self.info = {}

    @property
    def correspondence_set(self) -> np.ndarray:
        c = self._corr
        if c is None:
            return np.zeros((0, 2), np.int32)
        if callable(c):
            self._corr = c = c()
        return c

    @correspondence_set.setter
    def correspondence_set(self, v):
        self._corr = v

    def __bool__(self) -> bool:
        return True

    def __repr__(self) -> str:
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, "
                f"and correspondence_set size of {len(self.correspondence_set)}")
