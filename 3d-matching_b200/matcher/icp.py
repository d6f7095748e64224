"""Mirror of the reference's src/matcher/icp.py: refine_registration -> pcr_icp_point_to_plane.

Same positional signature (src, tgt, init_trans, voxel_size); full-resolution clouds (icp.py:43-44), threshold
0.4 * voxel (icp.py:41), point-to-plane estimator (icp.py:47) and — because the reference passes no criteria —
Open3D's defaults max_iteration 30, relative_fitness = relative_rmse = 1e-6.  Keyword-only extras
expose those three (config 2 of BASELINE.json asks for 50 iterations).
"""
from __future__ import annotations

import numpy as np

from pcr_b200.containers import PointCloud, RegistrationResult
from pcr_b200.engine import get_engine

from ._common import device_cloud, voxel_of


def _target_normals(tgt_pcd, eng):
    if isinstance(tgt_pcd, PointCloud):
        n = tgt_pcd.normals_xyzw
    else:
        nrm = np.asarray(getattr(tgt_pcd, "normals", np.zeros((0, 3))))
        n = eng.pack(nrm) if nrm.size else None
    if n is None or n.shape[0] == 0:
        # Open3D: "TransformationEstimationPointToPlane and TransformationEstimationColoredICP require
        # pre-computed normal vectors for target PointCloud."
        raise RuntimeError("point-to-plane ICP requires pre-computed normal vectors for the target point cloud")
    return n


def refine_registration(src, tgt, init_trans, voxel_size: float | None = None, *, max_iteration: int = 30,
                        relative_fitness: float = 1e-6, relative_rmse: float = 1e-6) -> RegistrationResult:
    voxel_size = voxel_of(src, voxel_size)
    eng = get_engine()
    dist_thresh = voxel_size * 0.4
    if hasattr(init_trans, "transformation"):
        init_trans = init_trans.transformation
    s, t = device_cloud(src.pcd, eng), device_cloud(tgt.pcd, eng)
    n = _target_normals(tgt.pcd, eng)
    r, corr = eng.icp_point_to_plane(s, t, n, dist_thresh, np.asarray(init_trans, np.float64), max_iteration,
                                     relative_fitness, relative_rmse, want_corr=True)

    def corr_set():
        c = corr.cpu().numpy()
        keep = np.nonzero(c >= 0)[0]
        return np.stack([keep.astype(np.int32), c[keep]], axis=1)

    res = RegistrationResult(r.transformation, r.fitness, r.inlier_rmse, corr_set)
    res.info = {"iterations": r.iterations, "converged": r.converged, "inlier_count": r.inlier_count}
    return res
