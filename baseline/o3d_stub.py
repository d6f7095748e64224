"""Container-only stand-in for the `open3d` module, so that the reference's NumPy functions
(src/matcher/ransac.py:104-277: compute_step_transformation, evaluate_inlier_ratio, evaluate_inlier_ratio_fast) import
and run UNMODIFIED where the 447 MB open3d wheel is absent.  It supplies the three container types those functions
touch (Vector2iVector / Vector3dVector as ndarray casts, RegistrationResult / Feature as plain classes) and a trivial
`ply` module for the reference's `from ply import Ply`; NO arithmetic is stubbed.  Test / baseline infrastructure only."""
import sys
import types

import numpy as np


def install_stub() -> None:
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
    ply = types.ModuleType("ply")

    class Ply:
        pass

    ply.Ply = Ply
    sys.modules["ply"] = ply
