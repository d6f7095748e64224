"""Headless counterpart of the reference's src/visualization/draw_registration_result.py:20-49.

The reference deep-copies the two down-sampled clouds, paints the source yellow [1, 0.706, 0] and the target cyan
[0, 0.651, 0.929], applies the transformation to the source and opens a blocking Open3D window.  A GPU box has no
display: the same scene is written as ONE coloured PLY (moved source followed by the target) plus a JSON file holding
the transformation, so it can be opened in any viewer afterwards.  Same name and positional arguments, so
src/main.py:35,39 runs unchanged; `transformation` may also be a RegistrationResult (main.py:39 passes one).

Output directory: `out_dir=` keyword, else $PCR_EXPORT_DIR, else ./registration_out; files are numbered per call
(registration_000.ply / .json, ...).  Returns the path of the PLY.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

from matcher._common import device_cloud
from pcr_b200.engine import get_engine
from pcr_b200.plyio import write_ply

SOURCE_COLOR = (1.0, 0.706, 0.0)    # draw_registration_result.py:37
TARGET_COLOR = (0.0, 0.651, 0.929)  # draw_registration_result.py:38
_calls = {"n": 0}


def _as_matrix(transformation) -> np.ndarray:
    T = getattr(transformation, "transformation", transformation)
    T = np.asarray(T, np.float64)
    if T.shape != (4, 4):
        raise ValueError(f"transformation must be 4x4, got {T.shape}")
    return T


def draw_registration_result(source, target, transformation, *, out_dir=None, full_resolution: bool = False) -> Path:
    T = _as_matrix(transformation)
    eng = get_engine()
    pick = (lambda p: p.pcd) if full_resolution else (lambda p: p.pcd_down)
    src = device_cloud(pick(source), eng)
    tgt = device_cloud(pick(target), eng)
    moved = eng.transform_points(src.contiguous(), T)  # pcd.transform(T) on a copy (:33-41); rule D7 rounding
    pts = np.concatenate([moved.cpu().numpy(), tgt.cpu().numpy()], axis=0)
    rgb = np.empty((len(pts), 3), np.uint8)
    rgb[: len(moved)] = np.rint(np.array(SOURCE_COLOR) * 255.0)
    rgb[len(moved):] = np.rint(np.array(TARGET_COLOR) * 255.0)
    out = Path(out_dir or os.environ.get("PCR_EXPORT_DIR") or "registration_out")
    out.mkdir(parents=True, exist_ok=True)
    k = _calls["n"]
    _calls["n"] = k + 1
    ply_path = out / f"registration_{k:03d}.ply"
    write_ply(ply_path, pts, colors=rgb)
    meta = {"transformation": T.tolist(), "n_source": int(len(moved)), "n_target": int(len(tgt)),
            "source_color": SOURCE_COLOR, "target_color": TARGET_COLOR,
            "fitness": getattr(transformation, "fitness", None), "inlier_rmse": getattr(transformation, "inlier_rmse", None)}
    ply_path.with_suffix(".json").write_text(json.dumps(meta, indent=1))
    return ply_path
