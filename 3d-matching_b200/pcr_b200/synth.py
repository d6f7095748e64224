"""Deterministic synthetic cloud pairs.  The reference ships no data
(3d_data/.gitignore:1-2), so every test and benchmark uses this generator.

Target: N points on a closed bumpy surface r(u) = R(1 + 0.15 sin(3ux)cos(4uy) + 0.10 sin(5uz)), R chosen so
the mean spacing is 0.25*voxel, plus isotropic jitter 0.1*spacing.  Source: an independent sample of the same
surface (seed+1) moved by a known SE(3) drawn from the reference's own test distribution (+-30 deg per axis,
+-0.1*extent translation; src/visualize_matcher/_visualize_matcher.py:190-191,302-323).  Coordinates are
quantised to fp32 (determinism rule D1).
"""
from __future__ import annotations

import numpy as np


def euler_zyx(ax: float, ay: float, az: float) -> np.ndarray:
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def surface(n: int, voxel: float, seed: int, spacing_ratio: float = 0.25, jitter: float = 0.1) -> np.ndarray:
    """(n,3) float64 points on the bumpy sphere; mean spacing = spacing_ratio*voxel."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = spacing_ratio * voxel
    big_r = np.sqrt(n * s * s / (4.0 * np.pi))
    u = rng.standard_normal((n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    r = big_r * (1.0 + 0.15 * np.sin(3.0 * u[:, 0]) * np.cos(4.0 * u[:, 1]) + 0.10 * np.sin(5.0 * u[:, 2]))
    p = u * r[:, None]
    p += (jitter * s) * rng.standard_normal((n, 3))
    return p


def make_pair(n: int, voxel: float, seed: int, max_angle: float = np.pi / 6, max_shift: float = 0.1,
              n_src: int | None = None):
    """Returns (source (n,3) f32, target (n,3) f32, T_true (4,4) f64) with target ~= T_true(source)."""
    tgt = surface(n, voxel, seed)
    src0 = surface(n_src or n, voxel, seed + 1)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    ang = rng.uniform(-max_angle, max_angle, 3)
    rot = euler_zyx(*ang)
    extent = float(np.max(tgt.max(0) - tgt.min(0)))
    shift = rng.uniform(-max_shift, max_shift, 3) * extent
    c = src0.mean(0)
    # target = R (src - c) + c + shift  =>  src = R^T (x - c - shift) + c applied to the independent sample
    src = (src0 - c - shift) @ rot + c
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = c + shift - rot @ c
    return src.astype(np.float32), tgt.astype(np.float32), T


def make_icp_pair(n: int, voxel: float, seed: int):
    """ICP-only configuration (cfg 3): perturbation 0.1 deg / 0.25*(0.4 voxel) so pairs exist inside the gate."""
    tgt = surface(n, voxel, seed)
    src0 = surface(n, voxel, seed + 1)
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    ang = rng.uniform(-1.0, 1.0, 3) * np.deg2rad(0.1)
    rot = euler_zyx(*ang)
    d = rng.standard_normal(3)
    shift = d / np.linalg.norm(d) * 0.25 * 0.4 * voxel
    c = src0.mean(0)
    src = (src0 - c - shift) @ rot + c
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = c + shift - rot @ c
    return src.astype(np.float32), tgt.astype(np.float32), T
