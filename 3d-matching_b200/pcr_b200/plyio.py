"""PLY reader/writer: ctypes binding of the native pcr_ply_* exports (include/pcr.h, csrc/pcr_ply.cu).

Replaces o3d.io.read_point_cloud at src/ply/ply.py:80 (and o3d.io.write_point_cloud, trim_ply.py:40) for the formats
the reference produces and consumes: ASCII (its converter's output, convert_stl-ply.py:8), binary_little_endian and
binary_big_endian; vertex properties x y z of any scalar type, optional nx ny nz; other properties and elements are
skipped.  `read_ply_xyzw` decodes straight into the packed float4 layout of the device kernels, into pinned memory
when a GPU is present, so a file reaches HBM with one copy and no pack kernel.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi

_ERR_CAP = 512


def _raise(rc: int, err, path) -> None:
    msg = err.value.decode("utf-8", "replace") or f"libpcr_b200 error {rc}"
    if rc == _capi.PCR_ERR_IO:
        if not os.path.exists(path):
            raise FileNotFoundError(f"Ply file not found: {path}")
        raise OSError(f"{path}: {msg}")
    if rc == _capi.PCR_ERR_OOM:
        raise MemoryError(msg)
    raise ValueError(f"{path}: {msg}")


def probe_ply(path) -> _capi.PlyInfo:
    """Header only: vertex count, format, whether normals / colours are present."""
    lib = _capi.load()
    info, err = _capi.PlyInfo(), C.create_string_buffer(_ERR_CAP)
    rc = lib.pcr_ply_probe(os.fsencode(path), C.byref(info), err, C.c_int(_ERR_CAP))
    if rc != 0:
        _raise(rc, err, path)
    return info


def _read(path, xyzw, nrm, xyz64, n, threads):
    lib = _capi.load()
    info, err = _capi.PlyInfo(), C.create_string_buffer(_ERR_CAP)
    ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
    rc = lib.pcr_ply_read(os.fsencode(path), C.c_int64(n), ptr(xyzw), ptr(nrm), ptr(xyz64), C.c_int(threads),
                          C.byref(info), err, C.c_int(_ERR_CAP))
    if rc != 0:
        _raise(rc, err, path)
    return info


def read_ply(path, threads: int = 0):
    """Returns (points (n,3) float64, normals (n,3) float64 or None) — the values Open3D's containers would hold."""
    info = probe_ply(path)
    n = int(info.n_vertex)
    pts = np.empty((n, 3), np.float64)
    nrm4 = np.empty((n, 4), np.float32) if info.has_normals else None
    if n:
        _read(path, None, nrm4, pts, n, threads)
    return pts, (nrm4[:, :3].astype(np.float64) if nrm4 is not None else None)


def read_ply_xyzw(path, pin: bool | None = None, with_normals: bool = False, threads: int = 0):
    """Returns (xyzw, normals_xyzw or None): (n,4) fp32 HOST torch tensors in the packed device layout (w = 0, points
    quantised to fp32 — rule D1).  pin=None pins the memory when CUDA is available."""
    import torch
    info = probe_ply(path)
    n = int(info.n_vertex)
    if pin is None:
        pin = torch.cuda.is_available()
    xyzw = torch.empty((n, 4), dtype=torch.float32, pin_memory=bool(pin and n))
    nrm = torch.empty((n, 4), dtype=torch.float32, pin_memory=bool(pin and n)) if (with_normals and info.has_normals) else None
    if n:
        _read(path, xyzw.numpy(), nrm.numpy() if nrm is not None else None, None, n, threads)
    return xyzw, nrm


def write_ply(path, points, normals=None, binary: bool = True, colors=None) -> None:
    """points (n,3) or packed (n,4); normals likewise; colors (n,3) uint8, or floats in [0,1] as Open3D keeps them."""
    def pack4(a):
        a = np.asarray(a, np.float32)
        if a.ndim != 2 or a.shape[1] not in (3, 4):
            raise ValueError(f"expected an (n,3) array, got {a.shape}")
        if a.shape[1] == 4:
            return np.ascontiguousarray(a)
        out = np.zeros((len(a), 4), np.float32)
        out[:, :3] = a
        return out
    p4 = pack4(points)
    n4 = pack4(normals) if normals is not None else None
    rgb = None
    if colors is not None:
        c = np.asarray(colors)
        if c.dtype != np.uint8:
            c = np.clip(np.rint(np.asarray(c, np.float64) * 255.0), 0, 255).astype(np.uint8)
        rgb = np.ascontiguousarray(c.reshape(-1, 3))
        if len(rgb) != len(p4):
            raise ValueError("colors and points differ in length")
    if n4 is not None and len(n4) != len(p4):
        raise ValueError("normals and points differ in length")
    lib = _capi.load()
    err = C.create_string_buffer(_ERR_CAP)
    ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
    rc = lib.pcr_ply_write(os.fsencode(path), ptr(p4), C.c_int64(len(p4)), ptr(n4), ptr(rgb), C.c_int(int(binary)),
                           err, C.c_int(_ERR_CAP))
    if rc != 0:
        _raise(rc, err, path)
