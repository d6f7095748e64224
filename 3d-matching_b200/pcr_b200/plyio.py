"""Minimal PLY reader/writer (replaces o3d.io.read_point_cloud at src/ply/ply.py:80 for the formats the
reference produces: its converter writes ASCII PLY, convert_stl-ply.py:8).  ASCII and binary_little_endian,
vertex properties x y z (float/double) and optional nx ny nz; other properties are skipped."""
from __future__ import annotations

import numpy as np

_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
          "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
          "double": "f8", "float64": "f8"}


def read_ply(path):
    """Returns (points (n,3) float64, normals (n,3) float64 or None)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"not a PLY file: {path}")
        fmt = None
        n_vertex = 0
        props = []
        in_vertex = False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("unexpected end of PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties on vertices are not supported")
                props.append((tok[2], _TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if not all(k in names for k in ("x", "y", "z")):
            raise ValueError("PLY vertex element lacks x/y/z")
        if n_vertex == 0:
            return np.zeros((0, 3)), None
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=n_vertex, ndmin=2)
            cols = {n: data[:, i] for i, n in enumerate(names)}
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            e = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(n, e + t) for n, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * n_vertex), dtype=dt, count=n_vertex)
            cols = {n: rec[n].astype(np.float64) for n in names}
        else:
            raise ValueError(f"unsupported PLY format {fmt}")
    pts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
    nrm = None
    if all(k in cols for k in ("nx", "ny", "nz")):
        nrm = np.stack([cols["nx"], cols["ny"], cols["nz"]], axis=1)
    return pts, nrm


def write_ply(path, points, normals=None, binary=True):
    points = np.asarray(points, np.float32)
    n = len(points)
    has_n = normals is not None
    hdr = ["ply", "format binary_little_endian 1.0" if binary else "format ascii 1.0", f"element vertex {n}",
           "property float x", "property float y", "property float z"]
    if has_n:
        hdr += ["property float nx", "property float ny", "property float nz"]
    hdr.append("end_header")
    data = points if not has_n else np.concatenate([points, np.asarray(normals, np.float32)], axis=1)
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        if binary:
            f.write(np.ascontiguousarray(data, "<f4").tobytes())
        else:
            np.savetxt(f, data, fmt="%.9g")
