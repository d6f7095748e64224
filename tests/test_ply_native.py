"""Native PLY reader / writer (pcr_ply_* in include/pcr.h) against an independent NumPy restatement.

The reference reads its inputs with o3d.io.read_point_cloud (src/ply/ply.py:80) and its converter writes ASCII PLY
(convert_stl-ply.py:8); Open3D is absent here, so the checker is a header parser + numpy.loadtxt / numpy.frombuffer
written below.  Host-only code: everything runs without a GPU.
"""
import os
import struct

import numpy as np
import pytest

from pcr_b200 import _capi
from pcr_b200.plyio import probe_ply, read_ply, read_ply_xyzw, write_ply

_NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
       "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
       "double": "f8", "float64": "f8"}


def numpy_read_ply(path):
    """(points f64 (n,3), normals f64 (n,3) | None) — the vertex element must come first (all the files below that are
    compared through this function are laid out like that)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            tok = f.readline().decode("ascii").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                props.append((tok[2], _NP[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=n, ndmin=2)
            cols = {k: data[:, i] for i, k in enumerate(names)}
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(k, e + t) for k, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
            cols = {k: rec[k].astype(np.float64) for k in names}
    pts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1)
    nrm = np.stack([cols["nx"], cols["ny"], cols["nz"]], axis=1) if "nx" in cols else None
    return pts, nrm


def header(fmt, n, props, extra_before="", extra_after=""):
    h = f"ply\nformat {fmt} 1.0\ncomment made by a test\n{extra_before}element vertex {n}\n"
    h += "".join(f"property {t} {k}\n" for k, t in props)
    return (h + extra_after + "end_header\n").encode("ascii")


def cloud(n, seed=0, cols=3):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, cols)) * np.array([1.0, 10.0, 0.01] * (cols // 3))


def same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n", [1, 7, 1000, 120_000])  # 120k rows take the line-parallel path
def test_ascii_matches_numpy(tmp_path, n):
    p = tmp_path / "a.ply"
    d = cloud(n, seed=n, cols=6)
    with open(p, "wb") as f:
        f.write(header("ascii", n, [(k, "float") for k in ("x", "y", "z", "nx", "ny", "nz")]))
        np.savetxt(f, d, fmt="%.9g")
    pts, nrm = read_ply(p)
    rp, rn = numpy_read_ply(p)
    assert same(pts, rp)
    assert same(nrm, rn.astype(np.float32).astype(np.float64))  # normals are kept in fp32
    xyzw, n4 = read_ply_xyzw(p, pin=False, with_normals=True)
    assert same(xyzw.numpy()[:, :3], rp.astype(np.float32)) and not xyzw.numpy()[:, 3].any()
    assert same(n4.numpy()[:, :3], rn.astype(np.float32))
    info = probe_ply(p)
    assert (info.n_vertex, info.format, info.has_normals, info.has_colors, info.n_props) == (n, 0, 1, 0, 6)


def test_ascii_thread_counts_agree(tmp_path):
    p = tmp_path / "a.ply"
    n = 200_000
    d = cloud(n, seed=5)
    with open(p, "wb") as f:
        f.write(header("ascii", n, [(k, "double") for k in "xyz"]))
        np.savetxt(f, d, fmt="%.17g")
    ref = read_ply(p, threads=1)[0]
    assert same(ref, d)  # 17 significant digits round-trip fp64 exactly
    for t in (2, 3, 5, 12):
        assert same(read_ply(p, threads=t)[0], ref)


def test_ascii_odd_layouts_fall_back_to_the_token_stream(tmp_path):
    n = 70_000
    d = cloud(n, seed=9)
    rows = ["%.9g %.9g %.9g" % tuple(r) for r in d]
    want = np.array([[float(x) for x in r.split()] for r in rows])
    variants = {
        "crlf": "\r\n".join(rows) + "\r\n",
        "blank_lines": "\n".join(r + ("\n" if i % 1000 == 0 else "") for i, r in enumerate(rows)) + "\n",
        "two_per_line": "\n".join(" ".join(rows[i:i + 2]) for i in range(0, n, 2)) + "\n",
        "no_final_newline": "\n".join(rows),
        "tabs_and_padding": "\n".join("  " + r.replace(" ", "\t ") + "  " for r in rows) + "\n",
        "faces_after": "\n".join(rows) + "\n3 0 1 2\n3 2 3 4\n",
    }
    for name, body in variants.items():
        p = tmp_path / f"{name}.ply"
        after = "element face 2\nproperty list uchar int vertex_indices\n" if name == "faces_after" else ""
        p.write_bytes(header("ascii", n, [(k, "float") for k in "xyz"], extra_after=after) + body.encode("ascii"))
        assert same(read_ply(p)[0], want), name


def test_ascii_number_forms(tmp_path):
    p = tmp_path / "n.ply"
    body = "+1.5 -0.0 1e-400\n1E3 .5 5.\ninf -inf nan\n0x10 1 2\n"
    p.write_bytes(header("ascii", 3, [(k, "float") for k in "xyz"]) + body.encode())
    pts = read_ply(p)[0]
    assert pts[0].tolist() == [1.5, 0.0, 0.0] and np.signbit(pts[0, 1])
    assert pts[1].tolist() == [1000.0, 0.5, 5.0]
    assert pts[2, 0] == np.inf and pts[2, 1] == -np.inf and np.isnan(pts[2, 2])
    p.write_bytes(header("ascii", 4, [(k, "float") for k in "xyz"]) + body.encode())
    with pytest.raises(ValueError, match="malformed number in vertex 3"):
        read_ply(p)


@pytest.mark.parametrize("fmt", ["binary_little_endian", "binary_big_endian"])
def test_binary_mixed_types_match_numpy(tmp_path, fmt):
    n = 50_000
    e = "<" if fmt.endswith("little_endian") else ">"
    props = [("intensity", "ushort"), ("x", "double"), ("red", "uchar"), ("y", "float"), ("z", "int"), ("nx", "float"),
             ("ny", "float"), ("nz", "double"), ("green", "uchar"), ("blue", "uchar"), ("flag", "char")]
    dt = np.dtype([(k, e + _NP[t]) for k, t in props])
    rng = np.random.default_rng(3)
    rec = np.zeros(n, dt)
    for k, t in props:
        rec[k] = (rng.standard_normal(n) * 100).astype(_NP[t]) if _NP[t][0] in "iu" else rng.standard_normal(n)
    p = tmp_path / "b.ply"
    p.write_bytes(header(fmt, n, props, extra_after="element face 1\nproperty list uchar int vertex_indices\n")
                  + rec.tobytes() + struct.pack("<Biii", 3, 0, 1, 2))
    pts, nrm = read_ply(p)
    rp, rn = numpy_read_ply(p)
    assert same(pts, rp) and same(nrm, rn.astype(np.float32).astype(np.float64))
    info = probe_ply(p)
    assert (info.has_normals, info.has_colors, info.vertex_stride, info.format) == (1, 1, dt.itemsize, 1 if e == "<" else 2)


def test_elements_in_front_of_the_vertices_are_skipped(tmp_path):
    d = cloud(5, seed=1).astype(np.float32)
    before = "element camera 2\nproperty float a\nproperty list uchar short tags\n"
    cam = struct.pack("<fBhh", 1.0, 2, 7, 8) + struct.pack("<fB", 2.0, 0)
    p = tmp_path / "b.ply"
    p.write_bytes(header("binary_little_endian", 5, [(k, "float") for k in "xyz"], extra_before=before) + cam + d.tobytes())
    assert same(read_ply(p)[0], d.astype(np.float64))
    p.write_bytes(header("ascii", 5, [(k, "float") for k in "xyz"], extra_before=before)
                  + b"1.0 2 7 8\n2.0 0\n" + "\n".join("%.9g %.9g %.9g" % tuple(r) for r in d).encode() + b"\n")
    assert same(read_ply(p)[0].astype(np.float32), d)  # 9 significant digits identify an fp32


def test_error_behaviour(tmp_path):
    p = tmp_path / "e.ply"
    with pytest.raises(FileNotFoundError):
        read_ply(tmp_path / "missing.ply")
    p.write_bytes(b"solid stl\n")
    with pytest.raises(ValueError, match="not a PLY"):
        read_ply(p)
    p.write_bytes(b"")
    with pytest.raises(ValueError, match="not a PLY"):
        read_ply(p)
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\n")
    with pytest.raises(ValueError, match="end of PLY header"):
        read_ply(p)
    p.write_bytes(header("ascii", 2, [("x", "float"), ("y", "float")]) + b"1 2\n3 4\n")
    with pytest.raises(ValueError, match="lacks x/y/z"):
        read_ply(p)
    p.write_bytes(header("ascii", 2, [("x", "float"), ("y", "float"), ("z", "float")],
                         extra_after="property list uchar int bad\n") + b"1 2 3 0\n")
    with pytest.raises(ValueError, match="list properties on vertices"):
        read_ply(p)
    p.write_bytes(header("ascii", 3, [(k, "float") for k in "xyz"]) + b"1 2 3\n4 5\n")
    with pytest.raises(ValueError, match="truncated"):
        read_ply(p)
    p.write_bytes(header("binary_little_endian", 3, [(k, "float") for k in "xyz"]) + b"\0" * 35)
    with pytest.raises(ValueError, match="truncated"):
        read_ply(p)
    p.write_bytes(header("binary_middle_endian", 3, [(k, "float") for k in "xyz"]))
    with pytest.raises(ValueError, match="unsupported PLY format"):
        read_ply(p)
    p.write_bytes(header("ascii", 1, [("x", "quad"), ("y", "float"), ("z", "float")]) + b"1 2 3\n")
    with pytest.raises(ValueError, match="unknown property type"):
        read_ply(p)
    # empty clouds read as n = 0 (the Ply mirror turns that into the reference's ValueError, src/ply/ply.py:81-84)
    p.write_bytes(header("ascii", 0, [(k, "float") for k in "xyz"]))
    assert read_ply(p)[0].shape == (0, 3)
    p.write_bytes(b"ply\nformat ascii 1.0\nend_header\n")
    assert read_ply(p)[0].shape == (0, 3) and probe_ply(p).n_vertex == 0
    # a buffer that is too small is refused, not overrun
    import ctypes as C
    p.write_bytes(header("ascii", 2, [(k, "float") for k in "xyz"]) + b"1 2 3\n4 5 6\n")
    buf, err = np.zeros((1, 4), np.float32), C.create_string_buffer(256)
    rc = _capi.load().pcr_ply_read(os.fsencode(p), C.c_int64(1), C.c_void_p(buf.ctypes.data), None, None, 0, None, err, 256)
    assert rc == _capi.PCR_ERR_INVALID and b"buffer holds 1" in err.value and not buf.any()
    with pytest.raises(OSError):
        write_ply(tmp_path / "no_such_dir" / "x.ply", np.zeros((1, 3)))


@pytest.mark.parametrize("binary", [True, False])
def test_write_read_round_trip_is_bit_exact(tmp_path, binary):
    n = 40_000
    pts = cloud(n, seed=11).astype(np.float32)
    pts[0] = [np.float32(1e-45), np.float32(3.4028235e38), -0.0]  # denormal, max, signed zero
    nrm = cloud(n, seed=12).astype(np.float32)
    rgb = np.random.default_rng(1).integers(0, 256, (n, 3)).astype(np.uint8)
    p = tmp_path / "w.ply"
    write_ply(p, pts, nrm, binary=binary, colors=rgb)
    rp, rn = read_ply(p)
    assert same(rp.astype(np.float32), pts) and same(rn.astype(np.float32), nrm)
    a, b = numpy_read_ply(p)
    assert same(a.astype(np.float32), pts) and same(b.astype(np.float32), nrm)
    info = probe_ply(p)
    assert (info.n_vertex, info.has_normals, info.has_colors, info.format) == (n, 1, 1, int(binary))
    # points only, packed (n,4) input, float colours as Open3D keeps them
    write_ply(p, np.concatenate([pts, np.ones((n, 1), np.float32)], axis=1), binary=binary, colors=np.tile([1.0, 0.706, 0.0], (n, 1)))
    rp, rn = read_ply(p)
    assert same(rp.astype(np.float32), pts) and rn is None
    write_ply(p, pts[:0], binary=binary)
    assert read_ply(p)[0].shape == (0, 3)


def test_ascii_reader_throughput_is_reported(tmp_path, capsys):
    """Not a pass/fail timing: records the native reader next to numpy.loadtxt on a 100k-vertex ASCII file (the a1
    row of SURVEY §8 — 791 ms of `ply_loading` in the reference's benchmark_results.txt:6 include this)."""
    import time
    n = 100_000
    p = tmp_path / "t.ply"
    write_ply(p, cloud(n, seed=2), binary=False)
    t0 = time.perf_counter(); a = read_ply_xyzw(p, pin=False)[0].numpy(); t1 = time.perf_counter()
    b = numpy_read_ply(p)[0]; t2 = time.perf_counter()
    assert same(a[:, :3], b.astype(np.float32))
    with capsys.disabled():
        print(f"\n[ply] 100k ASCII vertices: native {1e3 * (t1 - t0):.1f} ms, numpy.loadtxt {1e3 * (t2 - t1):.1f} ms")


def test_reader_survives_mutated_files(tmp_path):
    """Robustness: truncations, byte flips and spliced garbage in valid files either decode or return an error code —
    the mapped input is never read out of bounds and nothing throws across the C ABI (a crash would take pytest down)."""
    import ctypes as C
    lib = _capi.load()
    rng = np.random.default_rng(1234)
    n = 300
    pts, nrm = cloud(n, 1).astype(np.float32), cloud(n, 2).astype(np.float32)
    seeds = []
    for binary in (True, False):
        p = tmp_path / f"seed{int(binary)}.ply"
        write_ply(p, pts, nrm, binary=binary, colors=np.zeros((n, 3), np.uint8))
        seeds.append(p.read_bytes())
    be = header("binary_big_endian", n, [("x", "double"), ("y", "short"), ("z", "float")],
                extra_before="element cam 3\nproperty list uchar int k\nproperty float a\n")
    seeds.append(be + bytes(rng.integers(0, 256, 3 * 9 + n * 14, dtype=np.uint8)))
    buf = np.zeros((n + 8, 4), np.float32)
    nbuf = np.zeros((n + 8, 4), np.float32)
    err = C.create_string_buffer(256)
    info = _capi.PlyInfo()
    f = tmp_path / "m.ply"
    outcomes = {0: 0, _capi.PCR_ERR_INVALID: 0}
    for it in range(1500):
        data = bytearray(seeds[it % len(seeds)])
        kind = it % 5
        if kind == 0:
            data = data[: int(rng.integers(0, len(data)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 8))):
                data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
        elif kind == 2:  # flips confined to the header
            hdr_end = data.find(b"end_header") + 11
            for _ in range(int(rng.integers(1, 4))):
                data[int(rng.integers(0, hdr_end))] = int(rng.integers(32, 127))
        elif kind == 3:  # a count that no longer matches the body
            data = data.replace(b"element vertex 300", b"element vertex %d" % int(rng.integers(0, 10**12)))
        else:
            a = int(rng.integers(0, len(data)))
            data[a:a] = bytes(rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8))
        f.write_bytes(bytes(data))
        rc = lib.pcr_ply_read(os.fsencode(f), C.c_int64(len(buf)), C.c_void_p(buf.ctypes.data), C.c_void_p(nbuf.ctypes.data),
                              None, 0, C.byref(info), err, 256)
        assert rc in (0, _capi.PCR_ERR_INVALID), (it, rc, err.value)
        assert (rc == 0) or err.value, it
        outcomes[rc] += 1
    assert outcomes[0] > 50 and outcomes[_capi.PCR_ERR_INVALID] > 300, outcomes


def test_ascii_numbers_are_correctly_rounded(tmp_path):
    """Every decimal form goes to the double Python's float() gives (correct rounding) — the exact fast path (<= 15
    digits, |exponent| <= 22) and the general routine alike."""
    rng = np.random.default_rng(7)
    toks = []
    for _ in range(6000):
        kind = int(rng.integers(0, 8))
        x = float(rng.standard_normal() * 10.0 ** int(rng.integers(-8, 9)))
        if kind == 0:
            toks.append("%.9g" % x)
        elif kind == 1:
            toks.append("%.17g" % x)                      # 17 digits: general routine
        elif kind == 2:
            toks.append("%.6f" % x)
        elif kind == 3:
            toks.append("%.3e" % x)
        elif kind == 4:
            toks.append("%d" % int(x))
        elif kind == 5:
            toks.append("%.15g" % x)                      # 15 digits: the edge of the fast path
        elif kind == 6:
            toks.append(("%.5f" % rng.standard_normal()) + "e%+d" % int(rng.integers(-30, 31)))  # exponents around +-22
        else:
            toks.append("%se%d" % (int(rng.integers(1, 10**15)), int(rng.integers(-40, 40))))
    toks += ["0", "-0", "0.0", "-0.000", "000123.4500", ".5", "5.", "-.25", "1e22", "1e23", "1e-22", "1e-23",
             "9007199254740993", "999999999999999", "1000000000000000", "4.9e-324", "1.7976931348623157e308", "1e400",
             "123456789012345678901234567890", "0.1e1", "1E5", "+3.5"]
    while len(toks) % 3:
        toks.append("1")
    p = tmp_path / "n.ply"
    n = len(toks) // 3
    body = "\n".join(" ".join(toks[3 * i: 3 * i + 3]) for i in range(n)) + "\n"
    p.write_bytes(header("ascii", n, [(k, "double") for k in "xyz"]) + body.encode())
    got = read_ply(p)[0].reshape(-1)
    want = np.array([float(t) for t in toks])
    assert got.tobytes() == want.tobytes(), [(t, g, w) for t, g, w in zip(toks, got, want) if not (g == w and np.signbit(g) == np.signbit(w))][:5]


def test_reader_is_reentrant_across_threads(tmp_path):
    """ctypes releases the GIL: concurrent reads of different files from Python threads (the GUI calls the matcher from
    worker threads, _visualize_matcher.py:264,275,292) do not interfere."""
    import threading
    files = []
    for i in range(6):
        p = tmp_path / f"f{i}.ply"
        pts = cloud(40_000 + 1000 * i, seed=100 + i).astype(np.float32)
        write_ply(p, pts, binary=bool(i % 2))
        files.append((p, pts))
    bad = []

    def work(k):
        for rep in range(4):
            p, pts = files[(k + rep) % len(files)]
            got = read_ply_xyzw(p, pin=False)[0].numpy()[:, :3]
            if not same(got, pts):
                bad.append((k, rep))

    th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad, bad


def test_probe_rejects_a_vertex_count_the_file_cannot_hold(tmp_path):
    """readers allocate (pinned) memory from the header's vertex count; a count that the file's size rules out
    must fail in pcr_ply_probe, before anything is allocated."""
    import ctypes as C
    from pcr_b200 import _capi
    lib = _capi.load()
    info = _capi.PlyInfo()
    err = C.create_string_buffer(256)
    hdr = b"ply\nformat binary_little_endian 1.0\nelement vertex 4000000000\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
    f = tmp_path / "huge.ply"
    f.write_bytes(hdr + b"\0" * 48)
    assert lib.pcr_ply_probe(str(f).encode(), C.byref(info), err, 256) == _capi.PCR_ERR_INVALID and b"4000000000" in err.value
    g = tmp_path / "huge_ascii.ply"
    g.write_bytes(hdr.replace(b"binary_little_endian", b"ascii") + b"0 0 0\n")
    assert lib.pcr_ply_probe(str(g).encode(), C.byref(info), err, 256) == _capi.PCR_ERR_INVALID
    ok = tmp_path / "ok.ply"
    ok.write_bytes(hdr.replace(b"4000000000", b"4") + b"\0" * 48)
    assert lib.pcr_ply_probe(str(ok).encode(), C.byref(info), err, 256) == 0 and info.n_vertex == 4
