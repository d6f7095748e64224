"""The C-ABI library loads on a CPU-only host and exports every symbol include/pcr.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pcr.h")).read()
    return sorted(set(re.findall(r"PCR_API\s+[\w\s\*]+?\b(pcr_\w+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    from pcr_b200 import _capi
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_capi.EXPORTS) == syms  # the Python binding list is in sync with the header


def test_struct_layouts_match_header():
    from pcr_b200 import _capi
    assert ctypes.sizeof(_capi.HypRecord) == 128
    assert ctypes.sizeof(_capi.RegResult) == 16 * 8 + 2 * 8 + 2 * 8 + 4 * 4 + 4 * 8
    lib = _capi.load()
    assert lib.pcr_version() >= 100
    p = _capi.AlignParams()
    lib.pcr_align_default_params(ctypes.byref(p))
    # reference defaults: Ply voxel 0.3 (ply.py:32), iteration 30 / confidence 0.999 (ransac.py:24,58), ICP 30 (A.7)
    assert (p.voxel_size, p.ransac_max_iter, p.ransac_confidence, p.icp_max_iter) == (0.3, 30, 0.999, 30)
    assert p.icp_rel_fitness == 1e-6 and p.icp_rel_rmse == 1e-6


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, not fall back, when no CUDA device is present."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pcr_b200.engine import Engine, get_engine
    with pytest.raises(RuntimeError):
        Engine(0)
    with pytest.raises(RuntimeError):
        get_engine()
    from pcr_b200 import align
    import numpy as np
    with pytest.raises(RuntimeError):
        align(np.zeros((10, 3)), np.zeros((10, 3)), 0.1)


def test_product_never_imports_oracle():
    """Nothing under 3d-matching_b200/ may import, include, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "3d-matching_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(libpcr_oracle)|(CDLL\([^)]*oracle)", re.M)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert not bad.search(txt), os.path.join(d, f)


def test_header_is_plain_c_and_links(tmp_path):
    """include/pcr.h compiles as C11 (-pedantic) and a C program links libpcr_b200.so and calls the host-only exports
    (PLY I/O, RANSAC replay, defaults) — the drop-in boundary is a C ABI, not a C++ one."""
    import shutil
    import subprocess
    from pcr_b200 import _capi
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-L", libdir, "-lpcr_b200", f"-Wl,-rpath,{libdir}",
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path / "scratch.ply")], capture_output=True, text=True)
    assert r.returncode == 0 and "abi_check ok" in r.stdout, r.stderr
