"""EXPERIMENTAL candidate lists for RANSAC validation (3d-matching_b200/csrc/pcr_celllists.cuh, off unless
PCR_VAL_LISTS is set): the per-cell build logic is __host__ __device__, so the code the kernel runs is exercised here on
the CPU — grid laid out as pcr_grid.cu does, lists built for every fine cell, and 40,000 list-based radius-limited
nearest-neighbour queries compared with brute force (index and fp32 distance bits).  No GPU involved."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from pcr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else shutil.which("nvcc")


@pytest.mark.skipif(NVCC is None, reason="nvcc not found")
@pytest.mark.parametrize("div", [2, 3])
def test_candidate_lists_equal_brute_force_on_the_host(tmp_path, orc, div):
    exe = tmp_path / "celllists_check"
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([NVCC, "-ccbin", host_cxx, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                        "-I", os.path.join(ROOT, "3d-matching_b200", "csrc"), "-o", str(exe),
                        os.path.join(ROOT, "tests", "c", "celllists_host_check.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    v = 0.005
    _, tgt, _ = synth.make_pair(40000, v, 77)
    td = orc.voxel_downsample(tgt, v)
    pts = tmp_path / "td.f32"
    np.ascontiguousarray(td, np.float32).tofile(pts)
    r = subprocess.run([str(exe), str(pts), str(len(td)), repr(1.5 * v), str(div), "40000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
    import re
    hits, listed = (int(re.search(rf"{k} (\d+)", r.stdout).group(1)) for k in ("hits", "listed"))
    assert hits > 10000 and listed > 20000  # the lists, not the fallback, answered
