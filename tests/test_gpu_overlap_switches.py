"""pcr_align overlaps independent stages (the two voxel grids, normals + FPFH of the two down-sampled clouds, the two
directions of the descriptor matching, generation of the second RANSAC wave beside the first wave's validation, the
source normals beside the first ICP passes, the full-resolution normals on a helper context).  None of it may change a
bit of the result: every switch that serialises a stage must reproduce the default run and the oracle.  The switches are
read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
import numpy as np
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import Engine
orc.build()
eng = Engine(0)
v = 0.005
for n, iters, conf, seed in ((8000, 20000, 0.999, 5), (20000, 6000, 1.0, 11)):
    src, tgt, _ = synth.make_pair(n, v, 977 + n)
    p = eng.default_params(v); p.ransac_max_iter = iters; p.ransac_confidence = conf; p.seed = seed
    r = None
    for _ in range(3):  # repeated: the overlapped stages race differently every time
        r2 = eng.align_host(src, tgt, p)
        assert r is None or (tuple(r2.icp.transformation) == tuple(r.icp.transformation) and r2.ransac.best_hyp == r.ransac.best_hyp)
        r = r2
    S, G = orc.preprocess(src, v), orc.preprocess(tgt, v)
    ro = orc.global_registration(S, G, v, iters, conf, seed)
    io = orc.refine_registration(S, G, ro.transformation, v)
    assert r.ransac.best_hyp == ro.best_hyp and r.ransac.hyp_evaluated == ro.hyp_evaluated, (r.ransac.best_hyp, ro.best_hyp)
    assert r.ransac.inlier_count == ro.inlier_count and r.ransac.sum_d2_fixed == ro.sum_d2_fixed
    assert r.icp.inlier_count == io.inlier_count and r.icp.iterations == io.iterations
    assert np.array_equal(np.array(r.icp.transformation).reshape(4, 4), io.transformation)
print("overlap ok")
"""


@pytest.mark.parametrize("switch", ["", "PCR_MATCH_CONCURRENT=0", "PCR_RANSAC_SPECULATE=0", "PCR_RANSAC_SPECULATE=2", "PCR_ICP_EARLY=0",
                                    "PCR_PRE_CONCURRENT=0", "PCR_ALIGN_OVERLAP=0"])
def test_alignment_is_identical_with_a_stage_serialised(switch):
    """(PCR_RANSAC_SPECULATE=2 speculates also where the run may stop early, i.e. with the 0.999 case of the script.)"""
    env = dict(os.environ)
    if switch:
        k, val = switch.split("=")
        env[k] = val
    code = _SCRIPT.format(root=ROOT, pkg=os.path.join(ROOT, "3d-matching_b200"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "overlap ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
