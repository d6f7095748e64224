"""RANSAC validation through the per-fine-cell candidate lists (pcr_celllists.cu; the default since round 2) and through
the 27-cell grid walk it replaced (PCR_VAL_LISTS=0, still the fallback for over-long lists and oversized lattices):
both must reproduce the oracle bit for bit.  The switch is read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys, time
sys.path[:0] = [{root!r}, {pkg!r}]
import numpy as np, torch
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import get_engine
orc.build()
eng = get_engine(0)
v = 0.005
for n, seed, iters, conf in ((20000, 20241, 100000, 0.999), (20000, 5, 30000, 1.0), (3000, 9, 5000, 1.0)):
    src, tgt, T = synth.make_pair(n, v, seed)
    S, G = orc.preprocess(src, v, full_normals=False), orc.preprocess(tgt, v, full_normals=False)
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    want = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * v, iters, conf, seed=3)
    sd, td = eng.pack(S.pcd_down), eng.pack(G.pcd_down)
    dc = torch.from_numpy(np.ascontiguousarray(corr, np.int32)).to(eng.tdev)
    got = eng.ransac(sd, td, dc, 1.5 * v, iters, conf, 3, edge_sim=0.9)
    assert got.best_hyp == want.best_hyp and got.inlier_count == want.inlier_count, (got.best_hyp, want.best_hyp)
    assert got.sum_d2_fixed == want.sum_d2_fixed and got.hyp_evaluated == want.hyp_evaluated
    assert np.array_equal(got.transformation, want.transformation)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        eng.ransac(sd, td, dc, 1.5 * v, iters, conf, 3, edge_sim=0.9)
    torch.cuda.synchronize()
    print("n", n, "iters", iters, "ms per RANSAC", (time.perf_counter() - t0) / 5 * 1e3)
print("lists ok")
"""


@pytest.mark.parametrize("div", ["0", "1", "3"])
def test_ransac_through_candidate_lists_equals_the_oracle(div):
    """div 0 = the grid walk, 1 = lists with cells of v/2 (the default), 3 = cells of v/3."""
    env = dict(os.environ, PCR_VAL_LISTS=div)
    code = _SCRIPT.format(root=ROOT, pkg=os.path.join(ROOT, "3d-matching_b200"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=180)
    print(r.stdout)
    assert r.returncode == 0 and "lists ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
