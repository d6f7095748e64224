"""estimate_normals timing on the 100k cfg-2 cloud (full resolution) and its down-sampled cloud, standalone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(int(os.environ.get("N", "100000")), v, 20242)
d = eng.pack(src)
dd = eng.voxel_downsample(d, v).contiguous()
for name, x in (("full", d), ("down", dd)):
    eng.estimate_normals(x, 2 * v, 30)
    eng.set_profiling(True); eng.kernel_stats(reset=True)
    for _ in range(10):
        eng.estimate_normals(x, 2 * v, 30)
    ks = eng.kernel_stats(reset=True); eng.set_profiling(False)
    print(name, x.shape[0], {k: round(v_["ms"] / 10 * 1e3, 1) for k, v_ in ks.items()}, "us per call")
