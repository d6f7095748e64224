"""Poor man's timeline of ONE alignment (cfg2 pair, fixed work): PCR_TIMELINE=1 makes pcr_kernel_stats print the start offset and
duration of every timed kernel scope on the main and the helper context.  usage: PCR_TIMELINE=1 python tools/gpu_timeline.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
p = eng.default_params(v); p.ransac_max_iter = 100000; p.ransac_confidence = 1.0; p.seed = 7; p.icp_max_iter = 50; p.icp_rel_fitness = 0.0; p.icp_rel_rmse = 0.0
ds, dt = eng.pack(src), eng.pack(tgt)
for _ in range(3):
    r = eng.align_device(ds, dt, p)
eng.set_profiling(True); eng.kernel_stats(reset=True)
r = eng.align_device(ds, dt, p)
eng.kernel_stats(reset=True)
print("stage ms", list(r.stage_ms))
