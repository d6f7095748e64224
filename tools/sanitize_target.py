"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): 6k-point pair, short RANSAC, 8 ICP passes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(6000, v, 77)
p = eng.default_params(v); p.ransac_max_iter = 6000; p.ransac_confidence = 1.0; p.seed = 3; p.icp_max_iter = 8
p.icp_rel_fitness = 0.0; p.icp_rel_rmse = 0.0
r = eng.align_host(src, tgt, p)
print("align ok", r.icp.fitness, r.ransac.best_hyp)
