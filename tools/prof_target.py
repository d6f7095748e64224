"""Small workload for ncu captures: one fixed-work 100k alignment (+ optional 1M-point ICP passes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
p = eng.default_params(v); p.ransac_max_iter = 100000; p.ransac_confidence = 1.0; p.seed = 7; p.icp_max_iter = 4; p.icp_rel_fitness = 0.0; p.icp_rel_rmse = 0.0
for _ in range(2):
    r = eng.align_host(src, tgt, p)
print("align ok", r.icp.fitness)
if os.environ.get("ICP1M", "1") == "1":
    s1, t1, _ = synth.make_icp_pair(1000000, v, 20243)
    d1s, d1t = eng.pack(s1), eng.pack(t1)
    nrm = eng.estimate_normals(d1t, 2 * v, 30)
    g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), int(os.environ.get("ICP1M_ITERS", "3")), 0.0, 0.0)
    print("icp1m ok", g.fitness)
