"""ICP-only workload for ncu: launch 0 = 100k points, 50 iterations; launch 1 = 1M points, 50 iterations (-k regex:k_icp_persist -c 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
for n, iters in ((100000, 50), (1000000, 50)):
    s1, t1, _ = synth.make_icp_pair(n, v, 20243)
    d1s, d1t = eng.pack(s1), eng.pack(t1)
    nrm = eng.estimate_normals(d1t, 2 * v, 30)
    g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), iters, 0.0, 0.0)
    print("icp ok", n, iters, g.fitness)
