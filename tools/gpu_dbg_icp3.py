import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
n = int(os.environ.get("N", "1000000"))
s1, t1, _ = synth.make_icp_pair(n, v, 20243)
d1s, d1t = eng.pack(s1), eng.pack(t1)
nrm = eng.estimate_normals(d1t, 2 * v, 30)
for _ in range(2): eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 5, 0.0, 0.0)
eng.set_profiling(True); eng.kernel_stats(reset=True)
g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 50, 0.0, 0.0)
ks = eng.kernel_stats()["icp_pass"]; ks["launches"] = g.iterations + 1
print("dbg", os.environ.get("PCR_ICP_DBG"), "n", n, "pass us %.1f" % (ks["ms"] * 1e3 / ks["launches"]), "fitness", g.fitness)
