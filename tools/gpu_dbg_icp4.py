import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
n = int(os.environ.get("N", "2000")); iters = int(os.environ.get("ITERS", "50"))
s1, t1, _ = synth.make_icp_pair(n, v, 20243)
d1s, d1t = eng.pack(s1), eng.pack(t1)
nrm = eng.estimate_normals(d1t, 2 * v, 30)
for _ in range(2): eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 5, 0.0, 0.0)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), iters, 0.0, 0.0); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print({k: os.environ.get(k) for k in ("PCR_DBG_SKIPNN", "PCR_DBG_SKIPACC")}, "iters", iters, "n", n, "call ms %.3f -> %.1f us/pass" % (best, best * 1e3 / (iters + 1)))
