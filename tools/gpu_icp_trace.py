"""ICP per-pass timing on a synthetic pair (N points, ITERS fixed iterations); PCR_ICP_TRACE=1 adds the kernel's own cycle
counters (point loop / reduce + barrier / end-of-pass solve, slowest CTA per pass).  usage: N=1000000 ITERS=50 PCR_ICP_TRACE=1 python tools/gpu_icp_trace.py"""
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
print("iters", iters, "n", n, "call ms %.3f -> %.1f us/pass" % (best, best * 1e3 / (iters + 1)))
