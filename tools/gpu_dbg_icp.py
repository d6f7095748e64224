import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
s1, t1, _ = synth.make_icp_pair(1000000, v, 20243)
d1s, d1t = eng.pack(s1), eng.pack(t1)
nrm = eng.estimate_normals(d1t, 2 * v, 30)
torch.cuda.synchronize()
def run(tag, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 50, 0.0, 0.0, **kw)
    torch.cuda.synchronize(); print(tag, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), g.iterations, g.fitness, flush=True)
run("plain 1"); run("plain 2"); run("nocorr", want_corr=False)
eng.set_profiling(True); eng.kernel_stats(reset=True)
run("prof 1"); run("prof 2")
print(eng.kernel_stats())
eng.set_profiling(False)
run("plain 3")
