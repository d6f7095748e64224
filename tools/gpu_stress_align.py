"""Determinism / race stress: the same alignment many times (overlap on), sequentially and from 3 threads; every result must
be bit-identical to the first."""
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
v = 0.005
src, tgt, _ = synth.make_pair(int(os.environ.get("N", "30000")), v, 4711)
reps = int(os.environ.get("REPS", "150"))
def run(eng, out):
    ds, dt = eng.pack(src), eng.pack(tgt)
    p = eng.default_params(v); p.ransac_max_iter = 30000; p.seed = 5
    for _ in range(reps):
        r = eng.align_device(ds, dt, p)
        out.append((tuple(r.icp.transformation), r.icp.fitness, r.icp.inlier_rmse, r.ransac.best_hyp, r.icp.iterations))
a = []
run(Engine(0), a)
assert all(x == a[0] for x in a), "sequential runs differ"
outs = [[] for _ in range(3)]
ths = [threading.Thread(target=lambda k=k: (torch.cuda.set_device(0), run(Engine(0), outs[k]))) for k in range(3)]
[t.start() for t in ths]; [t.join() for t in ths]
for o in outs:
    assert len(o) == reps and all(x == a[0] for x in o), "threaded runs differ"
print("stress ok:", reps, "sequential +", 3 * reps, "threaded alignments identical; fitness", a[0][1], "best_hyp", a[0][3])
