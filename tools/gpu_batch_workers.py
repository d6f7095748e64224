"""pcr_align_batch throughput on one GPU for several worker counts (50k-point pairs, reference-default criteria)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
if os.environ.get("CORES"):  # emulate a box with fewer cores than host threads (several ranks sharing the CPUs)
    os.sched_setaffinity(0, set(range(int(os.environ["CORES"]))))
from pcr_b200 import synth
from pcr_b200.engine import get_engine
eng = get_engine(0); eng.comm_init()
v = 0.005
B = int(os.environ.get("B", "96"))
pairs = []
for i in range(B):
    s, t, _ = synth.make_pair(50000, v, 30000 + i)
    pairs.append((eng.pack(s), eng.pack(t)))
p = eng.default_params(v); p.ransac_max_iter = 100000; p.seed = 7
print("host cores", len(os.sched_getaffinity(0)))
ref = None
for w in [int(x) for x in os.environ.get("WORKERS", "1,2,3,4,6,8").split(",")]:
    eng.align_batch(pairs, p, B, workers=w)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tab = eng.align_batch(pairs, p, B, workers=w)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if ref is None: ref = tab
    print(f"workers {w}: {B / dt:.0f} pairs/s, identical {np.array_equal(tab, ref)}", flush=True)
