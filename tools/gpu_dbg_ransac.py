import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
ds, dt = eng.pack(src), eng.pack(tgt)
sd, td = eng.voxel_downsample(ds, v).contiguous(), eng.voxel_downsample(dt, v).contiguous()
sn, tn = eng.estimate_normals(sd, 2 * v, 30), eng.estimate_normals(td, 2 * v, 30)
sf, tf = eng.compute_fpfh(sd, sn, 5 * v, 100), eng.compute_fpfh(td, tn, 5 * v, 100)
corr = eng.match_features(sf, tf, True).contiguous()
H = int(os.environ.get("HYPS", "100000"))
for rep in range(2):
    print("--- run", rep, file=sys.stderr)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = eng.ransac(sd, td, corr, 1.5 * v, H, 1.0, 7); b.record(); torch.cuda.synchronize()
    print("ransac ms", a.elapsed_time(b), "best", r.best_hyp, "count", r.inlier_count, "surv", r.survivors, file=sys.stderr)
