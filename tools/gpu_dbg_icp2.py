import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import synth
from pcr_b200.engine import Engine
from pcr_b200.dist import ransac_multi_gpu
eng = Engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
ds, dt = eng.pack(src), eng.pack(tgt)
sd, td = eng.voxel_downsample(ds, v).contiguous(), eng.voxel_downsample(dt, v).contiguous()
sn, tn = eng.estimate_normals(sd, 2 * v, 30), eng.estimate_normals(td, 2 * v, 30)
sf, tf = eng.compute_fpfh(sd, sn, 5 * v, 100), eng.compute_fpfh(td, tn, 5 * v, 100)
corr = eng.match_features(sf, tf, True).contiguous()
t0=time.perf_counter(); r, st = ransac_multi_gpu(eng, sd, td, corr, 1.5 * v, 2000000, 1.0, 7); print("ransac", time.perf_counter()-t0, st, flush=True)
s1, t1, _ = synth.make_icp_pair(1000000, v, 20243)
d1s, d1t = eng.pack(s1), eng.pack(t1)
t0=time.perf_counter(); nrm = eng.estimate_normals(d1t, 2 * v, 30); torch.cuda.synchronize(); print("normals 1M", time.perf_counter()-t0, flush=True)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 5 if i == 0 else 50, 0.0, 0.0, want_corr=(i > 0))
    torch.cuda.synchronize(); print("icp", i, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
