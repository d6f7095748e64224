"""First GPU slice check: grid 1-NN and point-to-plane ICP vs the CPU oracle (run under gpurun)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import Engine

eng = Engine(0)
out = {}
for n, v in ((20000, 0.005), (100000, 0.005)):
    src, tgt, T = synth.make_icp_pair(n, v, 20243)
    r = 0.4 * v
    ds, dt = eng.pack(src), eng.pack(tgt)
    idx, d2 = eng.nn1(dt, ds, r)
    oi, od = orc.nn1(tgt, src, r)
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    print(n, "nn1 idx equal:", np.array_equal(idx, oi), "d2 equal:", np.array_equal(d2, od), "found", (oi >= 0).mean())
    tn = orc.estimate_normals(tgt, 2 * v, 30)
    dn = eng.pack(tn)
    t0 = time.time(); o = orc.icp_point_to_plane(src, tgt, tn, r, np.eye(4), 30); t_or = time.time() - t0
    g, corr = eng.icp_point_to_plane(ds, dt, dn, r, np.eye(4), 30)
    print(n, "icp oracle:", o.fitness, o.inlier_rmse, o.iterations, o.converged, "t=%.3f" % t_or)
    print(n, "icp device:", g.fitness, g.inlier_rmse, g.iterations, g.converged)
    print(n, "T bit-equal:", np.array_equal(g.transformation, o.transformation), "max|dT|", np.abs(g.transformation - o.transformation).max(),
          "count eq", g.inlier_count == o.inlier_count, "sumq eq", g.sum_d2_fixed == o.sum_d2_fixed,
          "corr eq", np.array_equal(corr.cpu().numpy(), o.correspondence))
    g2, _ = eng.icp_point_to_plane(ds, dt, dn, r, np.eye(4), 10, 0.0, 0.0)
    o2 = orc.icp_point_to_plane(src, tgt, tn, r, np.eye(4), 10, 0.0, 0.0)
    print(n, "fixed-10: T bit-equal:", np.array_equal(g2.transformation, o2.transformation), g2.iterations, o2.iterations,
          np.abs(g2.transformation - o2.transformation).max())

# timing at 1M, 50 fixed passes
n, v = 1000000, 0.005
src, tgt, T = synth.make_icp_pair(n, v, 20243)
ds, dt = eng.pack(src), eng.pack(tgt)
nrm = tgt / np.linalg.norm(tgt, axis=1, keepdims=True)   # radial pseudo-normals are fine for timing
dn = eng.pack(nrm.astype(np.float32))
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g, corr = eng.icp_point_to_plane(ds, dt, dn, 0.4 * v, np.eye(4), 50, 0.0, 0.0); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("1M icp 51 passes: %.3f ms total -> %.1f us/pass (incl. grid build), fitness %.4f it %d" % (ms, ms * 1e3 / 51, g.fitness, g.iterations))
print("launches", eng.launch_count())
