"""Bring-up check of the tcgen05 matching kernel against the exact CUDA-core kernel / oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
print("swap", os.environ.get("PCR_TC_SWAP_LBO_SBO"), "exact", os.environ.get("PCR_MATCH_EXACT"), flush=True)
rng = np.random.default_rng(0)
for nq, nb in ((128, 512), (1000, 3000), (9245, 9170)):
    fs = rng.uniform(0, 200, (nq, 33)).astype(np.float32); ft = rng.uniform(0, 200, (nb, 33)).astype(np.float32)
    ft[7] = fs[3]; ft[9] = fs[3]; fs[50] = 0; ft[60] = 0; ft[70] = 0; ft[80] = 0; ft[90] = 0; ft[95] = 0
    dfs, dft = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
    eng.set_profiling(True); eng.kernel_stats(reset=True)
    nn = eng.nn_features(dfs, dft).cpu().numpy()
    ks = eng.kernel_stats(); eng.set_profiling(False)
    ref = orc.nn_features(fs, ft)
    print(nq, nb, "equal", np.array_equal(nn, ref), "mismatch", int((nn != ref).sum()), {k: round(v["ms"], 3) for k, v in ks.items()}, flush=True)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
ds, dt = eng.pack(src), eng.pack(tgt)
sd, td = eng.voxel_downsample(ds, v).contiguous(), eng.voxel_downsample(dt, v).contiguous()
sf = eng.compute_fpfh(sd, eng.estimate_normals(sd, 2 * v, 30), 5 * v, 100); tf = eng.compute_fpfh(td, eng.estimate_normals(td, 2 * v, 30), 5 * v, 100)
for rep in range(3):
    eng.set_profiling(True); eng.kernel_stats(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    corr = eng.match_features(sf, tf, True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ks = eng.kernel_stats(); eng.set_profiling(False)
ref = orc.match_features(sf.cpu().numpy(), tf.cpu().numpy(), True)
print("fpfh mutual match equal", np.array_equal(corr.cpu().numpy(), ref), len(ref), "wall ms %.3f" % ((t1 - t0) * 1e3), {k: round(v["ms"], 3) for k, v in ks.items()})
