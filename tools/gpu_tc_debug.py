"""Bring-up aid for pcr_match_tc.cu at bench sizes: random and real FPFH descriptors against the oracle, with timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import Engine
eng = Engine(0)
orc.build()
mode = sys.argv[1] if len(sys.argv) > 1 else "random"
if mode == "random":
    rng = np.random.default_rng(0)
    fs = rng.uniform(0, 100, (int(os.environ.get("NQ", "9245")), 33)).astype(np.float32)
    ft = rng.uniform(0, 100, (int(os.environ.get("NB", "9170")), 33)).astype(np.float32)
else:
    v = 0.005
    src, tgt, _ = synth.make_pair(100000, v, 20242)
    S, G = orc.preprocess(src, v, full_normals=False), orc.preprocess(tgt, v, full_normals=False)
    fs, ft = S.pcd_fpfh, G.pcd_fpfh
dfs, dft = torch.from_numpy(fs).cuda(), torch.from_numpy(ft).cuda()
print("sizes", fs.shape, ft.shape, flush=True)
for a, b, x, y in ((dfs, dft, fs, ft), (dft, dfs, ft, fs)):
    nn = eng.nn_features(a, b)
    torch.cuda.synchronize()
    print("ran", flush=True)
    t0 = time.perf_counter()
    for _ in range(10):
        nn = eng.nn_features(a, b)
    torch.cuda.synchronize()
    print("ms per call", (time.perf_counter() - t0) * 100, "equal to oracle:", np.array_equal(nn.cpu().numpy(), orc.nn_features(x, y)), flush=True)
