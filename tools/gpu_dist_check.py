"""Multi-GPU check of the C-side sharding (torchrun, one process per GPU): pcr_ransac_multi must return the committed
oracle golden (tests/golden/cfg4_ransac_10m.json) on every rank for every world size; pcr_align_batch must return the
same table on every rank and equal a sequential single-context run of the rank's own pairs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch, torch.distributed as dist
from pcr_b200 import synth
from pcr_b200.engine import get_engine
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
assert eng.comm_init() == world
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20242)
ds, dt = eng.pack(src), eng.pack(tgt)
sd, td = eng.voxel_downsample(ds, v).contiguous(), eng.voxel_downsample(dt, v).contiguous()
sf = eng.compute_fpfh(sd, eng.estimate_normals(sd, 2 * v, 30), 5 * v, 100)
tf = eng.compute_fpfh(td, eng.estimate_normals(td, 2 * v, 30), 5 * v, 100)
corr = eng.match_features(sf, tf, True).contiguous()
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg4_ransac_10m.json")))
H = int(os.environ.get("H", str(gold["hypotheses"])))
eng.ransac_multi(sd, td, corr, 1.5 * v, H, 1.0, 7)
for first, growth in ((0, 0), (2048, 8), (16384, 4)):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r, waves = eng.ransac_multi(sd, td, corr, 1.5 * v, H, 1.0, 7, first_wave=first, growth=growth)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    ok = True
    if H == gold["hypotheses"]:
        Tg = np.array([float.fromhex(x) for x in gold["transformation_hex"]]).reshape(4, 4)
        ok = (r.best_hyp == gold["best_hyp"] and r.inlier_count == gold["inlier_count"] and r.sum_d2_fixed == gold["sum_d2_fixed"]
              and r.survivors == gold["survivors"] and r.hyp_evaluated == gold["hyp_evaluated"] and np.array_equal(r.transformation, Tg))
    print(f"rank {rank}/{world} schedule ({first},{growth}): {ms:.2f} ms, {H / ms / 1e3:.1f} M hyp/s, {waves} waves, golden {ok}", flush=True)
    assert ok
# batch
B = int(os.environ.get("B", "12"))
pairs = []
for i in range(rank, B * world, world):
    s, t, _ = synth.make_pair(50000, v, 30000 + i)
    pairs.append((eng.pack(s), eng.pack(t)))
p = eng.default_params(v); p.ransac_max_iter = 100000; p.seed = 7
tab1 = eng.align_batch(pairs, p, B * world, workers=1)
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
tab3 = eng.align_batch(pairs, p, B * world, workers=3)
torch.cuda.synchronize(); dt_ = time.perf_counter() - t0
assert np.array_equal(tab1, tab3)
digest = hashlib.sha256(tab3.tobytes()).hexdigest()[:16]
print(f"rank {rank}/{world} batch {B * world} pairs: {B * world / dt_:.0f} pairs/s (3 workers), table {digest}, min fitness {tab3[:, 16].min():.4f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
