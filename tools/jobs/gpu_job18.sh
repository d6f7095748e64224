#!/usr/bin/env bash
set -u
N=${1:-8}
echo "== dist check N=$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py 2>&1 | grep "rank 0/" | tail -6
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; tail -2 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_${N}gpu.json") if l.startswith("{")][-1]
print({k:d[k] for k in ("value","n_gpus","ransac_hyp_per_s","ransac_identical_to_oracle_golden_10m","batch_pairs_per_s","batch_identical_to_oracle","icp_iters_per_s_1m","icp_1m_identical_to_oracle")}); print(d["e2e"]); print(d["aux"]["batch"]); print(d["aux"]["ransac"])
PY
