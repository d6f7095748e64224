#!/usr/bin/env bash
set -u
echo "== test"; timeout 300 python -m pytest tests/test_gpu_boundary.py -m gpu -q -x -p no:cacheprovider -k c_side 2>&1 | tail -3
echo "== 2 ranks"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -12
echo "== bench 2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cfg1 > gpurun_out/j15_bench2.json 2> gpurun_out/j15_bench2.err; echo "rc=$?"; tail -3 gpurun_out/j15_bench2.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j15_bench2.json") if l.startswith("{")][-1]
print({k:d[k] for k in ("value","ransac_hyp_per_s","ransac_identical_to_oracle_golden_10m","batch_pairs_per_s","batch_identical_to_oracle","icp_iters_per_s_1m","icp_1m_identical_to_oracle")}); print(d["aux"]["batch"]); print(d["aux"]["ransac"])
PY
