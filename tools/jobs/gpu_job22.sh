#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -12
echo "== icp trace 100k"; N=100000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -8
echo "== icp 1M no trace"; N=1000000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
echo "== icp 100k no trace"; N=100000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
echo "== bench no-cpu no-aux"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux > gpurun_out/j22_bench.json 2> gpurun_out/j22_bench.err; echo "rc=$?"; tail -3 gpurun_out/j22_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j22_bench.json") if l.startswith("{")][-1]
a=d.pop("aux")
print("e2e", d["e2e"]["value"], "value", d["value"], "launches", d["gpu_launches"])
print({k:v for k,v in a.items() if "normals" in k})
PY
