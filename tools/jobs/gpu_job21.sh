#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "rc=$?"
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2_bench.json") if l.startswith("{")][-1]
a=d.pop("aux")
print(json.dumps({k:d[k] for k in d if k not in ("config","result","clocks")})[:2500])
print(a["kernel_ms_per_step"]); print(a["batch"]); print(a["ransac"]); print(a["icp_1m"])
r=[json.loads(l) for l in open("gpurun_out/r2_bench_reference.json") if l.startswith("{")][-1]
print("reference", r["value"], r["cpu_baseline"])
PY
