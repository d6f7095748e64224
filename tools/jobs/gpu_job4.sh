#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tc tests"; timeout 180 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -m gpu -q -x -p no:cacheprovider -k "tensor or match or nan or pipeline or feature" > gpurun_out/j4_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/j4_pytest.log
echo "== bench"; timeout 300 python bench.py --no-cpu --no-aux > gpurun_out/j4_bench.json 2> gpurun_out/j4_bench.err; echo "rc=$?"; tail -3 gpurun_out/j4_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j4_bench.json") if l.startswith("{")][-1]
print("e2e",d["e2e"]["value"],"value",d["value"]); print(d["aux"]["kernel_ms_per_step"]); print(d["aux"]["stage_ms_device"])
PY
