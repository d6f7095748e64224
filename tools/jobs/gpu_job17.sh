#!/usr/bin/env bash
set -u
timeout 100 python tools/gpu_dist_check.py 2>&1 | grep schedule
timeout 300 python bench.py --no-cpu --no-aux --steps 20 > gpurun_out/j17_bench.json 2> gpurun_out/j17_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j17_bench.json") if l.startswith("{")][-1]
print("e2e",d["e2e"]["value"],"value",d["value"]); print(d["aux"]["kernel_ms_per_step"]); print(d["aux"]["stage_ms_device"])
PY
timeout 200 python -m pytest tests/test_gpu_ransac_lists.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
