#!/usr/bin/env bash
set -u
echo "== tests"; timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j11_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/j11_pytest.log
echo "== bench"; timeout 300 python bench.py --no-cpu --no-aux --steps 20 > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err; echo "rc=$?"; tail -3 gpurun_out/j11_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j11_bench.json") if l.startswith("{")][-1]
print("e2e",d["e2e"]["value"],"value",d["value"]); print(d["aux"]["kernel_ms_per_step"]); print(d["aux"]["stage_ms_device"])
PY
