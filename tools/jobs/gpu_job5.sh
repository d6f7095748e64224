#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PCR_ALIGN_OVERLAP=0 timeout 300 python bench.py --no-cpu --no-aux > gpurun_out/j5_bench.json 2> gpurun_out/j5_bench.err; echo "rc=$?"; tail -3 gpurun_out/j5_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j5_bench.json") if l.startswith("{")][-1]
print("e2e",d["e2e"]["value"],"value",d["value"]); print(d["aux"]["kernel_ms_per_step"]); print(d["aux"]["stage_ms_device"])
PY
