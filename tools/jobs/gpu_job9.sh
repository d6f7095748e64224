#!/usr/bin/env bash
set -u
for i in 1 2 3 4 5 6; do
  timeout 120 python bench.py --no-cpu --no-aux --steps 20 > gpurun_out/j9_$i.json 2> gpurun_out/j9_$i.err; echo "bench $i rc=$?"; tail -2 gpurun_out/j9_$i.err
done
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/j9_1.json") if l.startswith("{")][-1]
print("e2e",d["e2e"]["value"],"value",d["value"]); print(d["aux"]["kernel_ms_per_step"]); print(d["aux"]["stage_ms_device"])
PY
