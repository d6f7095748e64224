#!/usr/bin/env bash
# One gpurun call that re-establishes the state on a fresh box and tries the prepared experimental path:
#   /usr/local/graft/bin/gpurun --timeout 240 -- 'bash tools/gpu_round_start.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).  Every step has its own timeout; nothing here changes clocks.
set -u
mkdir -p gpurun_out
echo "== gpu tests";        timeout 90 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/rs_pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/rs_pytest_gpu.log
echo "== smoke";            timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rs_smoke.log 2>&1; echo "rc=$?"
echo "== bench (default)";  timeout 120 python bench.py > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err; echo "rc=$?"
echo "== experimental: candidate lists vs oracle (div 0 / 2 / 3, with timings)"
PCR_RUN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -m gpu -q -s -p no:cacheprovider > gpurun_out/rs_experimental.log 2>&1; echo "rc=$?"; grep -E "ms per RANSAC|passed|failed|Error" gpurun_out/rs_experimental.log | tail -20
echo "== bench with the lists (only meaningful if the step above passed)"
PCR_VAL_LISTS=1 timeout 120 python bench.py --no-cpu > gpurun_out/rs_bench_lists.json 2> gpurun_out/rs_bench_lists.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/rs_bench.json", "gpurun_out/rs_bench_lists.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        k = d["aux"]["kernel_ms_per_step"]
        print(f, "e2e", round(d["e2e"]["value"], 3), "ms; ransac_validate", round(k.get("ransac_validate", 0), 3),
              "grid_build", round(k.get("grid_build", 0), 3), "; RANSAC 10M:", round(d["aux"]["ransac"]["hyp_per_s"] / 1e6, 1), "M hyp/s")
    except Exception as e:
        print(f, "unreadable:", e)
PY
