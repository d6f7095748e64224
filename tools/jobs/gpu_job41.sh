#!/usr/bin/env bash
set -u
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f default-criteria e2e %.3f'%(d['e2e']['value'],d['value'],d['aux']['align_ms_reference_default_criteria_e2e']), d['aux']['stage_ms_device'])"; }
echo "== concurrent"; b; b
echo "== serial"; PCR_PRE_CONCURRENT=0 b
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j41_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j41_pytest_gpu.log
PCR_TIMELINE=1 python tools/gpu_timeline.py 2>&1 | grep "timeline" | head -24
