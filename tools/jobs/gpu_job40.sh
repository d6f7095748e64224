#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j40_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j40_pytest_gpu.log
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f ransac stage %.3f validate %.3f generate %.3f default-criteria e2e %.3f'%(d['e2e']['value'],d['value'],d['aux']['stage_ms_device']['ransac'],k['ransac_validate'],k['ransac_generate'],d['aux']['align_ms_reference_default_criteria_e2e']), d['result']['ransac_survivors'])"; }
echo "== pipelined"; b; b
echo "== plain"; PCR_WAVE_PIPELINE=0 b
