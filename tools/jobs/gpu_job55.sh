#!/usr/bin/env bash
# RANSAC: small blind first wave evaluated part-wise (PCR_BLIND_PARTS, default 4; 1 = the general kernel)
set -u
mkdir -p gpurun_out
b() { timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e %.3f value %.3f default-criteria e2e %.3f'%(d['e2e']['value'],d['value'],d['aux']['align_ms_reference_default_criteria_e2e']), 'ransac stage %.3f'%d['aux']['stage_ms_device']['ransac'], 'validate %.3f'%d['aux']['kernel_ms_per_step']['ransac_validate'])"; }
for p in 4 1 2 3 6 4 1; do echo "== parts $p"; PCR_BLIND_PARTS=$p b; done
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j55_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j55_pytest_gpu.log
