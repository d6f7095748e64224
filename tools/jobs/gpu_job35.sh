#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j35_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/j35_pytest_gpu.log
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f launches %d bounds %.3f grid %.3f'%(d['e2e']['value'],d['value'],d['gpu_launches']/20,k['bounds'],k['grid_build']), d['aux']['stage_ms_device'])"; }
echo "== bench cluster"; b; b
echo "== bench multi-kernel"; export PCR_GRID_CLUSTER=0; b
