#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j36_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j36_pytest_gpu.log
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f launches %d'%(d['e2e']['value'],d['value'],d['gpu_launches']/20), d['aux']['stage_ms_device'])"; }
echo "== bench"; b; b
PCR_TIMELINE=1 python tools/gpu_timeline.py 2>&1 | grep "main" | sed -n 18,45p
