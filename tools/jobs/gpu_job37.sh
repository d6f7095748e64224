#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests (fpfh/parity)"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "fpfh or parity or golden" > gpurun_out/j37_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j37_pytest_gpu.log
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f launches %d fpfh %.3f spfh %.3f knn_list %.3f'%(d['e2e']['value'],d['value'],d['gpu_launches']/20,k['fpfh'],k['spfh'],k['knn_list']), d['aux']['stage_ms_device'])"; }
echo "== bench"; b; b
PCR_TIMELINE=1 python tools/gpu_timeline.py 2>&1 | grep "main" | sed -n 1,18p
