#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j32_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/j32_pytest_gpu.log
run() {
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep "slowest CTA loop\|search path" | tail -2 | cut -c1-230
echo "== icp 1M no trace"; N=1000000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
echo "== icp 100k no trace"; N=100000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
}
echo "#### compact"; run
echo "#### dense"; export PCR_GRID_COMPACT=0; run; unset PCR_GRID_COMPACT
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e',d['e2e']['value'],'value',d['value'], d['aux']['kernel_ms_per_step'])"
