#!/usr/bin/env bash
set -u
run() {
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep "slowest CTA loop\|search path" | tail -2 | cut -c1-230
echo "== icp trace 100k"; N=100000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep "slowest CTA loop\|search path" | tail -2 | cut -c1-230
echo "== icp 1M no trace"; N=1000000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
echo "== icp 100k no trace"; N=100000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
}
for cp in 1 0; do
echo "##### CERT_PASS=$cp"
(cd 3d-matching_b200/csrc && touch pcr_icp.cu && make EXTRA=-DPCR_ICP_CERT_PASS=$cp 2>&1 | grep -v nvcc | tail -2)
run
done
echo "== align e2e with CERT_PASS=0"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e',d['e2e']['value'],'value',d['value'], d['aux']['kernel_ms_per_step']['icp_pass'])"
