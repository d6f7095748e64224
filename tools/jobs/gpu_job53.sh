#!/usr/bin/env bash
# ICP: dynamic tail (PCR_ICP_DYN=1 default on clouds with >= 3 chunks per CTA) vs all static
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "icp" > gpurun_out/j53_pytest_icp.log 2>&1; echo "icp tests rc=$?"; tail -15 gpurun_out/j53_pytest_icp.log | cut -c1-300
tr() { N=$1 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | sed -n 2,4p | cut -c1-330; }
for dy in 1 0; do for n in 1000000 600000; do echo "== dyn $dy n $n"; PCR_ICP_DYN=$dy tr $n; done; done
for dy in 1 0 1 0; do for n in 1000000 600000 2000000; do echo "== no trace: dyn $dy n $n"; PCR_ICP_DYN=$dy N=$n ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1; done; done
