#!/usr/bin/env bash
# ICP: sqrt-free certificates + spread accumulators + hoisted group switch; certificate start pass 2 / 1 / 0
set -u
mkdir -p gpurun_out
tr() { N=$1 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | cut -c1-420; }
for cp in 2 1 0; do for n in 100000 1000000; do echo "== cert_pass $cp n $n"; PCR_ICP_CERT_PASS=$cp tr $n; done; done
for cp in 2 1 0; do for n in 100000 1000000; do echo "== no trace: cert_pass $cp n $n"; PCR_ICP_CERT_PASS=$cp N=$n ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1; done; done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j44_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j44_pytest_gpu.log
