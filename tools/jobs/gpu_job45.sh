#!/usr/bin/env bash
# ICP: warp-cooperative box search in the full-search passes (PCR_ICP_BOX=1 default) against the per-lane walk (=0)
set -u
mkdir -p gpurun_out
tr() { N=$1 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | cut -c1-330; }
for bx in 1 0; do for n in 100000 1000000; do echo "== box $bx n $n"; PCR_ICP_BOX=$bx tr $n; done; done
for bx in 1 0; do for n in 100000 1000000; do echo "== no trace: box $bx n $n"; PCR_ICP_BOX=$bx N=$n ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1; done; done
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j45_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j45_pytest_gpu.log
