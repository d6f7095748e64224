#!/usr/bin/env bash
# ICP: pass 1 seeded from pass 0's correspondence (PCR_ICP_SEEDED=1 default) vs unseeded; per-CTA loop cycles of pass 20 at 1M
set -u
mkdir -p gpurun_out
tr() { N=$1 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | cut -c1-330; }
for sd in 1 0; do for n in 100000 1000000; do echo "== seeded $sd n $n"; PCR_ICP_SEEDED=$sd tr $n | sed -n 2,4p; done; done
for sd in 1 0; do for n in 100000 1000000; do echo "== no trace: seeded $sd n $n"; PCR_ICP_SEEDED=$sd N=$n ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1; done; done
N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep "per CTA" | tail -1 > gpurun_out/j46_percta.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j46_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j46_pytest_gpu.log
