#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j2_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/j2_pytest.log
echo "== icp trace 100k"; N=100000 ITERS=50 PCR_ICP_TRACE=1 timeout 120 python tools/gpu_icp_trace.py 2>&1 | tail -8
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 120 python tools/gpu_icp_trace.py 2>&1 | tail -8
