#!/usr/bin/env bash
set -u
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -8 | cut -c1-300
echo "== icp trace 100k"; N=100000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -8 | cut -c1-300
