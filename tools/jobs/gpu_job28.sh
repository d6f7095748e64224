#!/usr/bin/env bash
set -u
(cd 3d-matching_b200/csrc && touch pcr_icp.cu && make EXTRA=-DICP_EAGER_ALL 2>&1 | grep -v nvcc | tail -2)
echo "== icp trace 1M eager"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -8 | cut -c1-300
echo "== icp 1M no trace"; N=1000000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
