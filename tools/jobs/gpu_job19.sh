#!/usr/bin/env bash
set -u
for n in 100000 1000000; do N=$n ITERS=50 PCR_ICP_TRACE=1 python tools/gpu_icp_trace.py 2>&1 | tail -5; done
for n in 100000 1000000; do N=$n ITERS=50 python tools/gpu_icp_trace.py 2>&1 | tail -1; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_matcher_api.py -m gpu -q -x -p no:cacheprovider -k "icp or ICP or refine or pipeline or align" 2>&1 | tail -2
