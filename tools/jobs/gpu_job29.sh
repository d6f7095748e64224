#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=1000000 ITERS=30 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py > gpurun_out/j29_trace1m.log 2>&1; tail -1 gpurun_out/j29_trace1m.log
N=100000 ITERS=30 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py > gpurun_out/j29_trace100k.log 2>&1; tail -1 gpurun_out/j29_trace100k.log
