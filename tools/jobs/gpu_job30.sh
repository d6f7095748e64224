#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() {
echo "== icp trace 1M"; N=1000000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | cut -c1-330
echo "== icp trace 100k"; N=100000 ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | tail -9 | cut -c1-330
echo "== icp 1M no trace"; N=1000000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
echo "== icp 100k no trace"; N=100000 ITERS=50 timeout 300 python tools/gpu_icp_trace.py 2>&1 | tail -1
}
echo "##### default (4 CTAs/SM, 64 regs)"; run
echo "== gpu tests (icp/parity)"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "icp or parity or golden or align" > gpurun_out/j30_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j30_pytest_gpu.log
echo "##### 3 CTAs/SM, 80 regs"
(cd 3d-matching_b200/csrc && touch pcr_icp.cu && make EXTRA=-DPCR_ICP_CTAS=3 2>&1 | grep -v nvcc | tail -2)
run
