#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/prof_icp.py > gpurun_out/prof_icp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_icp_persist' -c 2 -o gpurun_out/prof_icp python tools/prof_icp.py > gpurun_out/prof_icp_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_icp_ncu.log
