#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
ICP1M=0 PCR_ALIGN_OVERLAP=0 python tools/prof_target.py > gpurun_out/j6_plain.log 2>&1 && \
ICP1M=0 PCR_ALIGN_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:k_match_tc -s 2 -c 2 -o gpurun_out/prof_tc python tools/prof_target.py > gpurun_out/j6_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/j6_ncu.log
