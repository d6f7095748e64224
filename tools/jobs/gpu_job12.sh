#!/usr/bin/env bash
set -u
python tools/gpu_knn_time.py > gpurun_out/j12_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_knn_cov -s 2 -c 1 -o gpurun_out/prof_knn python tools/gpu_knn_time.py > gpurun_out/j12_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/j12_ncu.log
