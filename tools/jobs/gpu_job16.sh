#!/usr/bin/env bash
set -u
export ICP1M=0 PCR_ALIGN_OVERLAP=0
python tools/prof_target.py > gpurun_out/j16_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k 'regex:celllists|validate|k_cell_|k_scan|k_morton|k_bounds' -s 0 -c 200 --csv --log-file gpurun_out/j16.csv python tools/prof_target.py > gpurun_out/j16_ncu.log 2>&1
echo rc=$?
