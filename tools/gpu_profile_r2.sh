#!/usr/bin/env bash
# One gpurun call: (1) the ncu launch list of a short bench, (2) ONE ncu --set full capture of the second alignment's main
# kernels and the 1M-point normals + ICP (tools/prof_target.py).  Each under ncu only after the same command exited 0 plain.
# Read here with tools/summarize_launches.py / tools/ncu_table.py / tools/ncu_traffic.py.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_ncu.log 2>&1
echo "launch list rc=$?"
export PCR_ALIGN_OVERLAP=0 ICP1M_ITERS=20
python tools/prof_target.py > gpurun_out/prof_r2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_(knn_cov|ransac_validate|icp_persist|match_tc|match_final|fpfh|spfh|knn_list|celllists|ransac_generate)' \
    -s 25 -c 27 -o gpurun_out/prof_r2 python tools/prof_target.py > gpurun_out/prof_r2_ncu.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/prof_r2_ncu.log
