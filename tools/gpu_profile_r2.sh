#!/usr/bin/env bash
# One gpurun call: the plain run, then ONE ncu --set full capture of the second alignment's main kernels and the 1M-point
# normals + ICP (tools/prof_target.py).  Read here with tools/ncu_table.py / tools/ncu_traffic.py.
set -u
mkdir -p gpurun_out
export PCR_ALIGN_OVERLAP=0 ICP1M_ITERS=20
python tools/prof_target.py > gpurun_out/prof_r2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_(knn_cov|ransac_validate|icp_persist|match_tc|match_final|fpfh|spfh|knn_list|celllists_build|ransac_generate)' \
    -s 22 -c 24 -o gpurun_out/prof_r2 python tools/prof_target.py > gpurun_out/prof_r2_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_r2_ncu.log
