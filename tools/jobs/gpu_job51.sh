#!/usr/bin/env bash
# final: GPU tests, 51-pass ICP capture merged into profiles/r2_traffic.json, then the driver's default bench line + launch list
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j51_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j51_pytest_gpu.log
python tools/prof_icp.py > gpurun_out/prof_icp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_icp_persist' -c 2 -o gpurun_out/prof_r2_icp python tools/prof_icp.py > gpurun_out/prof_icp_ncu.log 2>&1
echo "icp capture rc=$?"
python tools/ncu_icp51.py gpurun_out/prof_r2_icp.ncu-rep profiles/r2_traffic.json > gpurun_out/ncu_icp51.log 2>&1; echo "merge rc=$?"
cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_ncu.log 2>&1
echo "launch list rc=$?"
