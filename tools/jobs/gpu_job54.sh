#!/usr/bin/env bash
# end of round: GPU tests, smoke, default bench line (final code), launch list
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j54_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j54_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/r2_launch_ncu.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/gpu_stress_align.py 2>&1 | tail -1
