#!/usr/bin/env bash
set -u
python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/j10_plain.json 2> gpurun_out/j10_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux > gpurun_out/j10_ncu.log 2>&1
echo "rc=$?"
