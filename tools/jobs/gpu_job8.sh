#!/usr/bin/env bash
set -u
echo "== no overlap"; PCR_ALIGN_OVERLAP=0 timeout 90 python bench.py --no-cpu --no-aux --steps 3 > gpurun_out/j8_a.json 2> gpurun_out/j8_a.err; echo "rc=$?"; tail -3 gpurun_out/j8_a.err
echo "== overlap, no priority"; PCR_ALIGN_PRIORITY=0 timeout 90 python bench.py --no-cpu --no-aux --steps 3 > gpurun_out/j8_b.json 2> gpurun_out/j8_b.err; echo "rc=$?"; tail -3 gpurun_out/j8_b.err
echo "== default"; timeout 90 python bench.py --no-cpu --no-aux --steps 3 > gpurun_out/j8_c.json 2> gpurun_out/j8_c.err; echo "rc=$?"; tail -3 gpurun_out/j8_c.err
nvidia-smi --query-gpu=name,memory.used --format=csv
