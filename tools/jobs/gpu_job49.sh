#!/usr/bin/env bash
# final round-2 measurements: the driver's default bench line, the reference arm, then tools/gpu_profile_r2.sh (launch list + full capture)
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; echo "reference rc=$?"
bash tools/gpu_profile_r2.sh
ls -la gpurun_out | head -30
