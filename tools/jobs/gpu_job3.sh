#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/j3_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/j3_pytest.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/j3_bench.json 2> gpurun_out/j3_bench.err; echo "rc=$?"; tail -5 gpurun_out/j3_bench.err
