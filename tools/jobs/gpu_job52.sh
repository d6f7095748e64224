#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "voxel" > gpurun_out/j52_pytest_voxel.log 2>&1; echo "voxel tests rc=$?"; tail -30 gpurun_out/j52_pytest_voxel.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j52_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j52_pytest_gpu.log
