#!/usr/bin/env bash
set -u
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_boundary.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "== dist check, 1 process"; timeout 200 python tools/gpu_dist_check.py 2>&1 | tail -6
