#!/usr/bin/env bash
# pcr_ransac_multi: next wave generated during validation + exchange (PCR_DIST_SPECULATE=0: old)
set -u
B=4 timeout 600 python tools/gpu_dist_check.py 2>&1 | grep -v "^\*\|OMP" | tail -5
PCR_DIST_SPECULATE=0 B=4 timeout 600 python tools/gpu_dist_check.py 2>&1 | grep "schedule" | tail -3
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "ransac or dist or multi or batch" 2>&1 | tail -3
