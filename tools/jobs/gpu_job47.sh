#!/usr/bin/env bash
# pcr_align: both voxel grids side by side; ICP waits for the target normals + its search structures only (PCR_ICP_EARLY=0: old)
set -u
mkdir -p gpurun_out
b() { timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e %.3f value %.3f default-criteria e2e %.3f lazy %.3f'%(d['e2e']['value'],d['value'],d['aux']['align_ms_reference_default_criteria_e2e'],d['aux']['align_ms_e2e_without_unused_source_normals']), d['aux']['stage_ms_device'])"; }
echo "== new"; b; b
echo "== icp waits for everything"; PCR_ICP_EARLY=0 b
echo "== serial preprocessing"; PCR_PRE_CONCURRENT=0 b
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j47_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j47_pytest_gpu.log
PCR_TIMELINE=1 python tools/gpu_timeline.py 2>&1 | grep -i "timeline\|stage" | grep -v "match_misc\|nn_features\|spfh\|fpfh\|knn_list" | head -60
