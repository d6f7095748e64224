#!/usr/bin/env bash
set -u
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f'%(d['e2e']['value'],d['value']), d['aux']['stage_ms_device'])"; }
echo "== default"; b; b
echo "== src normals at RANSAC"; export PCR_SRC_NORMALS_AT_RANSAC=1; b; b
PCR_TIMELINE=1 python tools/gpu_timeline.py 2>&1 | grep "timeline" | tail -32
