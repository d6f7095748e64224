#!/usr/bin/env bash
set -u
b() { timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f ransac stage %.3f validate %.3f generate %.3f'%(d['e2e']['value'],d['value'],d['aux']['stage_ms_device']['ransac'],k['ransac_validate'],k['ransac_generate']), d['result']['ransac_survivors'])"; }
echo "== default (2048 x64)"; b
for f in 131072 512 1024 4096; do echo "== first $f x256"; PCR_WAVE_FIRST=$f PCR_WAVE_GROWTH=256 b; done
