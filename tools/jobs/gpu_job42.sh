#!/usr/bin/env bash
set -u
b() { timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-aux 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['aux']['kernel_ms_per_step']; print('e2e %.3f value %.3f default-criteria e2e %.3f'%(d['e2e']['value'],d['value'],d['aux']['align_ms_reference_default_criteria_e2e']), d['aux']['stage_ms_device'])"; }
for y in 0 1 2 4 8; do echo "== yield $y"; PCR_HELPER_YIELD=$y b; done
