#!/usr/bin/env bash
# re-entry state check: GPU tests, the full default bench line, ICP trace at 100k / 1M
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/j43_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j43_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/j43_bench.json 2> gpurun_out/j43_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/j43_bench.json') if l.startswith('{')][-1])
print('e2e %.3f value %.3f'%(d['e2e']['value'],d['value']), 'launches', d['gpu_launches'], d['roofline'])
print(d['aux']['stage_ms_device']); print(d['aux']['kernel_ms_per_step']); print(d['aux'].get('icp_1m'))
print({k:d[k] for k in d if 'identical' in k or k.endswith('_per_s')})
P
for n in 100000 1000000; do N=$n ITERS=50 PCR_ICP_TRACE=1 timeout 300 python tools/gpu_icp_trace.py 2>&1 | grep -v "per CTA" | cut -c1-600; done
