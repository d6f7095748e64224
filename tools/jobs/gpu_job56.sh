#!/usr/bin/env bash
# pcr_align_batch: spinning vs blocking host waits when the host threads outnumber the cores (CORES=n restricts the affinity)
set -u
for cores in 2 4 ""; do for bl in 0 1 auto; do
  echo "== cores '${cores:-all}' blocking $bl"
  if [ "$bl" = auto ]; then CORES=$cores WORKERS=1,3,6 B=96 timeout 300 python tools/gpu_batch_workers.py 2>&1 | tail -4
  else CORES=$cores WORKERS=1,3,6 B=96 PCR_BATCH_BLOCKING=$bl timeout 300 python tools/gpu_batch_workers.py 2>&1 | tail -4; fi
done; done
