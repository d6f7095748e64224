#!/usr/bin/env bash
set -u
for n in 100000 1000000; do
echo "== default n=$n"; N=$n ITERS=50 python tools/gpu_icp_trace.py 2>&1 | tail -1
done
cp 3d-matching_b200/pcr_b200/libpcr_b200.so /tmp/orig.so; cp tools/_var/lib_768.so 3d-matching_b200/pcr_b200/libpcr_b200.so
for n in 100000 1000000; do
echo "== 768 threads n=$n"; N=$n ITERS=50 python tools/gpu_icp_trace.py 2>&1 | tail -1
done
cp /tmp/orig.so 3d-matching_b200/pcr_b200/libpcr_b200.so
