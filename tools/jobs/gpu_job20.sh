#!/usr/bin/env bash
set -u
python tools/sanitize_target.py > gpurun_out/san_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/san_memcheck.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/san_memcheck.log
