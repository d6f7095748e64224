#!/usr/bin/env bash
set -u
PCR_DEBUG=1 timeout 60 python tools/gpu_tc_debug.py random 2>&1 | tail -12; echo "rc=$?"
NQ=1300 NB=9170 PCR_DEBUG=1 timeout 60 python tools/gpu_tc_debug.py random 2>&1 | tail -12; echo "rc=$?"
PCR_DEBUG=1 timeout 90 python tools/gpu_tc_debug.py real 2>&1 | tail -12; echo "rc=$?"
