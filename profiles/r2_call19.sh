#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python profiles/c5_trace.py 2 > gpurun_out/r2_call19_c5_trace_l2.log 2>&1
head -90 gpurun_out/r2_call19_c5_trace_l2.log | cut -c1-150
