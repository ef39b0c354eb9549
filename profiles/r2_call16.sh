#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python profiles/c5_breakdown.py 16 > gpurun_out/r2_call16_c5_breakdown.log 2>&1
head -70 gpurun_out/r2_call16_c5_breakdown.log | cut -c1-160
