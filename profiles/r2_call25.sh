#!/usr/bin/env bash
# Round 2, GPU calls 25-26 (one B200): CRPS sorting network -- mixed
# compare-exchanges x resident CTAs x grid size (call 25 also had a cp.async
# staging variant: slower, removed).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python profiles/exp_crps_mix.py 10 2>&1 | tee gpurun_out/r2_call25_crps_mix.log | grep timing | cut -c1-200
