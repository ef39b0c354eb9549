#!/usr/bin/env bash
# Round 2, GPU call 24 (one B200): CRPS sorting network with work moved to the
# FMA pipe (last layer folded into the moment, mixed compare-exchanges), pipe
# microbenchmark with the compare-exchange forms.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pipes"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb profiles/microbench_fp32_pipes.cu \
  && timeout 120 /tmp/mb | tee gpurun_out/r2_call24_pipes.log
echo "== crps mix"
timeout 600 python profiles/exp_crps_mix.py 10 2>&1 | tee gpurun_out/r2_call24_crps_mix.log | cut -c1-400
