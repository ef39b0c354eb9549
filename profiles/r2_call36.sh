#!/usr/bin/env bash
# Round 2, GPU call 36 (one B200): the fixed-size sort kernels shift the members
# by the first one (not by the target) before the network: tests, timing.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== CRPS tests"
timeout 900 python -m pytest tests/test_gpu_crps.py tests/test_gpu_next_rows.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== timing"
EXP_ONLY=sort,sort+moments timeout 200 python profiles/exp_crps.py 10 2>&1 | grep kernel | tee gpurun_out/r2_call36_exp_crps.log | cut -c1-220
EXP_MEMBERS=51 EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | grep kernel | tee -a gpurun_out/r2_call36_exp_crps.log | cut -c1-220
