#!/usr/bin/env bash
# Round 2, GPU call 34 (one B200): ensemble moments alone on the
# register-resident skeleton of the sort kernel (no network) against the
# pair-kernel skeleton; CRPS tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== CRPS / next-rows tests"
timeout 900 python -m pytest tests/test_gpu_crps.py tests/test_gpu_next_rows.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== moments: register skeleton"
EXP_ONLY=moments timeout 200 python profiles/exp_crps.py 10 2>&1 | grep kernel | tee gpurun_out/r2_call34_moments.log | cut -c1-330
echo "== moments: pair skeleton"
WBX_EXP_MOMENTS_PAIR=1 EXP_ONLY=moments timeout 200 python profiles/exp_crps.py 10 2>&1 | grep kernel | tee -a gpurun_out/r2_call34_moments.log | cut -c1-330
