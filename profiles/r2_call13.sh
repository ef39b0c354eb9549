#!/usr/bin/env bash
# Round 2, GPU call 13 (one B200): CRPS sort kernel, two grid points per thread.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== CRPS tests"
timeout 900 python -m pytest tests/test_gpu_crps.py tests/test_gpu_next_rows.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_call13_tests.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r2_call13_tests.log | tail -8
echo "== CRPS sort kernel"
EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -1
EXP_MEMBERS=51 EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -1
echo "== ncu of the sort kernel"
EXP_ONLY=sort timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:crps_sort_kernel -s 2 -c 1 -o gpurun_out/r2_prof_crps_sort_pair \
    python profiles/exp_crps.py 3 > gpurun_out/r2_prof_crps_sort_pair.log 2>&1
tail -2 gpurun_out/r2_prof_crps_sort_pair.log
