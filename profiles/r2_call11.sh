#!/usr/bin/env bash
# Round 2, GPU call 11 (one B200): software-pipelined CRPS sort kernel (register
# cap variants), f64 weight sums back in the unbinned element-weight path.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== CRPS + det tests"
timeout 900 python -m pytest tests/test_gpu_crps.py tests/test_gpu_det.py tests/test_gpu_categorical.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_call11_tests.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r2_call11_tests.log | tail -8
for b in default 2 3 4 5; do
  echo "== CRPS sort kernel, resident CTAs per SM: $b"
  if [ "$b" = default ]; then
    EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -1
  else
    WBX_EXP_SORT_MINB=$b EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -1
  fi
done
echo "== ncu of the sort kernel (default)"
EXP_ONLY=sort timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:crps_sort_kernel -s 2 -c 1 -o gpurun_out/r2_prof_crps_sort_pipelined \
    python profiles/exp_crps.py 3 > gpurun_out/r2_prof_crps_sort_pipelined.log 2>&1
tail -2 gpurun_out/r2_prof_crps_sort_pipelined.log
