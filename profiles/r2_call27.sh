#!/usr/bin/env bash
# Round 2, GPU call 27 (one B200): the sorting-network CRPS kernel as shipped
# after calls 24-26 (MIXPCT 45, 4 resident CTAs, 24 CTAs per SM, last layer in
# the moment, CTA-per-cell finalize): GPU tests, timing at M = 50 / 51 with and
# without the moments, ncu capture, the bench suite.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_call27_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_call27_gpu_tests.log
echo "== CRPS kernels"
EXP_ONLY=sort,sort+moments,pair timeout 200 python profiles/exp_crps.py 10 2>&1 | tee gpurun_out/r2_call27_exp_crps.log | cut -c1-250
EXP_MEMBERS=51 EXP_ONLY=sort,sort+moments timeout 200 python profiles/exp_crps.py 10 2>&1 | tee -a gpurun_out/r2_call27_exp_crps.log | cut -c1-250
echo "== ncu of the sort kernel"
EXP_ONLY=sort timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:crps_sort_kernel -s 2 -c 1 -o gpurun_out/r2_prof_crps_sort_mix \
    python profiles/exp_crps.py 3 > gpurun_out/r2_prof_crps_sort_mix.log 2>&1
tail -2 gpurun_out/r2_prof_crps_sort_mix.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call27_bench.json 2> gpurun_out/r2_call27_bench.err
tail -3 gpurun_out/r2_call27_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call27_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
