#!/usr/bin/env bash
# Round 2, GPU call 3 (one B200): second-generation binned kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bins2 tests"
timeout 600 python -m pytest tests/test_gpu_bins2.py tests/test_gpu_fastpath.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call3_bins2.log 2>&1
tail -30 gpurun_out/r2_call3_bins2.log
echo "== full GPU suite"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call3_gpu_tests.log 2>&1
tail -8 gpurun_out/r2_call3_gpu_tests.log
echo "== bench (suite only matters)"
timeout 800 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call3_bench.json 2> gpurun_out/r2_call3_bench.err
tail -3 gpurun_out/r2_call3_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call3_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
echo "== ncu of the bins2 kernel"
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins -s 2 -c 1 -o gpurun_out/r2_prof_bins2 \
    python profiles/bins_once.py > gpurun_out/r2_prof_bins2.log 2>&1
tail -3 gpurun_out/r2_prof_bins2.log
echo "== CRPS sort kernel: 5 vs 6 resident CTAs per SM"
EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -2
WBX_EXP_SORT_MINB6=1 EXP_ONLY=sort timeout 200 python profiles/exp_crps.py 10 2>&1 | tail -2
