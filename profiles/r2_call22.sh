#!/usr/bin/env bash
# Round 2, GPU call 22 (one B200): element-weight launches without a mask routed
# to the binned kernel with one class.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_call22_gpu_tests.log 2>&1
tail -3 gpurun_out/r2_call22_gpu_tests.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call22_bench.json 2> gpurun_out/r2_call22_bench.err
tail -3 gpurun_out/r2_call22_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call22_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
