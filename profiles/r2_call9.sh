#!/usr/bin/env bash
# Round 2, GPU call 9 (one B200): host time of the class API on replayed chunks;
# bench suite with the new binned legs.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== host profile (bins)"
timeout 300 python profiles/host_profile.py bins 300 > gpurun_out/r2_call9_host_bins.log 2>&1
head -60 gpurun_out/r2_call9_host_bins.log | cut -c1-150
echo "== host profile (plain)"
timeout 300 python profiles/host_profile.py plain 300 > gpurun_out/r2_call9_host_plain.log 2>&1
head -12 gpurun_out/r2_call9_host_plain.log | cut -c1-150
echo "== lon-major bins tests"
timeout 600 python -m pytest tests/test_gpu_det.py tests/test_gpu_bins3.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call9_bench.json 2> gpurun_out/r2_call9_bench.err
tail -3 gpurun_out/r2_call9_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call9_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
