#!/usr/bin/env bash
# Round 2, GPU call 10 (one B200): f32 column weights in the unbinned
# element-weight path; BLAS class-to-bin product on the host.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call10_gpu_tests.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r2_call10_gpu_tests.log | tail -12
echo "== host profile (bins)"
timeout 300 python profiles/host_profile.py bins 300 2>&1 | head -3
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call10_bench.json 2> gpurun_out/r2_call10_bench.err
tail -3 gpurun_out/r2_call10_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call10_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
