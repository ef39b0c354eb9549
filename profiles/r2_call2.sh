#!/usr/bin/env bash
# Round 2, GPU call 2 (one B200): GPU tests of the new host logic (fastpath
# replay, device-row cache, label alignment), bench with value_api / ceiling /
# c5 legs.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== new GPU tests first"
timeout 600 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_pipeline.py \
    "tests/test_gpu_det.py::test_config_c1_rmse_32x64_10_init" \
    "tests/test_gpu_det.py::test_config_c2_rmse_acc_128x256_13_levels" \
    -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call2_newtests.log 2>&1
tail -25 gpurun_out/r2_call2_newtests.log
echo "== full GPU suite"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call2_gpu_tests.log 2>&1
tail -8 gpurun_out/r2_call2_gpu_tests.log
echo "== bench"
timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_call2_bench.json 2> gpurun_out/r2_call2_bench.err
tail -5 gpurun_out/r2_call2_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call2_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'])
print('value_api', line.get('value_api'))
print('e2e', line['e2e'])
print('c5', json.dumps(line.get('c5'), indent=1)[:3000])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
