#!/usr/bin/env bash
# Round 2, GPU call 35 (one B200): what the driver runs at round end -- full
# pytest -m gpu, smoke(), bench.py (default flags) and its reference arm.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_call35_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_call35_gpu_tests.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (default flags)"
timeout 1500 python bench.py > gpurun_out/r2_call35_bench.json 2> gpurun_out/r2_call35_bench.err

python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call35_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); print(open('gpurun_out/r2_call35_bench.err').read()[-3000:]); raise SystemExit
print({k: line[k] for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'dtype', 'gpu_launches')})
print('roofline', {k: v for k, v in line['roofline'].items() if k != 'secondary'})
print('secondary', line['roofline'].get('secondary'))
print('e2e', line['e2e']['value'], line['e2e']['frac_of_ceiling'], 'api', line['value_api']['value'])
print('cpu', line.get('cpu_baseline'))
print('clocks', line.get('clocks'))
print('c5', (line.get('c5') or {}).get('value'), (line.get('c5') or {}).get('error'))
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference > gpurun_out/r2_call35_bench_ref.json 2> gpurun_out/r2_call35_bench_ref.err
tail -c 700 gpurun_out/r2_call35_bench_ref.json
