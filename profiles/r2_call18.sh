#!/usr/bin/env bash
# Round 2, GPU call 18 (one B200): bench.py with default flags (what the driver runs).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SECONDS=0
timeout 1500 python bench.py > gpurun_out/r2_call18_bench.json 2> gpurun_out/r2_call18_bench.err
echo "bench wall seconds: $SECONDS"
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call18_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); print(open('gpurun_out/r2_call18_bench.err').read()[-3000:]); raise SystemExit
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
