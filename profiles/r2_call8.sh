#!/usr/bin/env bash
# Round 2, GPU call 8 (one B200): binned kernel (f32 element weights everywhere),
# full GPU suite, full bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== exp_bins"
timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call8_exp_bins.log 2>&1
cat gpurun_out/r2_call8_exp_bins.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call8_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_call8_gpu_tests.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_call8_bench.json 2> gpurun_out/r2_call8_bench.err
tail -3 gpurun_out/r2_call8_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call8_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('e2e', line['e2e']['value'], line['e2e']['frac_of_ceiling'])
print('cpu', line.get('cpu_baseline', {}).get('value'))
print('c5', (line.get('c5') or {}).get('value'), (line.get('c5') or {}).get('error'))
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
