#!/usr/bin/env bash
# Round 2, GPU call 14 (one B200): parked mbarrier waits in the binned kernel;
# quick identity lookup on the replay path.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_bins3.py tests/test_gpu_fastpath.py tests/test_gpu_det.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
echo "== exp_bins"
timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call14_exp_bins.log 2>&1
cat gpurun_out/r2_call14_exp_bins.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
echo "== host profile (bins)"
timeout 300 python profiles/host_profile.py bins 300 2>&1 | head -3
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call14_bench.json 2> gpurun_out/r2_call14_bench.err
tail -3 gpurun_out/r2_call14_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call14_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
