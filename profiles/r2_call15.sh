#!/usr/bin/env bash
# Round 2, GPU call 15 (one B200): evaluation lanes of the C5 leg.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for L in 1 2 3 4; do
  echo "== c5 leg, $L lanes"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-suite --no-cpu-baseline --c5-lanes $L \
      > gpurun_out/r2_call15_bench_l$L.json 2> gpurun_out/r2_call15_bench_l$L.err
  python - <<PY
import json
try:
  line = json.loads(open('gpurun_out/r2_call15_bench_l$L.json').read().strip().splitlines()[-1])
  c5 = line['c5']
  print('c5', c5.get('value'), {k: (round(v['value'] / 1e9, 2), round(v['h2d_GBps_per_rank'], 1), v['seconds']) for k, v in c5['suites'].items()})
except Exception as e:
  print('failed', e); print(open('gpurun_out/r2_call15_bench_l$L.err').read()[-1500:])
PY
done
