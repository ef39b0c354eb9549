#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for L in 2 3; do
  timeout 600 python profiles/c5_trace.py $L > gpurun_out/r2_call20_c5_trace_l$L.log 2>&1
  head -34 gpurun_out/r2_call20_c5_trace_l$L.log | cut -c1-150
done
for L in 2 3; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-suite --no-cpu-baseline --c5-lanes $L \
      > gpurun_out/r2_call20_bench_l$L.json 2> gpurun_out/r2_call20_bench_l$L.err
  python - <<PY
import json
line = json.loads(open('gpurun_out/r2_call20_bench_l$L.json').read().strip().splitlines()[-1])
c5 = line['c5']
print('lanes $L c5', c5.get('value'), {k: (round(v['value'] / 1e9, 2), round(v['h2d_GBps_per_rank'], 1), v['seconds']) for k, v in c5['suites'].items()})
PY
done
