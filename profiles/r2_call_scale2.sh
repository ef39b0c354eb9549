#!/usr/bin/env bash
# Round 2, second scaling call on ONE 8-GPU box (final code, ranks spread over
# the PCIe uplinks):  gpurun --gpus 8 --timeout 900 -- 'bash profiles/r2_call_scale2.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for N in 2 4 8; do
  echo "== bench at $N ranks"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 2964$N bench.py --gpus $N --steps 20 --warmup 5 \
      --no-suite --no-cpu-baseline \
      > gpurun_out/r2_bench2_box8_n$N.json 2> gpurun_out/r2_bench2_box8_n$N.err
  python - <<PY
import json
try:
  line = json.loads(open('gpurun_out/r2_bench2_box8_n$N.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); print(open('gpurun_out/r2_bench2_box8_n$N.err').read()[-2500:]); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'n', line['n_gpus'], 'devices', line['notes'].get('devices'))
print('value_api', line['value_api']['value'])
e = line['e2e']
print('e2e', e['value'], 'achieved', e['h2d_achieved_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_ceiling'])
c5 = line.get('c5') or {}
print('c5', c5.get('value'), {k: (v.get('value'), v.get('h2d_GBps_per_rank'), v.get('sharded_vs_monolithic_max_rel_diff')) for k, v in (c5.get('suites') or {}).items()}, c5.get('error'))
PY
done
