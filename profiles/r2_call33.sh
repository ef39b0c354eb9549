#!/usr/bin/env bash
# Round 2, GPU call 33 (two B200s of one box): the final code under torchrun at
# N = 2 (NCCL, C5 leg through pipeline.run_pipeline + all_reduce_state), then
# the default bench at N = 1 with the default-policy CRPS leg.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bench at 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r2_call33_bench_n2.json 2> gpurun_out/r2_call33_bench_n2.err
tail -c 400 gpurun_out/r2_call33_bench_n2.err
echo "== bench at 1 rank (default flags)"
timeout 900 python bench.py > gpurun_out/r2_call33_bench.json 2> gpurun_out/r2_call33_bench.err
tail -c 300 gpurun_out/r2_call33_bench.err
python - <<'PY'
import json
for name in ('gpurun_out/r2_call33_bench_n2.json', 'gpurun_out/r2_call33_bench.json'):
  try:
    line = json.loads(open(name).read().strip().splitlines()[-1])
  except Exception as e:
    print(name, 'no line', e); continue
  print(name)
  print(' ', {k: line.get(k) for k in ('value', 'n_gpus', 'ms_per_step', 'scaling', 'gpu_launches')})
  print('  frac', line['roofline']['frac'], 'e2e', line['e2e']['value'], line['e2e'].get('frac_of_ceiling'))
  c5 = line.get('c5') or {}
  print('  c5', c5.get('value'), c5.get('error'), {k: (v.get('value'), v.get('sharded_vs_monolithic_max_rel_diff')) for k, v in (c5.get('suites') or {}).items()})
  print('  suite_error', line.get('suite_error'))
  for k, v in (line.get('suite') or {}).items():
    print('  ', k, round(v.get('ms_per_step', 0), 4), round(v.get('kernel_ms_per_step', 0), 4), round(v.get('roofline', {}).get('frac', 0), 4), v.get('error'))
PY
