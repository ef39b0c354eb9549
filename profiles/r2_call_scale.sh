#!/usr/bin/env bash
# Round 2, scaling call on ONE 8-GPU box:
#   gpurun --gpus 8 --timeout 1500 -- 'bash profiles/r2_call_scale.sh'
# For N = 1, 2, 4, 8 ranks: the box's host->device ceiling with N processes
# copying at once (profiles/h2d_ceiling.py), then bench.py at N ranks
# (headline, value_api, e2e against the ceiling, C5 pipeline leg through
# pipeline.run_pipeline + distributed.all_reduce_state).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | head -25
  echo "== numa"; numactl -H 2>/dev/null || ls /sys/devices/system/node/
  echo "== pci"
  for d in /sys/bus/pci/devices/*; do
    if [ -f "$d/class" ] && grep -q '^0x0302' "$d/class"; then
      echo "$d numa=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"
    fi
  done
  echo "== memory"; free -g
} > gpurun_out/r2_topology_n8.txt 2>&1
for N in 1 2 4 8; do
  echo "== h2d ceiling, $N processes"
  H2D_REPS=4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 2953$N profiles/h2d_ceiling.py \
      > gpurun_out/r2_h2d_box8_n$N.json 2> gpurun_out/r2_h2d_box8_n$N.err
  python - <<PY
import json
try:
  d = json.loads(open('gpurun_out/r2_h2d_box8_n$N.json').read().strip().splitlines()[-1])
  for k, v in d['legs'].items():
    if k.startswith(('all', 'even', 'first4')) or k in ('solo_0',):
      print(f"{k:22s} {v['aggregate_GBps']:8.1f} GB/s aggregate {v['per_gpu_GBps']:7.1f} per GPU")
except Exception as e:
  print('h2d failed', e); print(open('gpurun_out/r2_h2d_box8_n$N.err').read()[-1500:])
PY
  echo "== bench at $N ranks"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 \
      --no-suite --no-cpu-baseline \
      > gpurun_out/r2_bench_box8_n$N.json 2> gpurun_out/r2_bench_box8_n$N.err
  python - <<PY
import json
try:
  line = json.loads(open('gpurun_out/r2_bench_box8_n$N.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); print(open('gpurun_out/r2_bench_box8_n$N.err').read()[-2500:]); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'n', line['n_gpus'])
print('value_api', line['value_api']['value'])
e = line['e2e']
print('e2e', e['value'], 'achieved', e['h2d_achieved_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_ceiling'])
c5 = line.get('c5') or {}
print('c5', c5.get('value'), {k: (v.get('value'), v.get('h2d_GBps_per_rank'), v.get('sharded_vs_monolithic_max_rel_diff')) for k, v in (c5.get('suites') or {}).items()}, c5.get('error'))
PY
done
echo "== reference arm (rank 0 only works)"
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 \
    > gpurun_out/r2_bench_ref_box8.json 2> gpurun_out/r2_bench_ref_box8.err
tail -c 400 gpurun_out/r2_bench_ref_box8.json
