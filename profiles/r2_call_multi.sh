#!/usr/bin/env bash
# Round 2, multi-GPU call:  gpurun --gpus N --timeout 1200 -- 'bash profiles/r2_call_multi.sh N'
# topology, the box's host->device ceiling with N processes copying at once
# (profiles/h2d_ceiling.py), then bench.py at N ranks (headline, value_api, e2e
# against the ceiling, C5 pipeline leg through all_reduce_state).
set -u
N=${1:-2}
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
} > gpurun_out/r2_topology_n$N.txt 2>&1
echo "== h2d ceiling, $N processes"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29531 profiles/h2d_ceiling.py \
    > gpurun_out/r2_h2d_n$N.json 2> gpurun_out/r2_h2d_n$N.err
python - <<PY
import json
try:
  d = json.loads(open('gpurun_out/r2_h2d_n$N.json').read().strip().splitlines()[-1])
  for k, v in d['legs'].items():
    print(f"{k:22s} {v['aggregate_GBps']:8.1f} GB/s aggregate {v['per_gpu_GBps']:7.1f} per GPU")
except Exception as e:
  print('h2d failed', e); print(open('gpurun_out/r2_h2d_n$N.err').read()[-2000:])
PY
echo "== bench at $N ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
  line = json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'n', line['n_gpus'])
print('value_api', line['value_api']['value'])
e = line['e2e']
print('e2e', e['value'], 'achieved', e['h2d_achieved_gbs'], 'ceiling', e['h2d_ceiling_gbs'], 'frac', e['frac_of_ceiling'])
print('c5', json.dumps(line.get('c5'), indent=1)[:2500])
PY
echo "== reference arm at $N ranks (rank 0 only works)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 5 --warmup 1 \
    > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err
tail -c 600 gpurun_out/r2_bench_ref_n$N.json
