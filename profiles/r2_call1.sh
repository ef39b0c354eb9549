#!/usr/bin/env bash
# Round 2, GPU call 1 (one B200):
#   gpurun --timeout 900 -- 'bash profiles/r2_call1.sh'
# host topology, H2D copy ceiling of one GPU, `ncu --set full` of the SHIPPED
# CRPS kernels (crps_sort_kernel<64,50,...>, crps_reduce_tma_kernel) on one
# variable of config[2], GPU test suite as a sanity check.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | head -30
  echo "== numa"; numactl -H 2>/dev/null || ls /sys/devices/system/node/
  echo "== pci numa nodes"
  for d in /sys/bus/pci/devices/*; do
    if [ -f "$d/class" ] && grep -q '^0x0302' "$d/class"; then
      echo "$d numa=$(cat $d/numa_node) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"
    fi
  done
  echo "== memory"; free -g
  nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
} > gpurun_out/r2_topology_n1.txt 2>&1
tail -5 gpurun_out/r2_topology_n1.txt

echo "== h2d ceiling, one GPU"
timeout 120 python profiles/h2d_ceiling.py > gpurun_out/r2_h2d_n1.json 2> gpurun_out/r2_h2d_n1.err
cat gpurun_out/r2_h2d_n1.json | head -c 3000; tail -3 gpurun_out/r2_h2d_n1.err

echo "== GPU tests"
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu_tests_call1.log 2>&1
tail -3 gpurun_out/r2_gpu_tests_call1.log

echo "== ncu full: shipped CRPS sort kernel"
EXP_ONLY=sort timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:crps_sort_kernel -s 2 -c 1 -o gpurun_out/r2_prof_crps_sort \
    python profiles/exp_crps.py 1 > gpurun_out/r2_prof_crps_sort.log 2>&1
tail -2 gpurun_out/r2_prof_crps_sort.log
echo "== ncu full: shipped CRPS pair (TMA) kernel"
EXP_ONLY=pair timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:crps_reduce -s 2 -c 1 -o gpurun_out/r2_prof_crps_pair \
    python profiles/exp_crps.py 1 > gpurun_out/r2_prof_crps_pair.log 2>&1
tail -2 gpurun_out/r2_prof_crps_pair.log
ls -la gpurun_out | tail -12
