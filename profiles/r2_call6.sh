#!/usr/bin/env bash
# Round 2, GPU call 6 (one B200): binned kernel with unpadded schedule (class
# remainders in lane-granular warps); FP64 / shuffle pipe microbenchmark.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bins3 tests"
timeout 600 python -m pytest tests/test_gpu_bins3.py tests/test_gpu_fastpath.py tests/test_gpu_det.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call6_bins3.log 2>&1
tail -5 gpurun_out/r2_call6_bins3.log
echo "== exp_bins"
timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call6_exp_bins.log 2>&1
cat gpurun_out/r2_call6_exp_bins.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
echo "== microbench f64 / shfl"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb2 profiles/microbench_f64_shfl.cu && /tmp/mb2 | tee gpurun_out/r2_microbench_f64_shfl.log
echo "== ncu of the bins3 kernel (lat-major, 1 job per cell, 3 statistics)"
EXP_ONLY=lat_major/1/e_ae_se timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins3 -s 3 -c 1 -o gpurun_out/r2_prof_bins3_lat1c \
    python profiles/exp_bins.py 2 > gpurun_out/r2_prof_bins3_lat1c.log 2>&1
tail -2 gpurun_out/r2_prof_bins3_lat1c.log
echo "== ncu of the bins3 kernel (lon-major, 20 jobs per cell, acc6)"
EXP_ONLY=lon_major/20/acc6 timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins3 -s 3 -c 1 -o gpurun_out/r2_prof_bins3_lon20acc \
    python profiles/exp_bins.py 2 > gpurun_out/r2_prof_bins3_lon20acc.log 2>&1
tail -2 gpurun_out/r2_prof_bins3_lon20acc.log
