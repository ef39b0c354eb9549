#!/usr/bin/env bash
# Round 2, GPU call 4 (one B200): binned kernel with the host-compiled
# reduction schedule (csrc/det_bins3.cuh).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bins3 tests"
timeout 600 python -m pytest tests/test_gpu_bins3.py tests/test_gpu_fastpath.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call4_bins3.log 2>&1
tail -15 gpurun_out/r2_call4_bins3.log
echo "== exp_bins (v3)"
timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call4_exp_bins_v3.log 2>&1
cat gpurun_out/r2_call4_exp_bins_v3.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
echo "== exp_bins (v1, subset)"
EXP_ONLY=/se timeout 600 python profiles/exp_bins.py 5 v1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call4_gpu_tests.log 2>&1
tail -8 gpurun_out/r2_call4_gpu_tests.log
echo "== ncu of the bins3 kernel (lat-major, 20 jobs per cell, SE)"
EXP_ONLY=lat_major/20/se timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins3 -s 3 -c 1 -o gpurun_out/r2_prof_bins3_lat20 \
    python profiles/exp_bins.py 2 > gpurun_out/r2_prof_bins3_lat20.log 2>&1
tail -2 gpurun_out/r2_prof_bins3_lat20.log
echo "== ncu of the bins3 kernel (lat-major, 1 job per cell, 3 statistics)"
EXP_ONLY=lat_major/1/e_ae_se timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins3 -s 3 -c 1 -o gpurun_out/r2_prof_bins3_lat1 \
    python profiles/exp_bins.py 2 > gpurun_out/r2_prof_bins3_lat1.log 2>&1
tail -2 gpurun_out/r2_prof_bins3_lat1.log
echo "== ncu of the bins3 kernel (lon-major, 20 jobs per cell, SE)"
EXP_ONLY=lon_major/20/se timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins3 -s 3 -c 1 -o gpurun_out/r2_prof_bins3_lon20 \
    python profiles/exp_bins.py 2 > gpurun_out/r2_prof_bins3_lon20.log 2>&1
tail -2 gpurun_out/r2_prof_bins3_lon20.log
