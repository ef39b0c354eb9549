#!/usr/bin/env bash
# Round 2, GPU call 7 (one B200): binned kernel with f32 element weights.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bins3 tests"
timeout 600 python -m pytest tests/test_gpu_bins3.py tests/test_gpu_fastpath.py tests/test_gpu_det.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call7_bins3.log 2>&1
tail -5 gpurun_out/r2_call7_bins3.log
echo "== exp_bins"
timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call7_exp_bins.log 2>&1
cat gpurun_out/r2_call7_exp_bins.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
