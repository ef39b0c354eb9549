#!/usr/bin/env bash
# Round 2, GPU call 21 (one B200): binned kernel with one slot per thread in 31
# consumer warps (WBX_BINS_SPT=1, the new default) against two slots in 16.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests (default shape)"
timeout 900 python -m pytest tests/test_gpu_bins3.py tests/test_gpu_det.py tests/test_gpu_fastpath.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
echo "== tests (two slots per thread)"
WBX_BINS_SPT=2 timeout 900 python -m pytest tests/test_gpu_bins3.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for SPT in 1 2; do
  echo "== exp_bins, $SPT slot(s) per thread"
  WBX_BINS_SPT=$SPT timeout 600 python profiles/exp_bins.py 10 > gpurun_out/r2_call21_exp_bins_spt$SPT.log 2>&1
  cat gpurun_out/r2_call21_exp_bins_spt$SPT.log | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
done
