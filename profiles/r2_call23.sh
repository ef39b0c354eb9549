#!/usr/bin/env bash
# Round 2, GPU call 23 (one B200): binned launches with FEW classes (is the
# finalize kernel, one warp per (cell, class, column), a bottleneck there?)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for C in 1 2 6; do
  echo "== $C classes"
  EXP_CLASSES=$C EXP_ONLY=/se timeout 300 python profiles/exp_bins.py 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(f\"{d['case']:34s} k{d['kernel']} {d['kernel_ms']:.4f} ms step {d['step_ms']:.4f} frac {d['hbm_frac']:.3f} ok {d['checked']}\")
"
done
