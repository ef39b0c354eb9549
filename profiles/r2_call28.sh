#!/usr/bin/env bash
# Round 2, GPU calls 28 and 30 (one B200): second fixed-shape spectrum kernel
# ((5, 12, 12) radices, first pass from registers, paired real-FFT split);
# call 30: + padded layout between its passes 2 and 3, + the two-pass 24 x 30 kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== spectrum tests"
timeout 600 python -m pytest tests/test_gpu_spectrum.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
echo "== A/B"
timeout 300 python profiles/exp_spectrum.py 20 2>&1 | tee gpurun_out/r2_call${CALL:-28}_exp_spectrum.log | cut -c1-300
