#!/usr/bin/env bash
# Round 2, GPU call 29 (one B200): ncu capture of the second fixed-shape
# spectrum kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:zonal_spectrum_fixed2 -s 2 -c 1 -o gpurun_out/r2_prof_spectrum_fixed2 \
    python profiles/exp_spectrum.py 3 > gpurun_out/r2_prof_spectrum_fixed2.log 2>&1
tail -3 gpurun_out/r2_prof_spectrum_fixed2.log
