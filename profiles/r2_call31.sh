#!/usr/bin/env bash
# Round 2, GPU call 31 (one B200): evidence for profiles/ -- ncu launch list of
# the bench command, ncu --set full of the headline kernel and of the two-pass
# spectrum kernel; CRPS / spectrum tests and sort-kernel timing after the
# mad.wide addressing change.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests (crps, spectrum)"
timeout 900 python -m pytest tests/test_gpu_crps.py tests/test_gpu_spectrum.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== CRPS sort kernel"
EXP_ONLY=sort,sort+moments timeout 200 python profiles/exp_crps.py 10 2>&1 | grep kernel | tee gpurun_out/r2_call31_exp_crps.log | cut -c1-200
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_r2_z.csv \
    python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_call31_bench_under_ncu.json 2> gpurun_out/r2_call31_bench_under_ncu.err
tail -c 300 gpurun_out/r2_call31_bench_under_ncu.err
wc -l gpurun_out/launches_r2_z.csv
echo "== ncu full: headline kernel"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_tma_kernel -s 4 -c 1 -o gpurun_out/prof_det_r2_z \
    python bench.py --steps 2 --warmup 3 --no-suite --no-c5 --no-cpu-baseline > gpurun_out/r2_call31_prof_det.log 2>&1
tail -2 gpurun_out/r2_call31_prof_det.log | cut -c1-300
echo "== ncu full: two-pass spectrum kernel"
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:zonal_spectrum_2pass -s 2 -c 1 -o gpurun_out/r2_prof_spectrum_2pass \
    python profiles/exp_spectrum.py 3 > gpurun_out/r2_prof_spectrum_2pass.log 2>&1
tail -2 gpurun_out/r2_prof_spectrum_2pass.log | cut -c1-300
