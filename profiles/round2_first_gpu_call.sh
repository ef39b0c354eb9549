#!/usr/bin/env bash
# First GPU call of round 2 (run from the repo root under gpurun, one B200):
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/round2_first_gpu_call.sh'
#
# Round 1 ended with the categorical path and SEEPS verified on hardware (391 +
# 23 tests) but not timed.  This script collects, in order of importance and
# with its own timeouts:
#   1. the SEEPS / late-case GPU tests alone,
#   2. the whole GPU suite,
#   3. bench.py (headline line + suite incl. the new `contingency_3thr` leg),
#   4. the ncu launch list of a short bench run and one `--set full` capture of
#      the XF reduction kernel and of the SEEPS kernel.
# Everything lands in gpurun_out/; copy what is to be judged into profiles/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

echo "== 1. SEEPS, late-case and unconfirmed-case tests (look for XPASS)" | tee gpurun_out/r2_seeps_tests.log
timeout 120 python -m pytest tests/test_zz_gpu_seeps.py -m gpu -q -rxX \
    -p no:cacheprovider >> gpurun_out/r2_seeps_tests.log 2>&1
tail -15 gpurun_out/r2_seeps_tests.log

echo "== 2. full GPU suite"
timeout 300 python -m pytest tests -m gpu -q -rxX -p no:cacheprovider \
    > gpurun_out/r2_gpu_suite.log 2>&1
tail -5 gpurun_out/r2_gpu_suite.log

echo "== 3. bench"
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
line = json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('value', line['value'], 'e2e', line['e2e']['value'],
      'frac', line['roofline']['frac'])
for k, v in line.get('suite', {}).items():
  print(k, v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'),
        v.get('error'))
PY

echo "== 4. ncu: launch list, then full captures of the new kernels"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 \
    --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2_bench_under_ncu.log 2>&1
cat > gpurun_out/_xf_once.py <<'PY'
import numpy as np, torch
from weatherbenchx_b200 import aggregation, weighting, xarray_lite as xl
from weatherbenchx_b200.metrics import categorical, wrappers
n_init, ny, nx = 20, 721, 1440
dims = ('init_time', 'latitude', 'longitude')
coords = {'init_time': np.arange(n_init), 'latitude': np.linspace(-90, 90, ny),
          'longitude': np.linspace(0, 360, nx, endpoint=False)}
t = torch.empty((n_init, ny, nx), device='cuda').exponential_(0.5)
p = (t + torch.randn_like(t)).clamp_(0)
P = {'rain': xl.DataArray(p, dims, coords=coords, name='rain')}
T = {'rain': xl.DataArray(t, dims, coords=coords, name='rain')}
both = [wrappers.ContinuousToBinary('both', [0.1, 1.0, 5.0], 'threshold')]
metrics = {'ets': wrappers.WrappedMetric(categorical.ETS(), both)}
agg = aggregation.Aggregator(reduce_dims=list(dims),
                             weigh_by=[weighting.GridAreaWeighting()])
for _ in range(4):
  out = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
print(out['ets.rain'].values)
PY
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_tma_kernel -s 2 -c 2 -o gpurun_out/r2_prof_xf \
    python gpurun_out/_xf_once.py > gpurun_out/r2_prof_xf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:seeps_elementwise_kernel -c 2 -o gpurun_out/r2_prof_seeps \
    python -m pytest tests/test_zz_gpu_seeps.py -m gpu -q -k bit_for_bit \
    -p no:cacheprovider > gpurun_out/r2_prof_seeps.log 2>&1
echo "== 5. job-order experiment for K thresholds (engine.XF_L2_BLOCK_BYTES)"
timeout 200 python profiles/exp_xf_l2.py > gpurun_out/r2_exp_xf_l2.log 2>&1
tail -14 gpurun_out/r2_exp_xf_l2.log
ls -la gpurun_out | tail -20
