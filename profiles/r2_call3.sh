#!/usr/bin/env bash
# Round 2, GPU call 3 (one B200): second-generation binned kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== bins2 tests"
timeout 600 python -m pytest tests/test_gpu_bins2.py tests/test_gpu_fastpath.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call3_bins2.log 2>&1
tail -30 gpurun_out/r2_call3_bins2.log
echo "== full GPU suite"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call3_gpu_tests.log 2>&1
tail -8 gpurun_out/r2_call3_gpu_tests.log
echo "== bench (suite only matters)"
timeout 800 python bench.py --steps 20 --warmup 5 --no-c5 --no-cpu-baseline > gpurun_out/r2_call3_bench.json 2> gpurun_out/r2_call3_bench.err
tail -3 gpurun_out/r2_call3_bench.err
python - <<'PY'
import json
try:
  line = json.loads(open('gpurun_out/r2_call3_bench.json').read().strip().splitlines()[-1])
except Exception as e:
  print('no line', e); raise SystemExit
print('value', line['value'], 'frac', line['roofline']['frac'], 'api', line['value_api']['value'])
print('suite_error', line.get('suite_error'))
for k, v in line.get('suite', {}).items():
  print(k, v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
echo "== ncu of the bins2 kernel"
cat > gpurun_out/_bins_once.py <<'PY'
import numpy as np, torch
from weatherbenchx_b200 import aggregation, binning, weighting, xarray_lite as xl
from weatherbenchx_b200.metrics import deterministic
NLAT, NLON = 721, 1440
lat = np.linspace(-90, 90, NLAT); lon = np.linspace(0, 360, NLON, endpoint=False)
rng = np.random.default_rng(7)
land = xl.DataArray(np.kron(rng.random((21, 24)) > 0.7, np.ones((35, 60), bool))[:NLAT],
                    ('latitude', 'longitude'), coords={'latitude': lat, 'longitude': lon})
regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
  'northern-hemisphere': ((20, 90), (0, 360)), 'southern-hemisphere': ((-90, -20), (0, 360)),
  'europe': ((35, 75), (-12.5, 42.5)), 'north-america': ((25, 60), (240, 285)),
  'north-atlantic': ((25, 65), (290, 350)), 'north-pacific': ((25, 60), (145, 230)),
  'east-asia': ((25, 60), (102.5, 150)), 'ausnz': ((-45, -12.5), (120, 175)),
  'arctic': ((60, 90), (0, 360)), 'antarctic': ((-90, -60), (0, 360)),
  'northern-africa': ((5, 32.5), (-12.5, 37.5)), 'southern-africa': ((-30, 5), (12.5, 37.5)),
  'south-america': ((-40, 5), (-75, -45)), 'west-asia': ((15, 60), (42.5, 102.5)),
  'south-east-asia': ((-12.5, 25), (95, 125))}
coords = {'init_time': np.arange(20), 'latitude': lat, 'longitude': lon}
dims = ('init_time', 'latitude', 'longitude')
P, T = {}, {}
for v in range(5):
  t = torch.empty((20, NLAT, NLON), device='cuda').normal_(280, 10)
  P[f'v{v}'] = xl.DataArray(t + torch.randn_like(t), dims, coords=coords, name=f'v{v}')
  T[f'v{v}'] = xl.DataArray(t, dims, coords=coords, name=f'v{v}')
agg = aggregation.Aggregator(reduce_dims=list(dims), weigh_by=[weighting.GridAreaWeighting()],
                             bin_by=[binning.Regions(regions, land_sea_mask=land)])
for _ in range(4):
  out = aggregation.compute_metric_values_for_single_chunk({'rmse': deterministic.RMSE()}, agg, P, T)
  print(float(out['rmse.v0'].values[0]))
PY
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:det_reduce_bins -s 2 -c 1 -o gpurun_out/r2_prof_bins2 \
    python gpurun_out/_bins_once.py > gpurun_out/r2_prof_bins2.log 2>&1
tail -3 gpurun_out/r2_prof_bins2.log
