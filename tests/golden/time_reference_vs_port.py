"""Times the reference's own RMSE path against the oracle's op-for-op port.

    python tests/golden/time_reference_vs_port.py

Build container only (needs /root/reference).  The bench's CPU baseline is the
port ``oracle.reference_path_rmse`` because the reference cannot travel to the
GPU box; this script shows, on the same machine and the same arrays, what the
unmodified ``weatherbenchX.aggregation.compute_metric_values_for_single_chunk``
costs (on the stand-in xarray of reference_runtime.py -- real xarray adds its
own alignment / indexing overhead on top), so the port is not a strawman.
"""

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle'))
import reference_runtime  # noqa: E402

xr = reference_runtime.install()
import wbx_oracle as oracle  # noqa: E402
from weatherbenchX import aggregation, weighting  # noqa: E402
from weatherbenchX.metrics import deterministic  # noqa: E402

n_init, nlat, nlon = 5, 721, 1440
rng = np.random.default_rng(0)
lat = np.linspace(-90, 90, nlat)
coords = {'init_time': np.arange(n_init), 'latitude': lat,
          'longitude': np.linspace(0, 360, nlon, endpoint=False)}
dims = ('init_time', 'latitude', 'longitude')
p = rng.standard_normal((n_init, nlat, nlon), dtype=np.float32)
t = rng.standard_normal((n_init, nlat, nlon), dtype=np.float32)
P = {'t2m': xr.DataArray(p, dims, coords=coords)}
T = {'t2m': xr.DataArray(t, dims, coords=coords)}
agg = aggregation.Aggregator(reduce_dims=list(dims),
                             weigh_by=[weighting.GridAreaWeighting()])
metrics = {'rmse': deterministic.RMSE()}
w = oracle.grid_area_weights(lat)


def best(fn, reps=5):
  fn()
  out = []
  for _ in range(reps):
    t0 = time.perf_counter()
    fn()
    out.append(time.perf_counter() - t0)
  return min(out)


ref = best(lambda: aggregation.compute_metric_values_for_single_chunk(
    metrics, agg, P, T))
port = best(lambda: oracle.reference_path_rmse(p, t, w))
value = float(aggregation.compute_metric_values_for_single_chunk(
    metrics, agg, P, T)['rmse.t2m'].values)
sws, sw = oracle.reference_path_rmse(p, t, w)
pts = n_init * nlat * nlon
print(json.dumps({
    'points': pts, 'threads': 1,
    'reference_code_ms': ref * 1e3, 'reference_code_Mpts_per_s': pts / ref / 1e6,
    'port_ms': port * 1e3, 'port_Mpts_per_s': pts / port / 1e6,
    'port_over_reference_speed': ref / port,
    'rmse_reference': value, 'rmse_port': float(np.sqrt(sws / sw))}))
