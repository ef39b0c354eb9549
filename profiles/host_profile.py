"""cProfile of the class API on a replayed chunk (host time per call):
  python profiles/host_profile.py bins|plain [calls]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from weatherbenchx_b200 import aggregation, binning, weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import deterministic
from exp_bins import REGIONS, NLAT, NLON  # (script directory)


def main():
  kind = sys.argv[1] if len(sys.argv) > 1 else 'bins'
  calls = int(sys.argv[2]) if len(sys.argv) > 2 else 300
  lat = np.linspace(-90, 90, NLAT)
  lon = np.linspace(0, 360, NLON, endpoint=False)
  rng = np.random.default_rng(7)
  land = xl.DataArray(
      np.kron(rng.random((21, 24)) > 0.7, np.ones((35, 60), bool))[:NLAT],
      ('latitude', 'longitude'), coords={'latitude': lat, 'longitude': lon})
  coords = {'init_time': np.arange(20), 'latitude': lat, 'longitude': lon}
  dims = ('init_time', 'latitude', 'longitude')
  P, T = {}, {}
  for v in range(5):
    t = torch.empty((20, NLAT, NLON), device='cuda').normal_(280, 10)
    P[f'v{v}'] = xl.DataArray(t + torch.randn_like(t), dims, coords=coords, name=f'v{v}')
    T[f'v{v}'] = xl.DataArray(t, dims, coords=coords, name=f'v{v}')
  agg = aggregation.Aggregator(
      reduce_dims=list(dims), weigh_by=[weighting.GridAreaWeighting()],
      bin_by=[binning.Regions(REGIONS, land_sea_mask=land)] if kind == 'bins' else None)
  metrics = {'rmse': deterministic.RMSE()}

  def loop(n):
    previous = None
    for _ in range(n):
      current = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
      if previous is not None:
        for k in previous:
          previous[k].values  # noqa: B018
      previous = current
    for k in previous:
      previous[k].values  # noqa: B018

  loop(5)
  torch.cuda.synchronize()
  import time
  t0 = time.perf_counter()
  loop(calls)
  torch.cuda.synchronize()
  print(f'{kind}: {(time.perf_counter() - t0) / calls * 1e6:.1f} us per call (wall)')
  prof = cProfile.Profile()
  prof.enable()
  loop(calls)
  prof.disable()
  stats = pstats.Stats(prof)
  stats.sort_stats('cumulative').print_stats(45)


if __name__ == '__main__':
  main()
