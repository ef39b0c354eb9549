"""Where the time of a C5 deterministic chunk goes (one GPU): per-chunk wall
time of the loader, of the evaluation (statistics + aggregation through the
host-space plan) and of the accumulation, then a cProfile of the same loop.

  python profiles/c5_breakdown.py [n_init]
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from weatherbenchx_b200 import aggregation, time_chunks, weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.data_loaders import array_loaders
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import deterministic

NLAT, NLON = 721, 1440


def main():
  n_init = int(sys.argv[1]) if len(sys.argv) > 1 else 16
  dev = torch.device('cuda', 0)
  torch.cuda.set_device(0)
  n_lead, six = 12, np.timedelta64(6, 'h')
  t0 = np.datetime64('2020-01-01T00', 'ns')
  init = t0 + np.arange(n_init) * 2 * six
  lead = (np.arange(n_lead) * six).astype('timedelta64[ns]')
  n_valid = 2 * n_init + n_lead
  valid = t0 + np.arange(n_valid) * six
  lat = np.linspace(-90, 90, NLAT)
  lon = np.linspace(0, 360, NLON, endpoint=False)
  grid = {'latitude': lat, 'longitude': lon}
  det_vars = ('2m_temperature', 'geopotential_500')

  def pinned(shape):
    host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    host.normal_()
    return host

  keep, analyses, blocks, clim = [], {}, {}, {}
  for name in det_vars:
    h = pinned((n_valid, NLAT, NLON)); keep.append(h)
    analyses[name] = xl.DataArray(h.numpy(), ('valid_time', 'latitude', 'longitude'),
                                  coords=dict(grid, valid_time=valid), name=name)
    h = pinned((4, n_lead, NLAT, NLON)); keep.append(h)
    blocks[name] = h.numpy()
    clim[name] = xl.DataArray(
        torch.empty((366, 4, NLAT, NLON), device=dev).normal_(0.0, 0.5),
        ('dayofyear', 'hour', 'latitude', 'longitude'),
        coords=dict(grid, dayofyear=np.arange(1, 367), hour=np.arange(0, 24, 6)),
        name=name)
  metrics = {'rmse': deterministic.RMSE(), 'mse': deterministic.MSE(),
             'mae': deterministic.MAE(), 'bias': deterministic.Bias(),
             'acc': deterministic.ACC(clim)}
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  times = time_chunks.TimeChunks(init, lead, init_time_chunk_size=1,
                                 lead_time_chunk_size=n_lead)
  forecasts = bench._RecyclingForecasts(blocks, init, lead, grid)

  def loop(timers):
    loader = array_loaders.TargetsFromArrays(
        {v: analyses[v] for v in det_vars}, device_cache=True)
    total = None
    for i in range(len(times)):
      a = time.perf_counter()
      init_times, lead_times = times[i]
      targets = loader.load_chunk(init_times, lead_times)
      predictions = forecasts.load_chunk(init_times, lead_times, targets)
      b = time.perf_counter()
      statistics = metrics_base.compute_unique_statistics_for_all_metrics(
          metrics, predictions, targets)
      c = time.perf_counter()
      state = aggregator.aggregate_statistics(statistics)
      d = time.perf_counter()
      total = state if total is None else total + state
      e = time.perf_counter()
      if timers is not None:
        timers.append((b - a, c - b, d - c, e - d))
    return total

  loop(None)
  torch.cuda.synchronize()
  timers = []
  w0 = time.perf_counter()
  loop(timers)
  torch.cuda.synchronize()
  wall = time.perf_counter() - w0
  t = np.array(timers) * 1e3
  print(f'{len(timers)} chunks, {wall * 1e3 / len(timers):.3f} ms per chunk '
        f'(100 MB of forecasts = 1.80 ms at 55.6 GB/s)')
  print('ms per chunk: load %.3f  statistics %.3f  aggregate %.3f  accumulate %.3f'
        % tuple(t[2:].mean(0)))
  prof = cProfile.Profile()
  prof.enable()
  loop(None)
  prof.disable()
  pstats.Stats(prof).sort_stats('cumulative').print_stats(40)


if __name__ == '__main__':
  main()
