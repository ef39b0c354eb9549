"""End-to-end evaluation of a synthetic forecast archive with the chunk driver.

  python examples/evaluate_synthetic.py --out /tmp/metrics.nc
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29512 \
      examples/evaluate_synthetic.py --out /tmp/metrics.nc

The counterpart of the reference's run_example_evaluation.py for one node:
TimeChunks -> array loaders -> RMSE / bias / wind-vector RMSE with latitude
weights and region bins -> NetCDF.  Every rank evaluates a contiguous share of
the chunks on its own GPU; one all-reduce combines the states.
"""

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from weatherbenchx_b200 import aggregation, binning, pipeline  # noqa: E402
from weatherbenchx_b200 import time_chunks, weighting  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402
from weatherbenchx_b200.data_loaders import array_loaders  # noqa: E402
from weatherbenchx_b200.metrics import deterministic  # noqa: E402

H = np.timedelta64(1, 'h')


def synthetic_archive(n_init, n_lead, nlat, nlon, variables, seed=0):
  rng = np.random.default_rng(seed)
  init = np.datetime64('2020-01-01T00', 'ns') + np.arange(n_init) * 12 * H
  lead = (np.arange(n_lead) * 12 * H).astype('timedelta64[ns]')
  valid = np.datetime64('2020-01-01T00', 'ns') + np.arange(
      n_init + n_lead) * 12 * H
  grid = {'latitude': np.linspace(-90, 90, nlat),
          'longitude': np.linspace(0, 360, nlon, endpoint=False)}
  forecasts, analyses = {}, {}
  for var in variables:
    truth = rng.standard_normal((len(valid), nlat, nlon), dtype=np.float32)
    analyses[var] = xl.DataArray(
        truth, ('valid_time', 'latitude', 'longitude'),
        coords=dict(grid, valid_time=valid), name=var)
    fc = np.empty((n_init, n_lead, nlat, nlon), np.float32)
    for j in range(n_lead):
      fc[:, j] = truth[j:j + n_init] + np.float32(0.1 * (j + 1)) * (
          rng.standard_normal((n_init, nlat, nlon), dtype=np.float32))
    forecasts[var] = xl.DataArray(
        fc, ('init_time', 'lead_time', 'latitude', 'longitude'),
        coords=dict(grid, init_time=init, lead_time=lead), name=var)
  return init, lead, forecasts, analyses


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--n-init', type=int, default=16)
  ap.add_argument('--n-lead', type=int, default=4)
  ap.add_argument('--nlat', type=int, default=181)
  ap.add_argument('--nlon', type=int, default=360)
  ap.add_argument('--init-chunk', type=int, default=2)
  ap.add_argument('--out', default='/tmp/wbx_metrics.nc')
  ap.add_argument('--state-out', default=None)
  ap.add_argument('--temporal', action='store_true',
                  help='keep init_time (per-init_time state)')
  args = ap.parse_args()

  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

  variables = ['u10', 'v10', 't2m']
  init, lead, forecasts, analyses = synthetic_archive(
      args.n_init, args.n_lead, args.nlat, args.nlon, variables)
  regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
             'nh': ((20, 90), (0, 360)), 'sh': ((-90, -20), (0, 360))}
  metrics = {'rmse': deterministic.RMSE(), 'bias': deterministic.Bias(),
             'wind': deterministic.WindVectorRMSE('u10', 'v10', 'wind10')}
  reduce_dims = (['latitude', 'longitude'] if args.temporal else
                 ['init_time', 'latitude', 'longitude'])
  aggregator = aggregation.Aggregator(
      reduce_dims=reduce_dims, weigh_by=[weighting.GridAreaWeighting()],
      bin_by=[binning.Regions(regions)])
  times = time_chunks.TimeChunks(init, lead,
                                 init_time_chunk_size=args.init_chunk)
  start = time.perf_counter()
  out = pipeline.run_pipeline(
      times, array_loaders.PredictionsFromArrays(forecasts),
      array_loaders.TargetsFromArrays(analyses), metrics, aggregator,
      out_path=args.out, aggregation_state_out_path=args.state_out)
  seconds = time.perf_counter() - start
  values = out[None][1]
  if int(os.environ.get('RANK', '0')) == 0:
    points = args.n_init * args.n_lead * args.nlat * args.nlon * len(variables)
    print(json.dumps({
        'ranks': world, 'chunks': len(times), 'seconds': seconds,
        'grid_points_per_s': points / seconds,
        'rmse.t2m[global]': values['rmse.t2m'].sel(region='global'
                                                   ).values.tolist(),
        'wind.wind10[tropics]': values['wind.wind10'].sel(
            region='tropics').values.tolist(),
        'out': args.out}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
