"""Round-2 experiment: K-threshold contingency launch, job order A/B.

    python profiles/exp_xf_l2.py            (one B200)

Times the XF reduction kernel (CUDA events around the kernel on its stream,
wbx_ctx_profile) for K = 1, 3, 6 thresholds on 20 x 721 x 1440 fields with
engine.XF_L2_BLOCK_BYTES = 0 (threshold-major job table: every threshold is a
pass over HBM) and with 32 / 64 / 96 MB blocks (each block of slabs is swept
for all thresholds, so the re-reads can hit L2).  Prints one JSON line per
configuration; the sums must agree between the orders.
"""

import json

import numpy as np
import torch

from weatherbenchx_b200 import _cabi, aggregation, engine, weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import categorical, wrappers

PEAK_GBS = 6490.8  # MEASURED_PEAKS.json hbm_gbs of round 1; re-read if it changed


def main():
  n_init, ny, nx = 20, 721, 1440
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': np.arange(n_init), 'lead_time': np.arange(1),
            'latitude': np.linspace(-90, 90, ny),
            'longitude': np.linspace(0, 360, nx, endpoint=False)}
  t = torch.empty((n_init, 1, ny, nx), device='cuda').exponential_(0.5)
  p = (t + torch.randn_like(t)).clamp_(0)
  P = {'rain': xl.DataArray(p, dims, coords=coords, name='rain')}
  T = {'rain': xl.DataArray(t, dims, coords=coords, name='rain')}
  # lead_time is kept, so every (init, threshold) pair is a job of one slab
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  ctx = _cabi.get_context()
  points = n_init * ny * nx
  for n_thr in (1, 3, 6):
    thresholds = list(np.linspace(0.1, 3.0, n_thr))
    both = [wrappers.ContinuousToBinary('both', thresholds, 'threshold')]
    metrics = {'ets': wrappers.WrappedMetric(categorical.ETS(), both)}
    reference = None
    for block_mb in (0, 32, 64, 96):
      engine.XF_L2_BLOCK_BYTES = block_mb << 20
      engine.clear_plan_cache()
      step = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
          metrics, aggregator, P, T)
      for _ in range(3):
        out = step()
      torch.cuda.synchronize()
      ctx.profile(True)
      ctx.kernel_time(reset=True)
      for _ in range(20):
        out = step()
      torch.cuda.synchronize()
      ms, n = ctx.kernel_time(reset=True)
      ctx.profile(False)
      values = out['ets.rain'].values
      if reference is None:
        reference = values
      np.testing.assert_allclose(values, reference, rtol=1e-9)
      per_launch = ms / max(n, 1)
      print(json.dumps({
          'thresholds': n_thr, 'l2_block_mb': block_mb,
          'kernel_ms': per_launch, 'launches': int(n),
          'points_per_s': points / (per_launch * 1e-3),
          'algorithmic_gbs_8B_per_point_per_threshold':
              points * 8 * n_thr / (per_launch * 1e-3) / 1e9,
          'frac_of_hbm_peak_single_pass_model':
              points * 8 / (per_launch * 1e-3) / 1e9 / PEAK_GBS}))
  engine.XF_L2_BLOCK_BYTES = 0


if __name__ == '__main__':
  main()
