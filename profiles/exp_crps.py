"""A/B timing of the CRPS kernels on one variable of config[2] (20 init x 50
members x 721 x 1440 f32 = 4.15 GB ensemble), CUDA events around the kernel on
its stream:  python profiles/exp_crps.py [steps]
Prints one JSON line per kernel policy (sort / pair / moments)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weatherbenchx_b200 import _cabi, aggregation, engine, weighting  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402
from weatherbenchx_b200.metrics import probabilistic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
members = int(os.environ.get('EXP_MEMBERS', '50'))
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
ctx = _cabi.get_context(0)
NLAT, NLON, n_init, m = 721, 1440, 20, members
lat = np.linspace(-90, 90, NLAT)
coords = {'init_time': np.arange(n_init), 'number': np.arange(m),
          'latitude': lat,
          'longitude': np.linspace(0, 360, NLON, endpoint=False)}
gen = torch.Generator(device=dev)
gen.manual_seed(3)
y = torch.empty((n_init, NLAT, NLON), device=dev).normal_(0, 1, generator=gen)
x = torch.empty((n_init, m, NLAT, NLON), device=dev).normal_(0, 1, generator=gen)
x += y[:, None]
preds = {'v': xl.DataArray(x, ('init_time', 'number', 'latitude', 'longitude'),
                           coords=coords, name='v')}
tgts = {'v': xl.DataArray(y, ('init_time', 'latitude', 'longitude'),
                          coords={k: coords[k] for k in
                                  ('init_time', 'latitude', 'longitude')},
                          name='v')}
agg = aggregation.Aggregator(reduce_dims=['init_time', 'latitude', 'longitude'],
                             weigh_by=[weighting.GridAreaWeighting()])
pts = n_init * NLAT * NLON
peak = 6534.5
try:
  peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:  # pylint: disable=broad-except
  pass


def timed(fn):
  fn(); fn()
  torch.cuda.synchronize()
  ctx.profile(True)
  ctx.kernel_time(reset=True)
  for _ in range(steps):
    out = fn()
  torch.cuda.synchronize()
  kms, kn = ctx.kernel_time(reset=True)
  ctx.profile(False)
  return kms / steps, out


results = {}
for name, policy, metrics in (
    ('sort', 'auto', {'crps': probabilistic.CRPSEnsemble(use_sort=True)}),
    ('sort+moments', 'auto', {
        'crps': probabilistic.CRPSEnsemble(use_sort=True),
        'ssr': probabilistic.UnbiasedSpreadSkillRatio()}),
    ('pair', 'pair', {'crps': probabilistic.CRPSEnsemble()}),
    # moments alone: register-resident members without the network (GPU call
    # 34 also ran the pair-kernel skeleton through a hook that is gone)
    ('moments', 'auto', {
        'ssr': probabilistic.UnbiasedSpreadSkillRatio(),
        'rmse': probabilistic.UnbiasedEnsembleMeanRMSE()})):
  if os.environ.get('EXP_ONLY') and name not in os.environ['EXP_ONLY'].split(','):
    continue
  engine.CRPS_KERNEL = policy
  kms, out = timed(lambda: aggregation.compute_metric_values_for_single_chunk(
      metrics, agg, preds, tgts))
  first = 'crps.v' if 'crps.v' in out else sorted(out)[0]
  results[name] = float(out[first].values)
  print(json.dumps({
      'kernel': name, 'members': m, 'kernel_ms': kms,
      'gpts_per_s': pts / kms / 1e6,
      'hbm_frac': pts * 4 * (m + 1) / (kms * 1e-3) / 1e9 / peak,
      first: results[name],
      'values': {k: float(out[k].values) for k in sorted(out)}}), flush=True)
engine.CRPS_KERNEL = 'auto'
if 'sort' in results and 'pair' in results:
  rel = abs(results['sort'] - results['pair']) / abs(results['pair'])
  print(json.dumps({'sort_vs_pair_rel_diff': rel}))
  assert rel < 1e-5, rel
