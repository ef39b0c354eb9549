"""A/B of the sorting-network CRPS kernel with work moved to the FMA pipe
(crps.cu, sort_ce_mixed / sorted_moment): the last layer folded into the
moment and MIXPCT per cent of the compare-exchanges with max = (a + b) - min.

    python profiles/exp_crps_mix.py [steps]

1. parity on a small case with NaN members, NaN / inf targets and identical
   members: per-point skill and spread of every variant against the NumPy
   oracle and against the shipped kernel;
2. timing on one variable of config[2] (20 init x 50 members x 721 x 1440),
   CUDA events around the kernel on its stream.
WBX_EXP_SORT_MIX (read by crps_launch at every launch) selected the variant,
WBX_EXP_SORT_GRID the CTAs per SM of the grid; unset = the shipped kernel.
Both hooks existed only for GPU calls 24-26 of round 2 (results:
profiles/exp_crps_r2_mix_call24.log, _call25.txt, _call26.log); the library
now ships MIXPCT 45 at 4 resident CTAs and 24 CTAs per SM, so with today's
library every "variant" of this script runs that kernel."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import wbx_oracle as oracle  # noqa: E402  (checker only)
from weatherbenchx_b200 import _cabi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
VARIANTS = [None] + [int(v) for v in os.environ.get(
    'EXP_MIX', '0,30,40,140,240,340,230,250').split(',')]
# CTAs per SM of the grid (0 = the library's default of 8)
GRIDS = [int(v) for v in os.environ.get('EXP_GRID', '0,3,4,6,12').split(',')]
torch.cuda.set_device(0)
ctx = _cabi.get_context(0)
ctx.use_torch_stream()


def select(v, grid=0):
  if grid:
    os.environ['WBX_EXP_SORT_GRID'] = str(grid)
  else:
    os.environ.pop('WBX_EXP_SORT_GRID', None)
  if v is None:
    os.environ.pop('WBX_EXP_SORT_MIX', None)
  else:
    os.environ['WBX_EXP_SORT_MIX'] = str(v)


def plan_for(xd, yd, members, n_init, ny, nx, w_y=None):
  return _cabi.CrpsPlan(
      ctx, space=_cabi.SPACE_DEVICE,
      flags=_cabi.CRPS_FAIR | _cabi.CRPS_USE_SORT, ny=ny, nx=nx,
      n_members=members, member_stride=ny * nx, point_stride=1,
      ens=np.array([xd.data_ptr() + i * members * ny * nx * 4
                    for i in range(n_init)], np.uint64),
      target=np.array([yd.data_ptr() + i * ny * nx * 4
                       for i in range(n_init)], np.uint64),
      cell=np.zeros(n_init, np.int32), n_cells=1, w_y=w_y, stat_mask=3)


# -- 1. parity ---------------------------------------------------------------
rng = np.random.default_rng(5)
members, n_init, ny, nx = 50, 2, 16, 64
x = (280 + 3 * rng.normal(size=(n_init, members, ny, nx))).astype(np.float32)
y = (280 + 3 * rng.normal(size=(n_init, ny, nx))).astype(np.float32)
x[0, 7, 3, 5] = np.nan            # NaN member
x[1, :, 2, :8] = x[1, :1, 2, :8]  # identical members
x[1, 49, 9, 9] = np.inf           # one infinite member
y[0, 4, 4] = np.nan               # NaN target
y[0, 5, 5] = np.inf               # infinite target
y[1, 6, :] = np.nan               # a masked row of the analysis
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
with np.errstate(invalid='ignore'):
  want = [oracle.crps_skill(x, y, 1), oracle.crps_spread(x, 1, fair=True)]
plan = plan_for(xd, yd, members, n_init, ny, nx)
shipped = None
for v in VARIANTS:
  select(v)
  fields = [torch.full((n_init, ny, nx), -1.0, device='cuda') for _ in range(2)]
  plan.run_fields([fields[0].data_ptr(), fields[1].data_ptr(), None, None])
  got = [f.cpu().numpy() for f in fields]
  if v is None:
    shipped = got
  rec = {'check': 'parity', 'variant': v}
  for k, name in enumerate(('skill', 'spread')):
    same_nan = bool((np.isnan(got[k]) == np.isnan(want[k])).all())
    ok = np.isfinite(want[k]) & np.isfinite(got[k])
    rel = np.abs(got[k][ok] - want[k][ok]) / np.maximum(np.abs(want[k][ok]),
                                                        1e-30)
    rel[(want[k][ok] == 0) & (got[k][ok] == 0)] = 0
    rec[name] = {
        'nan_pattern_equal_oracle': same_nan,
        'nan_pattern_equal_shipped': bool(
            (np.isnan(got[k]) == np.isnan(shipped[k])).all()),
        'inf_equal_oracle': bool(
            (np.isinf(got[k]) == np.isinf(want[k])).all()),
        'max_rel_vs_oracle': float(rel.max()),
        'identical_members_exact_zero': bool(
            (got[k][1, 2, :8] == 0).all()) if name == 'spread' else None}
  print(json.dumps(rec), flush=True)

# -- 2. timing ---------------------------------------------------------------
NLAT, NLON, n_init, m = 721, 1440, 20, 50
gen = torch.Generator(device='cuda')
gen.manual_seed(3)
yb = torch.empty((n_init, NLAT, NLON), device='cuda').normal_(0, 1, generator=gen)
xb = torch.empty((n_init, m, NLAT, NLON), device='cuda').normal_(0, 1,
                                                                 generator=gen)
xb += yb[:, None]
w_y = np.cos(np.deg2rad(np.linspace(-90, 90, NLAT)))
pts = n_init * NLAT * NLON
peak = 6534.5
try:
  peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:  # pylint: disable=broad-except
  pass
base = None
for rep in range(2):   # twice: order effects / clocks
  for v in VARIANTS:
    for grid in GRIDS:
      select(v, grid)
      big = plan_for(xb, yb, m, n_init, NLAT, NLON, w_y=w_y)  # grid: at build
      big.run_to_host(); big.run_to_host()
      torch.cuda.synchronize()
      ctx.profile(True)
      ctx.kernel_time(reset=True)
      for _ in range(steps):
        ws, w = big.run_to_host()
      torch.cuda.synchronize()
      kms, kn = ctx.kernel_time(reset=True)
      ctx.profile(False)
      big.close()
      kms /= steps
      val = ws[0, :2] / w[0, :2]
      if base is None:
        base = val
      print(json.dumps({
          'check': 'timing', 'rep': rep, 'variant': v, 'ctas_per_sm': grid,
          'kernel_ms': round(kms, 4), 'gpts_per_s': round(pts / kms / 1e6, 3),
          'hbm_frac': round(pts * 4 * (m + 1) / (kms * 1e-3) / 1e9 / peak, 4),
          'rel_vs_shipped': [float(abs(val[k] - base[k]) / abs(base[k]))
                             for k in range(2)]}), flush=True)
select(None)
