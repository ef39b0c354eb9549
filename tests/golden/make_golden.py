"""Generates tests/golden/hotpath_golden.npz.

The reference itself cannot be imported in the build container (no xarray /
jax), so these vectors are NOT reference outputs.  They are produced by
scalar, loop-by-loop float64 evaluation of the formulas the reference
implements (cited per block), written independently of both the vectorised
oracle (oracle/wbx_oracle.py) and the CUDA kernels, so that a shared
vectorisation mistake cannot hide.  Inputs are float32 (the field dtype of the
path); statistics are rounded to float32 after every operation, as NumPy does
for the reference, and sums are accumulated in float64.

Run:  python tests/golden/make_golden.py
"""

import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def lat_weights(lat_deg):
  """weighting.py:62-88,105-130 (increasing latitude), scalar loops."""
  n = len(lat_deg)
  x = [math.radians(float(v)) for v in lat_deg]
  bounds = [0.0] * (n + 1)
  bounds[0] = max(x[0] - (x[1] - x[0]) / 2, -math.pi / 2)
  bounds[n] = min(x[-1] + (x[-1] - x[-2]) / 2, math.pi / 2)
  for i in range(1, n):
    bounds[i] = (x[i - 1] + x[i]) / 2
  w = [math.sin(bounds[i + 1]) - math.sin(bounds[i]) for i in range(n)]
  mean = sum(w) / n
  return np.array([v / mean for v in w])


def main():
  rng = np.random.default_rng(20260924)
  n_init, n_lead, n_lat, n_lon = 3, 2, 7, 8
  lat = np.linspace(-90, 90, n_lat)
  w = lat_weights(lat)
  p = rng.normal(280, 10, (n_init, n_lead, n_lat, n_lon)).astype(f32)
  t = (p + rng.normal(0, 2, p.shape)).astype(f32)
  c = rng.normal(280, 5, (4, n_lat, n_lon)).astype(f32)   # 4 climatology rows
  clim_row = rng.integers(0, 4, (n_init, n_lead))
  mask = rng.random((n_init, n_lead, n_lat, n_lon)) > 0.3
  t_nan = t.copy()
  t_nan[0, 1, 2, 3] = np.nan
  t_nan[2, 0, 5, 1] = np.nan

  # --- deterministic statistics, reduce (init, lat, lon), keep lead --------
  # deterministic.py:94-123,225-259 ; aggregation.py:337-366
  names = ['Error', 'AbsoluteError', 'SquaredError',
           'SquaredPredictionAnomaly', 'SquaredTargetAnomaly',
           'AnomalyCovariance']

  def stats_at(pv, tv, cv):
    d = f32(pv - tv)
    a = f32(pv - cv)
    b = f32(tv - cv)
    return [d, f32(abs(d)), f32(d * d), f32(a * a), f32(b * b), f32(a * b)]

  def reduce(tt, mode):
    sws = np.zeros((n_lead, 6))
    sw = np.zeros((n_lead, 6))
    for i in range(n_init):
      for l in range(n_lead):
        for y in range(n_lat):
          for x in range(n_lon):
            vals = stats_at(p[i, l, y, x], tt[i, l, y, x],
                            c[clim_row[i, l], y, x])
            for s, v in enumerate(vals):
              valid = True
              if mode in ('masked', 'masked_skipna'):
                valid = bool(mask[i, l, y, x])
              if mode in ('skipna', 'masked_skipna'):
                valid = valid and not math.isnan(float(v))
              if valid:
                sws[l, s] += float(v) * w[y]
                sw[l, s] += w[y]
              elif mode == 'propagate':
                raise AssertionError
    return sws, sw

  out = dict(p=p, t=t, t_nan=t_nan, c=c, clim_row=clim_row, mask=mask,
             lat=lat, w_lat=w, stat_names=np.array(names))
  out['det_propagate_sws'], out['det_propagate_sw'] = reduce(t, 'propagate')
  out['det_masked_sws'], out['det_masked_sw'] = reduce(t, 'masked')
  out['det_skipna_sws'], out['det_skipna_sw'] = reduce(t_nan, 'skipna')
  out['det_masked_skipna_sws'], out['det_masked_skipna_sw'] = reduce(
      t_nan, 'masked_skipna')

  # --- CRPS (probabilistic.py:129-145,194-247), M = 4 and 5 ----------------
  for m in (4, 5):
    x = rng.normal(0, 1, (n_init, n_lat, n_lon, m)).astype(f32)
    y = rng.normal(0, 1, (n_init, n_lat, n_lon)).astype(f32)
    skill = np.zeros((n_init, n_lat, n_lon))
    pair = np.zeros((n_init, n_lat, n_lon))
    for i in range(n_init):
      for a in range(n_lat):
        for b in range(n_lon):
          s = 0.0
          for k in range(m):
            s += float(f32(abs(f32(x[i, a, b, k] - y[i, a, b]))))
          skill[i, a, b] = s / m
          q = 0.0
          for k in range(m):
            for j in range(m):
              q += float(f32(abs(f32(x[i, a, b, k] - x[i, a, b, j]))))
          pair[i, a, b] = q
    out[f'crps{m}_x'] = x
    out[f'crps{m}_y'] = y
    out[f'crps{m}_skill'] = skill
    out[f'crps{m}_spread_fair'] = pair / (m * (m - 1))
    out[f'crps{m}_spread_unfair'] = pair / (m * m)

  np.savez_compressed(os.path.join(HERE, 'hotpath_golden.npz'), **out)
  print('wrote', os.path.join(HERE, 'hotpath_golden.npz'))


if __name__ == '__main__':
  main()
