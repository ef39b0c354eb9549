"""Pins the CPU oracle (oracle/wbx_oracle.py) to the reference's own tests.

Every test names the reference test it re-expresses (paths relative to
/root/reference/weatherbenchX).  This file anchors the oracle on the inline
known answers of the reference's tests; tests/golden/hotpath_golden.npz adds
loop-by-loop float64 vectors, and tests/test_reference_golden.py pins the
oracle to outputs of the reference's own code (run on stand-in xarray).
"""

import itertools
import os

import numpy as np
import pytest

import wbx_oracle as oracle
import wbx_test_utils as utils

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'hotpath_golden.npz')
RTOL = 1e-5  # xr.testing.assert_allclose default used by the reference tests


def _data(**kw):
  preds = utils.mock_prediction_data(
      time_start='2020-01-01T00', time_stop='2020-01-03T00', lead_start=0,
      lead_stop=1, **kw)
  return utils.rename_all(preds, time='init_time',
                          prediction_timedelta='lead_time')


def _rmse(p, t, reduce_dims, **kw):
  res = oracle.aggregate(oracle.squared_error(p.to_numpy(), t.to_numpy()),
                         p.dims, reduce_dims, **kw)
  if res is None:
    return None
  sws, sw, dims = res
  return oracle.rmse_from_mean(oracle.mean_statistic(sws, sw)), dims, sws, sw


def test_expected_output_rmse_is_one():
  """aggregation_test.py:69-103."""
  template = _data()
  for var, da in template.items():
    p = da
    t = da.copy(data=np.ones_like(da.to_numpy()))
    rmse, dims, sws, sw = _rmse(p, t, ['init_time', 'latitude', 'longitude'])
    expected_dims = (('lead_time', 'level') if var == 'geopotential'
                     else ('lead_time',))
    assert dims == expected_dims
    np.testing.assert_allclose(rmse, 1.0, rtol=RTOL)
    # state + state gives the same value
    np.testing.assert_allclose(np.sqrt((2 * sws) / (2 * sw)), 1.0, rtol=RTOL)


def test_missing_reduce_dims_drops_variable():
  """aggregation_test.py:105-119."""
  d = _data()
  p2 = d['2m_temperature']
  assert _rmse(p2, p2, ['level', 'latitude', 'longitude']) is None
  p3 = d['geopotential']
  assert _rmse(p3, p3, ['level', 'latitude', 'longitude']) is not None


def test_nan_mask_skipna_semantics():
  """aggregation_test.py:121-169."""
  p = _data()['geopotential']
  t_np = np.ones_like(p.to_numpy())
  lat = p.coords['latitude'].to_numpy()
  lat_axis = p.dims.index('latitude')
  shape = [1] * p.ndim
  shape[lat_axis] = -1
  t_np = np.where(lat.reshape(shape) > 0, t_np, np.nan).astype(np.float32)
  stat = oracle.squared_error(p.to_numpy(), t_np)
  mask = ~np.isnan(t_np)
  rd = ['init_time', 'latitude', 'longitude']
  sws, sw, _ = oracle.aggregate(stat, p.dims, rd)
  assert np.isnan(sws / sw).all()
  sws, sw, _ = oracle.aggregate(stat, p.dims, rd, mask=mask,
                                mask_dims=p.dims, masked=True)
  assert np.isfinite(sws / sw).all()
  np.testing.assert_allclose(np.sqrt(sws / sw), 1.0, rtol=RTOL)
  sws, sw, _ = oracle.aggregate(stat, p.dims, rd, skipna=True)
  assert np.isfinite(sws / sw).all()
  # masked=True but no mask on the variable -> NaN propagates
  sws, sw, _ = oracle.aggregate(stat, p.dims, rd, masked=True)
  assert np.isnan(sws / sw).any()


def test_weights_multiply():
  """aggregation_test.py:171-221: two 2x weightings -> sums x4, mean same."""
  p = _data()['geopotential']
  t = np.ones_like(p.to_numpy())
  stat = oracle.squared_error(p.to_numpy(), t)
  rd = ['init_time', 'latitude', 'longitude']
  sws, sw, _ = oracle.aggregate(stat, p.dims, rd)
  two = np.full(stat.shape, 2.0, dtype=stat.dtype)
  sws4, sw4, _ = oracle.aggregate(
      stat, p.dims, rd, weights=[(two, p.dims), (two, p.dims)])
  np.testing.assert_allclose(sws * 4, sws4, rtol=RTOL)
  np.testing.assert_allclose(sw * 4, sw4, rtol=RTOL)
  np.testing.assert_allclose(sws / sw, sws4 / sw4, rtol=RTOL)


def test_binning_adds_dims():
  """aggregation_test.py:223-246."""
  p = _data()['geopotential']
  lat, lon = p.coords['latitude'].to_numpy(), p.coords['longitude'].to_numpy()
  m1, _ = oracle.regions_masks(
      lat, lon, {'north': ((0, 90), (0, 360)), 'south': ((-90, 0), (0, 360))})
  m2, _ = oracle.regions_masks(
      lat, lon, {'east': ((-90, 90), (0, 180)), 'west': ((-90, 90), (180, 360))})
  stat = oracle.squared_error(p.to_numpy(), np.ones_like(p.to_numpy()))
  _, _, dims = oracle.aggregate(
      stat, p.dims, ['init_time', 'latitude', 'longitude'],
      bin_masks=[(m1, ('bins1', 'latitude', 'longitude')),
                 (m2, ('bins2', 'latitude', 'longitude'))])
  assert set(dims) == {'bins1', 'bins2', 'lead_time', 'level'}


def test_latitude_weights():
  """weighting_test.py:24-46."""
  lat = np.linspace(-90, 90, 19)
  w = oracle.grid_area_weights(lat)
  assert w.shape == lat.shape
  np.testing.assert_allclose(w.mean(), 1.0, rtol=1e-12)
  full = oracle.grid_area_weights(lat, return_normalized=False)
  sel = (lat >= -30) & (lat <= 30)
  regional = oracle.grid_area_weights(lat[sel], return_normalized=False)
  np.testing.assert_allclose(regional, full[sel], rtol=RTOL)
  # descending latitude gives the reversed weights (weighting.py:118-126)
  np.testing.assert_allclose(oracle.grid_area_weights(lat[::-1]), w[::-1])


def test_squared_error_statistic_values():
  """metrics/metrics_test.py:44-98: pred = target + 1 -> SE == 1, shape kept."""
  t = _data()['geopotential'].to_numpy()
  p = t + 1
  se = oracle.squared_error(p, t)
  assert se.shape == p.shape
  assert se.mean() == 1.0
  assert se.dtype == np.float32


@pytest.mark.parametrize('ensemble_size,use_sort,fair', list(
    itertools.product([4, 5], [False, True], [True, False])))
def test_crps_equals_brute_force(ensemble_size, use_sort, fair):
  """metrics/metrics_test.py:603-660 (8 cases)."""
  targets = _data(random=True)
  preds = _data(random=True, ensemble_size=ensemble_size)
  for v in ('2m_temperature', 'geopotential'):
    x, y = preds[v], targets[v]
    ens_axis = x.dims.index('realization')
    assert ens_axis == x.ndim - 1
    skill = oracle.crps_skill(x.to_numpy(), y.to_numpy(), ens_axis)
    spread = oracle.crps_spread(x.to_numpy(), ens_axis, fair=fair,
                                use_sort=use_sort)
    rd = ['latitude', 'longitude']
    dims = y.dims
    s_ws, s_w, _ = oracle.aggregate(skill, dims, rd)
    p_ws, p_w, _ = oracle.aggregate(spread, dims, rd)
    score = oracle.crps_from_means(s_ws / s_w, p_ws / p_w)
    # the reference's brute force (metrics_test.py:603-633)
    axes = tuple(dims.index(d) for d in rd)
    bf_spread = oracle.crps_spread_brute_force(x.to_numpy(), ens_axis, fair)
    bf_skill = np.abs(y.to_numpy()[..., None] - x.to_numpy()).mean(-1)
    expected = bf_skill.mean(axes) - 0.5 * bf_spread.mean(axes)
    np.testing.assert_allclose(score, expected, rtol=RTOL)


def test_acc_is_one():
  """metrics/metrics_test.py:983-1006: pred == target, clim = target - 1."""
  preds = utils.rename_all(
      utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-02T00'),
      time='init_time', prediction_timedelta='lead_time')
  for da in preds.values():
    p = da.to_numpy()
    field = da.isel(init_time=0, lead_time=0).to_numpy()
    clim = np.broadcast_to(field - 1, (366, 4) + field.shape)
    cdims = ('dayofyear', 'hour') + da.dims[2:]
    aligned, adims = oracle.align_climatology(
        clim, cdims, {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)},
        da.coords['init_time'].to_numpy(), da.coords['lead_time'].to_numpy())
    aligned = np.transpose(aligned, [adims.index(d) for d in da.dims])
    rd = ['latitude', 'longitude']
    means = {}
    for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
      sws, sw, _ = oracle.aggregate(fn(p, p, aligned), da.dims, rd)
      means[name] = sws / sw
    acc = oracle.acc_from_means(means['AnomalyCovariance'],
                                means['SquaredPredictionAnomaly'],
                                means['SquaredTargetAnomaly'])
    np.testing.assert_allclose(acc, 1.0, rtol=RTOL)


@pytest.mark.parametrize('ensemble_size,fair', list(
    itertools.product([4, 5], [True, False])))
def test_crps_nan_is_missing_member(ensemble_size, fair):
  """metrics/metrics_test.py:1199-1274."""
  rng = np.random.default_rng(3)
  x = rng.random((2, 5, 6, ensemble_size))
  y = rng.random((2, 5, 6))
  x_nan = x.copy()
  x_nan[..., 0] = np.nan
  got_skill = oracle.crps_skill(x_nan, y, -1, skipna_ensemble=True)
  got_spread = oracle.crps_spread(x_nan, -1, fair=fair, skipna_ensemble=True)
  exp_skill = oracle.crps_skill(x[..., 1:], y, -1)
  exp_spread = oracle.crps_spread(x[..., 1:], -1, fair=fair)
  np.testing.assert_allclose(got_skill, exp_skill, rtol=RTOL)
  np.testing.assert_allclose(got_spread, exp_spread, rtol=RTOL)


def test_crps_needs_two_members_and_sort_rejects_skipna():
  """probabilistic.py:210-216."""
  with pytest.raises(ValueError):
    oracle.crps_spread(np.zeros((3, 1)), -1)
  with pytest.raises(ValueError):
    oracle.crps_spread(np.zeros((3, 4)), -1, use_sort=True,
                       skipna_ensemble=True)


def test_ensemble_averaged_rmse():
  """metrics/metrics_test.py:1276-1308: reducing 'realization' in the
  Aggregator == averaging the squared error over members first."""
  targets = _data(random=True)['geopotential']
  preds = _data(random=True, ensemble_size=5)['geopotential']
  se = oracle.squared_error(preds.to_numpy(), targets.to_numpy()[..., None])
  a_ws, a_w, _ = oracle.aggregate(
      se, preds.dims, ['latitude', 'longitude', 'realization'])
  b_ws, b_w, _ = oracle.aggregate(
      se.mean(-1), targets.dims, ['latitude', 'longitude'])
  np.testing.assert_allclose(np.sqrt(a_ws / a_w), np.sqrt(b_ws / b_w),
                             rtol=RTOL)


def test_rankdata_is_ordinal():
  """probabilistic.py:148-158."""
  x = np.array([[0.3, 0.1, 0.2], [5.0, 7.0, 6.0]])
  np.testing.assert_array_equal(oracle.rankdata(x, -1),
                                [[3, 1, 2], [1, 3, 2]])


# ---------------------------------------------------------------------------
# golden vectors (scalar float64 loops, tests/golden/make_golden.py)
# ---------------------------------------------------------------------------


@pytest.fixture(scope='module')
def golden():
  return np.load(GOLDEN)


@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna',
                                  'masked_skipna'])
def test_golden_deterministic(golden, mode):
  g = golden
  p = g['p']
  t = g['t_nan'] if 'skipna' in mode else g['t']
  c = g['c'][g['clim_row']]
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  fns = dict(oracle.DETERMINISTIC_STATISTICS)
  for i, name in enumerate(g['stat_names']):
    if name in fns:
      stat = fns[name](p, t)
    else:
      stat = oracle.CLIMATOLOGY_STATISTICS[str(name)](p, t, c)
    sws, sw, out_dims = oracle.aggregate(
        stat, dims, ['init_time', 'latitude', 'longitude'],
        weights=[(g['w_lat'], ('latitude',))], mask=g['mask'],
        mask_dims=dims, masked='masked' in mode, skipna='skipna' in mode)
    assert out_dims == ('lead_time',)
    np.testing.assert_allclose(sws, g[f'det_{mode}_sws'][:, i], rtol=1e-12)
    np.testing.assert_allclose(sw, g[f'det_{mode}_sw'][:, i], rtol=1e-12)


@pytest.mark.parametrize('m', [4, 5])
def test_golden_crps(golden, m):
  x, y = golden[f'crps{m}_x'], golden[f'crps{m}_y']
  np.testing.assert_allclose(oracle.crps_skill(x, y, -1),
                             golden[f'crps{m}_skill'], rtol=2e-6)
  for fair in (True, False):
    key = f'crps{m}_spread_{"fair" if fair else "unfair"}'
    for use_sort in (False, True):
      np.testing.assert_allclose(
          oracle.crps_spread(x, -1, fair=fair, use_sort=use_sort),
          golden[key], rtol=1e-5, atol=1e-6)


def test_golden_weights(golden):
  np.testing.assert_allclose(oracle.grid_area_weights(golden['lat']),
                             golden['w_lat'], rtol=1e-13)


# ---------------------------------------------------------------------------
# zonal energy spectrum: PARITY UNPINNED (no reference implementation); only
# numpy.fft identities are available.
# ---------------------------------------------------------------------------


def test_spectrum_identities():
  n, nlat = 48, 5
  lat = np.linspace(-60, 60, nlat)
  circ = 2 * np.pi * oracle.EARTH_RADIUS_M * np.cos(np.deg2rad(lat))
  lon = np.arange(n)
  # constant field -> only k = 0, S0 = C * c^2
  s = oracle.zonal_energy_spectrum(np.full((nlat, n), 3.0), lat)
  np.testing.assert_allclose(s[:, 0], circ * 9.0, rtol=1e-12)
  np.testing.assert_allclose(s[:, 1:], 0, atol=1e-6)
  # single sinusoid cos(2 pi k0 l / N) -> S[k0] = C / 2
  k0 = 5
  f = np.broadcast_to(np.cos(2 * np.pi * k0 * lon / n), (nlat, n))
  s = oracle.zonal_energy_spectrum(f, lat)
  np.testing.assert_allclose(s[:, k0], circ / 2, rtol=1e-10)
  # Parseval: sum_k S[k] = C / N * sum_l f^2   (N even: Nyquist counted twice
  # by the factor 2, so compare with the Nyquist term halved)
  rng = np.random.default_rng(0)
  f = rng.normal(size=(nlat, n))
  s = oracle.zonal_energy_spectrum(f, lat)
  total = s.sum(-1) - s[:, -1] / 2
  np.testing.assert_allclose(total, circ / n * (f ** 2).sum(-1), rtol=1e-10)
  # scaling by a -> a^2
  np.testing.assert_allclose(oracle.zonal_energy_spectrum(2 * f, lat), 4 * s,
                             rtol=1e-12)


# ---------------------------------------------------------------------------
# ensemble moments (probabilistic.py:250-336).  The reference's own
# known-answer test is statistical: metrics_test.py:947-983
# (test_spread_skill_ratio) draws targets and a 5-member ensemble iid from the
# same distribution and expects the unbiased spread-skill ratio to be 1 within
# 4 / sqrt(sample_size * ensemble_size).
# ---------------------------------------------------------------------------


def test_spread_skill_ratio_of_iid_ensemble_is_one():
  ensemble_size = 5
  shape = (2, 19, 36)       # time, latitude, longitude of the mock data
  # test_utils.py:43-46: uniform [0, 1) samples, seeds 0 (targets) and 1
  y = np.random.default_rng(0).random(size=shape).astype(np.float32)
  x = np.random.default_rng(1).random(
      size=(ensemble_size,) + shape).astype(np.float32)
  var = oracle.ensemble_variance(x, 0)
  umse = oracle.unbiased_ensemble_mean_squared_error(x, y, 0)
  ratio = np.sqrt(var.mean() / umse.mean())
  atol = 4 / np.sqrt(y.size * ensemble_size)
  assert abs(ratio - 1.0) < atol


def test_ensemble_moments_closed_form():
  # members 1, 2, 3, 6 -> mean 3, unbiased variance 14 / 3
  x = np.array([1.0, 2.0, 3.0, 6.0], np.float32).reshape(4, 1)
  y = np.array([1.0], np.float32)
  np.testing.assert_allclose(oracle.ensemble_variance(x, 0), [14 / 3],
                             rtol=1e-6)
  np.testing.assert_allclose(
      oracle.unbiased_ensemble_mean_squared_error(x, y, 0),
      [4.0 - 14 / 12], rtol=1e-6)
  # skipna_ensemble: a NaN member is a missing member (n = 3 -> mean 2, var 1)
  xn = np.array([1.0, 2.0, 3.0, np.nan], np.float32).reshape(4, 1)
  np.testing.assert_allclose(
      oracle.ensemble_variance(xn, 0, skipna_ensemble=True), [1.0], rtol=1e-6)
  np.testing.assert_allclose(
      oracle.unbiased_ensemble_mean_squared_error(xn, y, 0, True),
      [1.0 - 1 / 3], rtol=1e-6)
  assert np.isnan(oracle.ensemble_variance(xn, 0))
  # a single member has no unbiased variance
  assert np.isnan(oracle.ensemble_variance(x[:1], 0))
