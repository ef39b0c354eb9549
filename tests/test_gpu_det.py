"""GPU parity tests of the deterministic statistics + aggregation path.

CUDA path (through the class surface and through the C ABI) versus the CPU
oracle on the same seeded inputs, versus the reference's inline known answers,
versus tests/golden, and -- at BASELINE sizes -- through size-independent
properties.  Tolerance: 1e-5 relative (the reference's assert_allclose default
and the north_star's f32 tolerance); counts / unweighted sum_weights exact.
"""

import os

import numpy as np
import pytest

import wbx_oracle as oracle
import wbx_test_utils as utils
from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import binning
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import deterministic

pytestmark = pytest.mark.gpu
RTOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'hotpath_golden.npz')


def _test_data():
  template = utils.rename_all(
      utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-03T00', lead_start=0,
                                 lead_stop=1),
      time='init_time', prediction_timedelta='lead_time')
  predictions = {k: xl.zeros_like(v) for k, v in template.items()}
  targets = {k: xl.ones_like(v) for k, v in template.items()}
  return predictions, targets


def _aggregate(metrics, predictions, targets, **kwargs):
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  return aggregation.Aggregator(**kwargs).aggregate_statistics(statistics)


# ---------------------------------------------------------------------------
# reference known-answer tests, through the class surface
# ---------------------------------------------------------------------------


def test_expected_output():
  """aggregation_test.py:69-103."""
  predictions, targets = _test_data()
  metrics = {'rmse': deterministic.RMSE()}
  state = _aggregate(metrics, predictions, targets,
                     reduce_dims=['init_time', 'latitude', 'longitude'])
  actual = state.metric_values(metrics)
  summed = (state + state).metric_values(metrics)
  assert actual['rmse.2m_temperature'].dims == ('lead_time',)
  assert set(actual['rmse.geopotential'].dims) == {'lead_time', 'level'}
  for v in ('rmse.2m_temperature', 'rmse.geopotential'):
    np.testing.assert_allclose(actual[v].values, 1.0, rtol=RTOL)
    np.testing.assert_allclose(summed[v].values, 1.0, rtol=RTOL)
  np.testing.assert_array_equal(
      actual['rmse.geopotential'].coords['level'].values, [500, 700, 850])


def test_missing_reduce_dims():
  """aggregation_test.py:105-119."""
  predictions, targets = _test_data()
  metrics = {'rmse': deterministic.RMSE()}
  values = _aggregate(metrics, predictions, targets,
                      reduce_dims=['level', 'latitude', 'longitude']
                      ).metric_values(metrics)
  assert list(values) == ['rmse.geopotential']


def test_nan_handling():
  """aggregation_test.py:121-169."""
  predictions, targets = _test_data()
  targets = {k: v.where(v.coords['latitude'] > 0) for k, v in targets.items()}
  targets = utils.add_nan_mask_to_data(targets)
  metrics = {'rmse': deterministic.RMSE()}
  rd = ['init_time', 'latitude', 'longitude']
  actual = _aggregate(metrics, predictions, targets,
                      reduce_dims=rd).metric_values(metrics)
  for v in actual.values():
    assert np.isnan(v.values).all()
  actual = _aggregate(metrics, predictions, targets, reduce_dims=rd,
                      masked=True).metric_values(metrics)
  for v in actual.values():
    assert not np.isnan(v.values).any()
    np.testing.assert_allclose(v.values, 1.0, rtol=RTOL)
  actual = _aggregate(metrics, predictions, targets, reduce_dims=rd,
                      skipna=True).metric_values(metrics)
  for v in actual.values():
    assert not np.isnan(v.values).any()
  targets['2m_temperature'] = targets['2m_temperature'].drop_vars('mask')
  actual = _aggregate(metrics, predictions, targets, reduce_dims=rd,
                      masked=True).metric_values(metrics)
  assert not np.isnan(actual['rmse.geopotential'].values).any()
  assert np.isnan(actual['rmse.2m_temperature'].values).any()


def test_weighting():
  """aggregation_test.py:171-221."""
  predictions, targets = _test_data()
  metrics = {'rmse': deterministic.RMSE()}

  class TwoTimes(weighting.Weighting):

    def weights(self, statistic):
      return xl.DataArray(np.full(statistic.shape, 2.0, np.float32),
                          statistic.dims)

  rd = ['init_time', 'latitude', 'longitude']
  base = _aggregate(metrics, predictions, targets, reduce_dims=rd)
  four = _aggregate(metrics, predictions, targets, reduce_dims=rd,
                    weigh_by=[TwoTimes(), TwoTimes()])
  for stat in base.sum_weighted_statistics:
    for var in base.sum_weighted_statistics[stat]:
      xl.testing.assert_allclose(base.sum_weighted_statistics[stat][var] * 4,
                                 four.sum_weighted_statistics[stat][var])
      xl.testing.assert_allclose(base.sum_weights[stat][var] * 4,
                                 four.sum_weights[stat][var])
  a, b = base.metric_values(metrics), four.metric_values(metrics)
  for v in a:
    xl.testing.assert_allclose(a[v], b[v])


def test_binning():
  """aggregation_test.py:223-246 (+ values against the oracle)."""
  predictions, targets = _test_data()
  rng = np.random.default_rng(0)
  predictions = {k: v.copy(data=rng.normal(size=v.shape).astype(np.float32))
                 for k, v in predictions.items()}
  metrics = {'rmse': deterministic.RMSE()}
  r1 = {'north': ((0, 90), (0, 360)), 'south': ((-90, 0), (0, 360))}
  r2 = {'east': ((-90, 90), (0, 180)), 'west': ((-90, 90), (180, 360))}
  state = _aggregate(
      metrics, predictions, targets,
      reduce_dims=['init_time', 'latitude', 'longitude'],
      bin_by=[binning.Regions(r1, bin_dim_name='bins1'),
              binning.Regions(r2, bin_dim_name='bins2')])
  actual = state.metric_values(metrics)
  assert set(actual.dims) == {'bins1', 'bins2', 'lead_time', 'level'}
  p, t = predictions['geopotential'], targets['geopotential']
  lat, lon = p.coords['latitude'].values, p.coords['longitude'].values
  m1, _ = oracle.regions_masks(lat, lon, r1)
  m2, _ = oracle.regions_masks(lat, lon, r2)
  sws, sw, dims = oracle.aggregate(
      oracle.squared_error(p.values, t.values), p.dims,
      ['init_time', 'latitude', 'longitude'],
      bin_masks=[(m1, ('bins1', 'latitude', 'longitude')),
                 (m2, ('bins2', 'latitude', 'longitude'))])
  got = actual['rmse.geopotential'].transpose(*dims).values
  np.testing.assert_allclose(got, np.sqrt(sws / sw), rtol=RTOL)


def test_statistics_computation():
  """metrics/metrics_test.py:44-98 (materialised statistic values)."""
  target = utils.mock_prediction_data(time_start='2020-01-01T00',
                                      time_stop='2020-01-04T00')
  prediction = {k: v + 1 for k, v in target.items()}
  metrics = {'rmse': deterministic.RMSE()}
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, prediction, target)
  assert set(stats['SquaredError']) == set(target)
  se = stats['SquaredError']['geopotential']
  assert se.shape == prediction['geopotential'].shape
  assert se.values.mean() == 1.0
  assert se.values.dtype == np.float32


def test_acc_is_one():
  """metrics/metrics_test.py:983-1006."""
  prediction = utils.rename_all(
      utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-02T00'),
      time='init_time', prediction_timedelta='lead_time')
  target = dict(prediction)
  climatology = {}
  for k, v in target.items():
    field = v.isel(init_time=0, lead_time=0, drop=True) - 1
    climatology[k] = field.expand_dims(
        {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)})
  metrics = {'acc': deterministic.ACC(climatology=climatology)}
  state = _aggregate(metrics, prediction, target,
                     reduce_dims=['latitude', 'longitude'])
  results = state.metric_values(metrics)
  assert set(results) == {'acc.2m_temperature', 'acc.geopotential'}
  for v in results.values():
    np.testing.assert_allclose(v.values, 1.0, rtol=RTOL)


# ---------------------------------------------------------------------------
# fused kernel vs oracle on random data
# ---------------------------------------------------------------------------


def _random_case(seed, shape=(3, 4, 2, 24, 40), nan=False):
  rng = np.random.default_rng(seed)
  dims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  coords = {
      'init_time': (np.datetime64('2020-02-27T00', 'ns') +
                    np.arange(shape[0]) * np.timedelta64(12, 'h')),
      'lead_time': (np.arange(shape[1]) * np.timedelta64(6, 'h')
                    ).astype('timedelta64[ns]'),
      'level': np.arange(shape[2]) * 100 + 500,
      'latitude': np.linspace(-90, 90, shape[3]),
      'longitude': np.linspace(0, 360, shape[4], endpoint=False),
  }
  p = rng.normal(280, 10, shape).astype(np.float32)
  t = (p + rng.normal(0, 2, shape)).astype(np.float32)
  if nan:
    t[rng.random(shape) < 0.05] = np.nan
    p[rng.random(shape) < 0.01] = np.nan
  clim = rng.normal(280, 5, (366, 4) + shape[2:]).astype(np.float32)
  cdims = ('dayofyear', 'hour') + dims[2:]
  ccoords = {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
             **{d: coords[d] for d in dims[2:]}}
  P = xl.DataArray(p, dims, coords=coords, name='z')
  T = xl.DataArray(t, dims, coords=coords, name='z')
  C = xl.DataArray(clim, cdims, coords=ccoords, name='z')
  return P, T, C


def _oracle_states(P, T, C, reduce_dims, weights=(), mask=None, masked=False,
                   skipna=False):
  aligned, adims = oracle.align_climatology(
      C.values, C.dims, {k: C.coords[k].values for k in ('dayofyear', 'hour')},
      P.coords['init_time'].values, P.coords['lead_time'].values)
  aligned = np.transpose(aligned, [adims.index(d) for d in P.dims])
  out = {}
  for name, fn in oracle.DETERMINISTIC_STATISTICS.items():
    out[name] = fn(P.values, T.values)
  for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
    out[name] = fn(P.values, T.values, aligned)
  # The mask sits on the targets: (p - c)**2 does not inherit it and stays
  # unmasked (deterministic.py:225-232, aggregation.py:339).
  return {
      name: oracle.aggregate(
          stat, P.dims, reduce_dims, weights=weights,
          mask=None if name == 'SquaredPredictionAnomaly' else mask,
          mask_dims=P.dims, masked=masked, skipna=skipna)
      for name, stat in out.items()
  }


ALL_METRICS = lambda C: {  # noqa: E731
    'bias': deterministic.Bias(), 'mae': deterministic.MAE(),
    'mse': deterministic.MSE(), 'rmse': deterministic.RMSE(),
    'acc': deterministic.ACC({'z': C}),
}


def _check_state(state, expected, var='z', exact_weights=False):
  for name, (sws, sw, dims) in expected.items():
    got_ws = state.sum_weighted_statistics[name][var].transpose(*dims).values
    got_w = state.sum_weights[name][var].transpose(*dims).values
    np.testing.assert_allclose(got_ws, sws, rtol=RTOL,
                               atol=1e-6 * np.abs(sws).max(), equal_nan=True)
    if exact_weights:
      np.testing.assert_array_equal(got_w, sw)
    else:
      np.testing.assert_allclose(got_w, sw, rtol=1e-12)


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna',
                                  'masked_skipna'])
@pytest.mark.parametrize('reduce_dims', [
    ('init_time', 'latitude', 'longitude'),
    ('latitude', 'longitude'),
    ('init_time', 'lead_time', 'level', 'latitude', 'longitude'),
    ('longitude',),
])
def test_fused_matches_oracle(space, mode, reduce_dims):
  P, T, C = _random_case(1, nan='skipna' in mode)
  mask_np = np.random.default_rng(5).random(P.shape) > 0.25
  masked = 'masked' in mode
  skipna = 'skipna' in mode
  if masked:
    T = T.assign_coords(mask=xl.DataArray(mask_np, P.dims))
  w = oracle.grid_area_weights(P.coords['latitude'].values)
  expected = _oracle_states(
      P, T, C, reduce_dims, weights=[(w, ('latitude',))], mask=mask_np,
      masked=masked, skipna=skipna)
  if space == 'device':
    P, T, C = engine.to_device(P), engine.to_device(T), engine.to_device(C)
    if masked:
      T = T.assign_coords(mask=engine.to_device(xl.DataArray(mask_np, P.dims)))
  state = _aggregate(ALL_METRICS(C), {'z': P}, {'z': T},
                     reduce_dims=list(reduce_dims),
                     weigh_by=[weighting.GridAreaWeighting()],
                     masked=masked, skipna=skipna)
  _check_state(state, expected)
  # metric values too
  values = state.metric_values(ALL_METRICS(C))
  sws, sw, dims = expected['SquaredError']
  np.testing.assert_allclose(
      values['rmse.z'].transpose(*dims).values, np.sqrt(sws / sw), rtol=RTOL,
      equal_nan=True)


@pytest.mark.parametrize('flag', [_cabi.FLAG_FORCE_LDG, _cabi.FLAG_FORCE_TMA])
def test_tma_and_ldg_paths_agree_bitwise_on_weights(flag):
  P, T, C = _random_case(2)
  stats = [deterministic.SquaredError().compute({'z': P}, {'z': T})['z'],
           deterministic.AbsoluteError().compute({'z': P}, {'z': T})['z']]
  w = weighting.GridAreaWeighting().weights(stats[0])
  rd = ['init_time', 'latitude', 'longitude']
  ref = engine.aggregate_fused(stats, rd, [w])
  engine.clear_plan_cache()
  got = engine.aggregate_fused(stats, rd, [w], flags_extra=flag)
  engine.clear_plan_cache()
  for k in ref:
    np.testing.assert_allclose(got[k][0].values, ref[k][0].values, rtol=1e-12)
    np.testing.assert_array_equal(got[k][1].values, ref[k][1].values)


def test_unaligned_shapes_use_scalar_path():
  """19 x 37 slabs (703 elements, not a multiple of 4) and lat innermost."""
  rng = np.random.default_rng(3)
  dims = ('init_time', 'longitude', 'latitude')
  coords = {'init_time': np.arange(5), 'longitude': np.arange(37) * 5.0,
            'latitude': np.linspace(-90, 90, 19)}
  p = xl.DataArray(rng.normal(size=(5, 37, 19)).astype(np.float32), dims,
                   coords=coords, name='v')
  t = xl.DataArray(rng.normal(size=(5, 37, 19)).astype(np.float32), dims,
                   coords=coords, name='v')
  w = oracle.grid_area_weights(coords['latitude'])
  for rd in (['longitude', 'latitude'], ['init_time', 'longitude', 'latitude']):
    state = _aggregate({'mse': deterministic.MSE(), 'mae': deterministic.MAE()},
                       {'v': p}, {'v': t}, reduce_dims=rd,
                       weigh_by=[weighting.GridAreaWeighting()])
    for name, fn in (('SquaredError', oracle.squared_error),
                     ('AbsoluteError', oracle.absolute_error)):
      sws, sw, out_dims = oracle.aggregate(
          fn(p.values, t.values), dims, rd, weights=[(w, ('latitude',))])
      got = state.sum_weighted_statistics[name]['v']
      np.testing.assert_allclose(got.values, sws, rtol=RTOL)
      np.testing.assert_allclose(state.sum_weights[name]['v'].values, sw,
                                 rtol=1e-12)


def test_broadcast_targets_and_ensemble_reduce():
  """metrics/metrics_test.py:1276-1308: predictions carry 'realization',
  targets do not; reducing it in the Aggregator == member-averaged MSE."""
  d = utils.rename_all(
      utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-03T00', lead_start=0,
                                 lead_stop=1, random=True, seed=1),
      time='init_time', prediction_timedelta='lead_time')
  e = utils.rename_all(
      utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-03T00', lead_start=0,
                                 lead_stop=1, random=True, seed=2,
                                 ensemble_size=5),
      time='init_time', prediction_timedelta='lead_time')
  metrics = {'rmse': deterministic.RMSE()}
  state = _aggregate(metrics, e, d,
                     reduce_dims=['latitude', 'longitude', 'realization'])
  values = state.metric_values(metrics)
  for v in ('geopotential', '2m_temperature'):
    se = oracle.squared_error(e[v].values.astype(np.float32),
                              d[v].values.astype(np.float32)[..., None])
    sws, sw, dims = oracle.aggregate(
        se, e[v].dims, ['latitude', 'longitude', 'realization'])
    np.testing.assert_allclose(
        values[f'rmse.{v}'].transpose(*dims).values, np.sqrt(sws / sw),
        rtol=RTOL)


# ---------------------------------------------------------------------------
# golden fixtures through the C ABI
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna',
                                  'masked_skipna'])
def test_golden_through_cabi(mode):
  g = np.load(GOLDEN)
  p = np.ascontiguousarray(g['p'])
  t = np.ascontiguousarray(g['t_nan'] if 'skipna' in mode else g['t'])
  c = np.ascontiguousarray(g['c'])
  mask = np.ascontiguousarray(g['mask']).view(np.uint8)
  n_init, n_lead, ny, nx = p.shape
  slab = ny * nx
  # job order: lead (kept) major, init (reduced) minor
  jobs = [(i, l) for l in range(n_lead) for i in range(n_init)]
  addr = lambda a, i, l, item: a.ctypes.data + ((i * n_lead + l) * slab) * item
  plan = _cabi.DetPlan(
      _cabi.get_context(), space=_cabi.SPACE_HOST,
      flags=(_cabi.FLAG_MASKED if 'masked' in mode else 0) |
      (_cabi.FLAG_SKIPNA if 'skipna' in mode else 0),
      ny=ny, nx=nx,
      pred=np.array([addr(p, i, l, 4) for i, l in jobs], np.uint64),
      target=np.array([addr(t, i, l, 4) for i, l in jobs], np.uint64),
      clim=np.array([c.ctypes.data + int(g['clim_row'][i, l]) * slab * 4
                     for i, l in jobs], np.uint64),
      mask=(np.array([addr(mask, i, l, 1) for i, l in jobs], np.uint64)
            if 'masked' in mode else None),
      cell=np.array([l for _, l in jobs], np.int32), n_cells=n_lead,
      w_y=g['w_lat'])
  ws, w = plan.run_to_host()
  exp_ws, exp_w = g[f'det_{mode}_sws'], g[f'det_{mode}_sw']
  np.testing.assert_allclose(ws, exp_ws, rtol=RTOL,
                             atol=1e-6 * np.abs(exp_ws).max())
  for slot in range(6):
    np.testing.assert_allclose(w[:, _cabi.STAT_WCLASS[slot]], exp_w[:, slot],
                               rtol=1e-12)


# ---------------------------------------------------------------------------
# BASELINE-size properties (0.25 degree fields, device resident)
# ---------------------------------------------------------------------------


@pytest.fixture(scope='module')
def quarter_degree():
  import torch
  torch.manual_seed(0)
  n_init, ny, nx = 20, 721, 1440
  t = torch.randn(n_init, ny, nx, device='cuda') * 10 + 280
  p = t + torch.randn(n_init, ny, nx, device='cuda') * 2
  lat = np.linspace(-90, 90, ny)
  coords = {'init_time': np.arange(n_init), 'latitude': lat,
            'longitude': np.linspace(0, 360, nx, endpoint=False)}
  dims = ('init_time', 'latitude', 'longitude')
  return (xl.DataArray(p, dims, coords=coords, name='t2m'),
          xl.DataArray(t, dims, coords=coords, name='t2m'))


def _fused(P, T, reduce_dims, weights=True, kinds=('SquaredError',), **kw):
  stats = [engine.LazyStatistic(k, P, T) for k in kinds]
  w = [weighting.GridAreaWeighting().weights(stats[0])] if weights else []
  return engine.aggregate_fused(stats, list(reduce_dims), w, **kw)


def test_full_size_counts_are_exact(quarter_degree):
  P, T = quarter_degree
  res = _fused(P, T, P.dims, weights=False)
  sw = res['SquaredError'][1].values
  assert sw == 20 * 721 * 1440  # 20 764 800 > 2**24: exact only in f64/int
  res = _fused(P, T, P.dims, weights=False, skipna=True)
  assert res['SquaredError'][1].values == 20 * 721 * 1440


def test_full_size_matches_torch_f64_and_is_deterministic(quarter_degree):
  import torch
  P, T = quarter_degree
  res = _fused(P, T, P.dims, kinds=('Error', 'AbsoluteError', 'SquaredError'))
  w = torch.as_tensor(oracle.grid_area_weights(P.coords['latitude'].values),
                      device='cuda')
  d = (P.data - T.data)
  ref = {
      'Error': (d.double() * w[None, :, None]).sum().item(),
      'AbsoluteError': (d.abs().double() * w[None, :, None]).sum().item(),
      'SquaredError': ((d * d).double() * w[None, :, None]).sum().item(),
  }
  for k, v in ref.items():
    np.testing.assert_allclose(res[k][0].values, v, rtol=RTOL,
                               atol=1e-9 * 20 * 721 * 1440)
  np.testing.assert_allclose(res['SquaredError'][1].values, 20 * 721 * 1440,
                             rtol=1e-12)
  again = _fused(P, T, P.dims, kinds=('Error', 'AbsoluteError', 'SquaredError'))
  for k in ref:
    assert again[k][0].values.tobytes() == res[k][0].values.tobytes()


def test_full_size_chunk_combine_equals_monolithic(quarter_degree):
  """beam_pipeline_test.py:82-170 identity: per-chunk states summed == all."""
  P, T = quarter_degree
  whole = _fused(P, T, P.dims)['SquaredError']
  parts = []
  for lo in range(0, 20, 5):
    sl = {'init_time': slice(lo, lo + 5)}
    parts.append(_fused(P.isel(sl), T.isel(sl), P.dims)['SquaredError'])
  total_ws = sum(p[0].values for p in parts)
  total_w = sum(p[1].values for p in parts)
  np.testing.assert_allclose(total_ws, whole[0].values, rtol=1e-12)
  np.testing.assert_allclose(total_w, whole[1].values, rtol=1e-12)
  # keeping init_time gives per-init sums that add up to the total
  per_init = _fused(P, T, ('latitude', 'longitude'))['SquaredError']
  assert per_init[0].dims == ('init_time',)
  np.testing.assert_allclose(per_init[0].values.sum(), whole[0].values,
                             rtol=1e-12)


def test_full_size_known_answer_and_scaling(quarter_degree):
  P, T = quarter_degree
  plus_one = xl.DataArray(T.data + 1, T.dims, coords=T.coords, name='t2m')
  res = _fused(plus_one, T, P.dims, kinds=('Error', 'SquaredError'))
  # pred = target + 1: (p - t) is 1 up to f32 rounding of t + 1 at |t| ~ 280
  np.testing.assert_allclose(res['SquaredError'][0].values /
                             res['SquaredError'][1].values, 1.0, rtol=1e-4)
  # SE(2p, 2t) == 4 SE(p, t) exactly in binary floating point
  a = _fused(P, T, P.dims)['SquaredError'][0].values
  P2 = xl.DataArray(P.data * 2, P.dims, coords=P.coords, name='t2m')
  T2 = xl.DataArray(T.data * 2, T.dims, coords=T.coords, name='t2m')
  b = _fused(P2, T2, P.dims)['SquaredError'][0].values
  assert b == 4 * a


def test_nan_propagates_at_full_size(quarter_degree):
  P, T = quarter_degree
  t2 = T.data.clone()
  t2[7, 300, 700] = float('nan')
  T2 = xl.DataArray(t2, T.dims, coords=T.coords, name='t2m')
  res = _fused(P, T2, ('latitude', 'longitude'))['SquaredError'][0].values
  assert np.isnan(res[7]) and np.isfinite(np.delete(res, 7)).all()
  res = _fused(P, T2, ('latitude', 'longitude'), skipna=True)['SquaredError']
  assert np.isfinite(res[0].values).all()


def test_elementwise_is_bit_exact():
  P, T, C = _random_case(4, shape=(2, 3, 2, 16, 20))
  aligned, adims = oracle.align_climatology(
      C.values, C.dims, {k: C.coords[k].values for k in ('dayofyear', 'hour')},
      P.coords['init_time'].values, P.coords['lead_time'].values)
  for name, fn in oracle.DETERMINISTIC_STATISTICS.items():
    got = engine.LazyStatistic(name, P, T).values
    np.testing.assert_array_equal(got, fn(P.values, T.values))
  ac = engine.align_climatology(P, C)
  for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
    got = engine.LazyStatistic(name, P, T, ac).values
    np.testing.assert_array_equal(got, fn(P.values, T.values, aligned))


def test_errors_are_exceptions_not_aborts():
  ctx = _cabi.get_context()
  with pytest.raises(_cabi.WbxError):
    _cabi.DetPlan(ctx, space=0, flags=0, ny=4, nx=4,
                  pred=np.array([0], np.uint64), target=np.array([16], np.uint64),
                  cell=np.array([0], np.int32), n_cells=1)
  with pytest.raises(_cabi.WbxError):
    _cabi.DetPlan(ctx, space=0, flags=0, ny=4, nx=4,
                  pred=np.array([16, 16], np.uint64),
                  target=np.array([16, 16], np.uint64),
                  cell=np.array([1, 0], np.int32), n_cells=2)


# ---------------------------------------------------------------------------
# binned aggregation through the fused class-map kernel
# ---------------------------------------------------------------------------

REGIONS = {
    'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
    'nh-extratropics': ((20, 90), (0, 360)),
    'sh-extratropics': ((-90, -20), (0, 360)),
    'europe': ((35, 75), (-12.5, 42.5)), 'n-america': ((25, 60), (240, 285)),
    'east-asia': ((25, 60), (102.5, 150)), 'ausnz': ((-45, -12.5), (120, 175)),
}


def _bin_case(seed, shape=(3, 2, 24, 48), nan=False):
  rng = np.random.default_rng(seed)
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {
      'init_time': (np.datetime64('2020-03-01T00', 'ns') +
                    np.arange(shape[0]) * np.timedelta64(12, 'h')),
      'lead_time': (np.arange(shape[1]) * np.timedelta64(6, 'h')
                    ).astype('timedelta64[ns]'),
      'latitude': np.linspace(-90, 90, shape[2]),
      'longitude': np.linspace(0, 360, shape[3], endpoint=False)}
  p = rng.normal(280, 10, shape).astype(np.float32)
  t = (p + rng.normal(0, 2, shape)).astype(np.float32)
  if nan:
    t[1, 0, 5, 7] = np.nan
  c = rng.normal(280, 5, (366, 4) + shape[2:]).astype(np.float32)
  P = xl.DataArray(p, dims, coords=coords, name='z')
  T = xl.DataArray(t, dims, coords=coords, name='z')
  C = xl.DataArray(c, ('dayofyear', 'hour', 'latitude', 'longitude'),
                   coords={'dayofyear': np.arange(1, 367),
                           'hour': np.arange(0, 24, 6),
                           'latitude': coords['latitude'],
                           'longitude': coords['longitude']}, name='z')
  land = xl.DataArray(rng.random(shape[2:]) > 0.65, ('latitude', 'longitude'),
                      coords={'latitude': coords['latitude'],
                              'longitude': coords['longitude']})
  return P, T, C, land


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('masked', [False, True])
def test_fused_bins_match_oracle(space, masked, monkeypatch):
  P, T, C, land = _bin_case(3)
  mask_np = np.random.default_rng(9).random(P.shape) > 0.2
  if masked:
    T = T.assign_coords(mask=xl.DataArray(mask_np, P.dims))
  bin_by = [binning.Regions(REGIONS, land_sea_mask=land),
            binning.LandSea(xl.DataArray(land.values.astype(float), land.dims,
                                         coords=land.coords),
                            bin_dim_name='ls')]
  if space == 'device':
    P, T, C = engine.to_device(P), engine.to_device(T), engine.to_device(C)
    if masked:
      T = T.assign_coords(mask=engine.to_device(xl.DataArray(mask_np, P.dims)))
  # the fused class-map kernel must serve this request, not the generic one
  from weatherbenchx_b200 import generic
  monkeypatch.setattr(generic, 'aggregate', lambda *a, **k: (_ for _ in ()).throw(
      AssertionError('generic path used')))
  rd = ['init_time', 'latitude', 'longitude']
  state = _aggregate(ALL_METRICS(C), {'z': P}, {'z': T}, reduce_dims=rd,
                     weigh_by=[weighting.GridAreaWeighting()], bin_by=bin_by,
                     masked=masked)
  Ph, Th = P.to_host(), T.to_host()
  m1 = bin_by[0].create_bin_mask(Ph).values
  m2 = bin_by[1].create_bin_mask(Ph).values
  w = oracle.grid_area_weights(Ph.coords['latitude'].values)
  expected = _oracle_states_bins(Ph, Th, C.to_host(), rd, w, m1, m2, mask_np,
                                 masked)
  for name, (sws, sw, dims) in expected.items():
    got_ws = state.sum_weighted_statistics[name]['z']
    got_w = state.sum_weights[name]['z']
    assert set(got_ws.dims) == set(dims) == {'lead_time', 'region', 'ls'}
    np.testing.assert_allclose(got_ws.transpose(*dims).values, sws, rtol=RTOL,
                               atol=1e-6 * np.abs(sws).max())
    np.testing.assert_allclose(got_w.transpose(*dims).values, sw, rtol=1e-10)
  assert (state.sum_weights['SquaredError']['z'].coords['region'].values.tolist()
          == list(REGIONS) + [f'{r}_land' for r in REGIONS])


def _oracle_states_bins(P, T, C, rd, w, m1, m2, mask_np, masked):
  aligned, adims = oracle.align_climatology(
      C.values, C.dims, {k: C.coords[k].values for k in ('dayofyear', 'hour')},
      P.coords['init_time'].values, P.coords['lead_time'].values)
  aligned = np.transpose(aligned, [adims.index(d) for d in P.dims])
  vals = {n: f(P.values, T.values)
          for n, f in oracle.DETERMINISTIC_STATISTICS.items()}
  vals.update({n: f(P.values, T.values, aligned)
               for n, f in oracle.CLIMATOLOGY_STATISTICS.items()})
  return {
      n: oracle.aggregate(
          v, P.dims, rd, weights=[(w, ('latitude',))],
          bin_masks=[(m1, ('region', 'latitude', 'longitude')),
                     (m2, ('ls', 'latitude', 'longitude'))],
          mask=None if n == 'SquaredPredictionAnomaly' else mask_np,
          mask_dims=P.dims, masked=masked)
      for n, v in vals.items()}


def test_fused_bins_nan_poisons_every_bin_like_the_reference():
  """aggregation.py:272-277: a NaN outside a bin still makes that bin NaN."""
  P, T, C, land = _bin_case(4, nan=True)
  state = _aggregate({'mse': deterministic.MSE()}, {'z': P}, {'z': T},
                     reduce_dims=['init_time', 'latitude', 'longitude'],
                     bin_by=[binning.Regions(REGIONS)])
  got = state.sum_weighted_statistics['SquaredError']['z']
  lead = got.dims.index('lead_time')
  vals = np.moveaxis(got.values, lead, 0)
  assert np.isnan(vals[0]).all() and np.isfinite(vals[1]).all()


def test_fused_bins_full_size_public_benchmark_regions(quarter_degree):
  """0.25 degree, 2 x 17-ish regions with a land mask: fused == torch f64."""
  import torch
  P, T = quarter_degree
  rng = np.random.default_rng(5)
  lat, lon = P.coords['latitude'].values, P.coords['longitude'].values
  # blocky synthetic continents (coherent along longitude like real coasts)
  land_np = np.kron(rng.random((103, 96)) > 0.7, np.ones((7, 15), bool))
  land = xl.DataArray(land_np, ('latitude', 'longitude'),
                      coords={'latitude': lat, 'longitude': lon})
  regions = binning.Regions(REGIONS, land_sea_mask=land)
  agg = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'], bin_by=[regions],
      weigh_by=[weighting.GridAreaWeighting()])
  stats = {'SquaredError': {'t2m': engine.LazyStatistic('SquaredError', P, T)}}
  state = agg.aggregate_statistics(stats)
  got = state.sum_weighted_statistics['SquaredError']['t2m']
  got_w = state.sum_weights['SquaredError']['t2m']
  masks = torch.as_tensor(regions.create_bin_mask(P.isel(init_time=0)).values,
                          device='cuda')
  w = torch.as_tensor(oracle.grid_area_weights(lat), device='cuda')
  d = (P.data - T.data)
  se = (d * d).double().sum(0) * w[:, None]
  ref = (masks.double() * se[None]).sum((1, 2)).cpu().numpy()
  ref_w = (masks.double() * w[None, :, None]).sum((1, 2)).cpu().numpy() * 20
  np.testing.assert_allclose(got.values, ref, rtol=RTOL)
  np.testing.assert_allclose(got_w.values, ref_w, rtol=1e-10)
  again = agg.aggregate_statistics(stats)
  assert (again.sum_weighted_statistics['SquaredError']['t2m'].values.tobytes()
          == got.values.tobytes())


# ---------------------------------------------------------------------------
# longitude-major storage (latitude fastest): the latitude weight is w_x
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('nlat', [24, 19, 721])
@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna'])
def test_lon_major_layout_matches_oracle(space, nlat, mode):
  """[init, longitude, latitude] arrays as in the 1440x721 archives: rows of
  the slab are meridians, so the area weight varies along the row.  nlat = 24
  takes the vectorised w_x path, 19 / 721 (odd: groups straddle rows) the
  per-element one."""
  rng = np.random.default_rng(nlat)
  n_init, n_lead, nlon = 3, 2, 16 if nlat < 100 else 8
  dims = ('init_time', 'lead_time', 'longitude', 'latitude')
  coords = {'init_time': np.arange(n_init),
            'lead_time': (np.arange(n_lead) * np.timedelta64(6, 'h')
                          ).astype('timedelta64[ns]'),
            'longitude': np.linspace(0, 360, nlon, endpoint=False),
            'latitude': np.linspace(-90, 90, nlat)}
  shape = (n_init, n_lead, nlon, nlat)
  p = rng.normal(280, 5, shape).astype(np.float32)
  t = (p + rng.normal(0, 2, shape)).astype(np.float32)
  c = rng.normal(280, 3, (366, 4, nlon, nlat)).astype(np.float32)
  if mode != 'propagate':
    t[rng.random(shape) < 0.05] = np.nan
  coords['init_time'] = (np.datetime64('2020-02-27T00', 'ns') +
                         np.arange(n_init) * np.timedelta64(12, 'h'))
  P = xl.DataArray(p, dims, coords=coords, name='z')
  T = xl.DataArray(t, dims, coords=coords, name='z')
  C = xl.DataArray(c, ('dayofyear', 'hour', 'longitude', 'latitude'),
                   coords={'dayofyear': np.arange(1, 367),
                           'hour': np.arange(0, 24, 6),
                           'longitude': coords['longitude'],
                           'latitude': coords['latitude']}, name='z')
  mask_np = ~np.isnan(t)
  if mode == 'masked':
    T = T.assign_coords(mask=xl.DataArray(mask_np, dims))
  if space == 'device':
    P, C = engine.to_device(P), engine.to_device(C)
    Td = engine.to_device(T)
    if mode == 'masked':
      Td = Td.assign_coords(mask=engine.to_device(xl.DataArray(mask_np, dims)))
    T = Td
  rd = ['init_time', 'lead_time', 'longitude', 'latitude']
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
             'acc': deterministic.ACC({'z': C})}
  state = _aggregate(metrics, {'z': P}, {'z': T}, reduce_dims=rd,
                     weigh_by=[weighting.GridAreaWeighting()],
                     masked=mode == 'masked', skipna=mode == 'skipna')
  w = oracle.grid_area_weights(coords['latitude'])
  aligned, adims = oracle.align_climatology(
      c, ('dayofyear', 'hour', 'longitude', 'latitude'),
      {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)},
      coords['init_time'], coords['lead_time'])
  aligned = np.transpose(aligned, [adims.index(d) for d in dims])
  fields = {'SquaredError': oracle.squared_error(p, t),
            'AbsoluteError': oracle.absolute_error(p, t)}
  fields.update({n: f(p, t, aligned)
                 for n, f in oracle.CLIMATOLOGY_STATISTICS.items()})
  for name, field in fields.items():
    sws, sw, _ = oracle.aggregate(
        field, dims, rd, weights=[(w, ('latitude',))],
        mask=None if name == 'SquaredPredictionAnomaly' else mask_np,
        mask_dims=dims, masked=mode == 'masked', skipna=mode == 'skipna')
    got_ws = state.sum_weighted_statistics[name]['z'].values
    got_w = state.sum_weights[name]['z'].values
    np.testing.assert_allclose(got_ws, sws, rtol=RTOL, equal_nan=True)
    np.testing.assert_allclose(got_w, sw, rtol=1e-10)


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('masked', [False, True])
@pytest.mark.parametrize('nlat', [24, 19])
def test_fused_bins_lon_major(space, masked, nlat, monkeypatch):
  """Region / land-sea bins on [.., longitude, latitude] arrays: the binned
  kernel weighs every element (w_x; odd row lengths too) instead of falling
  back to the generic reduction."""
  P, T, C, land = _bin_case(11, shape=(3, 2, nlat, 48))
  order = ('init_time', 'lead_time', 'longitude', 'latitude')
  mask_np = np.random.default_rng(5).random(P.shape) > 0.2
  mask = xl.DataArray(mask_np, P.dims)
  P, T = P.transpose(*order), T.transpose(*order)
  P = xl.DataArray(np.ascontiguousarray(P.values), order, coords=P.coords,
                   name='z')
  T = xl.DataArray(np.ascontiguousarray(T.values), order, coords=T.coords,
                   name='z')
  if masked:
    T = T.assign_coords(mask=xl.DataArray(
        np.ascontiguousarray(mask.transpose(*order).values), order))
  if space == 'device':
    P, T = engine.to_device(P), engine.to_device(T)
    if masked:
      T = T.assign_coords(mask=engine.to_device(xl.DataArray(
          np.ascontiguousarray(mask.transpose(*order).values), order)))
  bin_by = [binning.Regions(REGIONS, land_sea_mask=land)]
  from weatherbenchx_b200 import generic
  monkeypatch.setattr(generic, 'aggregate', lambda *a, **k: (_ for _ in ()).throw(
      AssertionError('generic path used')))
  rd = ['init_time', 'latitude', 'longitude']
  metrics = {'rmse': deterministic.RMSE(), 'bias': deterministic.Bias()}
  state = _aggregate(metrics, {'z': P}, {'z': T}, reduce_dims=rd,
                     weigh_by=[weighting.GridAreaWeighting()], bin_by=bin_by,
                     masked=masked)
  Ph, Th = P.to_host(), T.to_host()
  m1 = bin_by[0].create_bin_mask(Ph).values    # (region, lat, lon) order below
  m1 = np.asarray(bin_by[0].create_bin_mask(
      Ph.transpose('init_time', 'lead_time', 'latitude', 'longitude')).values)
  w = oracle.grid_area_weights(Ph.coords['latitude'].values)
  pv = Ph.transpose('init_time', 'lead_time', 'latitude', 'longitude').values
  tv = Th.transpose('init_time', 'lead_time', 'latitude', 'longitude').values
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  for name, field in (('SquaredError', oracle.squared_error(pv, tv)),
                      ('Error', oracle.error(pv, tv))):
    sws, sw, odims = oracle.aggregate(
        field, dims, rd, weights=[(w, ('latitude',))],
        bin_masks=[(m1, ('region', 'latitude', 'longitude'))],
        mask=mask_np, mask_dims=dims, masked=masked)
    got_ws = state.sum_weighted_statistics[name]['z'].transpose(*odims).values
    got_w = state.sum_weights[name]['z'].transpose(*odims).values
    np.testing.assert_allclose(got_ws, sws, rtol=RTOL,
                               atol=1e-6 * np.abs(sws).max())
    # element weights are float32 inside the binned kernel: the masked sum of
    # weights carries their rounding (~3e-8), the constant one is exact
    np.testing.assert_allclose(got_w, sw, rtol=1e-7 if masked else 1e-10)


def test_latitude_and_longitude_band_bins(monkeypatch):
  """LatitudeBins x LongitudeBins (binning.py:204-298): 1-d band masks are
  broadcast over the slab and served by the fused class-map kernel."""
  P, T, C, _ = _bin_case(21)
  del C
  P, T = engine.to_device(P), engine.to_device(T)
  bin_by = [binning.LatitudeBins(degrees=30),
            binning.LongitudeBins(degrees=90, lon_range=(270, 90))]
  from weatherbenchx_b200 import generic
  monkeypatch.setattr(generic, 'aggregate', lambda *a, **k: (_ for _ in ()).throw(
      AssertionError('generic path used')))
  rd = ['init_time', 'latitude', 'longitude']
  metrics = {'rmse': deterministic.RMSE()}
  state = _aggregate(metrics, {'z': P}, {'z': T}, reduce_dims=rd,
                     weigh_by=[weighting.GridAreaWeighting()], bin_by=bin_by)
  Ph, Th = P.to_host(), T.to_host()
  lat, lon = Ph.coords['latitude'].values, Ph.coords['longitude'].values
  m_lat = np.stack([(lat >= s) & (lat <= s + 30) for s in range(-90, 90, 30)])
  lon_bands = [(270, 360), (360, 450)]
  m_lon = np.stack([
      ((np.mod(lon, 360) >= np.mod(a, 360)) | (np.mod(lon, 360) <= np.mod(b, 360)))
      if np.mod(b, 360) <= np.mod(a, 360) else
      ((np.mod(lon, 360) >= np.mod(a, 360)) & (np.mod(lon, 360) <= np.mod(b, 360)))
      for a, b in lon_bands])
  w = oracle.grid_area_weights(lat)
  sws, sw, odims = oracle.aggregate(
      oracle.squared_error(Ph.values, Th.values), Ph.dims, rd,
      weights=[(w, ('latitude',))],
      bin_masks=[(m_lat, ('latitude_bins', 'latitude')),
                 (m_lon, ('longitude_bins', 'longitude'))])
  got = state.sum_weighted_statistics['SquaredError']['z']
  assert set(got.dims) == set(odims) == {'lead_time', 'latitude_bins',
                                         'longitude_bins'}
  np.testing.assert_allclose(got.transpose(*odims).values, sws, rtol=RTOL,
                             atol=1e-6 * np.abs(sws).max())
  np.testing.assert_allclose(
      state.sum_weights['SquaredError']['z'].transpose(*odims).values, sw,
      rtol=1e-10)
  np.testing.assert_array_equal(got.coords['latitude_bins'].values,
                                np.arange(-90, 90, 30))
  np.testing.assert_array_equal(got.coords['longitude_bins'].values, [270, 0])


# ---------------------------------------------------------------------------
# BASELINE.json configs[0] and configs[1] at their exact shapes
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('space', ['host', 'device'])
def test_config_c1_rmse_32x64_10_init(space):
  """configs[0]: RMSE of 2m_temperature on the 5.625 degree grid (32 x 64),
  10 init times, against the oracle; host and device inputs."""
  from weatherbenchx_b200 import aggregation, engine, weighting
  from weatherbenchx_b200.metrics import deterministic
  rng = np.random.default_rng(100)
  ny, nx, n_init = 32, 64, 10
  lat = np.linspace(-90 + 5.625 / 2, 90 - 5.625 / 2, ny)
  lon = np.linspace(0, 360, nx, endpoint=False)
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': np.datetime64('2020-01-01', 'ns') +
                         np.arange(n_init) * np.timedelta64(1, 'D'),
            'lead_time': np.array([0], 'timedelta64[ns]'),
            'latitude': lat, 'longitude': lon}
  t = rng.normal(280, 12, (n_init, 1, ny, nx)).astype(np.float32)
  p = (t + rng.normal(0, 2, t.shape)).astype(np.float32)
  P = xl.DataArray(p, dims, coords=coords, name='2m_temperature')
  T = xl.DataArray(t, dims, coords=coords, name='2m_temperature')
  if space == 'device':
    P, T = engine.to_device(P), engine.to_device(T)
  w = oracle.grid_area_weights(lat)
  for rd in (['init_time', 'latitude', 'longitude'],
             ['init_time', 'lead_time', 'latitude', 'longitude'],
             ['latitude', 'longitude']):
    agg = aggregation.Aggregator(reduce_dims=rd,
                                 weigh_by=[weighting.GridAreaWeighting()])
    got = aggregation.compute_metric_values_for_single_chunk(
        {'rmse': deterministic.RMSE()}, agg, {'2m_temperature': P},
        {'2m_temperature': T})['rmse.2m_temperature']
    sws, sw, out_dims = oracle.aggregate(oracle.squared_error(p, t), dims, rd,
                                         weights=[(w, ('latitude',))])
    np.testing.assert_allclose(
        got.transpose(*out_dims).values if out_dims else got.values,
        np.sqrt(sws / sw), rtol=1e-6)


def test_config_c2_rmse_acc_128x256_13_levels():
  """configs[1]: RMSE + ACC of 6 pressure-level variables on the 1.4 degree
  grid (128 x 256), 13 levels, 10 lead times, a 4-init slice of the 40 init
  times, climatology [366, 4, 13, 128, 256], GridAreaWeighting, reducing
  init_time / latitude / longitude -- every value against the oracle."""
  from weatherbenchx_b200 import aggregation, engine, weighting
  from weatherbenchx_b200.metrics import deterministic
  rng = np.random.default_rng(200)
  n_var, n_init, n_lead, n_lev, ny, nx = 6, 4, 10, 13, 128, 256
  lat = np.linspace(-90, 90, ny)
  lon = np.linspace(0, 360, nx, endpoint=False)
  init = (np.datetime64('2020-02-27T00', 'ns') +
          np.arange(n_init) * np.timedelta64(12, 'h'))
  lead = (np.arange(n_lead) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
  dims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  coords = {'init_time': init, 'lead_time': lead,
            'level': np.array([50, 100, 150, 200, 250, 300, 400, 500, 600,
                               700, 850, 925, 1000]),
            'latitude': lat, 'longitude': lon}
  cdims = ('dayofyear', 'hour', 'level', 'latitude', 'longitude')
  ccoords = {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
             'level': coords['level'], 'latitude': lat, 'longitude': lon}
  names = ['geopotential', 'temperature', 'u_component_of_wind',
           'v_component_of_wind', 'specific_humidity', 'vertical_velocity']
  host, P, T, C = {}, {}, {}, {}
  # the climatology rows the slice touches (days 58..62) carry data; one shared
  # array keeps the host copy at 1.4 GB
  c = np.zeros((366, 4, n_lev, ny, nx), np.float32)
  c[55:66] = rng.normal(0, 0.5, (11, 4, n_lev, ny, nx)).astype(np.float32)
  Cdev = engine.to_device(xl.DataArray(c, cdims, coords=ccoords))
  for k, name in enumerate(names):
    t = rng.normal(k, 1, (n_init, n_lead, n_lev, ny, nx)).astype(np.float32)
    p = (t + rng.normal(0, 0.3, t.shape)).astype(np.float32)
    host[name] = (p, t)
    P[name] = engine.to_device(xl.DataArray(p, dims, coords=coords, name=name))
    T[name] = engine.to_device(xl.DataArray(t, dims, coords=coords, name=name))
    C[name] = Cdev.rename(name)
  rd = ['init_time', 'latitude', 'longitude']
  agg = aggregation.Aggregator(reduce_dims=rd,
                               weigh_by=[weighting.GridAreaWeighting()])
  got = aggregation.compute_metric_values_for_single_chunk(
      {'rmse': deterministic.RMSE(), 'acc': deterministic.ACC(C)}, agg, P, T)
  w = oracle.grid_area_weights(lat)
  aligned, adims = oracle.align_climatology(
      c, cdims, {'dayofyear': ccoords['dayofyear'], 'hour': ccoords['hour']},
      init, lead)
  assert adims == dims
  for name, (p, t) in host.items():
    sws, sw, out_dims = oracle.aggregate(oracle.squared_error(p, t), dims, rd,
                                         weights=[(w, ('latitude',))])
    assert out_dims == ('lead_time', 'level')
    np.testing.assert_allclose(
        got[f'rmse.{name}'].transpose(*out_dims).values, np.sqrt(sws / sw),
        rtol=1e-6, err_msg=name)
    means = {}
    for stat, fn in oracle.CLIMATOLOGY_STATISTICS.items():
      a, b, _ = oracle.aggregate(fn(p, t, aligned), adims, rd,
                                 weights=[(w, ('latitude',))])
      means[stat] = a / b
    acc = oracle.acc_from_means(means['AnomalyCovariance'],
                                means['SquaredPredictionAnomaly'],
                                means['SquaredTargetAnomaly'])
    np.testing.assert_allclose(
        got[f'acc.{name}'].transpose(*out_dims).values, acc, rtol=1e-5,
        err_msg=name)
