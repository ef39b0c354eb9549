"""GPU parity tests of the SURVEY.md section 8(f) rank-2 statistics:
WindVectorSquaredError / WindVectorRMSE, Prediction/TargetPassthrough and the
EnsembleMean input transform with the wrappers that apply it."""

import numpy as np
import pytest

import wbx_oracle as oracle
import wbx_test_utils as utils
from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import deterministic
from weatherbenchx_b200.metrics import wrappers

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DIMS = ('init_time', 'latitude', 'longitude')


def _aggregate(metrics, predictions, targets, **kwargs):
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  return aggregation.Aggregator(**kwargs).aggregate_statistics(statistics)


def _coords(n_init, nlat, nlon):
  return {'init_time': np.arange(n_init),
          'latitude': np.linspace(-85, 85, nlat),
          'longitude': np.linspace(0, 360, nlon, endpoint=False)}


def _fields(names, shape, seed, nan=False, device=False):
  rng = np.random.default_rng(seed)
  coords = _coords(*shape)
  out = {}
  for name in names:
    a = rng.normal(size=shape).astype(np.float32)
    if nan:
      a[rng.random(shape) < 0.03] = np.nan
    da = xl.DataArray(a, DIMS, coords=coords, name=name)
    out[name] = engine.to_device(da) if device else da
  return out


# ---------------------------------------------------------------------------
# wind vector
# ---------------------------------------------------------------------------


def test_wind_vector_rmse_known_answer():
  """metrics_test.py:501-543: targets = predictions + 1 -> sqrt(2)."""
  names = dict(variables_2d=['10m_u_component_of_wind',
                             '10m_v_component_of_wind'],
               variables_3d=['u_component_of_wind', 'v_component_of_wind'])
  prediction = utils.mock_prediction_data(
      time_start='2020-01-01T00', time_stop='2020-01-03T00', **names)
  target = {k: v + 1 for k, v in prediction.items()}
  metrics = {'vector_rmse': deterministic.WindVectorRMSE(
      ['u_component_of_wind', '10m_u_component_of_wind'],
      ['v_component_of_wind', '10m_v_component_of_wind'],
      ['wind', '10m_wind'])}
  state = _aggregate(metrics, prediction, target,
                     reduce_dims=['time', 'latitude', 'longitude'])
  results = state.metric_values(metrics)
  assert set(results) == {'vector_rmse.wind', 'vector_rmse.10m_wind'}
  assert set(results['vector_rmse.wind'].dims) == {'prediction_timedelta',
                                                  'level'}
  for v in results.values():
    np.testing.assert_allclose(v.values, np.sqrt(2), rtol=RTOL)


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('mode', ['propagate', 'skipna'])
@pytest.mark.parametrize('reduce_dims', [['latitude', 'longitude'],
                                         ['init_time', 'latitude', 'longitude']])
def test_wind_vector_matches_oracle(space, mode, reduce_dims):
  shape = (3, 12, 20)
  nan = mode == 'skipna'
  P = _fields(['u', 'v'], shape, 0, nan=nan, device=space == 'device')
  T = _fields(['u', 'v'], shape, 1, nan=nan, device=space == 'device')
  metrics = {'wind': deterministic.WindVectorRMSE('u', 'v', 'wind'),
             'rmse': deterministic.RMSE()}
  state = _aggregate(metrics, P, T, reduce_dims=reduce_dims,
                     weigh_by=[weighting.GridAreaWeighting()],
                     skipna=mode == 'skipna')
  values = state.metric_values(metrics)
  p = {k: v.to_numpy() for k, v in P.items()}
  t = {k: v.to_numpy() for k, v in T.items()}
  se = oracle.wind_vector_squared_error(p['u'], t['u'], p['v'], t['v'])
  w = oracle.grid_area_weights(_coords(*shape)['latitude'])
  ws, sw, _ = oracle.aggregate(se, DIMS, reduce_dims,
                               weights=[(w, ('latitude',))],
                               skipna=mode == 'skipna')
  np.testing.assert_allclose(values['wind.wind'].values, np.sqrt(ws / sw),
                             rtol=RTOL)
  for c in 'uv':
    ws, sw, _ = oracle.aggregate(oracle.squared_error(p[c], t[c]), DIMS,
                                 reduce_dims, weights=[(w, ('latitude',))],
                                 skipna=mode == 'skipna')
    np.testing.assert_allclose(values[f'rmse.{c}'].values, np.sqrt(ws / sw),
                               rtol=RTOL)


def test_wind_vector_shares_the_component_launch():
  """WindVectorRMSE next to the per-component RMSE costs no extra launch: the
  parts join the SquaredError pass of their operands, and u and v (same grid)
  are merged into one launch."""
  P = _fields(['u', 'v'], (2, 16, 32), 0, device=True)
  T = _fields(['u', 'v'], (2, 16, 32), 1, device=True)
  ctx = _cabi.get_context()
  kw = dict(reduce_dims=['latitude', 'longitude'])
  both = {'wind': deterministic.WindVectorRMSE('u', 'v', 'wind'),
          'rmse': deterministic.RMSE()}
  _aggregate(both, P, T, **kw)  # warm the plan cache
  n0 = ctx.kernel_launches()
  _aggregate({'rmse': deterministic.RMSE()}, P, T, **kw)
  n1 = ctx.kernel_launches()
  _aggregate(both, P, T, **kw)
  n2 = ctx.kernel_launches()
  assert n2 - n1 == n1 - n0


def test_wind_vector_field_is_numpy_exact():
  P = _fields(['u', 'v'], (2, 5, 8), 2, device=True)
  T = _fields(['u', 'v'], (2, 5, 8), 3)
  stat = deterministic.WindVectorSquaredError(['u'], ['v'], ['wind'])
  assert stat.unique_name == 'WindVectorSquaredError_wind'
  field = stat.compute(P, T)['wind']
  expected = oracle.wind_vector_squared_error(
      P['u'].to_numpy(), T['u'].to_numpy(), P['v'].to_numpy(),
      T['v'].to_numpy())
  assert field.values.tobytes() == expected.tobytes()
  with pytest.raises(ValueError, match='same length'):
    deterministic.WindVectorSquaredError(['u'], ['v'], ['a', 'b'])


# ---------------------------------------------------------------------------
# passthrough
# ---------------------------------------------------------------------------


def test_prediction_passthrough_known_answer():
  """metrics_test.py:1008-1030."""
  predictions = xl.DataArray(
      np.array([[1.0, 2.0], [np.nan, 4.0]], np.float32), ['x', 'y'])
  targets = xl.DataArray(
      np.array([[5.0, np.nan], [7.0, 8.0]], np.float32), ['x', 'y'])
  result = deterministic.PredictionPassthrough(
      copy_nans_from_targets=False)._compute_per_variable(predictions, targets)
  np.testing.assert_array_equal(result.values, [[1.0, 2.0], [np.nan, 4.0]])
  result = deterministic.PredictionPassthrough(
      copy_nans_from_targets=True)._compute_per_variable(predictions, targets)
  np.testing.assert_array_equal(result.values,
                                [[1.0, np.nan], [np.nan, 4.0]])
  result = deterministic.TargetPassthrough(
      copy_nans_from_predictions=True)._compute_per_variable(predictions,
                                                             targets)
  np.testing.assert_array_equal(result.values,
                                [[5.0, np.nan], [np.nan, 8.0]])


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('copy_nans', [False, True])
def test_average_metrics_match_oracle(space, copy_nans):
  shape = (3, 12, 20)
  P = _fields(['t'], shape, 0, nan=True, device=space == 'device')
  T = _fields(['t'], shape, 1, nan=True, device=space == 'device')
  metrics = {
      'pavg': deterministic.PredictionAverage(copy_nans_from_targets=copy_nans),
      'tavg': deterministic.TargetAverage(copy_nans_from_predictions=copy_nans),
      'bias': deterministic.Bias(),
  }
  rd = ['latitude', 'longitude']
  state = _aggregate(metrics, P, T, reduce_dims=rd, skipna=True,
                     weigh_by=[weighting.GridAreaWeighting()])
  values = state.metric_values(metrics)
  p, t = P['t'].to_numpy(), T['t'].to_numpy()
  w = oracle.grid_area_weights(_coords(*shape)['latitude'])
  for name, field in (('pavg', oracle.passthrough(p, t, copy_nans)),
                      ('tavg', oracle.passthrough(t, p, copy_nans)),
                      ('bias', oracle.error(p, t))):
    ws, sw, _ = oracle.aggregate(field, DIMS, rd, weights=[(w, ('latitude',))],
                                 skipna=True)
    np.testing.assert_allclose(values[f'{name}.t'].values, ws / sw, rtol=RTOL,
                               atol=1e-7)


def test_passthrough_keeps_coordinates_of_both_inputs():
  P = _fields(['t'], (2, 6, 8), 0)
  T = _fields(['t'], (2, 6, 8), 1)
  mask = xl.DataArray(T['t'].to_numpy() > 0, DIMS)
  T = {'t': T['t'].assign_coords(mask=mask)}
  stat = deterministic.PredictionPassthrough().compute(P, T)['t']
  assert 'mask' in stat.coords and stat.is_lazy
  state = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'], masked=True
  ).aggregate_statistics({'PredictionPassthrough': {'t': stat}})
  m = mask.to_numpy()
  expected = (P['t'].to_numpy() * m).sum((1, 2)) / m.sum((1, 2))
  np.testing.assert_allclose(
      state.mean_statistics()['PredictionPassthrough']['t'].values, expected,
      rtol=RTOL)


# ---------------------------------------------------------------------------
# EnsembleMean transform + wrappers
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('skipna', [True, False])
def test_mean_over_realization_dim(skipna):
  """wrappers_test.py:125-148: one realization NaN on one level."""
  forecast = utils.to_f32(utils.mock_target_data(
      random=True, ensemble_size=3, time_stop='2020-01-05'))
  x = forecast['geopotential']
  a = x.to_numpy().copy()                    # (time, lat, lon, level, real.)
  a[..., 0, 0] = np.nan
  x = xl.DataArray(a, x.dims, coords=x.coords, name=x.name)
  em = wrappers.EnsembleMean(which='both', ensemble_dim='realization',
                             skipna=skipna)
  assert em.unique_name_suffix == (
      f"ensemble_mean_self._ensemble_dim='realization'_self._skipna={skipna}")
  y = em.transform_fn(x)
  assert y.dims == ('time', 'latitude', 'longitude', 'level')
  np.testing.assert_allclose(y.values, oracle.ensemble_mean(a, -1, skipna),
                             rtol=2e-7, equal_nan=True)
  np.testing.assert_array_equal(y.coords['level'].values, [500, 700, 850])


@pytest.mark.parametrize('layout', ['member_major', 'member_last', 'odd'])
@pytest.mark.parametrize('space', ['host', 'device'])
def test_ensemble_mean_is_numpy_exact(layout, space):
  rng = np.random.default_rng(3)
  m, shape = 11, (2, 9, 16)
  if layout == 'odd':
    shape = (2, 9, 15)          # scalar (non-float4) path
  if layout == 'member_last':
    x = rng.normal(280, 5, size=shape + (m,)).astype(np.float32)
    dims, axis = DIMS + ('number',), 3
  else:
    x = rng.normal(280, 5, size=(shape[0], m) + shape[1:]).astype(np.float32)
    dims, axis = ('init_time', 'number', 'latitude', 'longitude'), 1
  x[rng.random(x.shape) < 0.02] = np.nan
  X = xl.DataArray(x, dims, name='t')
  if space == 'device':
    X = engine.to_device(X)
  for skipna in (False, True):
    y = engine.ensemble_mean(X, 'number', skipna=skipna).values
    ref = oracle.ensemble_mean(x, axis, skipna)
    if layout == 'member_last':   # NumPy sums a trailing axis pairwise
      np.testing.assert_allclose(y, ref, rtol=3e-7, equal_nan=True)
    else:                         # member-order float32 sum: identical
      np.testing.assert_array_equal(y, ref)
  assert engine.ensemble_mean(X, 'number', skipna=True) is engine.ensemble_mean(
      X, 'number', skipna=True)


def test_wrapped_ensemble_mean_metrics_match_oracle_in_one_launch():
  rng = np.random.default_rng(5)
  m, shape = 10, (3, 12, 20)
  coords = dict(_coords(*shape), number=np.arange(m))
  x = rng.normal(size=(shape[0], m) + shape[1:]).astype(np.float32)
  y = rng.normal(size=shape).astype(np.float32)
  X = engine.to_device(xl.DataArray(
      x, ('init_time', 'number', 'latitude', 'longitude'), coords=coords,
      name='t'))
  Y = engine.to_device(xl.DataArray(
      y, DIMS, coords={d: coords[d] for d in DIMS}, name='t'))
  em = [wrappers.EnsembleMean(which='predictions', ensemble_dim='number',
                              skip_if_ensemble_dim_missing=True)]
  metrics = {
      'rmse': wrappers.WrappedMetric(deterministic.RMSE(), em),
      'mae': wrappers.WrappedMetric(deterministic.MAE(), em),
      'bias': wrappers.WrappedMetric(deterministic.Bias(), em,
                                     unique_name_suffix='ens_mean'),
  }
  names = {s.unique_name for mt in metrics.values()
           for s in mt.statistics.values()}
  suffix = "predictions_ensemble_mean_self._ensemble_dim='number'_self._skipna=False"
  assert names == {f'SquaredError_{suffix}', f'AbsoluteError_{suffix}',
                   'Error_ens_mean'}
  rd = ['init_time', 'latitude', 'longitude']
  kw = dict(reduce_dims=rd, weigh_by=[weighting.GridAreaWeighting()])
  ctx = _cabi.get_context()
  values = _aggregate(metrics, {'t': X}, {'t': Y}, **kw).metric_values(metrics)
  n0 = ctx.kernel_launches()
  _aggregate(metrics, {'t': X}, {'t': Y}, **kw)
  # cached ensemble mean: one fused reduce + its finalize, nothing else
  assert ctx.kernel_launches() - n0 <= 2
  mean = oracle.ensemble_mean(x, 1)
  w = oracle.grid_area_weights(coords['latitude'])
  for name, field, fn in (
      ('rmse', oracle.squared_error(mean, y), np.sqrt),
      ('mae', oracle.absolute_error(mean, y), lambda v: v),
      ('bias', oracle.error(mean, y), lambda v: v)):
    ws, sw, _ = oracle.aggregate(field, DIMS, rd, weights=[(w, ('latitude',))])
    np.testing.assert_allclose(values[f'{name}.t'].values, fn(ws / sw),
                               rtol=RTOL)


def test_subselect_and_select_wrappers():
  P = _fields(['a', 'b'], (3, 6, 8), 0, device=True)
  T = _fields(['a', 'b'], (3, 6, 8), 1, device=True)
  metrics = {
      'rmse': wrappers.SubselectVariables(deterministic.RMSE(), ['a']),
      'first': wrappers.WrappedMetric(
          deterministic.RMSE(),
          [wrappers.Select(which='both', isel={'init_time': slice(0, 1)})]),
  }
  stat_names = [s.unique_name for s in metrics['rmse'].statistics.values()]
  assert stat_names == ['SquaredError_a']
  values = _aggregate(
      metrics, P, T, reduce_dims=['init_time', 'latitude', 'longitude']
  ).metric_values(metrics)
  assert set(values) == {'rmse.a', 'first.a', 'first.b'}
  p, t = P['a'].to_numpy(), T['a'].to_numpy()
  np.testing.assert_allclose(values['rmse.a'].values,
                             np.sqrt(np.mean((p - t) ** 2)), rtol=RTOL)
  np.testing.assert_allclose(values['first.a'].values,
                             np.sqrt(np.mean((p[:1] - t[:1]) ** 2)), rtol=RTOL)
  with pytest.raises(ValueError, match='Invalid value for `which`'):
    wrappers.EnsembleMean(which='nobody')


# ---------------------------------------------------------------------------
# EnsembleAveragedMetric (probabilistic.py:35-113)
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('layout', ['member_major', 'reference_mock'])
def test_ensemble_averaged_metric(layout):
  """metrics_test.py:1276-1308: per-member RMSE through the wrapper == RMSE
  with the ensemble dim among the aggregator's reduce_dims."""
  from weatherbenchx_b200.metrics import probabilistic
  if layout == 'reference_mock':
    targets = utils.to_f32(utils.mock_prediction_data(
        time_start='2020-01-01T00', time_stop='2020-01-03T00', random=True,
        seed=0))
    predictions = utils.to_f32(utils.mock_prediction_data(
        time_start='2020-01-01T00', time_stop='2020-01-03T00', random=True,
        ensemble_size=5, seed=1))
    rd = ['latitude', 'longitude']
  else:
    rng = np.random.default_rng(0)
    coords = dict(_coords(3, 12, 20), realization=np.arange(5))
    predictions = {'t': xl.DataArray(
        rng.normal(size=(3, 5, 12, 20)).astype(np.float32),
        ('init_time', 'realization', 'latitude', 'longitude'), coords=coords,
        name='t')}
    targets = {'t': xl.DataArray(
        rng.normal(size=(3, 12, 20)).astype(np.float32), DIMS,
        coords={d: coords[d] for d in DIMS}, name='t')}
    rd = ['init_time', 'latitude', 'longitude']
  kw = dict(weigh_by=[weighting.GridAreaWeighting()])
  explicit = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE()}
  expected = _aggregate(explicit, predictions, targets,
                        reduce_dims=rd + ['realization'], **kw
                        ).metric_values(explicit)
  wrapped = {k: probabilistic.EnsembleAveragedMetric(
      m, ensemble_dim='realization') for k, m in explicit.items()}
  names = {s.unique_name for m in wrapped.values()
           for s in m.statistics.values()}
  assert names == {'SquaredError_each_realization',
                   'AbsoluteError_each_realization'}
  state = _aggregate(wrapped, predictions, targets, reduce_dims=rd, **kw)
  actual = state.metric_values(wrapped)
  assert set(actual) == set(expected)
  for k in expected:
    assert actual[k].dims == expected[k].dims
    np.testing.assert_allclose(actual[k].values, expected[k].values, rtol=RTOL)
  if layout == 'member_major':
    # the state is the mean-over-members one (weights of one member, not five)
    w = oracle.grid_area_weights(predictions['t'].coords['latitude'].values)
    np.testing.assert_allclose(
        state.sum_weights['SquaredError_each_realization']['t'].values,
        3 * 20 * w.sum(), rtol=1e-12)
    p, t = predictions['t'].values, targets['t'].values
    field = oracle.squared_error(p, t[:, None]).mean(1)
    ws, sw, _ = oracle.aggregate(field, DIMS, rd,
                                 weights=[(w, ('latitude',))])
    np.testing.assert_allclose(actual['rmse.t'].values, np.sqrt(ws / sw),
                               rtol=RTOL)
    # NaN skipping needs the per-point member mean: generic path, same answer
    p2 = p.copy()
    p2[0, 2, 3, 4] = np.nan
    pn = {'t': xl.DataArray(p2, predictions['t'].dims,
                            coords=predictions['t'].coords, name='t')}
    skip = {'rmse': probabilistic.EnsembleAveragedMetric(
        deterministic.RMSE(), ensemble_dim='realization',
        skipna_ensemble=True)}
    got = _aggregate(skip, pn, targets, reduce_dims=rd, **kw
                     ).metric_values(skip)
    field = np.nanmean(oracle.squared_error(p2, t[:, None]), axis=1)
    ws, sw, _ = oracle.aggregate(field, DIMS, rd,
                                 weights=[(w, ('latitude',))])
    np.testing.assert_allclose(got['rmse.t'].values, np.sqrt(ws / sw),
                               rtol=RTOL)
    with pytest.raises(ValueError, match='Failed to compute'):
      _aggregate(wrapped, targets, targets, reduce_dims=rd)
