"""Operands are matched by coordinate LABEL, never by position (CPU).

xarray aligns the operands of arithmetic and of ``xr.dot`` by label: the
reference's ``predictions - climatology.sel(...)`` (metrics/base.py:397-403,
deterministic.py:225-259) and ``xr.dot(stat, *weights, *bin_masks)``
(aggregation.py:334-335) give the same result whatever order the climatology
or the bin mask store their latitudes in.  The host side above the C ABI is
exercised with plans interpreted by tests/wbx_emulator.py; the oracle is the
arbiter.  Also here: the member mean drops a mask that carries the ensemble dim
(probabilistic.py:56-69), NetCDF name collisions, checkpoint fingerprints and
the DataTree form of the AggregationState (aggregation.py:203-265).
"""

import numpy as np
import pytest

import wbx_emulator
import wbx_oracle as oracle
from weatherbenchx_b200 import aggregation, binning, io_netcdf, pipeline
from weatherbenchx_b200 import time_chunks, weighting, xarray_tree
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.data_loaders import array_loaders
from weatherbenchx_b200.metrics import deterministic, probabilistic

NY, NX = 8, 12
LAT = np.linspace(-87.5, 87.5, NY)
LON = np.linspace(0, 360, NX, endpoint=False)
INIT = (np.datetime64('2020-02-27T00', 'ns') +
        np.arange(3) * np.timedelta64(12, 'h'))
LEAD = (np.arange(2) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
DIMS = ('init_time', 'lead_time', 'latitude', 'longitude')
COORDS = {'init_time': INIT, 'lead_time': LEAD, 'latitude': LAT,
          'longitude': LON}
RD = ['init_time', 'latitude', 'longitude']


def _fields(seed=0):
  rng = np.random.default_rng(seed)
  p = rng.normal(280, 5, (3, 2, NY, NX)).astype(np.float32)
  t = (p + rng.normal(0, 2, p.shape)).astype(np.float32)
  c = rng.normal(280, 3, (366, 4, NY, NX)).astype(np.float32)
  return p, t, c


def _acc_oracle(p, t, c):
  aligned, adims = oracle.align_climatology(
      c, ('dayofyear', 'hour', 'latitude', 'longitude'),
      {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)}, INIT, LEAD)
  w = oracle.grid_area_weights(LAT)
  means = {}
  for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
    a, b, _ = oracle.aggregate(fn(p, t, aligned), adims, RD,
                               weights=[(w, ('latitude',))])
    means[name] = a / b
  return oracle.acc_from_means(means['AnomalyCovariance'],
                               means['SquaredPredictionAnomaly'],
                               means['SquaredTargetAnomaly'])


def _clim(c, lat):
  return xl.DataArray(
      c, ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords={'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
              'latitude': lat, 'longitude': LON}, name='v')


def test_acc_with_reversed_latitude_climatology(monkeypatch):
  """a climatology with the same latitude labels in reversed
  order must give the reference's (label-aligned) ACC, not a positional one."""
  wbx_emulator.installed(monkeypatch)
  p, t, c = _fields()
  P = {'v': xl.DataArray(p, DIMS, coords=COORDS, name='v')}
  T = {'v': xl.DataArray(t, DIMS, coords=COORDS, name='v')}
  agg = aggregation.Aggregator(reduce_dims=RD,
                               weigh_by=[weighting.GridAreaWeighting()])
  want = _acc_oracle(p, t, c)
  straight = aggregation.compute_metric_values_for_single_chunk(
      {'acc': deterministic.ACC({'v': _clim(c, LAT)})}, agg, P, T)
  np.testing.assert_allclose(straight['acc.v'].values, want, rtol=1e-6)
  flipped = _clim(np.ascontiguousarray(c[:, :, ::-1]), LAT[::-1].copy())
  got = aggregation.compute_metric_values_for_single_chunk(
      {'acc': deterministic.ACC({'v': flipped})}, agg, P, T)
  np.testing.assert_allclose(got['acc.v'].values, want, rtol=1e-6)
  # the re-ordered copy is made once and keeps its identity (plans stay shared)
  from weatherbenchx_b200 import engine
  a1 = engine.align_climatology(P['v'], flipped)
  a2 = engine.align_climatology(
      xl.DataArray(p, DIMS, coords=COORDS, name='v'), flipped)
  assert a1.climatology is a2.climatology


def test_climatology_with_other_labels_raises(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  p, t, c = _fields()
  P = {'v': xl.DataArray(p, DIMS, coords=COORDS, name='v')}
  T = {'v': xl.DataArray(t, DIMS, coords=COORDS, name='v')}
  agg = aggregation.Aggregator(reduce_dims=RD)
  shifted = _clim(c, LAT + 1.0)
  with pytest.raises(ValueError, match='climatology|Failed to compute'):
    aggregation.compute_metric_values_for_single_chunk(
        {'acc': deterministic.ACC({'v': shifted})}, agg, P, T)


def test_land_sea_mask_with_descending_latitude(monkeypatch):
  """LandSea returns its mask on the coordinates of the
  land-sea field; xr.dot aligns it with the statistic by label."""
  wbx_emulator.installed(monkeypatch)
  p, t, _ = _fields(1)
  rng = np.random.default_rng(5)
  land = rng.random((NY, NX)) < 0.4
  P = {'v': xl.DataArray(p, DIMS, coords=COORDS, name='v')}
  T = {'v': xl.DataArray(t, DIMS, coords=COORDS, name='v')}
  se = oracle.squared_error(p, t)
  w = oracle.grid_area_weights(LAT)
  want = []
  for m in (land, ~land):
    a, b, _ = oracle.aggregate(se * m[None, None], DIMS, RD,
                               weights=[(w, ('latitude',))])
    a2, b2, _ = oracle.aggregate(
        np.broadcast_to(m[None, None], se.shape).astype(np.float64), DIMS, RD,
        weights=[(w, ('latitude',))])
    want.append(a / a2)
  want = np.stack(want, axis=-1)
  for lat, field in ((LAT, land), (LAT[::-1].copy(), land[::-1].copy())):
    lsm = xl.DataArray(field.astype(np.float32), ('latitude', 'longitude'),
                       coords={'latitude': lat, 'longitude': LON})
    agg = aggregation.Aggregator(
        reduce_dims=RD, weigh_by=[weighting.GridAreaWeighting()],
        bin_by=[binning.LandSea(lsm)])
    got = aggregation.compute_metric_values_for_single_chunk(
        {'mse': deterministic.MSE()}, agg, P, T)['mse.v']
    np.testing.assert_allclose(
        got.transpose('lead_time', 'land_sea').values, want, rtol=1e-6)
  other = xl.DataArray(land.astype(np.float32), ('latitude', 'longitude'),
                       coords={'latitude': LAT * 0.5, 'longitude': LON})
  agg = aggregation.Aggregator(reduce_dims=RD, bin_by=[binning.LandSea(other)])
  with pytest.raises(ValueError, match='bin mask'):
    aggregation.compute_metric_values_for_single_chunk(
        {'mse': deterministic.MSE()}, agg, P, T)


def test_member_mean_drops_a_mask_that_carries_the_ensemble_dim(monkeypatch):
  """EnsembleAveragedMetric under Aggregator(masked=True)
  when the 'mask' coordinate has the ensemble dim (add_nan_mask_to_data on
  ensemble predictions): `.mean(ensemble_dim)` drops that coordinate in the
  reference, the averaged statistic is unmasked and a NaN member propagates."""
  wbx_emulator.installed(monkeypatch)
  rng = np.random.default_rng(2)
  edims = ('init_time', 'realization', 'latitude', 'longitude')
  x = rng.normal(size=(3, 4, NY, NX)).astype(np.float32)
  x[0, 1, 2, 3] = np.nan
  y = rng.normal(size=(3, NY, NX)).astype(np.float32)
  ecoords = {'init_time': INIT, 'realization': np.arange(4), 'latitude': LAT,
             'longitude': LON}
  X = xl.DataArray(x, edims, coords=ecoords, name='v')
  X = X.assign_coords(mask=xl.DataArray(~np.isnan(x), edims))
  Y = xl.DataArray(y, ('init_time', 'latitude', 'longitude'),
                   coords={k: ecoords[k] for k in
                           ('init_time', 'latitude', 'longitude')}, name='v')
  metric = {'mse': probabilistic.EnsembleAveragedMetric(
      deterministic.MSE(), ensemble_dim='realization', skipna_ensemble=False)}
  agg = aggregation.Aggregator(reduce_dims=RD, masked=True)
  got = aggregation.compute_metric_values_for_single_chunk(
      metric, agg, {'v': X}, {'v': Y})['mse.v']
  assert np.isnan(got.values).all()
  # a mask WITHOUT the ensemble dim survives the member mean and is applied
  ymask = np.ones(y.shape, bool)
  ymask[0, 2, 3] = False
  X2 = xl.DataArray(x, edims, coords=ecoords, name='v')
  Y2 = Y.assign_coords(mask=xl.DataArray(
      ymask, ('init_time', 'latitude', 'longitude')))
  got2 = aggregation.compute_metric_values_for_single_chunk(
      metric, agg, {'v': X2}, {'v': Y2})['mse.v']
  se = (x.astype(np.float64) - y[:, None]) ** 2
  want = np.where(ymask[:, None], se, 0).mean(axis=1).sum() / ymask.sum()
  np.testing.assert_allclose(got2.values, want, rtol=1e-5)


def test_netcdf_name_collisions(tmp_path):
  """names that collide after sanitising are kept apart;
  one coordinate name with two different value sets raises."""
  ds = xl.Dataset()
  ds['a#b'] = xl.DataArray([1.0, 2.0], ('x',), coords={'x': [0, 1]})
  ds['a=b'] = xl.DataArray([3.0, 4.0], ('x',), coords={'x': [0, 1]})
  path = str(tmp_path / 'names.nc')
  io_netcdf.to_netcdf(ds, path)
  back = io_netcdf.open_dataset(path)
  assert set(back) == {'a#b', 'a=b'}
  np.testing.assert_array_equal(back['a#b'].values, [1, 2])
  np.testing.assert_array_equal(back['a=b'].values, [3, 4])
  bad = xl.Dataset()
  bad['u'] = xl.DataArray([1.0, 2.0], ('level',), coords={'level': [500, 850]})
  bad['v'] = xl.DataArray([3.0, 4.0], ('level',), coords={'level': [700, 1000]})
  with pytest.raises(ValueError, match='different values'):
    io_netcdf.to_netcdf(bad, str(tmp_path / 'bad.nc'))


def test_state_data_tree_round_trip():
  """aggregation_test.py:248-270 of the reference (DataTree / Dataset forms)."""
  state = aggregation.AggregationState(
      sum_weighted_statistics={'stat_name': {
          'var1': xl.DataArray([1.0, 2.0], ('x',)),
          'var2': xl.DataArray([3.0, 4.0], ('x',))}},
      sum_weights={'stat_name': {
          'var1': xl.DataArray([5.0, 6.0], ('x',)),
          'var2': xl.DataArray([7.0, 8.0], ('x',))}})
  tree = state.to_data_tree()
  assert set(tree.children) == {'stat_name'}
  assert set(tree['stat_name'].children) == {'var1', 'var2'}
  np.testing.assert_array_equal(
      tree['/stat_name/var1'].dataset['sum_weights'].values, [5, 6])
  assert set(tree.to_dict()) == {'/', '/stat_name', '/stat_name/var1',
                                 '/stat_name/var2'}
  for back in (aggregation.AggregationState.from_data_tree(tree),
               aggregation.AggregationState.from_dataset(state.to_dataset())):
    xarray_tree.map_structure(
        xl.testing.assert_allclose,
        (state.sum_weighted_statistics, state.sum_weights),
        (back.sum_weighted_statistics, back.sum_weights))
    assert back.sum_weights['stat_name']['var2'].name == 'var2'
  assert set(state.to_dataset()) == {
      'stat_name#var1#sum_weighted_statistics', 'stat_name#var1#sum_weights',
      'stat_name#var2#sum_weighted_statistics', 'stat_name#var2#sum_weights'}
  # a single-leaf state stays a bare DataArray (aggregation.py:205-208,221-224)
  leaf = aggregation.AggregationState(xl.DataArray([1.0, 2.0], ('x',)),
                                      xl.DataArray([3.0, 4.0], ('x',)))
  assert set(leaf.to_dataset()) == {'#sum_weighted_statistics', '#sum_weights'}
  back = aggregation.AggregationState.from_dataset(leaf.to_dataset())
  assert isinstance(back.sum_weighted_statistics, xl.DataArray)
  np.testing.assert_array_equal(back.sum_weights.values, [3, 4])
  with pytest.raises(TypeError):
    aggregation.AggregationState(1.0, 2.0).to_data_tree()


class _CountingAggregator(aggregation.Aggregator):
  """Oracle-free stand-in: sums `predictions - targets` on the host."""

  def aggregate_statistics(self, statistics):
    sws, sw = {}, {}
    for name, per_var in statistics.items():
      sws[name], sw[name] = {}, {}
      for var, stat in per_var.items():
        arr = (stat.predictions.to_numpy().astype(np.float64) -
               stat.targets.to_numpy())
        sws[name][var] = xl.DataArray(arr.sum())
        sw[name][var] = xl.DataArray(float(arr.size))
    return aggregation.AggregationState(sws, sw)


def test_checkpoint_of_another_evaluation_is_refused(tmp_path):
  """resume only merges partial sums of the SAME evaluation
  (metrics, aggregators, times, chunking), not just the same chunk count."""
  rng = np.random.default_rng(0)
  init = (np.datetime64('2020-01-01T00', 'ns') +
          np.arange(4) * np.timedelta64(12, 'h'))
  lead = (np.arange(2) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
  valid = (np.datetime64('2020-01-01T00', 'ns') +
           np.arange(8) * np.timedelta64(6, 'h'))
  grid = {'latitude': LAT, 'longitude': LON}
  fc = {'v': xl.DataArray(
      rng.normal(size=(4, 2, NY, NX)).astype(np.float32), DIMS,
      coords=dict(grid, init_time=init, lead_time=lead), name='v')}
  an = {'v': xl.DataArray(
      rng.normal(size=(8, NY, NX)).astype(np.float32),
      ('valid_time', 'latitude', 'longitude'),
      coords=dict(grid, valid_time=valid), name='v')}
  times = time_chunks.TimeChunks(init, lead, init_time_chunk_size=1)
  agg = _CountingAggregator(reduce_dims=RD)
  ckpt = str(tmp_path / 'ckpt')

  def run(metrics):
    return pipeline.run_pipeline(
        times, array_loaders.PredictionsFromArrays(fc),
        array_loaders.TargetsFromArrays(an), metrics, agg,
        require_output=False, checkpoint_path=ckpt, checkpoint_every=1,
        prefetch=0)

  first = run({'bias': deterministic.Bias()})
  again = run({'bias': deterministic.Bias()})   # same evaluation: resumes
  np.testing.assert_allclose(again[None][1]['bias.v'].values,
                             first[None][1]['bias.v'].values)
  with pytest.raises(ValueError, match='different evaluation'):
    run({'mae': deterministic.MAE()})
