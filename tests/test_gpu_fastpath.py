"""Replay of planned chunk evaluations (weatherbenchx_b200/fastpath.py) on the
B200: a repeated `compute_metric_values_for_single_chunk` call over the same
device arrays must return exactly what the ordinary path returns, follow
in-place refills of the arrays, and never outlive or resurrect its inputs."""

import gc

import numpy as np
import pytest

import wbx_oracle as oracle
from weatherbenchx_b200 import aggregation, binning, engine, fastpath
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import deterministic, probabilistic

pytestmark = pytest.mark.gpu

NY, NX = 64, 128
LAT = np.linspace(-90, 90, NY)
LON = np.linspace(0, 360, NX, endpoint=False)
INIT = (np.datetime64('2020-01-01T00', 'ns') +
        np.arange(6) * np.timedelta64(12, 'h'))
LEAD = (np.arange(3) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
DIMS = ('init_time', 'lead_time', 'latitude', 'longitude')
COORDS = {'init_time': INIT, 'lead_time': LEAD, 'latitude': LAT,
          'longitude': LON}
RD = ['init_time', 'latitude', 'longitude']


def _device_data(seed, names=('a', 'b', 'c')):
  rng = np.random.default_rng(seed)
  P, T, host = {}, {}, {}
  for n in names:
    p = rng.normal(280, 5, (6, 3, NY, NX)).astype(np.float32)
    t = (p + rng.normal(0, 2, p.shape)).astype(np.float32)
    host[n] = (p, t)
    P[n] = engine.to_device(xl.DataArray(p, DIMS, coords=COORDS, name=n))
    T[n] = engine.to_device(xl.DataArray(t, DIMS, coords=COORDS, name=n))
  return P, T, host, rng


def _same(a, b):
  assert list(a) == list(b)
  for k in a:
    assert a[k].dims == b[k].dims
    np.testing.assert_array_equal(a[k].values, b[k].values, err_msg=k)


def test_replay_equals_ordinary_and_follows_refills():
  fastpath.clear()
  P, T, host, rng = _device_data(0)
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE()}
  agg = aggregation.Aggregator(reduce_dims=RD,
                               weigh_by=[weighting.GridAreaWeighting()])
  before = dict(fastpath.STATS)
  first = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
  assert not isinstance(first, fastpath.LazyDataset)
  second = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
  assert isinstance(second, fastpath.LazyDataset) and second.is_pending
  _same(first, second)
  assert not second.is_pending
  assert fastpath.STATS['compiled'] == before['compiled'] + 1
  assert fastpath.STATS['replayed'] == before['replayed'] + 1
  w = oracle.grid_area_weights(LAT)
  for n, (p, t) in host.items():
    sws, sw, _ = oracle.aggregate(oracle.squared_error(p, t), DIMS, RD,
                                  weights=[(w, ('latitude',))])
    np.testing.assert_allclose(second[f'rmse.{n}'].values, np.sqrt(sws / sw),
                               rtol=1e-5)
  # refill one array in place: the replay reads the new numbers
  import torch
  new_p = rng.normal(280, 5, (6, 3, NY, NX)).astype(np.float32)
  P['b'].data.copy_(torch.from_numpy(new_p))
  third = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
  assert isinstance(third, fastpath.LazyDataset)
  sws, sw, _ = oracle.aggregate(oracle.squared_error(new_p, host['b'][1]),
                                DIMS, RD, weights=[(w, ('latitude',))])
  np.testing.assert_allclose(third['rmse.b'].values, np.sqrt(sws / sw),
                             rtol=1e-5)
  np.testing.assert_array_equal(third['rmse.a'].values, first['rmse.a'].values)
  fastpath.ENABLED = False
  try:
    slow = aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
  finally:
    fastpath.ENABLED = True
  assert not isinstance(slow, fastpath.LazyDataset)
  _same(slow, third)


def test_pipelined_reads_keep_their_own_results():
  """Results are read one call late (the intended use): every pending Dataset
  owns its pinned slot, later launches do not overwrite it."""
  fastpath.clear()
  import torch
  P, T, host, rng = _device_data(1, names=('a',))
  metrics = {'mse': deterministic.MSE()}
  agg = aggregation.Aggregator(reduce_dims=RD)
  aggregation.compute_metric_values_for_single_chunk(metrics, agg, P, T)
  fills = [rng.normal(280, 5, (6, 3, NY, NX)).astype(np.float32)
           for _ in range(5)]
  pending = []
  for f in fills:
    P['a'].data.copy_(torch.from_numpy(f), non_blocking=False)
    pending.append(aggregation.compute_metric_values_for_single_chunk(
        metrics, agg, P, T))
  assert all(isinstance(x, fastpath.LazyDataset) and x.is_pending
             for x in pending)
  for f, res in zip(fills, pending):
    want = ((f.astype(np.float64) - host['a'][1]) ** 2).mean(axis=(0, 2, 3))
    np.testing.assert_allclose(res['mse.a'].values, want, rtol=1e-5)


def test_bins_climatology_and_ensemble_replays():
  fastpath.clear()
  P, T, host, rng = _device_data(2, names=('a', 'b'))
  clim = {n: engine.to_device(xl.DataArray(
      rng.normal(280, 3, (366, 4, NY, NX)).astype(np.float32),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords={'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
              'latitude': LAT, 'longitude': LON}, name=n)) for n in P}
  land = xl.DataArray(rng.random((NY, NX)) < 0.4, ('latitude', 'longitude'),
                      coords={'latitude': LAT, 'longitude': LON})
  regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
             'box': ((10, 80), (30, 200))}
  metrics = {'rmse': deterministic.RMSE(), 'acc': deterministic.ACC(clim)}
  for agg in (
      aggregation.Aggregator(reduce_dims=RD,
                             weigh_by=[weighting.GridAreaWeighting()]),
      aggregation.Aggregator(
          reduce_dims=RD, weigh_by=[weighting.GridAreaWeighting()],
          bin_by=[binning.Regions(regions, land_sea_mask=land)]),
      aggregation.Aggregator(
          reduce_dims=RD, bin_by=[binning.ByTimeUnit('hour', 'init_time')])):
    first = aggregation.compute_metric_values_for_single_chunk(
        metrics, agg, P, T)
    second = aggregation.compute_metric_values_for_single_chunk(
        metrics, agg, P, T)
    assert isinstance(second, fastpath.LazyDataset)
    _same(first, second)
  x = rng.normal(size=(6, 10, NY, NX)).astype(np.float32)
  y = rng.normal(size=(6, NY, NX)).astype(np.float32)
  ecoords = {'init_time': INIT, 'number': np.arange(10), 'latitude': LAT,
             'longitude': LON}
  X = {'e': engine.to_device(xl.DataArray(
      x, ('init_time', 'number', 'latitude', 'longitude'), coords=ecoords,
      name='e'))}
  Y = {'e': engine.to_device(xl.DataArray(
      y, ('init_time', 'latitude', 'longitude'),
      coords={k: ecoords[k] for k in ('init_time', 'latitude', 'longitude')},
      name='e'))}
  ens = {'crps': probabilistic.CRPSEnsemble(ensemble_dim='number'),
         'ssr': probabilistic.UnbiasedSpreadSkillRatio(ensemble_dim='number')}
  agg = aggregation.Aggregator(reduce_dims=RD,
                               weigh_by=[weighting.GridAreaWeighting()])
  first = aggregation.compute_metric_values_for_single_chunk(ens, agg, X, Y)
  second = aggregation.compute_metric_values_for_single_chunk(ens, agg, X, Y)
  assert isinstance(second, fastpath.LazyDataset)
  _same(first, second)


def test_identity_rules():
  """New arrays, edited coordinates or a dropped input never hit a stale
  replay; inputs are not kept alive by the cache."""
  fastpath.clear()
  import weakref
  P, T, host, rng = _device_data(3, names=('a',))
  metrics = {'mse': deterministic.MSE()}
  agg = aggregation.Aggregator(reduce_dims=['latitude', 'longitude'])
  run = lambda: aggregation.compute_metric_values_for_single_chunk(  # noqa: E731
      metrics, agg, P, T)
  run()
  assert isinstance(run(), fastpath.LazyDataset)
  # edited coordinate: the kept init_time labels must show up in the result
  shifted = INIT + np.timedelta64(1, 'D')
  P['a'].coords['init_time'] = shifted
  T['a'].coords['init_time'] = shifted
  out = run()
  assert not isinstance(out, fastpath.LazyDataset)
  np.testing.assert_array_equal(out['mse.a'].coords['init_time'].values,
                                shifted)
  # other aggregator settings: own evaluation
  agg2 = aggregation.Aggregator(reduce_dims=['latitude', 'longitude'],
                                skipna=True)
  assert not isinstance(
      aggregation.compute_metric_values_for_single_chunk(metrics, agg2, P, T),
      fastpath.LazyDataset)
  # the cache holds its inputs weakly
  ref = weakref.ref(P['a'].data)
  del P['a'], out
  P['a'] = engine.to_device(xl.DataArray(
      host['a'][0], DIMS, coords=dict(COORDS, init_time=shifted), name='a'))
  gc.collect()
  assert ref() is None
  fresh = run()
  assert not isinstance(fresh, fastpath.LazyDataset)
  want = ((host['a'][0].astype(np.float64) - host['a'][1]) ** 2
          ).mean(axis=(2, 3))
  np.testing.assert_allclose(
      fresh['mse.a'].transpose('init_time', 'lead_time').values, want,
      rtol=1e-5)
