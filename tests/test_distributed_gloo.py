"""world_size-2 gloo tests (CPU) of the multi-rank layer: unit sharding,
AggregationState packing, the packed all-reduce, the outer-join general path,
and the sharded evaluation driver (chunk-combine == monolithic, the identity
beam_pipeline_test.py:82-170 tests for the reference's Beam pipeline).

The per-rank states are produced by the oracle here (tests may use it); the
code under test is weatherbenchx_b200.distributed.
"""

import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.dirname(__file__)):
  if p not in sys.path:
    sys.path.insert(0, p)

import wbx_oracle as oracle  # noqa: E402
from weatherbenchx_b200 import aggregation  # noqa: E402
from weatherbenchx_b200 import distributed  # noqa: E402
from weatherbenchx_b200 import xarray_lite as xl  # noqa: E402
from weatherbenchx_b200.metrics import deterministic  # noqa: E402


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _run(fn, world_size=2):
  port = _free_port()
  with tempfile.TemporaryDirectory() as tmp:
    mp.spawn(_entry, args=(world_size, port, fn.__name__, tmp),
             nprocs=world_size, join=True)
    return [np.load(os.path.join(tmp, f'rank{r}.npz'), allow_pickle=True)
            for r in range(world_size)]


def _entry(rank, world_size, port, fn_name, tmp):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world_size)
  try:
    out = globals()[fn_name](rank, world_size)
    np.savez(os.path.join(tmp, f'rank{rank}.npz'), **out)
  finally:
    dist.destroy_process_group()


# --------------------------------------------------------------------------
# data + an oracle-backed stand-in for the (GPU) Aggregator
# --------------------------------------------------------------------------

N_INIT, NLAT, NLON = 6, 9, 12
LAT = np.linspace(-80, 80, NLAT)


def _fields(var, seed):
  rng = np.random.default_rng(seed)
  coords = {'init_time': np.arange(N_INIT), 'latitude': LAT,
            'longitude': np.arange(NLON) * 30.0}
  dims = ('init_time', 'latitude', 'longitude')
  p = xl.DataArray(rng.normal(size=(N_INIT, NLAT, NLON)).astype(np.float32),
                   dims, coords=coords, name=var)
  t = xl.DataArray(rng.normal(size=(N_INIT, NLAT, NLON)).astype(np.float32),
                   dims, coords=coords, name=var)
  return p, t


class OracleAggregator:
  """Same contract as aggregation.Aggregator.aggregate_statistics, evaluated
  with the CPU oracle (test double; the real one launches CUDA kernels)."""

  def __init__(self, reduce_dims):
    self.reduce_dims = reduce_dims

  def aggregate_statistics(self, statistics):
    w = oracle.grid_area_weights(LAT)
    sws, sw = {}, {}
    for name, per_var in statistics.items():
      sws[name], sw[name] = {}, {}
      for var, lazy in per_var.items():
        fn = oracle.DETERMINISTIC_STATISTICS[lazy.kind]
        a, b, dims = oracle.aggregate(
            fn(lazy.predictions.values, lazy.targets.values), lazy.dims,
            self.reduce_dims, weights=[(w, ('latitude',))])
        coords = {d: lazy.coords[d] for d in dims if d in lazy.coords}
        sws[name][var] = xl.DataArray(a, dims, coords=coords, name=var)
        sw[name][var] = xl.DataArray(b, dims, coords=coords, name=var)
    return aggregation.AggregationState(sws, sw)


def _monolithic(reduce_dims):
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE()}
  from weatherbenchx_b200.metrics import base as metrics_base
  preds = {v: _fields(v, i)[0] for i, v in enumerate(('u', 'v', 'z'))}
  tgts = {v: _fields(v, i)[1] for i, v in enumerate(('u', 'v', 'z'))}
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, preds, tgts)
  return OracleAggregator(reduce_dims).aggregate_statistics(
      stats).metric_values(metrics)


# --------------------------------------------------------------------------
# worker bodies (run in every rank)
# --------------------------------------------------------------------------


def _w_packed_allreduce(rank, world_size):
  p, t = _fields('u', 0)
  sl = slice(rank * 3, rank * 3 + 3)        # init_time split, init reduced
  stat = oracle.squared_error(p.values[sl], t.values[sl])
  sws, sw, _ = oracle.aggregate(stat, p.dims, ['init_time', 'latitude',
                                                'longitude'])
  if rank == 1:
    sws = sws + np.nan                      # NaN must survive the reduction
  state = aggregation.AggregationState(
      {'SquaredError': {'u': xl.DataArray(sws, ()),
                        'nan': xl.DataArray(np.float64(rank), ())}},
      {'SquaredError': {'u': xl.DataArray(sw, ()),
                        'nan': xl.DataArray(np.float64(1.0), ())}})
  total = distributed.all_reduce_state(state)
  return {'sw': total.sum_weights['SquaredError']['u'].values,
          'sws': total.sum_weighted_statistics['SquaredError']['u'].values,
          'other': total.sum_weighted_statistics['SquaredError']['nan'].values}


def _w_sharded_eval_reduced(rank, world_size):
  return _sharded(['init_time', 'latitude', 'longitude'])


def _w_sharded_eval_kept_init(rank, world_size):
  return _sharded(['latitude', 'longitude'])


def _sharded(reduce_dims):
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE()}
  units = [(v, c) for v in ('u', 'v', 'z') for c in range(3)]  # 2 inits/chunk

  def load_unit(unit):
    var, chunk = unit
    p, t = _fields(var, ('u', 'v', 'z').index(var))
    sl = {'init_time': slice(2 * chunk, 2 * chunk + 2)}
    return {var: p.isel(sl)}, {var: t.isel(sl)}

  values = distributed.evaluate_sharded(
      metrics, OracleAggregator(reduce_dims), units, load_unit)
  return {k: v.values for k, v in values.items()}


class _Patch:
  """monkeypatch stand-in for the spawned workers."""

  def setattr(self, obj, name, value):
    setattr(obj, name, value)


def _categorical_setup():
  """Thresholded contingency metrics, error exceedance and SEEPS on rain-like
  fields; the REAL Aggregator, its plans interpreted by tests/wbx_emulator.py
  (host side only -- the kernels are GPU-tested)."""
  import wbx_emulator
  from weatherbenchx_b200 import weighting
  from weatherbenchx_b200.metrics import categorical, wrappers
  wbx_emulator.installed(_Patch())
  init = np.datetime64('2020-02-27T00', 'ns') + np.arange(
      N_INIT) * np.timedelta64(12, 'h')
  lead = (np.arange(2) * np.timedelta64(12, 'h')).astype('timedelta64[ns]')
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': init, 'lead_time': lead, 'latitude': LAT,
            'longitude': np.arange(NLON) * 30.0}
  rng = np.random.default_rng(77)
  shape = (N_INIT, 2, NLAT, NLON)
  rain = lambda: (np.round(rng.gamma(0.8, 2.0, shape) * 4) / 4 *  # noqa: E731
                  (rng.random(shape) < 0.6)).astype(np.float32)
  preds = {'rain': xl.DataArray(rain(), dims, coords=coords, name='rain')}
  tgts = {'rain': xl.DataArray(rain(), dims, coords=coords, name='rain')}
  cdims = ('hour', 'dayofyear', 'latitude', 'longitude')
  ccoords = {'hour': np.array([0, 12]), 'dayofyear': np.arange(1, 367),
             'latitude': LAT, 'longitude': coords['longitude']}
  clim = xl.Dataset({
      'rain_seeps_threshold': xl.DataArray(
          (np.round(rng.uniform(0.5, 3, (2, 366, NLAT, NLON)) * 4) / 4
           ).astype(np.float32), cdims, coords=ccoords),
      'rain_seeps_dry_fraction': xl.DataArray(
          np.broadcast_to(rng.uniform(0.0, 1.0, (NLAT, NLON)),
                          (2, 366, NLAT, NLON)).astype(np.float32),
          cdims, coords=ccoords)})
  both = [wrappers.ContinuousToBinary('both', [0.0, 0.5, 2.0], 'threshold')]
  metrics = {
      'csi': wrappers.WrappedMetric(categorical.CSI(), both),
      'ets': wrappers.WrappedMetric(categorical.ETS(), both),
      'exceed': deterministic.ErrorExceedance([0.25, 1.0]),
      'seeps': categorical.SEEPS(['rain'], clim, dry_threshold_mm=250.0)}
  make_aggregator = lambda rd: aggregation.Aggregator(  # noqa: E731
      reduce_dims=rd, weigh_by=[weighting.GridAreaWeighting()], masked=True)
  return metrics, make_aggregator, preds, tgts


def _categorical_sharded(reduce_dims):
  metrics, make_aggregator, preds, tgts = _categorical_setup()
  units = list(range(3))                                  # 2 init_times each

  def load_unit(chunk):
    sl = {'init_time': slice(2 * chunk, 2 * chunk + 2)}
    return ({'rain': preds['rain'].isel(sl)}, {'rain': tgts['rain'].isel(sl)})

  values = distributed.evaluate_sharded(
      metrics, make_aggregator(reduce_dims), units, load_unit)
  return {k: v.values for k, v in values.items()}


def _w_categorical_reduced(rank, world_size):
  return _categorical_sharded(['init_time', 'latitude', 'longitude'])


def _w_categorical_kept_init(rank, world_size):
  return _categorical_sharded(['latitude', 'longitude'])


def _categorical_monolithic(reduce_dims):
  metrics, make_aggregator, preds, tgts = _categorical_setup()
  return aggregation.compute_metric_values_for_single_chunk(
      metrics, make_aggregator(reduce_dims), preds, tgts)


# --------------------------------------------------------------------------
# tests
# --------------------------------------------------------------------------


def test_shard_units_partitions_everything_once():
  units = list(range(11))
  for world_size in (1, 2, 3, 8, 16):
    got = [distributed.shard_units(units, r, world_size)
           for r in range(world_size)]
    assert sum(got, []) == units
    assert max(map(len, got)) - min(map(len, got)) <= 1


def test_ranks_are_spread_over_the_visible_devices():
  """Fewer ranks than GPUs: every other (fourth, ...) device, so that as few
  ranks as possible share a host uplink."""
  place = distributed.device_for_local_rank
  assert [place(r, 1, 8) for r in range(1)] == [0]
  assert [place(r, 2, 8) for r in range(2)] == [0, 4]
  assert [place(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]
  assert [place(r, 8, 8) for r in range(8)] == list(range(8))
  assert [place(r, 3, 8) for r in range(3)] == [0, 2, 4]
  assert [place(r, 4, 4) for r in range(4)] == [0, 1, 2, 3]   # nothing spare
  assert [place(r, 4, 6) for r in range(4)] == [0, 1, 2, 3]
  assert place(0, 1, 1) == 0


def test_pack_unpack_round_trip():
  state = aggregation.AggregationState(
      {'s': {'a': xl.DataArray(np.arange(6.0).reshape(2, 3), ('x', 'y'),
                               coords={'x': [1, 2]}),
             'b': xl.DataArray(np.float64(7.0), ())}},
      {'s': {'a': xl.DataArray(np.ones((2, 3)), ('x', 'y'),
                               coords={'x': [1, 2]}),
             'b': xl.DataArray(np.float64(2.0), ())}})
  layout = distributed.state_layout(state)
  flat = distributed.pack_state(state, layout)
  assert flat.shape == (14,)
  back = distributed.unpack_state(flat, layout)
  np.testing.assert_array_equal(
      back.sum_weighted_statistics['s']['a'].values, np.arange(6.0).reshape(2, 3))
  np.testing.assert_array_equal(back.sum_weights['s']['b'].values, 2.0)
  np.testing.assert_array_equal(
      back.sum_weights['s']['a'].coords['x'].values, [1, 2])


def test_combine_states_unions_keys_and_concatenates():
  a = aggregation.AggregationState(
      {'s': {'u': xl.DataArray([1.0], ('init_time',), coords={'init_time': [0]})}},
      {'s': {'u': xl.DataArray([2.0], ('init_time',), coords={'init_time': [0]})}})
  b = aggregation.AggregationState(
      {'s': {'u': xl.DataArray([3.0], ('init_time',), coords={'init_time': [1]}),
             'v': xl.DataArray([5.0], ('init_time',), coords={'init_time': [1]})}},
      {'s': {'u': xl.DataArray([4.0], ('init_time',), coords={'init_time': [1]}),
             'v': xl.DataArray([6.0], ('init_time',), coords={'init_time': [1]})}})
  c = distributed.combine_states([a, aggregation.AggregationState.zero(), b])
  np.testing.assert_array_equal(
      c.sum_weighted_statistics['s']['u'].values, [1.0, 3.0])
  np.testing.assert_array_equal(c.sum_weights['s']['v'].values, [6.0])


def test_packed_allreduce_gloo():
  res = _run(_w_packed_allreduce)
  p, t = _fields('u', 0)
  _, sw, _ = oracle.aggregate(oracle.squared_error(p.values, t.values), p.dims,
                              ['init_time', 'latitude', 'longitude'])
  for r in res:
    np.testing.assert_allclose(r['sw'], sw, rtol=1e-13)
    assert np.isnan(r['sws'])
    assert r['other'] == 1.0


def test_sharded_evaluation_equals_monolithic_gloo():
  """N-rank chunked evaluation == single evaluation of everything."""
  res = _run(_w_sharded_eval_reduced)
  mono = _monolithic(['init_time', 'latitude', 'longitude'])
  for r in res:
    assert set(r.files) == set(mono)
    for k in mono:
      np.testing.assert_allclose(r[k], mono[k].values, rtol=1e-12)


def test_sharded_evaluation_kept_init_time_gloo():
  """init_time kept: ranks hold disjoint init_times -> concatenation."""
  res = _run(_w_sharded_eval_kept_init)
  mono = _monolithic(['latitude', 'longitude'])
  for r in res:
    for k in mono:
      assert r[k].shape == (N_INIT,)
      np.testing.assert_allclose(r[k], mono[k].values, rtol=1e-12)


@pytest.mark.parametrize('worker,reduce_dims', [
    ('_w_categorical_reduced', ['init_time', 'latitude', 'longitude']),
    ('_w_categorical_kept_init', ['latitude', 'longitude'])])
def test_sharded_categorical_evaluation_equals_monolithic_gloo(worker,
                                                               reduce_dims):
  """The thresholded contingency metrics, error exceedance and SEEPS through
  the real Aggregator on two ranks (init_time chunks sharded, one combine of
  the states) == one evaluation of everything."""
  res = _run(globals()[worker])
  mono = _categorical_monolithic(reduce_dims)
  assert set(mono) == {'csi.rain', 'ets.rain', 'exceed.rain', 'seeps.rain'}
  for r in res:
    assert set(r.files) == set(mono)
    for k in mono:
      assert r[k].shape == mono[k].values.shape, k
      np.testing.assert_allclose(r[k], mono[k].values, rtol=1e-12,
                                 equal_nan=True, err_msg=k)
  if 'init_time' not in reduce_dims:
    assert res[0]['csi.rain'].shape == (N_INIT, 2, 3)
