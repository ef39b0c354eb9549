"""CPU tests of the host side: labelled arrays, AggregationState arithmetic,
weighting / binning classes, the fused-launch planner (job tables are
interpreted with NumPy here and compared with the oracle), and the C ABI
exports.  No kernels run in this file.
"""

import ctypes
import pickle
import re
import os

import numpy as np
import pytest

import wbx_oracle as oracle
import wbx_test_utils as utils
from weatherbenchx_b200 import _build
from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import binning
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200 import xarray_tree
from weatherbenchx_b200.lazy import LazyStatistic
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import deterministic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------
# xarray_lite
# ---------------------------------------------------------------------------


def test_broadcast_by_name_and_coords():
  a = xl.DataArray(np.arange(6.0).reshape(2, 3), ('x', 'y'),
                   coords={'x': [10, 20], 'y': [1, 2, 3]})
  b = xl.DataArray(np.array([1.0, 2.0, 3.0]), ('y',), coords={'y': [1, 2, 3]})
  c = a * b
  assert c.dims == ('x', 'y')
  np.testing.assert_array_equal(c.values, a.values * b.values[None, :])
  d = b - a  # dims follow the first operand, then new dims
  assert d.dims == ('y', 'x')
  with pytest.raises(ValueError):
    _ = a + xl.DataArray(np.zeros(3), ('y',), coords={'y': [1, 2, 4]})
  assert np.sqrt(a).dims == a.dims
  assert hasattr(a, 'x') and not hasattr(a, 'mask')


def test_where_isnull_reductions():
  a = xl.DataArray(np.array([[1.0, np.nan], [3.0, 4.0]], np.float32),
                   ('x', 'y'))
  assert a.isnull().values.tolist() == [[False, True], [False, False]]
  filled = a.where(~a.isnull(), 0)
  assert filled.dtype == np.float32
  assert filled.sum().item() == 8.0
  assert np.isnan(a.sum('y', skipna=False).values[0])
  assert a.sum('y').values.tolist() == [1.0, 7.0]
  assert a.mean(('x', 'y'), skipna=True).item() == pytest.approx(8 / 3)
  assert a.count('y').values.tolist() == [1, 2]


def test_dot_and_align():
  s = xl.DataArray(np.arange(24.0).reshape(2, 3, 4), ('t', 'y', 'x'))
  w = xl.DataArray(np.array([1.0, 2.0, 3.0]), ('y',))
  m = xl.DataArray(np.ones((2, 3, 4), bool), ('r', 'y', 'x'))
  out = xl.dot(s, w, m, dim={'y', 'x'})
  assert out.dims == ('t', 'r')
  np.testing.assert_allclose(
      out.values, np.einsum('tyx,y,ryx->tr', s.values, w.values, m.values))
  a = xl.DataArray([1.0, 2.0], ('x',), coords={'x': [0, 1]})
  b = xl.DataArray([10.0, 20.0], ('x',), coords={'x': [1, 2]})
  a2, b2 = xl.align(a, b, join='outer', fill_value=0)
  np.testing.assert_array_equal(a2.coords['x'].values, [0, 1, 2])
  np.testing.assert_array_equal((a2 + b2).values, [1, 12, 20])


def test_isel_sel_expand_transpose():
  d = utils.mock_prediction_data(time_start='2020-01-01', time_stop='2020-01-04')
  g = d['geopotential']
  assert g.isel(time=0, drop=True).dims == (
      'prediction_timedelta', 'latitude', 'longitude', 'level')
  assert g.sel(level=700).shape == g.shape[:-1]
  assert g.sel(latitude=slice(-30, 30)).sizes['latitude'] == 7
  assert g.transpose('level', 'time', 'prediction_timedelta', 'longitude',
                     'latitude').shape[0] == 3
  e = g.isel(time=0, prediction_timedelta=0, drop=True).expand_dims(
      {'dayofyear': np.arange(1, 367), 'hour': [0, 6, 12, 18]})
  assert e.dims[:2] == ('dayofyear', 'hour') and e.shape[:2] == (366, 4)


# ---------------------------------------------------------------------------
# AggregationState / xarray_tree (aggregation_test.py:40-56,248-270)
# ---------------------------------------------------------------------------


def _example_state():
  return aggregation.AggregationState(
      sum_weighted_statistics={'stat_name': {
          'var1': xl.DataArray([1.0, 2.0], ('x',)),
          'var2': xl.DataArray([3.0, 4.0], ('x',))}},
      sum_weights={'stat_name': {
          'var1': xl.DataArray([5.0, 6.0], ('x',)),
          'var2': xl.DataArray([7.0, 8.0], ('x',))}})


def test_state_sum_mean_and_zero():
  s = _example_state()
  total = aggregation.AggregationState.sum(
      [s, aggregation.AggregationState.zero(), s])
  np.testing.assert_array_equal(
      total.sum_weighted_statistics['stat_name']['var1'].values, [2, 4])
  mean = (s + s).mean_statistics()
  np.testing.assert_allclose(mean['stat_name']['var2'].values, [3 / 7, 4 / 8])
  zero = aggregation.AggregationState.zero() + aggregation.AggregationState.zero()
  assert zero.sum_weighted_statistics is None
  assert s.sum_along_dims(['x']).sum_weights['stat_name']['var1'].item() == 11
  with pytest.raises(ValueError):
    aggregation.AggregationState.zero().map(lambda x: x)


def test_state_round_trip_dataset():
  s = _example_state()
  rt = aggregation.AggregationState.from_dataset(s.to_dataset())
  xarray_tree.map_structure(
      xl.testing.assert_allclose,
      (s.sum_weighted_statistics, s.sum_weights),
      (rt.sum_weighted_statistics, rt.sum_weights))


def test_combining_sum_outer_join():
  a = xl.DataArray([1.0, 2.0], ('init_time',), coords={'init_time': [0, 1]})
  b = xl.DataArray([5.0], ('init_time',), coords={'init_time': [2]})
  out = aggregation.combining_sum([a, b])
  np.testing.assert_array_equal(out.values, [1, 2, 5])


def test_metric_values_naming_and_rmse():
  state = aggregation.AggregationState(
      {'SquaredError': {'t': xl.DataArray([8.0, 18.0], ('lead_time',))}},
      {'SquaredError': {'t': xl.DataArray([2.0, 2.0], ('lead_time',))}})
  values = state.metric_values({'rmse': deterministic.RMSE(),
                                'mse': deterministic.MSE()})
  assert set(values) == {'rmse.t', 'mse.t'}
  np.testing.assert_allclose(values['rmse.t'].values, [2.0, 3.0])
  np.testing.assert_allclose(values['mse.t'].values, [4.0, 9.0])


def test_metrics_and_aggregator_pickle():
  """Beam pickles DoFn members (beam_pipeline.py:150-159)."""
  agg = aggregation.Aggregator(reduce_dims=['latitude'],
                               weigh_by=[weighting.GridAreaWeighting()])
  for obj in (agg, deterministic.RMSE(), deterministic.SquaredError()):
    assert type(pickle.loads(pickle.dumps(obj))) is type(obj)


def test_unique_statistics_are_deduplicated_and_lazy():
  p = utils.mock_prediction_data(time_start='2020-01-01', time_stop='2020-01-03')
  metrics = {'rmse': deterministic.RMSE(), 'mse': deterministic.MSE(),
             'mae': deterministic.MAE(), 'bias': deterministic.Bias()}
  stats = metrics_base.compute_unique_statistics_for_all_metrics(metrics, p, p)
  assert set(stats) == {'SquaredError', 'AbsoluteError', 'Error'}
  for per_var in stats.values():
    for s in per_var.values():
      assert isinstance(s, LazyStatistic) and s.is_lazy
  # variables only in predictions are dropped (base_test.py:24-90)
  only_p = dict(p, extra=p['2m_temperature'])
  stats = deterministic.SquaredError().compute(only_p, p)
  assert set(stats) == set(p)


def test_failed_statistic_is_wrapped_in_value_error():
  """metrics/base.py:263-269."""
  class Broken(metrics_base.PerVariableStatistic):

    def _compute_per_variable(self, predictions, targets):
      raise RuntimeError('boom')

  p = utils.mock_target_data(time_start='2020-01-01', time_stop='2020-01-02')
  with pytest.raises(ValueError, match='Failed to compute statistic'):
    metrics_base.compute_unique_statistics_for_all_metrics(
        {'b': Broken()}, p, p)


# ---------------------------------------------------------------------------
# weighting / binning classes against the oracle
# ---------------------------------------------------------------------------


def test_grid_area_weighting_class():
  """weighting_test.py:24-46 through the class surface."""
  d = utils.mock_prediction_data(time_start='2020-01-01T00',
                                 time_stop='2020-01-03T00')
  stat = d['2m_temperature']
  w = weighting.GridAreaWeighting().weights(stat)
  assert w.dims == ('latitude',)
  assert w.values.mean() == pytest.approx(1.0)
  np.testing.assert_allclose(
      w.values, oracle.grid_area_weights(stat.coords['latitude'].values),
      rtol=1e-13)
  raw = weighting.GridAreaWeighting(return_normalized=False)
  regional = raw.weights(stat.sel(latitude=slice(-30, 30)))
  np.testing.assert_allclose(
      regional.values, raw.weights(stat).sel(latitude=slice(-30, 30)).values)
  no_lat = xl.DataArray(np.zeros(3), ('x',))
  assert weighting.GridAreaWeighting().weights(no_lat).ndim == 0
  desc = stat.isel(latitude=slice(None, None, -1))
  np.testing.assert_allclose(
      weighting.GridAreaWeighting().weights(desc).values, w.values[::-1])


def test_regions_class_matches_oracle():
  d = utils.mock_target_data(time_start='2020-01-01', time_stop='2020-01-02',
                             spatial_resolution_in_degrees=5.0)
  stat = d['2m_temperature']
  regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
             'europe': ((35, 75), (-12.5, 42.5))}
  lat, lon = stat.coords['latitude'].values, stat.coords['longitude'].values
  rng = np.random.default_rng(0)
  land = xl.DataArray(rng.random((len(lat), len(lon))) > 0.7,
                      ('latitude', 'longitude'),
                      coords={'latitude': lat, 'longitude': lon})
  mask = binning.Regions(regions, land_sea_mask=land).create_bin_mask(stat)
  exp, names = oracle.regions_masks(lat, lon, regions, land.values)
  assert mask.dims == ('region', 'latitude', 'longitude')
  np.testing.assert_array_equal(mask.values, exp)
  assert mask.coords['region'].values.tolist() == names
  with pytest.raises(ValueError):
    binning.Regions({'bad': ((10, 0), (0, 10))}).create_bin_mask(stat)


# ---------------------------------------------------------------------------
# the planner: job tables interpreted on the CPU == oracle
# ---------------------------------------------------------------------------


def _read(addr, n, dtype):
  buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(addr))
  return np.frombuffer(buf, dtype=dtype, count=n)


def _interpret(spec):
  """NumPy execution of a FusedSpec (what the kernel is specified to do)."""
  slab = spec.ny * spec.nx
  ws = np.zeros((spec.n_cells, 6))
  w = np.zeros((spec.n_cells, 4))
  wy = spec.w_y if spec.w_y is not None else np.ones(spec.ny)
  wx = spec.w_x if spec.w_x is not None else np.ones(spec.nx)
  wgt = (wy[:, None] * wx[None, :]).reshape(-1)
  skipna = bool(spec.flags & _cabi.FLAG_SKIPNA)
  for j in range(len(spec.pred)):
    p = _read(spec.pred[j], slab, np.float32)
    t = _read(spec.target[j], slab, np.float32)
    c = (_read(spec.clim[j], slab, np.float32) if spec.clim is not None
         else np.zeros(slab, np.float32))
    m = (_read(spec.mask[j], slab, np.uint8) != 0 if spec.mask is not None
         else np.ones(slab, bool))
    wo = spec.w_outer[j] if spec.w_outer is not None else 1.0
    vals = [p - t, np.abs(p - t), (p - t) ** 2, (p - c) ** 2, (t - c) ** 2,
            (p - c) * (t - c)]
    for s, v in enumerate(vals):
      valid = m & ~np.isnan(v) if skipna else m
      v = np.where(valid, v, 0).astype(np.float64)
      ws[spec.cell[j], s] += wo * (v * wgt).sum()
      w[spec.cell[j], _cabi.STAT_WCLASS[s]] = (
          w[spec.cell[j], _cabi.STAT_WCLASS[s]] +
          (wo * (valid * wgt).sum() if s in (0, 3, 4, 5) else 0.0))
  return ws * spec.scalar, w * spec.scalar


def _case(order, seed=0, shape=(3, 2, 2, 6, 8)):
  rng = np.random.default_rng(seed)
  dims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  coords = {
      'init_time': np.datetime64('2020-12-30T00', 'ns') +
                   np.arange(shape[0]) * np.timedelta64(1, 'D'),
      'lead_time': (np.arange(shape[1]) * np.timedelta64(12, 'h')
                    ).astype('timedelta64[ns]'),
      'level': [500, 850], 'latitude': np.linspace(-75, 75, shape[3]),
      'longitude': np.arange(shape[4]) * 45.0}
  p = xl.DataArray(rng.normal(size=shape).astype(np.float32), dims,
                   coords=coords, name='z')
  t = xl.DataArray(rng.normal(size=shape).astype(np.float32), dims,
                   coords=coords, name='z')
  t.data[rng.random(shape) < 0.1] = np.nan
  cdims = ('dayofyear', 'hour', 'level', 'latitude', 'longitude')
  c = xl.DataArray(
      rng.normal(size=(366, 4) + shape[2:]).astype(np.float32), cdims,
      coords={'dayofyear': np.arange(1, 367), 'hour': [0, 6, 12, 18],
              **{d: coords[d] for d in dims[2:]}}, name='z')
  mask = xl.DataArray(rng.random(shape) > 0.3, dims)
  if order != dims:
    p, t, mask = p.transpose(*order), t.transpose(*order), mask.transpose(*order)
    p = p.copy(data=np.ascontiguousarray(p.values))
    t = t.copy(data=np.ascontiguousarray(t.values))
    mask = mask.copy(data=np.ascontiguousarray(mask.values))
    # the climatology is stored in the same spatial layout as the data
    corder = ('dayofyear', 'hour') + tuple(
        d for d in order if d in ('level', 'latitude', 'longitude'))
    c = c.transpose(*corder)
    c = c.copy(data=np.ascontiguousarray(c.values))
  # the same mask on both inputs: all six statistics carry it (a mask on the
  # targets alone would leave SquaredPredictionAnomaly unmasked)
  return p.assign_coords(mask=mask), t.assign_coords(mask=mask), c


@pytest.mark.parametrize('order', [
    ('init_time', 'lead_time', 'level', 'latitude', 'longitude'),
    ('lead_time', 'init_time', 'level', 'longitude', 'latitude'),
    ('level', 'init_time', 'lead_time', 'latitude', 'longitude'),
])
@pytest.mark.parametrize('reduce_dims', [
    ('init_time', 'latitude', 'longitude'), ('latitude', 'longitude'),
    ('longitude', 'latitude', 'init_time', 'lead_time', 'level')])
@pytest.mark.parametrize('masked,skipna', [(False, False), (True, False),
                                           (False, True), (True, True)])
def test_planner_tables_reproduce_oracle(order, reduce_dims, masked, skipna):
  p, t, c = _case(order)
  stats = [LazyStatistic(k, p, t, engine.align_climatology(p, c))
           for k in ('SquaredPredictionAnomaly', 'SquaredTargetAnomaly',
                     'AnomalyCovariance')]
  stats += [LazyStatistic(k, p, t) for k in ('Error', 'AbsoluteError',
                                             'SquaredError')]
  w = weighting.GridAreaWeighting().weights(stats[0])
  two = xl.DataArray(np.array(2.0))
  lead_w = xl.DataArray(np.array([1.0, 3.0]), ('lead_time',))
  spec = engine.build_fused_spec(stats, reduce_dims, [w, two, lead_w],
                                 masked=masked, skipna=skipna)
  assert spec.space == _cabi.SPACE_HOST
  assert np.all(np.diff(spec.cell) >= 0) and spec.cell[-1] == spec.n_cells - 1
  ws, wsum = _interpret(spec)
  aligned, adims = oracle.align_climatology(
      c.values, c.dims, {k: c.coords[k].values for k in ('dayofyear', 'hour')},
      p.coords['init_time'].values, p.coords['lead_time'].values)
  aligned = np.transpose(aligned, [adims.index(d) for d in p.dims])
  fns = dict(oracle.DETERMINISTIC_STATISTICS)
  weights = [(w.values, ('latitude',)), (np.array(2.0), ()),
             (lead_w.values, ('lead_time',))]
  for s in stats:
    if s.kind in fns:
      val = fns[s.kind](p.values, t.values)
    else:
      val = oracle.CLIMATOLOGY_STATISTICS[s.kind](p.values, t.values, aligned)
    sws, sw, out_dims = oracle.aggregate(
        val, p.dims, reduce_dims, weights=weights,
        mask=t.coords['mask'].values, mask_dims=p.dims, masked=masked,
        skipna=skipna)
    assert tuple(spec.kept) == out_dims
    slot = _cabi.STAT_SLOT[s.kind]
    np.testing.assert_allclose(ws[:, slot].reshape(spec.kept_shape), sws,
                               rtol=1e-6, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(
        wsum[:, _cabi.STAT_WCLASS[slot]].reshape(spec.kept_shape), sw,
        rtol=1e-12)


def test_planner_rejects_what_the_slab_kernel_cannot_do():
  p, t, _ = _case(('init_time', 'lead_time', 'level', 'latitude', 'longitude'))
  stat = LazyStatistic('SquaredError', p, t)
  # trailing dim kept -> no slab
  with pytest.raises(engine.FastPathUnavailable):
    engine.build_fused_spec([stat], ['init_time'], [])
  # N-d weights
  nd = xl.DataArray(np.ones(p.shape, np.float32), p.dims)
  with pytest.raises(engine.FastPathUnavailable):
    engine.build_fused_spec([stat], ['latitude', 'longitude'], [nd])
  # reduce dim missing -> not applicable (None), like aggregation.py:305-309
  assert engine.build_fused_spec([stat], ['realization'], []) is None


def test_climatology_alignment_indices():
  """metrics/base.py:383-403: dayofyear/hour of init + lead, leap year."""
  p, _, c = _case(('init_time', 'lead_time', 'level', 'latitude', 'longitude'))
  ac = engine.align_climatology(p, c)
  assert ac.time_dims == ('init_time', 'lead_time')
  # 2020-12-30, 12-31, 2021-01-01 at +0h / +12h
  np.testing.assert_array_equal(ac.positions['dayofyear'] + 1,
                                [[365, 365], [366, 366], [1, 1]])
  np.testing.assert_array_equal(ac.positions['hour'], [[0, 2]] * 3)
  doy, hour = oracle.dayofyear_and_hour(
      p.coords['init_time'].values[:, None] +
      p.coords['lead_time'].values[None, :])
  np.testing.assert_array_equal(doy, ac.positions['dayofyear'] + 1)
  np.testing.assert_array_equal(hour // 6, ac.positions['hour'])


# ---------------------------------------------------------------------------
# C ABI: the library builds, loads and exports what the header declares
# ---------------------------------------------------------------------------


def test_library_exports_every_declared_symbol():
  path = _build.build_library()
  lib = ctypes.CDLL(str(path))
  header = open(os.path.join(ROOT, 'include', 'wbx_b200.h')).read()
  declared = set(re.findall(r'\b(wbx_[a-z0-9_]+)\s*\(', header))
  assert declared, 'no declarations found'
  assert declared == set(_cabi.SIGNATURES), (
      declared ^ set(_cabi.SIGNATURES))
  for name in declared:
    assert hasattr(lib, name), name
  lib.wbx_abi_version.restype = ctypes.c_int
  assert lib.wbx_abi_version() == _cabi.ABI_VERSION == 2


def test_struct_layouts_match_header_sizes():
  """Every ctypes mirror has the size and member offsets the library was
  compiled with (wbx_struct_layout: sizeof / offsetof from the C side)."""
  mirrors = [_cabi.DetDesc, _cabi.CrpsDesc, _cabi.CrpsPointDesc,
             _cabi.SpectrumDesc, _cabi.GenericDesc]
  for which, mirror in enumerate(mirrors):
    layout = _cabi.struct_layout(which)
    assert layout[0] == ctypes.sizeof(mirror), mirror.__name__
    offsets = [getattr(mirror, name).offset for name, _ in mirror._fields_]
    assert layout[1:] == offsets, mirror.__name__
  assert ctypes.sizeof(_cabi.DetDesc) == 8 + 4 * 8 + 8 * 8 + 8 + 8 + 8 + 16
  with pytest.raises(_cabi.WbxError):
    _cabi.struct_layout(99)


def test_no_gpu_means_loud_failure():
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  with pytest.raises(_cabi.WbxError, match='NO_DEVICE|no CUDA device'):
    _cabi.Context(0)


# ---------------------------------------------------------------------------
# bin masks folded into a class map (fused binned kernel, host side)
# ---------------------------------------------------------------------------


def test_fold_bin_masks_reproduces_every_mask():
  rng = np.random.default_rng(0)
  lat = np.linspace(-90, 90, 24)
  lon = np.arange(40) * 9.0
  stat = xl.DataArray(np.zeros((24, 40), np.float32), ('latitude', 'longitude'),
                      coords={'latitude': lat, 'longitude': lon})
  land = xl.DataArray(rng.random((24, 40)) > 0.6, ('latitude', 'longitude'),
                      coords={'latitude': lat, 'longitude': lon})
  regions = {'global': ((-90, 90), (0, 360)), 'tropics': ((-20, 20), (0, 360)),
             'nh': ((20, 90), (0, 360)), 'europe': ((35, 75), (-12.5, 42.5))}
  m1 = binning.Regions(regions, land_sea_mask=land).create_bin_mask(stat)
  m2 = binning.LandSea(xl.DataArray(land.values.astype(float), land.dims,
                                    coords=land.coords),
                       include_global_mask=True).create_bin_mask(stat)
  sizes = {'latitude': 24, 'longitude': 40}
  cls = engine.fold_bin_masks([m1, m2], ['region', 'land_sea'],
                              ['latitude', 'longitude'], sizes)
  assert cls.class_map.dtype == np.uint8 and cls.class_map.shape == (960,)
  assert cls.n_classes == cls.class_map.max() + 1 <= 256
  for mask, member in ((m1, cls.membership[0]), (m2, cls.membership[1])):
    rebuilt = member[:, cls.class_map].reshape(mask.shape)
    np.testing.assert_array_equal(rebuilt.astype(bool), mask.values)
  # class sums -> bin sums == direct masked sums
  field = rng.normal(size=960)
  per_class = np.bincount(cls.class_map, weights=field,
                          minlength=cls.n_classes)[None, :]
  got = cls.to_bins(per_class)[0]
  exp = np.einsum('s,as,bs->ab', field, m1.values.reshape(8, -1).astype(float),
                  m2.values.reshape(3, -1).astype(float))
  np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-12)
  # a latitude-only mask broadcasts over longitude
  band = xl.DataArray(np.stack([lat > 0, lat <= 0]), ('band', 'latitude'),
                      coords={'band': np.array(['n', 's'])})
  cls2 = engine.fold_bin_masks([band], ['band'], ['latitude', 'longitude'], sizes)
  assert cls2.n_classes == 2
  # masks that depend on an outer dim cannot be folded
  outer = xl.DataArray(np.ones((2, 3, 24), bool), ('b', 'lead_time', 'latitude'))
  with pytest.raises(engine.FastPathUnavailable):
    engine.fold_bin_masks([outer], ['b'], ['latitude', 'longitude'], sizes)


# ---------------------------------------------------------------------------
# plan memoisation (engine.build_fused_spec)
# ---------------------------------------------------------------------------


def test_spec_cache_recognises_repeats_and_retains_nothing():
  import gc
  import weakref
  from weatherbenchx_b200 import weighting
  from weatherbenchx_b200.lazy import LazyStatistic
  rng = np.random.default_rng(0)
  coords = {'init_time': np.arange(3), 'latitude': np.linspace(-80, 80, 8),
            'longitude': np.arange(16) * 22.5}
  dims = ('init_time', 'latitude', 'longitude')
  P = xl.DataArray(rng.normal(size=(3, 8, 16)).astype(np.float32), dims,
                   coords=coords, name='t')
  T = xl.DataArray(rng.normal(size=(3, 8, 16)).astype(np.float32), dims,
                   coords=coords, name='t')
  weigher = weighting.GridAreaWeighting()

  def plan(p, t, scale=1.0):
    stat = LazyStatistic('SquaredError', p, t)
    w = weigher.weights(stat)
    if scale != 1.0:
      w = w * scale
    return engine.build_fused_spec([stat], ['latitude', 'longitude'], [w])

  a, b = plan(P, T), plan(P, T)
  assert a is b                       # same arrays, same weights -> same plan
  assert weigher.weights(LazyStatistic('Error', P, T)) is weigher.weights(
      LazyStatistic('Error', P, T))
  c = plan(P, T, scale=2.0)           # other weight values -> other plan
  assert c is not a and c.w_y[0] == 2 * a.w_y[0]
  assert plan(P, T.isel(init_time=slice(0, 3))) is not a   # other operand
  # a float64 operand is converted while planning: such a plan owns the copy
  # and is rebuilt every time (its content may have changed)
  P64 = xl.DataArray(P.values.astype(np.float64), dims, coords=coords, name='t')
  d, e = plan(P64, T), plan(P64, T)
  assert d is not e and len(d.keepalive) == 1 and not a.keepalive
  # the cache holds no strong reference to the operands
  ref = weakref.ref(P.data)
  del P, a, b, c, d, e
  gc.collect()
  assert ref() is None


def test_combining_sum_blocks_with_unsorted_labels_and_kept_time():
  """Per-chunk blocks of a state that keeps init_time and carries an unsorted
  string-labelled bin dim (region names): blocks land at their init_time
  coordinates, the region axis is left as it is."""
  regions = np.array(['global', 'tropics', 'nh', 'sh'])
  rng = np.random.default_rng(0)
  full = rng.normal(size=(6, 4))
  blocks = [
      xl.DataArray(full[i:i + 2], ('init_time', 'region'),
                   coords={'init_time': np.arange(i, i + 2), 'region': regions})
      for i in (4, 0, 2)]
  total = aggregation.combining_sum(blocks)
  assert total.dims == ('init_time', 'region')
  np.testing.assert_array_equal(total.coords['init_time'].values, np.arange(6))
  np.testing.assert_array_equal(total.coords['region'].values, regions)
  np.testing.assert_array_equal(total.values, full)
  # overlapping blocks add; transposed blocks are aligned by name
  again = aggregation.combining_sum([total, blocks[1].transpose(
      'region', 'init_time')])
  expect = full.copy()
  expect[0:2] *= 2
  np.testing.assert_array_equal(again.values, expect)
  # different label sets on the string axis: sorted union, zero fill
  other = xl.DataArray(np.ones((2, 2)), ('init_time', 'region'),
                       coords={'init_time': np.arange(2),
                               'region': np.array(['europe', 'nh'])})
  mixed = aggregation.combining_sum([blocks[1], other])
  labels = mixed.coords['region'].values.tolist()
  assert labels == sorted(set(regions.tolist()) | {'europe'})
  np.testing.assert_array_equal(
      mixed.sel(region='nh').values, full[0:2, 2] + 1)
  np.testing.assert_array_equal(mixed.sel(region='europe').values, [1, 1])


# ---------------------------------------------------------------------------
# 'mask' coordinate per statistic (what the reference's expressions carry)
# ---------------------------------------------------------------------------


def _masked_inputs():
  rng = np.random.default_rng(3)
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': np.datetime64('2020-01-01', 'ns') +
                         np.arange(2) * np.timedelta64(1, 'D'),
            'lead_time': (np.arange(2) * np.timedelta64(6, 'h')
                          ).astype('timedelta64[ns]'),
            'latitude': np.linspace(-90, 90, 5),
            'longitude': np.linspace(0, 360, 8, endpoint=False)}
  shape = (2, 2, 5, 8)
  p = xl.DataArray(rng.normal(size=shape).astype(np.float32), dims,
                   coords=coords)
  holes = rng.random(shape) < 0.2
  t_values = rng.normal(size=shape).astype(np.float32)
  t_values[holes] = np.nan
  t = xl.DataArray(t_values, dims, coords=dict(coords, mask=(dims, ~holes)))
  clim = xl.DataArray(
      rng.normal(size=(366, 4, 5, 8)).astype(np.float32),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords={'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
              'latitude': coords['latitude'],
              'longitude': coords['longitude']})
  return p, t, clim


def test_mask_coordinate_follows_the_operands_of_each_statistic():
  """deterministic.py:225-259: (p - c)**2 has no targets' mask; the other
  statistics inherit it (aggregation.py:339 tests hasattr(stat, 'mask'))."""
  from weatherbenchx_b200.metrics import deterministic, probabilistic
  p, t, clim = _masked_inputs()
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      {'acc': deterministic.ACC({'x': clim}), 'rmse': deterministic.RMSE()},
      {'x': p}, {'x': t})
  has_mask = {name: 'mask' in per_var['x'].coords
              for name, per_var in stats.items()}
  assert has_mask == {'SquaredPredictionAnomaly': False,
                      'SquaredTargetAnomaly': True, 'AnomalyCovariance': True,
                      'SquaredError': True}
  # a mask on the predictions instead: the mirror image
  p2 = p.assign_coords(mask=t.coords['mask'])
  t2 = t.drop_vars('mask')
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      {'acc': deterministic.ACC({'x': clim})}, {'x': p2}, {'x': t2})
  assert {n: 'mask' in v['x'].coords for n, v in stats.items()} == {
      'SquaredPredictionAnomaly': True, 'SquaredTargetAnomaly': False,
      'AnomalyCovariance': True}
  # conflicting masks are dropped by the coordinate merge
  flipped = xl.DataArray(~t.coords['mask'].values, t.dims)
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      {'rmse': deterministic.RMSE()}, {'x': p.assign_coords(mask=flipped)},
      {'x': t})
  assert 'mask' not in stats['SquaredError']['x'].coords
  # ensembles: spread and variance are functions of the predictions alone
  ens = xl.DataArray(
      np.zeros((2, 2, 5, 8, 3), np.float32), p.dims + ('realization',),
      coords={d: p.coords[d].values for d in p.dims})
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      {'crps': probabilistic.CRPSEnsemble(ensemble_dim='realization'),
       'ssr': probabilistic.UnbiasedSpreadSkillRatio(
           ensemble_dim='realization')}, {'x': ens}, {'x': t})
  has_mask = {name.split('_')[0]: 'mask' in per_var['x'].coords
              for name, per_var in stats.items()}
  assert has_mask == {'CRPSSkill': True, 'CRPSSpread': False,
                      'EnsembleVariance': False,
                      'UnbiasedEnsembleMeanSquaredError': True}


def test_masked_aggregation_launches_unmasked_statistics_separately(
    monkeypatch):
  """Aggregator(masked=True): one launch per distinct mask, and statistics
  without a mask coordinate are not masked (engine stubbed: no GPU here)."""
  from weatherbenchx_b200.metrics import deterministic
  p, t, clim = _masked_inputs()
  metrics = {'acc': deterministic.ACC({'x': clim}),
             'rmse': deterministic.RMSE()}
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, {'x': p}, {'x': t})
  launches = []

  def build(stat_list, reduce_dims, weights, masked=False, skipna=False,
            **kwargs):
    launches.append((sorted(s.kind for s in stat_list), masked))
    return ('spec', [s.kind for s in stat_list])

  def run(pairs, leaves=None):
    return [{kind: (xl.DataArray(1.0), xl.DataArray(1.0))
             for kind in spec[1]} for spec, _ in pairs]

  monkeypatch.setattr(engine, 'build_fused_spec', build)
  monkeypatch.setattr(engine, 'run_fused_specs', run)
  rd = ['init_time', 'latitude', 'longitude']
  aggregation.Aggregator(reduce_dims=rd, masked=True).aggregate_statistics(
      stats)
  assert sorted(launches) == [
      (['AnomalyCovariance', 'SquaredError', 'SquaredTargetAnomaly'], True),
      (['SquaredPredictionAnomaly'], False)]
  launches.clear()
  aggregation.Aggregator(reduce_dims=rd).aggregate_statistics(stats)
  assert launches == [(['AnomalyCovariance', 'SquaredError',
                        'SquaredPredictionAnomaly', 'SquaredTargetAnomaly'],
                       False)]


def test_coordinate_views_and_assignment():
  """binning.py:135,181,190-199 rely on ``stat.latitude.latitude`` and on
  ``masks.coords[name] = labels``."""
  da = xl.DataArray(np.zeros((2, 3)), ('region', 'latitude'),
                    coords={'latitude': [10., 20., 30.]})
  lat = da.latitude
  assert lat.dims == ('latitude',)
  np.testing.assert_array_equal(lat.latitude.values, [10., 20., 30.])
  np.testing.assert_array_equal(da.coords['latitude'].coords['latitude'].values,
                                [10., 20., 30.])
  da.coords['region'] = np.array(['a', 'b'])
  assert list(da.coords) == ['latitude', 'region']
  assert da.coords['region'].dims == ('region',)
  assert list(da.sel(region='b').coords['region'].values.ravel()) == ['b']
  with pytest.raises(ValueError):
    da.coords['region'] = np.array(['a', 'b', 'c'])
  assert 'region' in da.coords and len(da.coords) == 2
  assert dict(da.coords).keys() == {'latitude', 'region'}


def test_outer_bin_masks_fold_into_the_cell_table():
  """Bin masks over outer dims (time units, level sets) need no per-point
  operand: the planner sorts the jobs into (kept cell, outer class) launch
  cells and maps the class sums to bins on the host."""
  p, t, c = _case(('init_time', 'lead_time', 'level', 'latitude', 'longitude'))
  p, t = p.drop_vars('mask'), t.drop_vars('mask')
  stat = LazyStatistic('SquaredError', p, t)
  sets = binning.ByTimeUnitSets({'a': [0], 'b': [0, 12]}, 'hour', 'lead_time')
  levels = binning.BySets({'low': [850]}, 'level', bin_dim_name='level_set',
                          add_set_complements=True)
  masks = [sets.create_bin_mask(stat), levels.create_bin_mask(stat)]
  names = [sets.bin_dim_name, levels.bin_dim_name]
  rd = ('init_time', 'lead_time', 'latitude', 'longitude')
  spec = engine.build_fused_spec([stat], rd, [], bin_masks=masks,
                                 bin_dim_names=names)
  assert spec.classes is None and spec.outer is not None
  assert spec.outer.bin_dims == names and spec.bin_order == tuple(names)
  # jobs: level (kept) x init x lead; lead hours {0, 12} -> 1 outer class for
  # lead (both in 'a' and 'b' differ: hour 0 in a+b, hour 12 in b only) x 2
  # level classes, all combined with the kept level cell
  assert np.all(np.diff(spec.cell) >= 0)
  assert spec.cell[-1] == spec.n_cells - 1 == len(spec.outer.dense_index) - 1
  ws, wsum = _interpret(spec)
  slot = _cabi.STAT_SLOT['SquaredError']
  got = spec.outer.to_bins(ws[:, slot].reshape(spec.n_cells))
  got_w = spec.outer.to_bins(
      wsum[:, _cabi.STAT_WCLASS[slot]].reshape(spec.n_cells))
  val = oracle.squared_error(p.values, t.values)
  sws, sw, out_dims = oracle.aggregate(
      val, p.dims, rd,
      bin_masks=[(masks[0].values, masks[0].dims),
                 (masks[1].values, masks[1].dims)])
  assert out_dims == ('level',) + tuple(names)
  np.testing.assert_allclose(got, sws, rtol=1e-6, equal_nan=True)
  np.testing.assert_allclose(got_w, sw, rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir('/root/reference/weatherbenchX'),
                    reason='the reference tree is only present in the build '
                           'container')
def test_state_algebra_equals_the_reference_class():
  """AggregationState sum / zero / outer join / mean_statistics /
  sum_along_dims / dot / map side by side with the reference's own class
  (aggregation.py:63-202), in a subprocess because importing the reference
  needs stand-in modules registered under the names xarray / jax."""
  import subprocess
  import sys
  script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                        'compare_state_algebra.py')
  proc = subprocess.run([sys.executable, script], capture_output=True,
                        text=True, timeout=300)
  assert proc.returncode == 0, proc.stdout + proc.stderr
  assert 'state algebra ok' in proc.stdout
