"""Host side of the categorical path on the GPU-less box.

The oracle's categorical functions against the reference's inline known
answers; the float32 threshold rule; and everything above the C ABI (lazy
handles, launch grouping, job / threshold tables, result labelling) with the
plans executed by the NumPy interpreter of tests/wbx_emulator.py.  The kernels
themselves are validated by tests/test_gpu_categorical.py and the ``cat/*``
cases of tests/test_reference_golden.py on a B200.
"""

import numpy as np
import pytest

import wbx_emulator
import wbx_oracle as oracle
from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import LazyBinarized
from weatherbenchx_b200.lazy import LazyCategoricalStatistic
from weatherbenchx_b200.lazy import threshold_f32
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import categorical
from weatherbenchx_b200.metrics import deterministic
from weatherbenchx_b200.metrics import wrappers

KINDS = ('TruePositives', 'FalsePositives', 'FalseNegatives', 'TrueNegatives')
DIMS = ('init_time', 'latitude', 'longitude')


def _da(values, name='rain'):
  n, ny, nx = values.shape
  return xl.DataArray(values, DIMS, name=name, coords={
      'init_time': np.arange(n), 'latitude': np.linspace(-90, 90, ny),
      'longitude': np.linspace(0, 360, nx, endpoint=False)})


# ---------------------------------------------------------------------------
# oracle pins (reference inline known answers)
# ---------------------------------------------------------------------------


def test_oracle_error_exceedance_known_answer():
  """metrics/metrics_test.py:1031-1049."""
  p = np.array([0, -1, 1, np.nan])
  t = np.zeros(4)
  got = oracle.error_exceedance(p, t, [0, 0.5, 1, np.nan])
  expected = np.array([[0, 0, 0, np.nan], [1, 1, 0, np.nan],
                       [1, 1, 0, np.nan], [np.nan] * 4])
  np.testing.assert_array_equal(got, expected)


def _oracle_metric(name, p, t):
  table = oracle.contingency_table(p, t)
  mean = {k: v.mean() for k, v in table.items()}   # skipna=False
  return oracle.categorical_metric(
      name, mean['TruePositives'], mean['FalsePositives'],
      mean['FalseNegatives'], mean['TrueNegatives'])


def test_oracle_far_and_csi_known_answers():
  """metrics/metrics_test.py:100-170 on binary fields."""
  zeros, ones = np.zeros((2, 4, 6), np.float32), np.ones((2, 4, 6), np.float32)
  half = zeros.copy()
  half[0] = 1
  nan = ones.copy()
  nan[0] = np.nan
  assert np.isnan(_oracle_metric('far', zeros, zeros))
  assert _oracle_metric('far', ones, ones) == 0
  assert _oracle_metric('far', ones, zeros) == 1
  assert _oracle_metric('far', ones, half) == 0.5
  assert np.isnan(_oracle_metric('far', zeros, nan))
  assert np.isnan(_oracle_metric('csi', zeros, zeros))
  assert _oracle_metric('csi', ones, ones) == 1
  assert _oracle_metric('csi', ones, zeros) == 0
  assert _oracle_metric('csi', ones, half) == 0.5
  assert np.isnan(_oracle_metric('csi', zeros, nan))


# ---------------------------------------------------------------------------
# float32 thresholds
# ---------------------------------------------------------------------------


def test_threshold_f32_reproduces_the_float64_comparison():
  """x > t (float32 field, float64 threshold: NumPy promotes the field) must
  equal x > threshold_f32(t) evaluated in float32, also AT the boundary."""
  rng = np.random.default_rng(0)
  thr = np.concatenate([
      [0.0, 0.1, 0.25, 1e-45, -0.1, 1e9, 3.4e38, 1e39, -1e39, np.inf, -np.inf],
      rng.normal(0, 3, 200), rng.random(200) * 1e-3])
  t32 = threshold_f32(thr)
  assert t32.dtype == np.float32
  for t64, t in zip(thr, t32):
    with np.errstate(over='ignore'):
      near = np.float32(t64)
    x = np.array([near, np.nextafter(near, np.float32(np.inf)),
                  np.nextafter(near, np.float32(-np.inf)), 0, -0.0, 1, -1,
                  np.inf, -np.inf], np.float32)
    np.testing.assert_array_equal(x > t, x.astype(np.float64) > t64,
                                  err_msg=repr(t64))
  assert np.isnan(threshold_f32([np.nan]))[0]
  # exact float32 numbers stay as they are
  np.testing.assert_array_equal(threshold_f32([0.25, 2.0, -8.5]),
                                np.array([0.25, 2.0, -8.5], np.float32))
  assert threshold_f32([0.1])[0] < np.float32(0.1)


# ---------------------------------------------------------------------------
# handles
# ---------------------------------------------------------------------------


def test_binarized_handle_metadata_follows_the_reference():
  x = _da(np.zeros((2, 4, 8), np.float32))
  b = wrappers.ContinuousToBinary('both', [0.5, 1], 'thr').transform_fn(x)
  assert isinstance(b, LazyBinarized) and b.is_lazy
  assert b.dims == DIMS + ('thr',) and b.shape == (2, 4, 8, 2)
  np.testing.assert_array_equal(b.coords['thr'].values, [0.5, 1.0])
  assert b.name == 'rain' and b.dtype == np.float32
  # a scalar threshold becomes a length-one dim (wrappers.py:248-252)
  one = wrappers.ContinuousToBinary('predictions', 0.5, 'thr').transform_fn(x)
  assert one.shape == (2, 4, 8, 1)
  # unique names (wrappers.py:260-265, 984-986)
  t = wrappers.ContinuousToBinary('both', [0.5, 1], 'thr')
  assert t.unique_name_suffix == 'thr=0.5,1'
  wrapped = wrappers.WrappedMetric(categorical.CSI(), [t]).statistics
  assert wrapped['TruePositives'].unique_name == 'TruePositives_both_thr=0.5,1'
  with pytest.raises(ValueError):
    wrappers.ContinuousToBinary(
        'both', xl.DataArray([0.5], dims=['thr']), 'thr')
  named = wrappers.ContinuousToBinary(
      'both', xl.DataArray([0.5, 2.0], dims=['thr']), 'thr',
      unique_name_suffix='fixed')
  assert named.unique_name_suffix == 'thr=fixed'
  handle = named.transform_fn(x)
  assert handle.shape == (2, 4, 8, 2) and 'thr' not in handle.coords
  with pytest.raises(NotImplementedError):
    wrappers.ContinuousToBinary(
        'both', xl.DataArray(np.zeros((2, 4)), dims=['thr', 'latitude']),
        'thr', unique_name_suffix='x').transform_fn(x)


def test_categorical_handle_dims_and_grouping():
  x, y = _da(np.zeros((2, 4, 8), np.float32)), _da(np.ones((2, 4, 8), np.float32))
  t = wrappers.ContinuousToBinary('both', [0.5, 1, 2], 'thr')
  bp, bt = t.transform_fn(x), t.transform_fn(y)
  tp = categorical.TruePositives().compute({'v': bp}, {'v': bt})['v']
  fn = categorical.FalseNegatives().compute({'v': bp}, {'v': bt})['v']
  assert isinstance(tp, LazyCategoricalStatistic) and tp.is_lazy
  assert tp.dims == DIMS + ('thr',) and tp.plan_dims == ('thr',) + DIMS
  assert tp.xform == _cabi.XF_CONTINGENCY
  assert tp.group_key() == fn.group_key()
  # other thresholds, other operands or an untransformed input: other launches
  other = wrappers.ContinuousToBinary('both', [0.5, 1, 3], 'thr')
  tp2 = categorical.TruePositives().compute(
      {'v': other.transform_fn(x)}, {'v': other.transform_fn(y)})['v']
  assert tp2.group_key() != tp.group_key()
  raw = categorical.TruePositives().compute({'v': x}, {'v': y})['v']
  assert raw.dims == DIMS and raw.threshold_dim is None
  assert raw.xform == (_cabi.XF_CONTINGENCY | _cabi.XF_PRED_NONZERO |
                       _cabi.XF_TARGET_NONZERO)
  half = categorical.TruePositives().compute({'v': bp}, {'v': y})['v']
  assert half.xform == _cabi.XF_CONTINGENCY | _cabi.XF_TARGET_NONZERO
  assert half.thr_target is None and half.thr_pred is not None
  se = deterministic.SquaredError().compute({'v': x}, {'v': y})['v']
  assert se.group_key()[:2] != raw.group_key()[:2]
  with pytest.raises(ValueError, match='Failed to compute'):
    metrics_base.compute_unique_statistics_for_all_metrics(
        {'tp': categorical.TruePositives()}, {'v': bp},
        {'v': other.transform_fn(y)})


def test_planner_threshold_tables():
  """Jobs enumerate (threshold, kept dims, reduced outer dims); each job
  carries the float32 threshold of its index."""
  rng = np.random.default_rng(1)
  p = _da(rng.random((3, 4, 8)).astype(np.float32))
  t = _da(rng.random((3, 4, 8)).astype(np.float32))
  thresholds = [0.1, 0.5, 0.9]
  tr = wrappers.ContinuousToBinary('both', thresholds, 'thr')
  stats = [cls().compute({'v': tr.transform_fn(p)}, {'v': tr.transform_fn(t)})
           ['v'] for cls in (categorical.TruePositives,
                             categorical.TrueNegatives)]
  spec = engine.build_fused_spec(stats, ['init_time', 'latitude', 'longitude'])
  assert spec.xform == _cabi.XF_CONTINGENCY
  assert spec.stat_mask == 0b1001 and spec.n_cells == 3
  # init_time is reduced and contiguous with the grid: it joins the slab
  assert (spec.ny, spec.nx) == (12, 8) and len(spec.pred) == 3
  np.testing.assert_array_equal(spec.cell, np.arange(3))
  np.testing.assert_array_equal(spec.thr_pred, threshold_f32(thresholds))
  np.testing.assert_array_equal(spec.thr_pred, spec.thr_target)
  assert spec.thr_pred.dtype == np.float32
  # the three thresholds read the same slab
  assert len(set(spec.pred.tolist())) == 1
  assert spec.kept == ['thr'] and spec.kept_order == ('thr',)
  # keeping init_time: jobs = (threshold, init_time); the result has the
  # reference's dim order (init_time, thr)
  spec = engine.build_fused_spec(stats, ['latitude', 'longitude'])
  assert (spec.ny, spec.nx) == (4, 8) and len(spec.pred) == 9
  assert spec.kept == ['thr', 'init_time'] and spec.n_cells == 9
  assert spec.kept_order == ('init_time', 'thr')
  np.testing.assert_array_equal(
      spec.thr_pred, np.repeat(threshold_f32(thresholds), 3))
  np.testing.assert_array_equal(spec.pred[:3], spec.pred[3:6])
  # region bins need the class-map kernel, which has no categorical variant
  land = xl.DataArray(np.ones((2, 4, 8), bool), ('region',) + DIMS[1:])
  with pytest.raises(engine.FastPathUnavailable):
    engine.build_fused_spec(stats, ['init_time', 'latitude', 'longitude'],
                            bin_masks=[land], bin_dim_names=['region'])


# ---------------------------------------------------------------------------
# class surface with interpreted plans
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna'])
def test_table_metrics_with_interpreted_plans(mode, monkeypatch):
  wbx_emulator.installed(monkeypatch)
  launches = []
  real = _cabi.DetPlan

  class Counting(real):

    def __init__(self, ctx, **desc):
      launches.append(desc)
      super().__init__(ctx, **desc)

  monkeypatch.setattr(_cabi, 'DetPlan', Counting)
  engine.clear_plan_cache()
  rng = np.random.default_rng(4)
  shape = (3, 6, 8)
  p = (rng.gamma(1.0, 1.0, shape) * (rng.random(shape) < 0.7)).astype(np.float32)
  t = (rng.gamma(1.0, 1.0, shape) * (rng.random(shape) < 0.7)).astype(np.float32)
  p[rng.random(shape) < 0.1] = np.float32(0.1)
  P, T = _da(p), _da(t)
  if mode != 'propagate':
    holes = rng.random(shape) < 0.1
    t = np.where(holes, np.nan, t).astype(np.float32)
    T = _da(t).assign_coords(mask=xl.DataArray(~holes, DIMS))
  thresholds = [0.0, 0.1, 1.0]
  both = [wrappers.ContinuousToBinary('both', thresholds, 'threshold')]
  names = ('csi', 'accuracy', 'recall', 'far', 'precision', 'f1',
           'frequency_bias', 'hss', 'ets', 'sedi')
  classes = (categorical.CSI, categorical.Accuracy, categorical.Recall,
             categorical.FalseAlarmRate, categorical.Precision,
             categorical.F1Score, categorical.FrequencyBias, categorical.HSS,
             categorical.ETS, categorical.SEDI)
  metrics = {n: wrappers.WrappedMetric(c(), both)
             for n, c in zip(names, classes)}
  metrics['rmse'] = deterministic.RMSE()
  aggregator = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()], masked=mode == 'masked',
      skipna=mode == 'skipna')
  values = aggregation.compute_metric_values_for_single_chunk(
      metrics, aggregator, {'rain': P}, {'rain': T})
  # the contingency table of all thresholds is ONE plan, RMSE another
  assert len(launches) == 2
  assert sorted(d.get('xform', 0) for d in launches) == [0, _cabi.XF_CONTINGENCY]
  w = oracle.grid_area_weights(np.linspace(-90, 90, shape[1]))
  table = oracle.contingency_table(oracle.binarize_thresholds(p, thresholds),
                                   oracle.binarize_thresholds(t, thresholds))
  means = {}
  for kind in KINDS:
    sws, sw, dims = oracle.aggregate(
        table[kind], DIMS + ('threshold',), ['latitude', 'longitude'],
        weights=[(w, ('latitude',))],
        mask=~np.isnan(t) if mode == 'masked' else None,
        mask_dims=DIMS if mode == 'masked' else None, masked=mode == 'masked',
        skipna=mode == 'skipna')
    assert tuple(dims) == ('init_time', 'threshold')
    means[kind] = oracle.mean_statistic(sws, sw)
  for name in names:
    got = values[f'{name}.rain']
    assert got.dims == ('init_time', 'threshold')
    np.testing.assert_allclose(
        got.values,
        oracle.categorical_metric(
            name, means['TruePositives'], means['FalsePositives'],
            means['FalseNegatives'], means['TrueNegatives']),
        rtol=1e-9, err_msg=name)


def test_known_answers_with_interpreted_plans(monkeypatch):
  """metrics/metrics_test.py:100-170 through the class surface (binary inputs,
  no threshold transform: the non-zero test of `.astype(bool)`)."""
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  zeros = _da(np.zeros((2, 4, 8), np.float32))
  ones = _da(np.ones((2, 4, 8), np.float32))
  half_values = np.zeros((2, 4, 8), np.float32)
  half_values[0] = 1
  half = _da(half_values)
  nan_values = np.ones((2, 4, 8), np.float32)
  nan_values[0] = np.nan
  nan = _da(nan_values)

  def value(name, metric, p, t):
    out = aggregation.compute_metric_values_for_single_chunk(
        {name: metric}, aggregation.Aggregator(reduce_dims=list(DIMS)),
        {'rain': p}, {'rain': t})
    return float(out[f'{name}.rain'].values)

  far, csi = categorical.FalseAlarmRate(), categorical.CSI()
  assert np.isnan(value('far', far, zeros, zeros))
  assert value('far', far, ones, ones) == 0
  assert value('far', far, ones, zeros) == 1
  assert value('far', far, ones, half) == 0.5
  assert np.isnan(value('far', far, zeros, nan))
  assert np.isnan(value('csi', csi, zeros, zeros))
  assert value('csi', csi, ones, ones) == 1
  assert value('csi', csi, ones, zeros) == 0
  assert value('csi', csi, ones, half) == 0.5
  assert np.isnan(value('csi', csi, zeros, nan))


def test_error_exceedance_with_interpreted_plans(monkeypatch):
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  rng = np.random.default_rng(6)
  p = rng.normal(0, 1, (3, 4, 8)).astype(np.float32)
  t = rng.normal(0, 1, (3, 4, 8)).astype(np.float32)
  t[0, 1, 2] = np.nan
  thresholds = xl.Dataset({'rain': xl.DataArray(
      [0.1, 1.0, np.nan], dims=['level_of_error'],
      coords={'level_of_error': ['small', 'large', 'undefined']})})
  stat = deterministic.ErrorExceedance(thresholds)
  aggregator = aggregation.Aggregator(reduce_dims=['latitude', 'longitude'],
                                      skipna=True)
  state = aggregator.aggregate_statistics(
      {'ErrorExceedance': stat.compute({'rain': _da(p)}, {'rain': _da(t)})})
  got = state.mean_statistics()['ErrorExceedance']['rain']
  assert got.dims == ('init_time', 'level_of_error')
  assert list(got.coords['level_of_error'].values) == ['small', 'large',
                                                       'undefined']
  field = oracle.error_exceedance(p, t, [0.1, 1.0, np.nan])
  with np.errstate(invalid='ignore'):
    expected = np.nanmean(field, axis=(1, 2))
  np.testing.assert_allclose(got.values[:, :2], expected[:, :2], rtol=1e-12)
  assert np.isnan(got.values[:, 2]).all()   # 0 / 0: nothing valid


# ---------------------------------------------------------------------------
# SEEPS: constructor, names, host-side parameters
# ---------------------------------------------------------------------------


def test_seeps_names_parameters_and_p1_cache():
  from weatherbenchx_b200.metrics import categorical as cat
  variables = ['total_precipitation_6hr', 'total_precipitation_24hr']
  seeps = cat.SEEPS(variables=variables, climatology={})
  assert seeps.unique_name == (
      'SEEPS_total_precipitation_6hr_total_precipitation_24hr_'
      'dry_threshold_mm_0.25_0.25_min_p1_0.1_0.1_max_p1_0.85_0.85')
  seeps = cat.SEEPS(variables=variables[:1], climatology={},
                    dry_threshold_mm=[0.1], min_p1=[0.2], max_p1=[0.7])
  assert seeps.unique_name == (
      'SEEPS_total_precipitation_6hr_dry_threshold_mm_0.1_min_p1_0.2_'
      'max_p1_0.7')
  with pytest.raises(AssertionError):
    cat.SEEPS(variables=variables, climatology={}, min_p1=[0.1])
  # p1 = nanmean over (hour, dayofyear) in the input dtype, memoised per array
  rng = np.random.default_rng(0)
  values = rng.random((4, 6, 5, 3)).astype(np.float32)
  values[1, 2] = np.nan                      # a missing day is skipped
  values[:, :, 4, 2] = np.nan                # an all-NaN point stays NaN
  frac = xl.DataArray(values, ('hour', 'dayofyear', 'longitude', 'latitude'))
  p1 = cat._dry_fraction_mean(frac)
  assert p1.dims == ('longitude', 'latitude') and p1.dtype == np.float32
  np.testing.assert_array_equal(
      p1.values, oracle.seeps_p1(values, (0, 1)))
  assert np.isnan(p1.values[4, 2]) and not np.isnan(p1.values[0, 0])
  assert cat._dry_fraction_mean(frac) is p1
  with pytest.raises(ValueError):
    cat._dry_fraction_mean(xl.DataArray(values[0], ('dayofyear', 'longitude',
                                                    'latitude')))


def test_seeps_mask_combination(monkeypatch):
  """categorical.py:296-304: the p1 range mask, combined with the mask of the
  predictions OR the targets; both is an error."""
  from weatherbenchx_b200.metrics import categorical as cat
  wbx_emulator.installed(monkeypatch)
  lat, lon = np.linspace(-80, 80, 4), np.arange(8) * 45.0
  init = np.datetime64('2021-03-01T00', 'ns') + np.arange(2) * np.timedelta64(
      1, 'D')
  lead = (np.arange(2) * np.timedelta64(12, 'h')).astype('timedelta64[ns]')
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': init, 'lead_time': lead, 'latitude': lat,
            'longitude': lon}
  rng = np.random.default_rng(2)
  field = lambda: xl.DataArray(  # noqa: E731
      rng.random((2, 2, 4, 8)).astype(np.float32), dims, coords=coords,
      name='rain')
  cdims = ('hour', 'dayofyear', 'latitude', 'longitude')
  ccoords = {'hour': [0, 12], 'dayofyear': np.arange(1, 367), 'latitude': lat,
             'longitude': lon}
  frac = np.broadcast_to(np.linspace(0, 1, 32, dtype=np.float32).reshape(4, 8),
                         (2, 366, 4, 8))
  clim = xl.Dataset({
      'rain_seeps_dry_fraction': xl.DataArray(frac, cdims, coords=ccoords),
      'rain_seeps_threshold': xl.DataArray(
          np.full((2, 366, 4, 8), 0.5, np.float32), cdims, coords=ccoords)})
  seeps = cat.SEEPS(variables=['rain'], climatology=clim, dry_threshold_mm=100)
  in_range = (frac[0, 0] >= np.float32(0.1)) & (frac[0, 0] <= np.float32(0.85))
  p, t = field(), field()
  stat = seeps.compute({'rain': p}, {'rain': t})['rain']
  assert stat.dims == dims and stat.is_lazy
  assert stat.coords['mask'].dims == ('latitude', 'longitude')
  np.testing.assert_array_equal(stat.coords['mask'].values, in_range)
  holes = rng.random((2, 2, 4, 8)) < 0.3
  t_masked = t.assign_coords(mask=xl.DataArray(~holes, dims))
  stat = seeps.compute({'rain': p}, {'rain': t_masked})['rain']
  assert stat.coords['mask'].dims == dims
  np.testing.assert_array_equal(stat.coords['mask'].values, ~holes & in_range)
  p_masked = p.assign_coords(mask=xl.DataArray(~holes, dims))
  stat = seeps.compute({'rain': p_masked}, {'rain': t})['rain']
  np.testing.assert_array_equal(stat.coords['mask'].values, ~holes & in_range)
  with pytest.raises(ValueError, match='Both predictions and targets'):
    seeps.compute({'rain': p_masked}, {'rain': t_masked})


# ---------------------------------------------------------------------------
# chunk driver: categorical metrics + SEEPS through run_pipeline
# ---------------------------------------------------------------------------


def test_pipeline_with_categorical_metrics_and_seeps(tmp_path, monkeypatch):
  """Chunked evaluation (states combined by the zero-filled outer join along
  init / lead / threshold) == one monolithic call; the files round-trip the
  threshold dimension.  Real Aggregator, plans interpreted by the emulator."""
  import test_pipeline as tp
  from weatherbenchx_b200 import io_netcdf, pipeline, time_chunks
  from weatherbenchx_b200.data_loaders import array_loaders
  from weatherbenchx_b200.metrics import categorical as cat
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  preds, tgts = tp._datasets(('rain',))
  cdims = ('dayofyear', 'hour', 'latitude', 'longitude')
  ccoords = {'dayofyear': np.arange(1, 367), 'hour': np.array([0, 12]),
             'latitude': tp.LAT, 'longitude': tp.LON}
  rng = np.random.default_rng(9)
  clim = xl.Dataset({
      'rain_seeps_threshold': xl.DataArray(
          rng.uniform(0.3, 1.5, (366, 2, tp.NLAT, tp.NLON)).astype(np.float32),
          cdims, coords=ccoords),
      'rain_seeps_dry_fraction': xl.DataArray(
          np.broadcast_to(rng.uniform(0.05, 0.95, (tp.NLAT, tp.NLON)),
                          (366, 2, tp.NLAT, tp.NLON)).astype(np.float32),
          cdims, coords=ccoords)})
  both = [wrappers.ContinuousToBinary('both', [-0.5, 0.0, 0.75], 'threshold')]
  metrics = {
      'csi': wrappers.WrappedMetric(categorical.CSI(), both),
      'ets': wrappers.WrappedMetric(categorical.ETS(), both),
      'exceed': deterministic.ErrorExceedance([0.1, 0.3]),
      'seeps': cat.SEEPS(['rain'], clim, dry_threshold_mm=-200.0),
      'rmse': deterministic.RMSE()}
  aggregators = {
      'time_mean': aggregation.Aggregator(
          reduce_dims=['init_time', 'latitude', 'longitude'],
          weigh_by=[weighting.GridAreaWeighting()], masked=True),
      'per_init': aggregation.Aggregator(
          reduce_dims=['latitude', 'longitude'], masked=True)}
  out_path = str(tmp_path / 'metrics.nc')
  state_path = str(tmp_path / 'state.nc')

  def run(init_chunk, lead_chunk, **kw):
    times = time_chunks.TimeChunks(tp.INIT, tp.LEAD,
                                   init_time_chunk_size=init_chunk,
                                   lead_time_chunk_size=lead_chunk)
    return pipeline.run_pipeline(
        times, array_loaders.PredictionsFromArrays(preds),
        array_loaders.TargetsFromArrays(tgts), metrics, aggregators,
        prefetch=0, **kw)

  mono = run(None, None, require_output=False)
  chunked = run(2, 3, out_path=out_path, aggregation_state_out_path=state_path)
  names = {'csi.rain', 'ets.rain', 'exceed.rain', 'seeps.rain', 'rmse.rain'}
  for agg_name in aggregators:
    values, expected = chunked[agg_name][1], mono[agg_name][1]
    assert set(values) == set(expected) == names
    for k in names:
      assert values[k].dims == expected[k].dims, k
      np.testing.assert_allclose(values[k].values, expected[k].values,
                                 rtol=1e-12, equal_nan=True, err_msg=k)
    assert values['csi.rain'].dims[-1] == 'threshold'
    assert values['exceed.rain'].dims[-1] == 'error_exceedance_thresholds'
    assert np.isfinite(values['seeps.rain'].values).all()
    stored = io_netcdf.open_dataset(str(tmp_path / f'metrics_{agg_name}.nc'))
    for k in names:
      np.testing.assert_array_equal(stored[k].values, values[k].values)
    np.testing.assert_array_equal(
        stored['csi.rain'].coords['threshold'].values, [-0.5, 0.0, 0.75])
    state = pipeline.load_aggregation_state(
        str(tmp_path / f'state_{agg_name}.nc'))
    again = state.metric_values(metrics)
    for k in names:
      np.testing.assert_allclose(again[k].values, values[k].values,
                                 rtol=1e-15, equal_nan=True)
  assert chunked['per_init'][1]['csi.rain'].dims == (
      'init_time', 'lead_time', 'threshold')


def test_ensemble_error_exceedance_with_interpreted_plans(monkeypatch):
  """probabilistic.py:836-861 without NaN members: ONE fused launch over
  reduce_dims + [ensemble dim], divided by the member count."""
  from weatherbenchx_b200.metrics import probabilistic
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  launches = []
  real = _cabi.DetPlan

  class Counting(real):

    def __init__(self, ctx, **desc):
      launches.append(desc)
      super().__init__(ctx, **desc)

  monkeypatch.setattr(_cabi, 'DetPlan', Counting)
  rng = np.random.default_rng(12)
  y = rng.normal(0, 1, (3, 6, 8)).astype(np.float32)
  x = (y[:, None] + rng.normal(0, 1, (3, 5, 6, 8))).astype(np.float32)
  dims = ('init_time', 'number', 'latitude', 'longitude')
  coords = {'init_time': np.arange(3), 'number': np.arange(5),
            'latitude': np.linspace(-75, 75, 6), 'longitude': np.arange(8) * 45.0}
  X = xl.DataArray(x, dims, coords=coords, name='t')
  Y = xl.DataArray(y, ('init_time', 'latitude', 'longitude'),
                   coords={k: v for k, v in coords.items() if k != 'number'},
                   name='t')
  thresholds = [0.5, 1.5]
  metric = probabilistic.EnsembleErrorExceedance(thresholds)
  assert metric.unique_name == 'EnsembleErrorExceedance'
  stat = metric.compute({'t': X}, {'t': Y})['t']
  assert stat.dims == ('init_time', 'latitude', 'longitude',
                       'error_exceedance_thresholds')
  aggregator = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  values = aggregation.compute_metric_values_for_single_chunk(
      {'ee': metric}, aggregator, {'t': X}, {'t': Y})['ee.t']
  assert len(launches) == 1 and launches[0]['xform'] == _cabi.XF_ERROR_EXCEEDANCE
  field = oracle.error_exceedance(x, y[:, None], thresholds).mean(axis=1)
  w = oracle.grid_area_weights(coords['latitude'])
  sws, sw, out_dims = oracle.aggregate(
      field, stat.dims, ['latitude', 'longitude'],
      weights=[(w, ('latitude',))])
  assert values.dims == tuple(out_dims)
  np.testing.assert_allclose(values.values, sws / sw, rtol=1e-12)


@pytest.mark.parametrize('masked', [False, True])
def test_relative_intensity_with_interpreted_plans(masked, monkeypatch):
  """deterministic.py:28-88: the spatial means are fused-kernel reductions."""
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  rng = np.random.default_rng(21)
  shape = (3, 2, 6, 8)
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': np.arange(3), 'lead_time': np.arange(2),
            'latitude': np.linspace(-75, 75, 6), 'longitude': np.arange(8) * 45.0}
  p = rng.gamma(1.0, 2.0, shape).astype(np.float32)
  t = rng.gamma(1.0, 2.0, shape).astype(np.float32)
  P = xl.DataArray(p, dims, coords=coords, name='rain')
  T = xl.DataArray(t, dims, coords=coords, name='rain')
  mask = None
  if masked:
    mask = rng.random(shape) < 0.7
    mask[1, 0] = False                       # a slab with no valid point
    t = np.where(mask, t, np.nan).astype(np.float32)
    T = xl.DataArray(t, dims, coords=coords, name='rain').assign_coords(
        mask=xl.DataArray(mask, dims))
  stat = deterministic.RelativeIntensity().compute({'rain': P}, {'rain': T})
  got = stat['rain']
  assert got.dims == ('init_time', 'lead_time') and got.dtype == np.float32
  want, want_mask = oracle.relative_intensity(p, t, (2, 3), mask)
  np.testing.assert_allclose(got.values, want, rtol=1e-4, atol=2e-6)
  if masked:
    np.testing.assert_array_equal(got.coords['mask'].values, want_mask)
    assert got.values[1, 0] == 0 and want_mask[1, 0] == 0
  else:
    assert 'mask' not in got.coords
  # NaN propagates through the unmasked means (skipna=False)
  p_nan = p.copy()
  p_nan[2, 1, 3, 3] = np.nan
  P_nan = xl.DataArray(p_nan, dims, coords=coords, name='rain')
  got = deterministic.RelativeIntensity().compute({'rain': P_nan},
                                                  {'rain': T})['rain']
  want, _ = oracle.relative_intensity(p_nan, t, (2, 3), mask)
  np.testing.assert_array_equal(np.isnan(got.values), np.isnan(want))
  assert masked or np.isnan(got.values[2, 1])
  with pytest.raises(ValueError, match='Failed to compute'):
    metrics_base.compute_unique_statistics_for_all_metrics(
        {'ri': deterministic.RelativeIntensity(('x', 'y'))}, {'rain': P},
        {'rain': T})


@pytest.mark.parametrize('reduce_dims', [
    ['init_time', 'lead_time', 'latitude', 'longitude'],
    ['init_time', 'latitude', 'longitude'],
    ['latitude', 'longitude']])
@pytest.mark.parametrize('block_bytes', [1, 6 * 8 * 8 * 2, 10 ** 9])
def test_l2_blocked_job_order_gives_the_same_sums(reduce_dims, block_bytes,
                                                  monkeypatch):
  """engine.XF_L2_BLOCK_BYTES (an experiment knob, off by default): blocks of
  slabs are swept for all thresholds; a result cell is split into one launch
  cell per block and folded on the host.  Same numbers, cells still dense and
  non-decreasing, same slabs."""
  wbx_emulator.installed(monkeypatch)
  rng = np.random.default_rng(30)
  shape = (5, 3, 6, 8)
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': np.arange(5), 'lead_time': np.arange(3),
            'latitude': np.linspace(-75, 75, 6), 'longitude': np.arange(8) * 45.0}
  p = rng.gamma(1.0, 1.0, shape).astype(np.float32)
  t = rng.gamma(1.0, 1.0, shape).astype(np.float32)
  t[rng.random(shape) < 0.05] = np.nan
  # lead_time first in memory order would make init_time part of the slab; keep
  # (init, lead) as outer dims by making the arrays non-contiguous there
  P = xl.DataArray(p, dims, coords=coords, name='v')
  T = xl.DataArray(t, dims, coords=coords, name='v')
  both = [wrappers.ContinuousToBinary('both', [0.2, 0.5, 1.0, 2.0], 'thr')]
  metrics = {'acc': wrappers.WrappedMetric(categorical.Accuracy(), both)}
  aggregator = aggregation.Aggregator(
      reduce_dims=reduce_dims, weigh_by=[weighting.GridAreaWeighting()],
      skipna=True)

  def run():
    engine.clear_plan_cache()
    statistics = metrics_base.compute_unique_statistics_for_all_metrics(
        metrics, {'v': P}, {'v': T})
    return aggregator.aggregate_statistics(statistics)

  base = run()
  monkeypatch.setattr(engine, 'XF_L2_BLOCK_BYTES', block_bytes)
  stats = [c().compute({'v': both[0].transform_fn(P)},
                       {'v': both[0].transform_fn(T)})['v']
           for c in (categorical.TruePositives, categorical.TrueNegatives)]
  spec = engine.build_fused_spec(
      stats, reduce_dims, [weighting.GridAreaWeighting().weights(stats[0])],
      skipna=True)
  if spec.cell_fold is not None:
    steps = np.diff(spec.cell)
    assert spec.cell[0] == 0 and set(steps.tolist()) <= {0, 1}
    assert spec.cell[-1] == len(spec.cell_fold) - 1 == spec.n_cells - 1
    assert sorted(set(spec.cell_fold.tolist())) == list(
        range(int(np.prod(spec.kept_shape))))
  blocked = run()
  for name, per_var in base.sum_weighted_statistics.items():
    for var, da in per_var.items():
      other = blocked.sum_weighted_statistics[name][var]
      assert other.dims == da.dims
      np.testing.assert_allclose(other.values, da.values, rtol=1e-12)
      np.testing.assert_allclose(blocked.sum_weights[name][var].values,
                                 base.sum_weights[name][var].values,
                                 rtol=1e-12)


def test_crps_distance_to_an_ensemble_of_targets(monkeypatch):
  """probabilistic.py:135-145,199-204,691-782 with interpreted plans: the skill
  is the mean over target members of the usual skill (launches merged into
  one), the target spread is the spread kernel on the targets."""
  from weatherbenchx_b200.metrics import probabilistic
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  launches = []
  real = _cabi.CrpsPlan

  class Counting(real):

    def __init__(self, ctx, **desc):
      launches.append(desc)
      super().__init__(ctx, **desc)

  monkeypatch.setattr(_cabi, 'CrpsPlan', Counting)
  rng = np.random.default_rng(40)
  n_init, m_pred, m_tgt, ny, nx = 3, 6, 4, 5, 8
  truth = rng.normal(0, 1, (n_init, 1, ny, nx))
  x = (truth + rng.normal(0, 1, (n_init, m_pred, ny, nx))).astype(np.float32)
  y = (truth + rng.normal(0, 0.5, (n_init, m_tgt, ny, nx))).astype(np.float32)
  dims = ('init_time', 'number', 'latitude', 'longitude')
  grid = {'init_time': np.arange(n_init),
          'latitude': np.linspace(-60, 60, ny), 'longitude': np.arange(nx) * 45.0}
  X = xl.DataArray(x, dims, coords=dict(grid, number=np.arange(m_pred)),
                   name='t')
  Y = xl.DataArray(y, dims, coords=dict(grid, number=np.arange(m_tgt)),
                   name='t')
  metrics = {'distance': probabilistic.CRPSEnsembleDistance()}
  aggregator = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, {'t': X}, {'t': Y})
  assert set(statistics) == {'CRPSSkill_number',
                             'CRPSSpread_number_fair_predictions',
                             'CRPSSpread_number_fair_targets'}
  state = aggregator.aggregate_statistics(statistics)
  values = state.metric_values(metrics)['distance.t']
  # skill of the 4 target members in ONE merged launch, + the two spreads
  assert len(launches) == 3
  assert sorted(len(d['ens']) for d in launches) == [3, 3, 12]
  w = oracle.grid_area_weights(grid['latitude'])
  rd = ['latitude', 'longitude']
  out_dims = ('init_time', 'latitude', 'longitude')

  def mean(field):
    sws, sw, _ = oracle.aggregate(field, out_dims, rd,
                                  weights=[(w, ('latitude',))])
    return sws / sw

  skill = np.abs(x[:, :, None] - y[:, None, :]).mean(axis=(1, 2))
  expected = (mean(skill) - 0.5 * mean(oracle.crps_spread(x, 1, fair=True))
              - 0.5 * mean(oracle.crps_spread(y, 1, fair=True)))
  np.testing.assert_allclose(values.values, expected, rtol=1e-5)
  with pytest.raises(ValueError, match='Failed to compute'):
    metrics_base.compute_unique_statistics_for_all_metrics(
        {'s': probabilistic.CRPSSkill(skipna_ensemble=True)}, {'t': X},
        {'t': Y})


@pytest.mark.parametrize('ensemble_size,use_sort,fair', [
    (4, False, True), (5, True, True), (4, True, False), (5, False, False)])
def test_crps_ensemble_distance_reference_identities(ensemble_size, use_sort,
                                                     fair, monkeypatch):
  """metrics/metrics_test.py:662-752 with interpreted plans: predictions and
  targets from the same distribution give a distance near zero (fair), and
  targets whose members are all equal give the standard CRPS."""
  import wbx_test_utils as utils
  from weatherbenchx_b200.metrics import probabilistic
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  # (the 3-d variable of the reference's fixture stores `level` behind the grid
  # dims: that layout goes through the generic kernel and is a GPU test)
  kw = dict(time_start='2020-01-01T00', time_stop='2020-01-03T00', random=True,
            variables_3d=())
  targets = utils.to_f32(utils.mock_prediction_data(
      ensemble_size=ensemble_size + 1, seed=0, lead_stop=2, **kw))
  predictions = utils.to_f32(utils.mock_prediction_data(
      ensemble_size=ensemble_size, seed=1, lead_stop=2, **kw))
  # C-contiguous member-last arrays (the fixture's lead dim is a broadcast
  # view): the layouts the host-space CRPS plan streams without staging
  contiguous = lambda d: {k: v._replace(data=np.ascontiguousarray(v.values))  # noqa: E731
                          for k, v in d.items()}
  targets, predictions = contiguous(targets), contiguous(predictions)
  n_target = ensemble_size + 1
  no_spread = {}
  for k, v in targets.items():
    first = v.isel(realization=0, drop=True)
    stacked = np.ascontiguousarray(np.broadcast_to(
        first.values[..., None], first.shape + (n_target,)))
    no_spread[k] = xl.DataArray(
        stacked, first.dims + ('realization',),
        coords=dict({d: first.coords[d] for d in first.dims},
                    realization=np.arange(n_target)), name=k)
  distance = {'crps': probabilistic.CRPSEnsembleDistance(
      ensemble_dim='realization', use_sort=use_sort, fair=fair)}
  standard = {'crps': probabilistic.CRPSEnsemble(
      ensemble_dim='realization', use_sort=use_sort, fair=fair)}
  reduce_dims = ['latitude', 'longitude', 'time']
  aggregator = aggregation.Aggregator(reduce_dims=reduce_dims)

  def compute(metrics, p, t):
    return aggregation.compute_metric_values_for_single_chunk(
        metrics, aggregator, p, t)

  vs_targets = compute(distance, predictions, targets)
  vs_no_spread = compute(distance, predictions, no_spread)
  vs_no_ensemble = compute(standard, predictions, no_spread)
  sizes = predictions['2m_temperature'].sizes
  stderr = 1 / np.sqrt(np.prod([ensemble_size * sizes[d]
                                for d in reduce_dims]))
  for v in ('2m_temperature',):
    if fair:
      np.testing.assert_allclose(vs_targets[f'crps.{v}'].values, 0,
                                 atol=5 * stderr)
    a, b = vs_no_spread[f'crps.{v}'], vs_no_ensemble[f'crps.{v}']
    np.testing.assert_allclose(a.values, b.transpose(*a.dims).values,
                               rtol=1e-5, atol=5 * stderr)
    # all target members equal: their spread is exactly zero
    assert a.shape == b.shape


def _relative_intensity(predictions, targets, mask=None, dims=('latitude',
                                                               'longitude')):
  """The statistic through the class surface (fused reductions interpreted)
  and through the oracle, for the reference's inline cases."""
  coords = {d: np.arange(n) for d, n in zip(dims, np.shape(predictions))}
  P = xl.DataArray(np.asarray(predictions, np.float32), dims, coords=coords,
                   name='var')
  T = xl.DataArray(np.asarray(targets, np.float32), dims, coords=coords,
                   name='var')
  if mask is not None:
    T = T.assign_coords(mask=xl.DataArray(np.asarray(mask), dims))
  stat = deterministic.RelativeIntensity(
      spatial_dims=['latitude', 'longitude']).compute({'var': P}, {'var': T})
  want, want_mask = oracle.relative_intensity(
      np.asarray(predictions, np.float32), np.asarray(targets, np.float32),
      (len(dims) - 2, len(dims) - 1),
      None if mask is None else np.asarray(mask))
  got = stat['var']
  np.testing.assert_allclose(got.values, want, atol=1e-5, equal_nan=True)
  if mask is not None:
    np.testing.assert_array_equal(got.coords['mask'].values, want_mask)
  return got


def test_relative_intensity_reference_known_answers(monkeypatch):
  """metrics/deterministic_test.py:27-218 (integer 0/1 masks as there)."""
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  nan = np.nan
  # regular: mean 25 against 10 -> |2.5 - 1|
  got = _relative_intensity([[10, 20], [30, 40]], [[10, 10], [10, 10]])
  np.testing.assert_allclose(got.values, 1.5, atol=1e-5)
  assert 'mask' not in got.coords
  # NaNs that are masked out
  got = _relative_intensity([[10, nan], [30, 40]], [[10, nan], [10, 10]],
                            [[1, 0], [1, 1]])
  np.testing.assert_allclose(got.values, abs((80 / 3) / 10 - 1), atol=1e-5)
  assert got.coords['mask'].item() == 1
  # a kept time dim; the second time is masked out completely
  got = _relative_intensity(
      [[[10, 20], [30, 40]], [[100, 200], [300, 400]]],
      [[[10, 10], [10, 10]], [[10, 10], [10, 10]]],
      [[[1, 1], [1, 1]], [[0, 0], [0, 0]]],
      dims=('time', 'latitude', 'longitude'))
  np.testing.assert_allclose(got.values, [1.5, 0.0], atol=1e-5)
  np.testing.assert_array_equal(got.coords['mask'].values, [1, 0])
  # everything NaN and masked
  got = _relative_intensity([[nan, nan], [nan, nan]], [[nan, nan], [nan, nan]],
                            [[0, 0], [0, 0]])
  np.testing.assert_allclose(got.values, 0)
  assert got.coords['mask'].item() == 0
  # a NaN inside the valid region propagates (skipna=False)
  got = _relative_intensity([[10, nan], [30, 40]], [[10, 10], [10, 10]],
                            [[1, 1], [1, 1]])
  assert np.isnan(got.values) and got.coords['mask'].item() == 1
  # only mask == 1 counts as valid
  got = _relative_intensity([[10, 20], [30, 40]], [[10, 10], [10, 10]],
                            [[1, 2], [1, 1]])
  np.testing.assert_allclose(got.values, abs((80 / 3) / 10 - 1), atol=1e-5)


def test_repeated_categorical_requests_reuse_their_plan(monkeypatch):
  """The steady state of an evaluation loop: same arrays, same transform ->
  the planner returns the previous spec (payload identities, incl. the
  threshold coordinate) and the library plan is reused."""
  wbx_emulator.installed(monkeypatch)
  engine.clear_plan_cache()
  built = []
  real_build = engine._build_fused_spec

  def counting(*args, **kwargs):
    built.append(1)
    return real_build(*args, **kwargs)

  monkeypatch.setattr(engine, '_build_fused_spec', counting)
  plans = []
  real_plan = _cabi.DetPlan

  class Counting(real_plan):

    def __init__(self, ctx, **desc):
      plans.append(1)
      super().__init__(ctx, **desc)

  monkeypatch.setattr(_cabi, 'DetPlan', Counting)
  rng = np.random.default_rng(50)
  P = _da(rng.random((3, 6, 8)).astype(np.float32))
  T = _da(rng.random((3, 6, 8)).astype(np.float32))
  both = [wrappers.ContinuousToBinary('both', [0.25, 0.5], 'threshold')]
  metrics = {'csi': wrappers.WrappedMetric(categorical.CSI(), both),
             'exceed': deterministic.ErrorExceedance([0.1, 0.2])}
  aggregator = aggregation.Aggregator(
      reduce_dims=['latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  first = aggregation.compute_metric_values_for_single_chunk(
      metrics, aggregator, {'rain': P}, {'rain': T})
  n_built, n_plans = len(built), len(plans)
  assert n_built == 2 and n_plans == 2      # contingency table + exceedance
  for _ in range(3):
    again = aggregation.compute_metric_values_for_single_chunk(
        metrics, aggregator, {'rain': P}, {'rain': T})
  assert len(built) == n_built and len(plans) == n_plans
  for k in first:
    np.testing.assert_array_equal(again[k].values, first[k].values)
  # other thresholds on the same arrays are another plan
  other = {'csi': wrappers.WrappedMetric(categorical.CSI(), [
      wrappers.ContinuousToBinary('both', [0.25, 0.75], 'threshold')])}
  aggregation.compute_metric_values_for_single_chunk(
      other, aggregator, {'rain': P}, {'rain': T})
  assert len(built) == n_built + 1 and len(plans) == n_plans + 1


def test_host_climatology_gather_touches_only_the_needed_rows():
  """engine.gather_aligned_host == the oracle's alignment, for a climatology
  stored (hour, dayofyear, longitude, latitude) and valid times across 29 Feb."""
  init = np.datetime64('2020-02-28T00', 'ns') + np.arange(3) * np.timedelta64(
      12, 'h')
  lead = (np.arange(4) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
  rng = np.random.default_rng(60)
  clim = rng.normal(size=(4, 366, 5, 3)).astype(np.float32)
  C = xl.DataArray(clim, ('hour', 'dayofyear', 'longitude', 'latitude'),
                   coords={'hour': np.arange(0, 24, 6),
                           'dayofyear': np.arange(1, 367),
                           'longitude': np.arange(5) * 72.0,
                           'latitude': np.linspace(-60, 60, 3)})
  P = xl.DataArray(np.zeros((3, 4, 3, 5), np.float32),
                   ('init_time', 'lead_time', 'latitude', 'longitude'),
                   coords={'init_time': init, 'lead_time': lead,
                           'latitude': np.linspace(-60, 60, 3),
                           'longitude': np.arange(5) * 72.0})
  aligned = engine.align_climatology(P, C)
  got, dims = engine.gather_aligned_host(aligned)
  assert dims == ('init_time', 'lead_time', 'longitude', 'latitude')
  want, wdims = oracle.align_climatology(
      clim, ('hour', 'dayofyear', 'longitude', 'latitude'),
      {'hour': np.arange(0, 24, 6), 'dayofyear': np.arange(1, 367)}, init, lead)
  want = np.transpose(want, [list(wdims).index(d) for d in dims])
  np.testing.assert_array_equal(got, want)
  assert got.shape == (3, 4, 5, 3)


def test_new_metric_classes_pickle():
  """Beam pickles Metric objects to its workers (beam_pipeline.py:150-159):
  nothing on the instances may hold a library handle or a lazy field."""
  import pickle
  from weatherbenchx_b200.metrics import categorical as cat
  from weatherbenchx_b200.metrics import probabilistic
  clim = xl.Dataset({'rain_seeps_threshold': xl.DataArray(
      np.ones((2, 366, 3, 4), np.float32),
      ('hour', 'dayofyear', 'latitude', 'longitude'))})
  metrics = {
      'csi': wrappers.WrappedMetric(cat.CSI(), [
          wrappers.ContinuousToBinary('both', [0.1, 1.0], 'threshold')]),
      'sedi': cat.SEDI(), 'seeps': cat.SEEPS(['rain'], clim),
      'exceed': deterministic.ErrorExceedance([0.5]),
      'ens_exceed': probabilistic.EnsembleErrorExceedance([0.5]),
      'intensity': deterministic.RelativeIntensity(),
      'distance': probabilistic.CRPSEnsembleDistance()}
  clone = pickle.loads(pickle.dumps(metrics))
  assert set(clone) == set(metrics)
  for name, metric in metrics.items():
    assert (sorted(s.unique_name for s in clone[name].statistics.values()) ==
            sorted(s.unique_name for s in metric.statistics.values())), name
  transform = clone['csi'].transforms[0]
  np.testing.assert_array_equal(transform._labels, [0.1, 1.0])
