"""GPU parity tests of the categorical statistics inside the fused reduction.

``wbx_det_desc.xform`` (thresholds applied on load, contingency table / error
exceedance accumulated in the same pass) through the C ABI and through the
class surface, against the CPU oracle (``oracle.binarize_thresholds``,
``contingency_table``, ``error_exceedance`` -- pinned to the reference's own
categorical.py by tests/test_reference_golden.py), against the reference's
inline known answers, and -- at 0.25 degree size -- through size-independent
properties.  Counts are integers: unweighted sums are compared EXACTLY.
"""

import numpy as np
import pytest

import wbx_oracle as oracle
from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import threshold_f32
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import categorical
from weatherbenchx_b200.metrics import deterministic
from weatherbenchx_b200.metrics import wrappers

pytestmark = pytest.mark.gpu
RTOL = 1e-5
KINDS = ('TruePositives', 'FalsePositives', 'FalseNegatives', 'TrueNegatives')


def _rain(rng, shape, nan_frac=0.0):
  """Precipitation-like float32 field: zeros, quarter-mm values, 0.1f hits."""
  x = (np.round(rng.gamma(0.8, 2.0, shape) * 4) / 4 *
       (rng.random(shape) < 0.6)).astype(np.float32)
  x[rng.random(shape) < 0.05] = np.float32(0.1)
  if nan_frac:
    x[rng.random(shape) < nan_frac] = np.nan
  return x


def _table_oracle(p, t, thresholds):
  """{kind: field [..., K]} from the oracle (float64 threshold compare)."""
  bp = oracle.binarize_thresholds(p, thresholds)
  bt = oracle.binarize_thresholds(t, thresholds)
  return oracle.contingency_table(bp, bt)


# ---------------------------------------------------------------------------
# the C ABI directly: every kernel path, every NaN mode
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('mode', ['propagate', 'masked', 'skipna',
                                  'masked_skipna'])
@pytest.mark.parametrize('path', ['tma', 'ldg4', 'scalar', 'w_x'])
def test_xform_through_cabi(path, mode, space):
  """Jobs = (threshold, init); cells = thresholds.  Weighted sums of the four
  table entries and sum_weights against the oracle."""
  import torch
  rng = np.random.default_rng(11)
  ny, nx = (19, 37) if path == 'scalar' else (24, 40)
  n_init = 5
  thresholds = [0.0, 0.1, 0.5, 2.0]
  nan = 0.08 if mode != 'propagate' else 0.0
  p = _rain(rng, (n_init, ny, nx), nan_frac=nan / 2)
  t = _rain(rng, (n_init, ny, nx), nan_frac=nan)
  mask = np.ascontiguousarray(~np.isnan(t))
  if mode == 'masked':  # unmasked NaN would (correctly) poison every sum
    p = np.where(np.isnan(p), np.float32(1.0), p)
  w_y = oracle.grid_area_weights(np.linspace(-90, 90, ny))
  w_x = rng.random(nx) + 0.5 if path == 'w_x' else None
  thr32 = threshold_f32(thresholds)
  keep = [p, t, mask]
  if space == 'device':
    dp, dt = torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda()
    dm = torch.from_numpy(mask.view(np.uint8)).cuda()
    keep += [dp, dt, dm]
    base = (dp.data_ptr(), dt.data_ptr(), dm.data_ptr())
  else:
    base = (p.ctypes.data, t.ctypes.data, mask.ctypes.data)
  slab = ny * nx
  jobs = [(k, i) for k in range(len(thresholds)) for i in range(n_init)]
  flags = ((_cabi.FLAG_MASKED if 'masked' in mode else 0) |
           (_cabi.FLAG_SKIPNA if 'skipna' in mode else 0) |
           (_cabi.FLAG_FORCE_LDG if path == 'ldg4' else 0) |
           (_cabi.FLAG_FORCE_TMA if path == 'tma' else 0))
  ctx = _cabi.get_context()
  plan = _cabi.DetPlan(
      ctx, space=_cabi.SPACE_DEVICE if space == 'device' else _cabi.SPACE_HOST,
      flags=flags, ny=ny, nx=nx,
      pred=np.array([base[0] + i * slab * 4 for _, i in jobs], np.uint64),
      target=np.array([base[1] + i * slab * 4 for _, i in jobs], np.uint64),
      mask=(np.array([base[2] + i * slab for _, i in jobs], np.uint64)
            if 'masked' in mode else None),
      cell=np.array([k for k, _ in jobs], np.int32), n_cells=len(thresholds),
      w_y=w_y, w_x=w_x, xform=_cabi.XF_CONTINGENCY,
      thr_pred=np.array([thr32[k] for k, _ in jobs], np.float32),
      thr_target=np.array([thr32[k] for k, _ in jobs], np.float32))
  ws, w = plan.run_to_host()
  del keep
  table = _table_oracle(p, t, thresholds)
  weights = [(w_y, ('y',))] + ([(w_x, ('x',))] if w_x is not None else [])
  for kind in KINDS:
    sws, sw, dims = oracle.aggregate(
        table[kind], ('init', 'y', 'x', 'thr'), ['init', 'y', 'x'],
        weights=weights, mask=mask if 'masked' in mode else None,
        mask_dims=('init', 'y', 'x') if 'masked' in mode else None,
        masked='masked' in mode, skipna='skipna' in mode)
    assert tuple(dims) == ('thr',)
    slot = _cabi.XF_SLOT[kind]
    np.testing.assert_allclose(ws[:, slot], sws, rtol=RTOL, equal_nan=True,
                               err_msg=kind)
    for k in range(_cabi.NUM_DET_WCLASSES):
      np.testing.assert_allclose(w[:, k], sw, rtol=1e-12, err_msg=kind)
  # the four entries partition the valid points
  total = ws[:, :4].sum(axis=1)
  np.testing.assert_allclose(total, w[:, 0], rtol=1e-12, equal_nan=True)


def test_xform_errors_are_exceptions():
  ctx = _cabi.get_context()
  x = np.zeros((4, 4), np.float32)
  args = dict(space=_cabi.SPACE_HOST, flags=0, ny=4, nx=4,
              pred=np.array([x.ctypes.data], np.uint64),
              target=np.array([x.ctypes.data], np.uint64),
              cell=np.zeros(1, np.int32), n_cells=1)
  thr = np.zeros(1, np.float32)
  with pytest.raises(_cabi.WbxError):   # thresholds without a request
    _cabi.DetPlan(ctx, thr_pred=thr, **args)
  with pytest.raises(_cabi.WbxError):   # contingency needs both thresholds
    _cabi.DetPlan(ctx, xform=_cabi.XF_CONTINGENCY, thr_pred=thr, **args)
  with pytest.raises(_cabi.WbxError):   # NONZERO flag and a threshold table
    _cabi.DetPlan(ctx, xform=_cabi.XF_CONTINGENCY | _cabi.XF_PRED_NONZERO,
                  thr_pred=thr, thr_target=thr, **args)
  with pytest.raises(_cabi.WbxError):   # not together with a climatology
    _cabi.DetPlan(ctx, xform=_cabi.XF_CONTINGENCY, thr_pred=thr,
                  thr_target=thr, clim=np.array([x.ctypes.data], np.uint64),
                  **args)
  with pytest.raises(_cabi.WbxError):   # unknown request bits
    _cabi.DetPlan(ctx, xform=7, thr_pred=thr, thr_target=thr, **args)
  # the context is still usable afterwards
  plan = _cabi.DetPlan(ctx, xform=_cabi.XF_ERROR_EXCEEDANCE,
                       thr_pred=np.array([-1.0], np.float32), **args)
  ws, w = plan.run_to_host()
  assert ws[0, 0] == 16.0 and w[0, 0] == 16.0


# ---------------------------------------------------------------------------
# per-point fields (materialised handles): bit-exact
# ---------------------------------------------------------------------------


def _da(values, name='precip'):
  dims = ('init_time', 'latitude', 'longitude')
  n, ny, nx = values.shape
  return xl.DataArray(values, dims, name=name, coords={
      'init_time': np.arange(n), 'latitude': np.linspace(-90, 90, ny),
      'longitude': np.linspace(0, 360, nx, endpoint=False)})


def test_materialised_fields_are_bit_exact():
  rng = np.random.default_rng(5)
  p = _rain(rng, (3, 19, 36), nan_frac=0.03)
  t = _rain(rng, (3, 19, 36), nan_frac=0.03)
  thresholds = [0.1, 0.0, 2.0, np.nan]
  transform = wrappers.ContinuousToBinary('both', thresholds, 'threshold')
  bp, bt = transform.transform_fn(_da(p)), transform.transform_fn(_da(t))
  assert bp.dims == ('init_time', 'latitude', 'longitude', 'threshold')
  np.testing.assert_array_equal(bp.values,
                                oracle.binarize_thresholds(p, thresholds))
  bp, bt = transform.transform_fn(_da(p)), transform.transform_fn(_da(t))
  table = _table_oracle(p, t, thresholds)
  for kind in KINDS:
    stat = getattr(categorical, kind)().compute({'v': bp}, {'v': bt})['v']
    assert stat.dims == ('init_time', 'latitude', 'longitude', 'threshold')
    assert stat.is_lazy
    np.testing.assert_array_equal(stat.values, table[kind], err_msg=kind)


def test_error_exceedance_known_answer():
  """metrics/metrics_test.py:1031-1049, incl. the NaN threshold and input."""
  predictions = xl.DataArray(
      np.array([0, -1, 1, np.nan], np.float32), dims=['x'], name='v')
  targets = xl.DataArray(np.zeros(4, np.float32), dims=['x'], name='v')
  result = deterministic.ErrorExceedance(
      thresholds=xl.DataArray([0, 0.5, 1, np.nan], dims=['y'])
  )._compute_per_variable(predictions, targets)
  expected = np.array([[0, 0, 0, np.nan], [1, 1, 0, np.nan],
                       [1, 1, 0, np.nan], [np.nan] * 4])
  assert result.dims == ('x', 'y')
  np.testing.assert_array_equal(result.values, expected)


def _precipitation_metric(metric_name, metric, predictions, targets):
  """metrics/metrics_test_utils.py:69-83 (compute_precipitation_metric): the
  inputs are binary fields already -- no threshold transform -- and the
  statistics are averaged over every dim."""
  metrics = {metric_name: metric}
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'])
  values = aggregation.compute_metric_values_for_single_chunk(
      metrics, aggregator, {'rain': predictions}, {'rain': targets})
  return float(values[f'{metric_name}.rain'].values.reshape(-1)[0])


def test_far_and_csi_known_answers():
  """metrics/metrics_test.py:100-170."""
  zeros = _da(np.zeros((2, 19, 36), np.float32), 'rain')
  ones = _da(np.ones((2, 19, 36), np.float32), 'rain')
  half = np.zeros((2, 19, 36), np.float32)
  half[0] = 1
  half = _da(half, 'rain')
  nan = np.ones((2, 19, 36), np.float32)
  nan[0] = np.nan
  nan = _da(nan, 'rain')
  far, csi = categorical.FalseAlarmRate(), categorical.CSI()
  assert np.isnan(_precipitation_metric('far', far, zeros, zeros))
  assert _precipitation_metric('far', far, ones, ones) == 0
  assert _precipitation_metric('far', far, ones, zeros) == 1
  assert _precipitation_metric('far', far, ones, half) == 0.5
  assert np.isnan(_precipitation_metric('far', far, zeros, nan))
  assert np.isnan(_precipitation_metric('csi', csi, zeros, zeros))
  assert _precipitation_metric('csi', csi, ones, ones) == 1
  assert _precipitation_metric('csi', csi, ones, zeros) == 0
  assert _precipitation_metric('csi', csi, ones, half) == 0.5
  assert np.isnan(_precipitation_metric('csi', csi, zeros, nan))


# ---------------------------------------------------------------------------
# class surface: one launch per variable for the whole table
# ---------------------------------------------------------------------------


@pytest.mark.parametrize('space', ['host', 'device'])
def test_whole_table_is_one_launch_and_matches_oracle(space):
  rng = np.random.default_rng(8)
  p, t = _rain(rng, (4, 24, 48)), _rain(rng, (4, 24, 48))
  P, T = _da(p), _da(t)
  if space == 'device':
    P, T = engine.to_device(P), engine.to_device(T)
  thresholds = [0.25, 0.1, 1.0]
  both = [wrappers.ContinuousToBinary('both', thresholds, 'threshold')]
  metrics = {'ets': wrappers.WrappedMetric(categorical.ETS(), both),
             'csi': wrappers.WrappedMetric(categorical.CSI(), both),
             'sedi': wrappers.WrappedMetric(categorical.SEDI(), both)}
  aggregator = aggregation.Aggregator(
      reduce_dims=['init_time', 'latitude', 'longitude'],
      weigh_by=[weighting.GridAreaWeighting()])
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, {'v': P}, {'v': T})
  ctx = _cabi.get_context()
  engine.clear_plan_cache()
  before = ctx.kernel_launches()
  state = aggregator.aggregate_statistics(statistics)
  assert ctx.kernel_launches() - before == 2  # reduction + finalize
  values = state.metric_values(metrics)
  w = oracle.grid_area_weights(np.linspace(-90, 90, 24))
  table = _table_oracle(p, t, thresholds)
  means = {}
  for kind in KINDS:
    sws, sw, _ = oracle.aggregate(
        table[kind], P.dims + ('threshold',), aggregator.reduce_dims,
        weights=[(w, ('latitude',))])
    means[kind] = sws / sw
  for name in metrics:
    np.testing.assert_allclose(
        values[f'{name}.v'].values,
        oracle.categorical_metric(
            name, means['TruePositives'], means['FalsePositives'],
            means['FalseNegatives'], means['TrueNegatives']),
        rtol=RTOL, err_msg=name)
    np.testing.assert_array_equal(
        values[f'{name}.v'].coords['threshold'].values, thresholds)


# ---------------------------------------------------------------------------
# BASELINE-size properties (0.25 degree)
# ---------------------------------------------------------------------------


def test_full_size_counts_are_exact_and_monotone():
  """721 x 1440 x 4 fields, unweighted: the four entries are integer counts
  that add up to the number of points for every threshold; the number of
  predicted events never grows with the threshold; a threshold above every
  value gives only true negatives."""
  import torch
  gen = torch.Generator(device='cuda').manual_seed(3)
  shape = (4, 721, 1440)
  p = torch.rand(shape, generator=gen, device='cuda') * 4 - 1
  t = p + torch.randn(shape, generator=gen, device='cuda') * 0.5
  dims = ('init_time', 'latitude', 'longitude')
  P, T = xl.DataArray(p, dims, name='v'), xl.DataArray(t, dims, name='v')
  thresholds = [-2.0, 0.0, 0.5, 1.5, 10.0]
  transform = wrappers.ContinuousToBinary('both', thresholds, 'threshold')
  stats = {k: getattr(categorical, k)().compute(
      {'v': transform.transform_fn(P)}, {'v': transform.transform_fn(T)})
           for k in KINDS}
  state = aggregation.Aggregator(reduce_dims=list(dims)).aggregate_statistics(
      stats)
  counts = {k: state.sum_weighted_statistics[k]['v'].values for k in KINDS}
  n = float(np.prod(shape))
  total = sum(counts.values())
  np.testing.assert_array_equal(total, np.full(len(thresholds), n))
  for k in KINDS:
    np.testing.assert_array_equal(counts[k], np.round(counts[k]))
    np.testing.assert_array_equal(state.sum_weights[k]['v'].values,
                                  np.full(len(thresholds), n))
  predicted = counts['TruePositives'] + counts['FalsePositives']
  assert (np.diff(predicted) <= 0).all()
  assert predicted[0] == n and predicted[-1] == 0
  assert counts['TrueNegatives'][-1] == n
  # against torch on the device for one threshold
  k = 2
  bp, bt = p > thresholds[k], t > thresholds[k]
  assert counts['TruePositives'][k] == float((bp & bt).sum())
  assert counts['FalseNegatives'][k] == float((~bp & bt).sum())
