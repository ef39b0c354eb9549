"""Property tests of the host side against the oracle (no GPU).

Random evaluation requests -- dim orders, reduce dims, weights, NaN modes, a
mask on either input, grid bins, outer-dim bins, several variables -- go through
the real class surface (statistics, Aggregator, planner, merged launches,
result unpacking); the plans are executed by the NumPy interpreter of the plan
contract (tests/wbx_emulator.py) instead of the CUDA library, and every
AggregationState entry is compared with ``oracle.aggregate`` on the same
arrays.  Requests the fused kernels cannot express fall back to the generic
CUDA kernel, which needs a GPU: those draws are rejected here
(``hypothesis.assume``) and are covered by the ``-m gpu`` tests.
"""

import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings
from hypothesis import strategies as st

import wbx_emulator
import wbx_oracle as oracle
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import binning
from weatherbenchx_b200 import generic
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import deterministic

RTOL = 1e-5
OUTER = ('init_time', 'lead_time', 'level')
GRID = ('latitude', 'longitude')
SIZES = {'init_time': 3, 'lead_time': 2, 'level': 2, 'latitude': 6,
         'longitude': 8}
COORDS = {
    'init_time': np.datetime64('2020-02-28T00', 'ns') +
                 np.arange(3) * np.timedelta64(12, 'h'),
    'lead_time': (np.arange(2) * np.timedelta64(6, 'h')
                  ).astype('timedelta64[ns]'),
    'level': np.array([500, 850]),
    'latitude': np.linspace(-75, 75, 6),
    'longitude': np.arange(8) * 45.0,
}
REGIONS = {'all': ((-90, 90), (0, 360)), 'north': ((0, 90), (0, 360)),
           'box': ((-50, 20), (300, 100))}


class _GenericNeedsGpu(Exception):
  pass


@st.composite
def requests(draw):
  outer = list(draw(st.permutations(OUTER)))
  outer = outer[:draw(st.integers(1, 3))]
  grid = list(draw(st.permutations(GRID)))
  dims = tuple(outer + grid)
  reduce_outer = [d for d in outer if draw(st.booleans())]
  reduce_dims = reduce_outer + list(GRID)
  mode = draw(st.sampled_from(['propagate', 'masked', 'skipna',
                               'masked_skipna']))
  mask_on = draw(st.sampled_from(['targets', 'predictions', 'both']))
  weighted = draw(st.booleans())
  bins = draw(st.lists(st.sampled_from(
      ['regions', 'regions_land', 'lat_bands', 'init_hour', 'lead_sets',
       'level_sets']), max_size=2, unique=True))
  n_vars = draw(st.integers(1, 2))
  seed = draw(st.integers(0, 2**16))
  with_acc = draw(st.booleans()) and {'init_time', 'lead_time'} <= set(dims)
  return dict(dims=dims, reduce_dims=reduce_dims, mode=mode, mask_on=mask_on,
              weighted=weighted, bins=bins, n_vars=n_vars, seed=seed,
              with_acc=with_acc)


def _make_bins(names, land):
  out = []
  for name in names:
    if name == 'regions':
      out.append(binning.Regions(REGIONS))
    elif name == 'regions_land':
      out.append(binning.Regions(REGIONS, bin_dim_name='region_l',
                                 land_sea_mask=land))
    elif name == 'lat_bands':
      out.append(binning.LatitudeBins(60))
    elif name == 'init_hour':
      out.append(binning.ByTimeUnit('hour', 'init_time', add_global_bin=True))
    elif name == 'lead_sets':
      out.append(binning.ByTimeUnitSets({'zero': 0, 'any': [0, 6]}, 'hour',
                                        'lead_time'))
    elif name == 'level_sets':
      out.append(binning.BySets({'low': [850]}, 'level',
                                bin_dim_name='level_set',
                                add_set_complements=True))
  return out


@settings(max_examples=300, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.filter_too_much,
                                 HealthCheck.too_slow,
                                 HealthCheck.function_scoped_fixture])
@given(req=requests())
def test_random_requests_match_the_oracle(req, monkeypatch):
  wbx_emulator.installed(monkeypatch)

  def no_generic(*args, **kwargs):
    raise _GenericNeedsGpu()
  monkeypatch.setattr(generic, 'aggregate', no_generic)

  dims, rng = req['dims'], np.random.default_rng(req['seed'])
  shape = tuple(SIZES[d] for d in dims)
  coords = {d: COORDS[d] for d in dims}
  for needed, bin_name in (('init_time', 'init_hour'),
                           ('lead_time', 'lead_sets'),
                           ('level', 'level_sets')):
    assume(needed in dims or bin_name not in req['bins'])
  masked, skipna = 'masked' in req['mode'], 'skipna' in req['mode']
  land = xl.DataArray(rng.random((6, 8)) < 0.5, GRID,
                      coords={d: COORDS[d] for d in GRID})
  predictions, targets, arrays, climatology, clims = {}, {}, {}, {}, {}
  clim_dims = ('dayofyear', 'hour') + tuple(
      d for d in dims if d not in ('init_time', 'lead_time'))
  for v in range(req['n_vars']):
    if req['with_acc']:
      c = rng.normal(size=(366, 4) + tuple(SIZES[d] for d in clim_dims[2:])
                     ).astype(np.float32)
      climatology[f'v{v}'] = xl.DataArray(
          c, clim_dims, coords=dict(
              {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)},
              **{d: COORDS[d] for d in clim_dims[2:]}), name=f'v{v}')
      clims[f'v{v}'] = c
    p = rng.normal(size=shape).astype(np.float32)
    t = rng.normal(size=shape).astype(np.float32)
    if req['mode'] != 'propagate':
      t[rng.random(shape) < 0.1] = np.nan
    mask = rng.random(shape) > 0.3
    P = xl.DataArray(p, dims, coords=coords, name=f'v{v}')
    T = xl.DataArray(t, dims, coords=coords, name=f'v{v}')
    if masked and req['mask_on'] in ('predictions', 'both'):
      P = P.assign_coords(mask=xl.DataArray(mask, dims))
    if masked and req['mask_on'] in ('targets', 'both'):
      T = T.assign_coords(mask=xl.DataArray(mask, dims))
    predictions[f'v{v}'], targets[f'v{v}'] = P, T
    arrays[f'v{v}'] = (p, t, mask)
  metrics = {'rmse': deterministic.RMSE(), 'mae': deterministic.MAE(),
             'bias': deterministic.Bias()}
  if req['with_acc']:
    metrics['acc'] = deterministic.ACC(climatology)
  bin_by = _make_bins(req['bins'], land)
  aggregator = aggregation.Aggregator(
      reduce_dims=req['reduce_dims'],
      weigh_by=[weighting.GridAreaWeighting()] if req['weighted'] else None,
      bin_by=bin_by or None, masked=masked, skipna=skipna)
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  try:
    state = aggregator.aggregate_statistics(statistics)
  except _GenericNeedsGpu:
    assume(False)

  weights = []
  if req['weighted']:
    weights.append((oracle.grid_area_weights(COORDS['latitude']),
                    ('latitude',)))
  probe = predictions['v0']
  bin_masks = []
  for b in bin_by:
    m = b.create_bin_mask(probe)
    bin_masks.append((m.values, m.dims))
  fields = {'SquaredError': oracle.squared_error,
            'AbsoluteError': oracle.absolute_error, 'Error': oracle.error}
  # which inputs a statistic's expression touches decides whether it carries
  # the mask coordinate (deterministic.py:225-259, aggregation.py:339)
  touches = {'SquaredPredictionAnomaly': {'predictions'},
             'SquaredTargetAnomaly': {'targets'}}
  for var, (p, t, mask) in arrays.items():
    values = {name: fn(p, t) for name, fn in fields.items()}
    if req['with_acc']:
      aligned, adims = oracle.align_climatology(
          clims[var], clim_dims,
          {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6)},
          COORDS['init_time'], COORDS['lead_time'])
      aligned = np.transpose(aligned, [adims.index(d) for d in dims])
      for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
        values[name] = fn(p, t, aligned)
    for name, value in values.items():
      carries = masked and (
          req['mask_on'] == 'both' or
          req['mask_on'] in touches.get(name, {'predictions', 'targets'}))
      sws, sw, out_dims = oracle.aggregate(
          value, dims, req['reduce_dims'], weights=weights,
          bin_masks=bin_masks, mask=mask if carries else None,
          mask_dims=dims if carries else None, masked=carries, skipna=skipna)
      got_ws = state.sum_weighted_statistics[name][var]
      got_w = state.sum_weights[name][var]
      assert tuple(got_ws.dims) == tuple(out_dims), (got_ws.dims, out_dims)
      np.testing.assert_array_equal(np.isnan(got_ws.values), np.isnan(sws))
      scale = np.abs(sws[np.isfinite(sws)]).max() if np.isfinite(sws).any() else 0
      np.testing.assert_allclose(got_ws.values, sws, rtol=RTOL,
                                 atol=RTOL * scale, equal_nan=True)
      np.testing.assert_allclose(got_w.values, sw, rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------------------
# ensemble statistics
# ---------------------------------------------------------------------------


@st.composite
def ensemble_requests(draw):
  outer = list(draw(st.permutations(('init_time', 'lead_time'))))
  outer = outer[:draw(st.integers(1, 2))]
  layout = draw(st.sampled_from(['member_last', 'member_major',
                                 'member_first']))
  reduce_outer = [d for d in outer if draw(st.booleans())]
  return dict(
      outer=outer, layout=layout, reduce_dims=reduce_outer + list(GRID),
      members=draw(st.sampled_from([2, 3, 7, 10])),
      mode=draw(st.sampled_from(['propagate', 'masked', 'skipna'])),
      skipna_ensemble=draw(st.booleans()),
      use_sort=draw(st.booleans()), fair=draw(st.booleans()),
      weighted=draw(st.booleans()), n_vars=draw(st.integers(1, 2)),
      seed=draw(st.integers(0, 2**16)))


@settings(max_examples=200, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.filter_too_much,
                                 HealthCheck.too_slow,
                                 HealthCheck.function_scoped_fixture])
@given(req=ensemble_requests())
def test_random_ensemble_requests_match_the_oracle(req, monkeypatch):
  from weatherbenchx_b200.metrics import probabilistic
  wbx_emulator.installed(monkeypatch)

  def no_generic(*args, **kwargs):
    raise _GenericNeedsGpu()
  monkeypatch.setattr(generic, 'aggregate', no_generic)
  # the sort estimator rejects skipna_ensemble (probabilistic.py:215-216)
  assume(not (req['use_sort'] and req['skipna_ensemble']))

  rng = np.random.default_rng(req['seed'])
  tdims = tuple(req['outer']) + GRID
  m = req['members']
  edims = {'member_last': tdims + ('realization',),
           'member_major': tuple(req['outer']) + ('realization',) + GRID,
           'member_first': ('realization',) + tdims}[req['layout']]
  sizes = dict(SIZES, realization=m)
  coords = dict(COORDS, realization=np.arange(m))
  masked, skipna = req['mode'] == 'masked', req['mode'] == 'skipna'
  predictions, targets, arrays = {}, {}, {}
  for v in range(req['n_vars']):
    y = rng.normal(size=tuple(sizes[d] for d in tdims)).astype(np.float32)
    x = rng.normal(size=tuple(sizes[d] for d in edims)).astype(np.float32)
    if req['skipna_ensemble']:
      holes = rng.random(x.shape) < 0.2
      axis = edims.index('realization')
      keep = [slice(None)] * x.ndim
      keep[axis] = slice(0, 2)        # at least two members everywhere
      holes[tuple(keep)] = False
      x[holes] = np.nan
    if req['mode'] != 'propagate':
      y[rng.random(y.shape) < 0.1] = np.nan
    mask = rng.random(y.shape) > 0.3
    X = xl.DataArray(x, edims, coords={d: coords[d] for d in edims},
                     name=f'v{v}')
    Y = xl.DataArray(y, tdims, coords={d: coords[d] for d in tdims},
                     name=f'v{v}')
    if masked:
      Y = Y.assign_coords(mask=xl.DataArray(mask, tdims))
    predictions[f'v{v}'], targets[f'v{v}'] = X, Y
    arrays[f'v{v}'] = (x, y, mask)
  kw = dict(ensemble_dim='realization', skipna_ensemble=req['skipna_ensemble'])
  metrics = {
      'crps': probabilistic.CRPSEnsemble(
          use_sort=req['use_sort'], fair=req['fair'], **kw),
      'ssr': probabilistic.UnbiasedSpreadSkillRatio(**kw)}
  aggregator = aggregation.Aggregator(
      reduce_dims=req['reduce_dims'],
      weigh_by=[weighting.GridAreaWeighting()] if req['weighted'] else None,
      masked=masked, skipna=skipna)
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  try:
    state = aggregator.aggregate_statistics(statistics)
  except _GenericNeedsGpu:
    assume(False)

  weights = []
  if req['weighted']:
    weights.append((oracle.grid_area_weights(COORDS['latitude']),
                    ('latitude',)))
  axis = edims.index('realization')
  skip = req['skipna_ensemble']
  fair = 'fair' if req['fair'] else 'unfair'
  for var, (x, y, mask) in arrays.items():
    # per-point fields in the statistic's dim order (ensemble dim dropped)
    sdims = tuple(d for d in edims if d != 'realization')
    y_s = np.transpose(y, [tdims.index(d) for d in sdims])
    mask_s = np.transpose(mask, [tdims.index(d) for d in sdims])
    expected = {
        'CRPSSkill_realization': (oracle.crps_skill(x, y_s, axis, skip), True),
        f'CRPSSpread_realization_{fair}_predictions': (
            oracle.crps_spread(x, axis, fair=req['fair'],
                               use_sort=req['use_sort'],
                               skipna_ensemble=skip), False),
        f'EnsembleVariance_realization_skipna_ensemble_{skip}': (
            oracle.ensemble_variance(x, axis, skip), False),
        f'UnbiasedEnsembleMeanSquaredError_realization_skipna_ensemble_{skip}':
            (oracle.unbiased_ensemble_mean_squared_error(x, y_s, axis, skip),
             True),
    }
    assert set(expected) == set(state.sum_weighted_statistics)
    for name, (field, touches_targets) in expected.items():
      carries = masked and touches_targets
      sws, sw, out_dims = oracle.aggregate(
          field, sdims, req['reduce_dims'], weights=weights,
          mask=mask_s if carries else None,
          mask_dims=sdims if carries else None, masked=carries, skipna=skipna)
      got_ws = state.sum_weighted_statistics[name][var]
      got_w = state.sum_weights[name][var]
      assert set(got_ws.dims) == set(out_dims)
      got_ws = got_ws.transpose(*out_dims).values
      got_w = got_w.transpose(*out_dims).values
      np.testing.assert_array_equal(np.isnan(got_ws), np.isnan(sws),
                                    err_msg=name)
      finite = np.abs(sws[np.isfinite(sws)])
      np.testing.assert_allclose(
          got_ws, sws, rtol=1e-4, equal_nan=True, err_msg=name,
          atol=1e-4 * (finite.max() if finite.size else 0))
      np.testing.assert_allclose(got_w, sw, rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------------------
# categorical path: thresholds applied inside the reduction
# ---------------------------------------------------------------------------


@st.composite
def categorical_requests(draw):
  outer = list(draw(st.permutations(OUTER)))
  outer = outer[:draw(st.integers(1, 3))]
  grid = list(draw(st.permutations(GRID)))
  dims = tuple(outer + grid)
  reduce_dims = [d for d in outer if draw(st.booleans())] + list(GRID)
  return dict(
      dims=dims, reduce_dims=reduce_dims,
      mode=draw(st.sampled_from(['propagate', 'masked', 'skipna',
                                 'masked_skipna'])),
      mask_on=draw(st.sampled_from(['targets', 'predictions'])),
      weighted=draw(st.booleans()),
      which=draw(st.sampled_from(['both', 'predictions', 'targets', 'none'])),
      thresholds=draw(st.lists(st.sampled_from(
          [-0.5, 0.0, 0.1, 0.25, 1.0, 7.0]), min_size=1, max_size=3,
          unique=True)),
      bins=draw(st.lists(st.sampled_from(['init_hour', 'lead_sets',
                                          'level_sets', 'regions']),
                         max_size=1)),
      exceedance=draw(st.booleans()),
      n_vars=draw(st.integers(1, 2)), seed=draw(st.integers(0, 2**16)))


@settings(max_examples=200, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.filter_too_much,
                                 HealthCheck.too_slow,
                                 HealthCheck.function_scoped_fixture])
@given(req=categorical_requests())
def test_random_categorical_requests_match_the_oracle(req, monkeypatch):
  """Contingency tables (inputs thresholded on either side, or binary already)
  and error exceedance through the class surface, any dim order, NaN mode and
  outer-dim binning, against the oracle's restatement of categorical.py /
  wrappers.binarize_thresholds / deterministic.ErrorExceedance."""
  from weatherbenchx_b200.metrics import categorical, wrappers
  wbx_emulator.installed(monkeypatch)

  def no_generic(*args, **kwargs):
    raise _GenericNeedsGpu()
  monkeypatch.setattr(generic, 'aggregate', no_generic)

  dims, rng = req['dims'], np.random.default_rng(req['seed'])
  shape = tuple(SIZES[d] for d in dims)
  coords = {d: COORDS[d] for d in dims}
  for needed, bin_name in (('init_time', 'init_hour'),
                           ('lead_time', 'lead_sets'),
                           ('level', 'level_sets')):
    assume(needed in dims or bin_name not in req['bins'])
  masked, skipna = 'masked' in req['mode'], 'skipna' in req['mode']
  which, thresholds = req['which'], req['thresholds']
  land = xl.DataArray(rng.random((6, 8)) < 0.5, GRID,
                      coords={d: COORDS[d] for d in GRID})

  def rain():
    x = (np.round(rng.gamma(0.8, 1.5, shape) * 4) / 4 *
         (rng.random(shape) < 0.6)).astype(np.float32)
    x[rng.random(shape) < 0.05] = np.float32(0.1)
    return x

  predictions, targets, arrays = {}, {}, {}
  for v in range(req['n_vars']):
    p, t = rain(), rain()
    # inputs that skip the transform are binary events already
    if which in ('targets', 'none'):
      p = (p > 0.25).astype(np.float32)
    if which in ('predictions', 'none'):
      t = (t > 0.25).astype(np.float32)
    if req['mode'] != 'propagate':
      t[rng.random(shape) < 0.1] = np.nan
      if not masked:
        p[rng.random(shape) < 0.05] = np.nan
    mask = ~np.isnan(t) & (rng.random(shape) > 0.2)
    P = xl.DataArray(p, dims, coords=coords, name=f'v{v}')
    T = xl.DataArray(t, dims, coords=coords, name=f'v{v}')
    if masked and req['mask_on'] == 'predictions':
      P = P.assign_coords(mask=xl.DataArray(mask, dims))
    if masked and req['mask_on'] == 'targets':
      T = T.assign_coords(mask=xl.DataArray(mask, dims))
    predictions[f'v{v}'], targets[f'v{v}'] = P, T
    arrays[f'v{v}'] = (p, t, mask)
  table = categorical.Accuracy()          # all four entries
  if which == 'none':
    metrics = {'accuracy': table}
  else:
    metrics = {'accuracy': wrappers.WrappedMetric(table, [
        wrappers.ContinuousToBinary(which, thresholds, 'thr')])}
  if req['exceedance']:
    metrics['exceedance'] = deterministic.ErrorExceedance(thresholds)
  bin_by = _make_bins(req['bins'], land)
  aggregator = aggregation.Aggregator(
      reduce_dims=req['reduce_dims'],
      weigh_by=[weighting.GridAreaWeighting()] if req['weighted'] else None,
      bin_by=bin_by or None, masked=masked, skipna=skipna)
  statistics = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  try:
    state = aggregator.aggregate_statistics(statistics)
  except _GenericNeedsGpu:
    assume(False)   # grid bins: per-point fields + generic kernel (GPU tests)

  weights = []
  if req['weighted']:
    weights.append((oracle.grid_area_weights(COORDS['latitude']),
                    ('latitude',)))
  bin_masks = []
  for b in bin_by:
    m = b.create_bin_mask(predictions['v0'])
    bin_masks.append((m.values, m.dims))
  suffix = '' if which == 'none' else (
      f'_{which}_thr=' + ','.join(str(x) for x in thresholds))
  for var, (p, t, mask) in arrays.items():
    bp = (oracle.binarize_thresholds(p, thresholds)
          if which in ('both', 'predictions') else p)
    bt = (oracle.binarize_thresholds(t, thresholds)
          if which in ('both', 'targets') else t)
    if which in ('predictions', 'targets'):
      # the untransformed input broadcasts against the threshold dim
      bp, bt = (bp, bt[..., None]) if which == 'predictions' else (
          bp[..., None], bt)
    bp, bt = np.broadcast_arrays(bp, bt)
    fields = {f'{k}{suffix}': (v, dims + (('thr',) if which != 'none' else ()))
              for k, v in oracle.contingency_table(bp, bt).items()}
    if req['exceedance']:
      fields['ErrorExceedance'] = (
          oracle.error_exceedance(p, t, thresholds),
          dims + ('error_exceedance_thresholds',))
    for name, (value, vdims) in fields.items():
      sws, sw, out_dims = oracle.aggregate(
          value, vdims, req['reduce_dims'], weights=weights,
          bin_masks=bin_masks, mask=mask if masked else None,
          mask_dims=dims if masked else None, masked=masked, skipna=skipna)
      got_ws = state.sum_weighted_statistics[name][var]
      got_w = state.sum_weights[name][var]
      assert sorted(got_ws.dims) == sorted(out_dims), (got_ws.dims, out_dims)
      order = [got_ws.dims.index(d) for d in out_dims]
      np.testing.assert_array_equal(
          np.isnan(got_ws.values.transpose(order)), np.isnan(sws), err_msg=name)
      np.testing.assert_allclose(got_ws.values.transpose(order), sws,
                                 rtol=1e-9, atol=1e-9, equal_nan=True,
                                 err_msg=name)
      np.testing.assert_allclose(got_w.values.transpose(order), sw,
                                 rtol=1e-12, atol=1e-12, err_msg=name)
