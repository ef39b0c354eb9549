"""Parity against vectors produced by the reference's own code.

``tests/golden/reference_golden.npz`` holds, for 68 evaluation cases, the
AggregationState (sum_weighted_statistics / sum_weights per statistic and
variable) and the metric values that the UNMODIFIED modules of
/root/reference/weatherbenchX returned in the build container
(``tests/golden/make_reference_golden.py``; xarray / jax are not installable
there, the reference ran on the stand-in modules of
``tests/golden/reference_runtime.py`` -- read its docstring for the exact
split between reference code and stand-in).

* not gpu: the NumPy oracle reproduces every stored array -- this is what pins
  ``oracle/wbx_oracle.py`` to the reference;
* gpu: the very same case definitions (``tests/golden/reference_cases.py``)
  are run with this package's modules in place of the reference's, i.e. the
  user code is identical and only the import changes; state and values must
  agree with the stored ones.

Tolerance: 1e-5 relative (north_star; the reference's assert_allclose default)
plus an absolute term for sums that cancel (Error / AnomalyCovariance sums of
mixed sign); weight sums that are counts are compared exactly.
"""

import os
import sys
import warnings

import numpy as np
import pytest

import wbx_oracle as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import reference_cases as cases  # noqa: E402  pylint: disable=g-import-not-at-top

GOLDEN = os.path.join(HERE, 'golden', 'reference_golden.npz')
RTOL = 1e-5


@pytest.fixture(scope='module')
def golden():
  with np.load(GOLDEN) as data:
    return {k: data[k] for k in data.files}


@pytest.fixture(scope='module')
def inputs(golden):
  return {k[3:]: v for k, v in golden.items() if k.startswith('in/')}


def _case_keys(golden, case, part):
  prefix = f'{case}/{part}/'
  return [k for k in golden if k.startswith(prefix) and '@' not in k]


def _abs_tolerance(expected):
  """Sums of mixed sign cancel: allow RTOL of the typical magnitude."""
  finite = np.abs(expected[np.isfinite(expected)])
  return RTOL * float(finite.max()) if finite.size else 0.0


def _assert_close(actual, expected, what, exact=False):
  actual = np.asarray(actual)
  expected = np.asarray(expected)
  assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
  np.testing.assert_array_equal(np.isnan(actual), np.isnan(expected),
                                err_msg=f'{what}: NaN pattern')
  if exact:
    np.testing.assert_array_equal(actual, expected, err_msg=what)
  else:
    np.testing.assert_allclose(actual, expected, rtol=RTOL,
                               atol=_abs_tolerance(expected), err_msg=what)


def _transpose_to(values, dims, want_dims):
  dims, want_dims = list(dims), list(want_dims)
  assert sorted(dims) == sorted(want_dims), (dims, want_dims)
  return np.transpose(values, [dims.index(d) for d in want_dims])


# ---------------------------------------------------------------------------
# the oracle, case by case
# ---------------------------------------------------------------------------


def _oracle_bins(spec, inputs):
  """Stacked boolean masks [(mask, dims), ...] and their labels."""
  out = []
  for name in spec['bins']:
    land = inputs['ens_land' if name.startswith('ens_') else 'land']
    if name in ('regions', 'regions_land', 'ens_regions_land'):
      regions = cases.ENS_REGIONS if name.startswith('ens_') else cases.REGIONS
      masks, labels = oracle.regions_masks(
          cases.LAT, cases.LON, regions,
          land_sea_mask=None if name == 'regions' else land)
      out.append(((masks, ('region', 'latitude', 'longitude')), labels))
    elif name == 'landsea_global':
      # binning.py:92-144: land = fraction >= 0.5, sea = 1 - land, global.
      land_mask = land.astype(np.float32) >= 0.5
      masks = np.stack([land_mask, ~land_mask, np.ones_like(land_mask)])
      out.append(((masks, ('land_sea', 'latitude', 'longitude')),
                  ['land', 'sea', 'global']))
    elif name in ('init_hour', 'init_hour_global'):
      # binning.py:394-442 + 301-332: one bin per unique hour of init_time;
      # the optional 'global' bin comes first and turns the labels into str.
      hours = (cases.INIT.astype('datetime64[h]') -
               cases.INIT.astype('datetime64[D]').astype('datetime64[h]')
               ).astype(np.int64)
      unique = np.unique(hours)
      masks = hours[None, :] == unique[:, None]
      labels = list(unique)
      if name.endswith('global'):
        masks = np.concatenate([np.ones((1, len(hours)), bool), masks])
        labels = ['global'] + [str(u) for u in unique]
      out.append(((masks, ('init_time_hour', 'init_time')), labels))
    elif name == 'valid_month':
      valid = cases.INIT[:, None] + cases.LEAD[None, :]
      month = valid.astype('datetime64[M]').astype(np.int64) % 12 + 1
      unique = np.unique(month)
      masks = month[None] == unique[:, None, None]
      out.append(((masks, ('valid_time_month', 'init_time', 'lead_time')),
                  list(unique)))
    elif name == 'lead_sets':
      # binning.py:445-515: named, possibly overlapping sets; 'global' last.
      hours = cases.LEAD.astype('timedelta64[s]').astype(np.int64) // 3600
      masks = [np.isin(hours, np.atleast_1d(v))
               for v in cases.LEAD_SETS.values()]
      masks.append(np.ones(len(hours), bool))
      out.append(((np.stack(masks), ('lead_time_hour_sets', 'lead_time')),
                  list(cases.LEAD_SETS) + ['global']))
    elif name == 'level_sets':
      # binning.py:640-704 with add_set_complements.
      masks, labels = [], []
      for key, values in cases.LEVEL_SETS.items():
        m = np.isin(cases.LEVEL, values)
        masks += [m, ~m]
        labels += [key, f'not_in_{key}']
      out.append(((np.stack(masks), ('level_set', 'level')), labels))
    elif name == 'lat30':
      # binning.py:204-243: closed bands [start, start + degrees].
      starts = np.arange(-90, 90 + 30, 30)[:-1]
      masks = np.stack([(cases.LAT >= s) & (cases.LAT <= s + 30)
                        for s in starts])
      out.append(((masks, ('latitude_bins', 'latitude')), list(starts)))
    elif name == 'lon90':
      # binning.py:246-298 with the wrap-around rule of :63-76.
      starts = np.arange(0, 360 + 90, 90)[:-1]
      masks = []
      for s in starts:
        west, east = np.mod(s, 360), np.mod(s + 90, 360)
        lon = np.mod(cases.LON, 360)
        masks.append((lon >= west) & (lon <= east) if east > west
                     else (lon <= east) | (lon >= west))
      out.append(((np.stack(masks), ('longitude_bins', 'longitude')),
                  list(starts)))
    else:
      raise KeyError(name)
  return out


def _oracle_fields(spec, inputs):
  """{(statistic unique_name, variable): (values, dims, mask or None)}."""
  family = spec['family']
  out = {}
  if family in ('det', 'acc'):
    for var, p, t, holes, dims, rows in (
        ('2m_temperature', inputs['p2'], inputs['t2'], inputs['holes2'],
         cases.D2, inputs['c2_rows']),
        ('geopotential', inputs['p3'], inputs['t3'], inputs['holes3'],
         cases.D3, inputs['c3_rows'])):
      mask = None
      if spec['nan_targets']:
        t = cases.with_nan(t, holes)
        mask = ~holes
      out[('Error', var)] = (oracle.error(p, t), dims, mask)
      out[('AbsoluteError', var)] = (oracle.absolute_error(p, t), dims, mask)
      out[('SquaredError', var)] = (oracle.squared_error(p, t), dims, mask)
      if family == 'acc':
        clim = cases.full_climatology(rows)
        clim_dims = ('dayofyear', 'hour') + tuple(dims[2:])
        aligned, adims = oracle.align_climatology(
            clim, clim_dims,
            {'dayofyear': np.arange(1, 367), 'hour': cases.HOURS},
            cases.INIT, cases.LEAD)
        assert tuple(adims) == tuple(dims)
        assert not np.isnan(aligned).any()
        for name, fn in oracle.CLIMATOLOGY_STATISTICS.items():
          # (p - c)**2 never touches the targets, so it does not inherit
          # their 'mask' coordinate (deterministic.py:225-232).
          out[(name, var)] = (
              fn(p, t, aligned), dims,
              None if name == 'SquaredPredictionAnomaly' else mask)
  elif family == 'cat':
    var = 'total_precipitation_6hr'
    p, t, mask = inputs['rain_p'], inputs['rain_t'], None
    if spec['nan_predictions']:
      p = cases.with_nan(p, inputs['rain_p_holes'])
    if spec['nan_targets']:
      t = cases.with_nan(t, inputs['rain_holes'])
      mask = ~inputs['rain_holes']
    if spec['kind'] == 'relative_intensity':
      field, field_mask = oracle.relative_intensity(p, t, (2, 3), mask)
      out[('RelativeIntensity', var)] = (field, cases.D2[:2], field_mask)
    elif spec['kind'] == 'seeps':
      # climatology rows are stored (hour, dayofyear, longitude, latitude);
      # gather the wet threshold of every valid time, p1 = nanmean over time
      doy, hour = oracle.dayofyear_and_hour(
          cases.INIT[:, None] + cases.LEAD[None, :])
      rows = inputs['seeps_threshold_rows']
      doy_pos = np.searchsorted(cases.DOY_USED, doy)
      assert np.array_equal(cases.DOY_USED[doy_pos], doy)
      wet = rows[np.searchsorted(cases.HOURS, hour), doy_pos]  # [i, l, lon, lat]
      wet = np.swapaxes(wet, -1, -2)
      p1 = oracle.seeps_p1(inputs['seeps_dry_fraction_rows'], (0, 1)).T
      field, p1_mask = oracle.seeps(
          p, t, wet, p1, dry_threshold_mm=cases.SEEPS_DRY_THRESHOLD_MM)
      full_mask = np.broadcast_to(p1_mask, field.shape)
      if mask is not None:   # categorical.py:296-302
        full_mask = full_mask & mask
      name = ('SEEPS_total_precipitation_6hr_dry_threshold_mm_'
              f'{cases.SEEPS_DRY_THRESHOLD_MM}_min_p1_0.1_max_p1_0.85')
      out[(name, var)] = (field, cases.D2, full_mask)
    elif spec['kind'] == 'exceedance':
      dims = cases.D2 + ('error_exceedance_thresholds',)
      field = oracle.error_exceedance(p, t, cases.EXCEEDANCE_THRESHOLDS)
      out[('ErrorExceedance', var)] = (field, dims, mask)
    else:
      dims = cases.D2 + ('threshold',)
      if spec['kind'] == 'pred_only':
        thresholds = [0.5, 2.0]
        bt = (inputs['rain_t'] > 0.5).astype(np.float32)[..., None]
        suffix = 'predictions_threshold=0.5,2.0'
      else:
        thresholds = cases.RAIN_THRESHOLDS
        bt = oracle.binarize_thresholds(t, thresholds)
        suffix = 'both_threshold=' + ','.join(str(v) for v in thresholds)
      bp = oracle.binarize_thresholds(p, thresholds)
      bp, bt = np.broadcast_arrays(bp, bt)
      for name, field in oracle.contingency_table(bp, bt).items():
        out[(f'{name}_{suffix}', var)] = (field, dims, mask)
  elif family == 'wind':
    se = oracle.wind_vector_squared_error(
        inputs['u_p'], inputs['u_t'], inputs['v_p'], inputs['v_t'])
    out[('WindVectorSquaredError_wind_vector', 'wind_vector')] = (
        se, cases.D3, None)
  else:
    member_major_input = bool(spec.get('x_key'))   # 50 / 51-member cases
    x = inputs[spec.get('x_key') or 'x_last']
    if spec.get('member_nan'):
      x = cases.with_nan(x, inputs[spec.get('holes_key') or 'member_holes'])
    y, mask = inputs['y'], None
    if spec.get('nan_targets'):
      y = cases.with_nan(y, inputs['y_holes'])
      mask = ~inputs['y_holes']
    if member_major_input:
      axis = 1
    elif spec.get('layout', 'member_major') == 'member_major':
      x = np.ascontiguousarray(np.moveaxis(x, -1, 1))
      axis = 1
    else:
      axis = x.ndim - 1
    dims = cases.D_ENS_T
    if family == 'ens_averaged':
      se = oracle.squared_error(x, np.expand_dims(y, axis))
      out[('SquaredError_each_realization', 't2m')] = (
          se.mean(axis=axis), dims, None)
    elif family == 'ens_distance':
      # probabilistic.py:135-145: mean of |x_m - y_k| over both member dims;
      # :199-247 for the two spreads
      y_ens = inputs['y_ens']                              # [init, K, lat, lon]
      diff = np.abs(x[:, :, None] - y_ens[:, None, :])     # [init, M, K, ...]
      out[('CRPSSkill_realization', 't2m')] = (diff.mean(axis=(1, 2)), dims,
                                               None)
      for which, data in (('predictions', x), ('targets', y_ens)):
        out[(f'CRPSSpread_realization_fair_{which}', 't2m')] = (
            oracle.crps_spread(data, 1, fair=True,
                               use_sort=spec.get('use_sort', False)), dims,
            None)
    elif family == 'ens_exceedance':
      # probabilistic.py:855-861: exceedance of every member, then xarray's
      # NaN-skipping mean over the members
      field = oracle.error_exceedance(x, np.expand_dims(y, axis),
                                      [1.0, 2.5, 6.0])
      with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out[('EnsembleErrorExceedance', 't2m')] = (
            np.nanmean(field, axis=axis),
            dims + ('error_exceedance_thresholds',), None)
    elif family == 'ens_mean':
      name = ("SquaredError_predictions_ensemble_mean_self._ensemble_dim="
              "'realization'_self._skipna=False")
      out[(name, 't2m')] = (
          oracle.squared_error(oracle.ensemble_mean(x, axis), y), dims, None)
    else:
      for skip in (False, True):
        if skip and not spec.get('member_nan'):
          continue
        # functions of the predictions alone carry no target mask
        out[(f'EnsembleVariance_realization_skipna_ensemble_{skip}', 't2m')] = (
            oracle.ensemble_variance(x, axis, skip), dims, None)
        out[('UnbiasedEnsembleMeanSquaredError_realization_skipna_ensemble_'
             f'{skip}', 't2m')] = (
                 oracle.unbiased_ensemble_mean_squared_error(x, y, axis, skip),
                 dims, mask)
      out['skill'] = lambda skip: (oracle.crps_skill(x, y, axis, skip), dims,
                                   mask)
      out['spread'] = lambda fair, sort, skip: (
          oracle.crps_spread(x, axis, fair=fair, use_sort=sort,
                             skipna_ensemble=skip), dims, None)
  return out


def _oracle_state(case, spec, fields, stat, var, inputs):
  if (stat, var) in fields:
    values, dims, mask = fields[(stat, var)]
  elif stat == 'CRPSSkill_realization':
    values, dims, mask = fields['skill'](case == 'ens/skipna_ensemble')
  elif stat.startswith('CRPSSpread_realization_'):
    values, dims, mask = fields['spread'](
        '_fair_' in stat, case.endswith('/use_sort'),
        case == 'ens/skipna_ensemble')
  else:
    values, dims, mask = fields[(stat, var)]
  weights = []
  if spec['weighted']:
    weights.append((oracle.grid_area_weights(cases.LAT), ('latitude',)))
  bins = _oracle_bins(spec, inputs)
  result = oracle.aggregate(
      values, dims, spec['reduce_dims'], weights=weights,
      bin_masks=[b for b, _ in bins], mask=mask,
      mask_dims=(dims[:mask.ndim] if mask is not None else None),
      masked=spec['masked'],
      skipna=spec['skipna'])
  assert result is not None
  return result


def test_fixture_is_complete(golden):
  names = [str(n) for n in golden['cases']]
  assert len(names) == 68 and len(set(names)) == 68
  for case in names:
    assert _case_keys(golden, case, 'sws'), case
    assert _case_keys(golden, case, 'value'), case


def test_inputs_are_reproducible(inputs):
  """The stored inputs are exactly what make_inputs() produces (seeded)."""
  fresh = cases.make_inputs()
  assert set(fresh) == set(inputs)
  for name, values in fresh.items():
    np.testing.assert_array_equal(values, inputs[name], err_msg=name)


def _case_table(inputs_dict):
  ns = cases.namespace(
      xr=_PlainNamespace(), aggregation=_PlainNamespace(),
      binning=_PlainNamespace(), weighting=_PlainNamespace(),
      base=_PlainNamespace(), deterministic=_PlainNamespace(),
      probabilistic=_PlainNamespace(), wrappers=_PlainNamespace(),
      categorical=_PlainNamespace())
  return {name: spec for name, spec, *_ in cases.build_cases(ns, inputs_dict)}


class _PlainNamespace:
  """Accepts any attribute access or call: only the ``spec`` of each case is
  used by the oracle tests."""

  def __getattr__(self, name):
    return self

  def __call__(self, *args, **kwargs):
    return self


def test_oracle_reproduces_reference_states(golden, inputs):
  """Every sum_weighted_statistics / sum_weights array of every case."""
  table = _case_table(inputs)
  checked = 0
  for case, spec in table.items():
    fields = _oracle_fields(spec, inputs)
    for key in _case_keys(golden, case, 'sws'):
      _, stat, var = key[len(case) + 1:].split('/', 2)
      sws, sw, out_dims = _oracle_state(case, spec, fields, stat, var, inputs)
      want_dims = [str(d) for d in golden[key + '@dims']]
      _assert_close(_transpose_to(sws, out_dims, want_dims), golden[key], key)
      sw_key = f'{case}/sw/{stat}/{var}'
      _assert_close(_transpose_to(sw, out_dims, want_dims), golden[sw_key],
                    sw_key, exact=not spec['weighted'])
      checked += 1
  assert checked == len([k for k in golden if '/sws/' in k and '@' not in k])


def test_oracle_reproduces_reference_values(golden, inputs):
  """RMSE / ACC / CRPS ... values from the oracle's means."""
  table = _case_table(inputs)

  def mean(case, spec, fields, stat, var):
    sws, sw, dims = _oracle_state(case, spec, fields, stat, var, inputs)
    return oracle.mean_statistic(sws, sw), dims

  for case, spec in table.items():
    fields = _oracle_fields(spec, inputs)
    for key in _case_keys(golden, case, 'value'):
      metric, var = key[len(case) + len('/value/'):].split('.', 1)
      want_dims = [str(d) for d in golden[key + '@dims']]
      skip = case == 'ens/skipna_ensemble'
      if metric in ('rmse', 'mse', 'rmse_members', 'rmse_mean', 'wind_rmse'):
        stat = {
            'rmse_members': 'SquaredError_each_realization',
            'rmse_mean': ("SquaredError_predictions_ensemble_mean_self."
                          "_ensemble_dim='realization'_self._skipna=False"),
            'wind_rmse': 'WindVectorSquaredError_wind_vector',
        }.get(metric, 'SquaredError')
        value, dims = mean(case, spec, fields, stat, var)
        value = value if metric == 'mse' else oracle.rmse_from_mean(value)
      elif metric == 'mae':
        value, dims = mean(case, spec, fields, 'AbsoluteError', var)
      elif metric == 'bias':
        value, dims = mean(case, spec, fields, 'Error', var)
      elif metric == 'acc':
        cov, dims = mean(case, spec, fields, 'AnomalyCovariance', var)
        spa, _ = mean(case, spec, fields, 'SquaredPredictionAnomaly', var)
        sta, _ = mean(case, spec, fields, 'SquaredTargetAnomaly', var)
        value = oracle.acc_from_means(cov, spa, sta)
      elif metric == 'crps_distance':
        skill, dims = mean(case, spec, fields, 'CRPSSkill_realization', var)
        spread_p, _ = mean(case, spec, fields,
                           'CRPSSpread_realization_fair_predictions', var)
        spread_t, _ = mean(case, spec, fields,
                           'CRPSSpread_realization_fair_targets', var)
        value = skill - 0.5 * spread_p - 0.5 * spread_t
      elif metric == 'ens_exceedance':
        value, dims = mean(case, spec, fields, 'EnsembleErrorExceedance', var)
      elif spec['family'] == 'cat':
        if metric == 'exceedance':
          value, dims = mean(case, spec, fields, 'ErrorExceedance', var)
        elif metric == 'relative_intensity':
          value, dims = mean(case, spec, fields, 'RelativeIntensity', var)
        elif metric == 'seeps':
          stat = next(k[0] for k in fields if k[0].startswith('SEEPS_'))
          value, dims = mean(case, spec, fields, stat, var)
        else:
          suffix = next(k[0] for k in fields if k[0].startswith(
              'TruePositives_'))[len('TruePositives_'):]
          parts = {}
          for name in ('TruePositives', 'FalsePositives', 'FalseNegatives',
                       'TrueNegatives'):
            parts[name], dims = mean(case, spec, fields, f'{name}_{suffix}',
                                     var)
          value = oracle.categorical_metric(
              metric, parts['TruePositives'], parts['FalsePositives'],
              parts['FalseNegatives'], parts['TrueNegatives'])
      elif metric in ('crps_fair', 'crps_unfair'):
        fair = metric.split('_')[1]
        skill, dims = mean(case, spec, fields, 'CRPSSkill_realization', var)
        spread, _ = mean(case, spec, fields,
                         f'CRPSSpread_realization_{fair}_predictions', var)
        value = oracle.crps_from_means(skill, spread)
      else:
        var_stat = f'EnsembleVariance_realization_skipna_ensemble_{skip}'
        mse_stat = ('UnbiasedEnsembleMeanSquaredError_realization_'
                    f'skipna_ensemble_{skip}')
        if metric == 'ens_var':          # probabilistic.py EnsembleRootMeanVariance
          value, dims = mean(case, spec, fields, var_stat, var)
          value = np.sqrt(value)
        elif metric == 'unbiased_rmse':  # UnbiasedEnsembleMeanRMSE
          value, dims = mean(case, spec, fields, mse_stat, var)
          value = np.sqrt(value)
        elif metric == 'unbiased_ssr':   # UnbiasedSpreadSkillRatio
          spread, dims = mean(case, spec, fields, var_stat, var)
          skill, _ = mean(case, spec, fields, mse_stat, var)
          value = np.sqrt(spread / skill)
        else:
          raise KeyError(metric)
      _assert_close(_transpose_to(value, dims, want_dims), golden[key], key)


def test_oracle_weights_match_reference(golden):
  """GridAreaWeighting.weights on five latitude grids (weighting.py:45-130)."""
  for name in ('poles_ascending', 'poles_descending', 'no_poles',
               'quarter_degree', 'float32_coord'):
    lat = golden[f'weights/{name}/latitude']
    np.testing.assert_allclose(oracle.grid_area_weights(lat),
                               golden[f'weights/{name}/weights'], rtol=1e-12,
                               atol=1e-15, err_msg=name)
  assert golden['weights/no_latitude_dim'] == 1


def test_binning_masks_match_reference(golden):
  """create_bin_mask of the coordinate-value binnings (ByExactCoord,
  ByTimeUnit, ByTimeUnitSets, ByTimeUnitFromSeconds, ByCoordBins, BySets):
  mask, bin labels and label dtype kind equal the reference's
  (binning.py:301-704), on a sparse-style statistic."""
  ns = _product_namespace()
  built = {name: (instance, stat)
           for name, instance, stat in cases.mask_cases(ns)}
  assert list(built) == [str(n) for n in golden['mask_cases']]
  for name, (instance, stat) in built.items():
    mask = instance.create_bin_mask(stat)
    bdim = instance.bin_dim_name
    want_dims = [str(d) for d in golden[f'masks/{name}/dims']]
    assert want_dims[0] == bdim and set(mask.dims) == set(want_dims), name
    got = mask.transpose(*want_dims)
    assert got.dtype == bool, name
    np.testing.assert_array_equal(got.values, golden[f'masks/{name}/mask'],
                                  err_msg=name)
    labels = mask.coords[bdim].values
    assert [str(v) for v in labels] == [
        str(v) for v in golden[f'masks/{name}/labels']], name
    assert labels.dtype.kind == str(golden[f'masks/{name}/label_kind']), name


def test_chunk_combine_equals_monolithic_in_the_reference(golden):
  """AggregationState.__add__ over init_time chunks (aggregation.py:84-110):
  the reference's chunked values equal its monolithic ones."""
  for key in _case_keys(golden, 'det/chunked', 'value'):
    whole = key.replace('det/chunked', 'det/weighted')
    np.testing.assert_allclose(golden[key], golden[whole], rtol=1e-12)


# ---------------------------------------------------------------------------
# the CUDA path: same case code, this package's modules
# ---------------------------------------------------------------------------


def _product_namespace():
  from weatherbenchx_b200 import aggregation, binning, weighting
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.metrics import base, categorical, deterministic
  from weatherbenchx_b200.metrics import probabilistic, wrappers
  return cases.namespace(
      xr=xl, aggregation=aggregation, binning=binning, weighting=weighting,
      base=base, deterministic=deterministic, probabilistic=probabilistic,
      wrappers=wrappers, categorical=categorical)


def _labels_match(golden, key, da):
  for d in da.dims:
    stored = golden.get(f'{key}@labels/{d}')
    if stored is not None:
      assert [str(v) for v in da.coords[d].values] == [str(v) for v in stored]


CASE_NAMES = [
    'det/weighted', 'det/unweighted', 'det/keep_init',
    'det/reduce_all_but_level', 'det/nan_default', 'det/nan_masked',
    'det/nan_skipna', 'det/nan_masked_skipna', 'det/regions',
    'det/regions_x_landsea', 'det/regions_nan_masked',
    'det/regions_nan_default', 'det/lat_lon_bands', 'det/by_init_hour',
    'det/by_valid_month', 'det/lead_sets_x_regions',
    'det/regions_x_init_hour_global', 'det/by_level_sets',
    'det/by_init_hour_nan_default', 'det/by_init_hour_nan_masked',
    'det/by_valid_month_skipna', 'acc/weighted',
    'acc/nan_skipna', 'acc/nan_masked', 'wind/weighted', 'ens/member_last',
    'ens/member_major', 'ens/use_sort', 'ens/unweighted_keep_init',
    'ens/skipna_ensemble', 'ens/nan_members_propagate', 'ens/regions',
    'ens/nan_targets_default', 'ens/nan_targets_masked',
    'ens/nan_targets_skipna', 'ens/regions_nan_targets_masked',
    'ens/ensemble_averaged_rmse', 'ens/ensemble_mean_rmse',
    'ens50/all_metrics', 'ens51/use_sort', 'ens50/moments_only',
    'ens50/nan_targets_masked', 'ens51/nan_members_propagate',
    'ens50/regions',
    'cat/table_weighted', 'cat/table_unweighted_keep_init',
    'cat/table_nan_default', 'cat/table_nan_masked',
    'cat/table_nan_masked_nan_predictions', 'cat/table_nan_skipna',
    'cat/table_by_init_hour', 'cat/table_regions',
    'cat/predictions_thresholded_binary_targets', 'cat/error_exceedance',
    'cat/error_exceedance_nan_skipna',
    'cat/error_exceedance_nan_default_keep_init',
    'cat/relative_intensity', 'cat/relative_intensity_masked',
    'cat/ensemble_error_exceedance',
    'cat/ensemble_error_exceedance_nan_members',
    'ens/distance_to_target_ensemble',
    'ens/distance_to_target_ensemble_sorted',
    'seeps/masked_weighted', 'seeps/nan_targets_masked',
    'seeps/nan_both_masked_keep_init', 'seeps/regions_masked',
    'seeps/default_propagates', 'seeps/skipna_unweighted']


def test_case_names_cover_the_fixture(golden):
  assert CASE_NAMES == [str(n) for n in golden['cases']]


def _run_product_case(golden, inputs, case, space):
  from weatherbenchx_b200 import engine
  ns = _product_namespace()
  for name, spec, metrics, aggregator, predictions, targets in (
      cases.build_cases(ns, inputs)):
    if name == case:
      break
  else:
    raise KeyError(case)
  if space == 'device':
    predictions = {k: engine.to_device(v) for k, v in predictions.items()}
    targets = {k: engine.to_device(v) for k, v in targets.items()}
  statistics = ns.base.compute_unique_statistics_for_all_metrics(
      metrics, predictions, targets)
  state = aggregator.aggregate_statistics(statistics)
  sws_keys = _case_keys(golden, case, 'sws')
  stored = {tuple(k[len(case) + 5:].split('/', 1)) for k in sws_keys}
  produced = {(s, v) for s, per_var in state.sum_weighted_statistics.items()
              for v in per_var}
  assert produced == stored  # same statistic unique_names, same variables
  for stat, var in sorted(stored):
    for part, tree in (('sws', state.sum_weighted_statistics),
                       ('sw', state.sum_weights)):
      key = f'{case}/{part}/{stat}/{var}'
      da = tree[stat][var]
      want_dims = [str(d) for d in golden[key + '@dims']]
      _labels_match(golden, key, da)
      _assert_close(_transpose_to(da.values, da.dims, want_dims), golden[key],
                    key, exact=part == 'sw' and not spec['weighted'])
  values = state.metric_values(metrics)
  value_keys = _case_keys(golden, case, 'value')
  assert {k[len(case) + len('/value/'):] for k in value_keys} == set(values)
  for key in value_keys:
    da = values[key[len(case) + len('/value/'):]]
    want_dims = [str(d) for d in golden[key + '@dims']]
    _assert_close(_transpose_to(da.values, da.dims, want_dims), golden[key],
                  key)


# Cases added late in round 1 run from tests/test_zz_gpu_seeps.py (SEEPS, the
# ensemble error exceedance, relative intensity); they passed on the B200 in
# their own run (profiles/gpu_tests_seeps_late_cases_r1.log).
SEEPS_CASES = [c for c in CASE_NAMES if c.startswith('seeps/')]
# Added after the GPU budget of round 1 was spent: compositions of kernels that
# passed on the B200 (merged CRPS launches over strided target-member views),
# CPU-verified through the interpreted plans, not yet run on hardware.
UNCONFIRMED_CASES = ['ens/distance_to_target_ensemble',
                     'ens/distance_to_target_ensemble_sorted']
LATE_CASES = SEEPS_CASES + ['cat/ensemble_error_exceedance',
                            'cat/ensemble_error_exceedance_nan_members',
                            'cat/relative_intensity',
                            'cat/relative_intensity_masked']


@pytest.mark.gpu
@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('case',
                         [c for c in CASE_NAMES
                          if c not in LATE_CASES + UNCONFIRMED_CASES])
def test_cuda_path_reproduces_reference(golden, inputs, case, space):
  """State and values of the reference, from the CUDA path, for every case."""
  _run_product_case(golden, inputs, case, space)


# Cases whose host-space path needs device memory even before a kernel runs
# (per-point fields of the CRPS launch, the EnsembleMean transform).
NEEDS_DEVICE = {'ens/regions', 'ens/regions_nan_targets_masked',
                'ens50/regions', 'ens/ensemble_mean_rmse',
                # region bins + thresholds: per-point fields, generic kernel
                'cat/table_regions',
                # NaN members: the exact route through the member-mean field
                'cat/ensemble_error_exceedance_nan_members',
                # the statistic is a small host array: generic kernel
                'cat/relative_intensity', 'cat/relative_intensity_masked'}


@pytest.mark.parametrize('case',
                         [c for c in CASE_NAMES if c not in NEEDS_DEVICE])
def test_host_side_with_interpreted_plans_reproduces_reference(
    golden, inputs, case, monkeypatch):
  """Everything above the C ABI (statistic classes, launch grouping, planner
  tables, result unpacking) on the GPU-less box: the plans are executed by the
  NumPy interpreter of tests/wbx_emulator.py instead of the CUDA library."""
  import wbx_emulator
  wbx_emulator.installed(monkeypatch)
  _run_product_case(golden, inputs, case, 'host')


@pytest.mark.gpu
def test_cuda_path_chunk_combine(golden, inputs):
  """Chunked evaluation + AggregationState.__add__ == the reference's."""
  ns = _product_namespace()
  metrics, total = cases.chunked_case(ns, inputs)
  values = total.metric_values(metrics)
  for key in _case_keys(golden, 'det/chunked', 'value'):
    da = values[key[len('det/chunked/value/'):]]
    want_dims = [str(d) for d in golden[key + '@dims']]
    _assert_close(_transpose_to(da.values, da.dims, want_dims), golden[key],
                  key)


def test_host_weights_match_reference(golden):
  """GridAreaWeighting of this package (host float64, uploaded as w_y)."""
  from weatherbenchx_b200 import weighting
  from weatherbenchx_b200 import xarray_lite as xl
  for name in ('poles_ascending', 'poles_descending', 'no_poles',
               'quarter_degree', 'float32_coord'):
    lat = golden[f'weights/{name}/latitude']
    stat = xl.DataArray(np.zeros((len(lat), 4), np.float32),
                        ('latitude', 'longitude'),
                        coords={'latitude': lat,
                                'longitude': np.arange(4) * 90.0})
    w = weighting.GridAreaWeighting().weights(stat)
    np.testing.assert_allclose(w.values, golden[f'weights/{name}/weights'],
                               rtol=1e-12, atol=1e-15, err_msg=name)
