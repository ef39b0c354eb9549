"""Evaluation cases shared by the golden generator and its consumers.

TEST INFRASTRUCTURE.  ``build_cases(ns, inputs)`` is written once against a
namespace of modules with the reference's names (``aggregation``, ``binning``,
``weighting``, ``deterministic``, ``probabilistic``, ``wrappers``, ``xr``):

* ``make_reference_golden.py`` passes the reference's own modules
  (/root/reference/weatherbenchX, build container only) and stores what they
  return;
* ``tests/test_reference_golden.py`` passes ``weatherbenchx_b200``'s modules
  (CUDA path, ``-m gpu``) and compares with the stored results -- the same
  user code runs on both sides, which is the drop-in claim;
* the CPU oracle test uses the ``spec`` of each case (plain data).

Inputs are seeded NumPy arrays produced by ``make_inputs`` and stored in the
fixture, so consumers never regenerate them.
"""

from __future__ import annotations

import types

import numpy as np

NLAT, NLON = 19, 32  # 608 points: a multiple of 16, so the fused class-map kernel applies
LAT = np.linspace(-90, 90, NLAT)
LON = np.linspace(0, 360, NLON, endpoint=False)
# Valid times cross 29 February of a leap year (dayofyear 58..62).
INIT = np.datetime64('2020-02-27T00', 'ns') + np.arange(4) * np.timedelta64(
    12, 'h')
LEAD = (np.arange(3) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
LEVEL = np.array([500, 850])
MEMBERS = np.arange(7)
ENS = 'realization'
D2 = ('init_time', 'lead_time', 'latitude', 'longitude')
D3 = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
D_ENS_T = ('init_time', 'latitude', 'longitude')
D_ENS_LAST = D_ENS_T + (ENS,)
D_ENS_MAJOR = ('init_time', ENS, 'latitude', 'longitude')
RD = ['init_time', 'latitude', 'longitude']
COORDS = {'init_time': INIT, 'lead_time': LEAD, 'level': LEVEL,
          'latitude': LAT, 'longitude': LON, ENS: MEMBERS}
REGIONS = {
    'global': ((-90, 90), (0, 360)),
    'tropics': ((-20, 20), (0, 360)),
    'nh': ((20, 90), (0, 360)),
    'europe': ((35, 75), (-12.5, 42.5)),   # wraps across 0 degrees
    'box': ((-45, 10), (100, 250)),
}
ENS_REGIONS = {'global': ((-90, 90), (0, 360)), 'sh': ((-90, -20), (0, 360)),
               'box': ((-10, 60), (300, 60))}
LEAD_SETS = {'analysis': 0, 'short': [0, 6], 'late': [6, 12]}   # overlapping
LEVEL_SETS = {'low': [850], 'both': [500, 850]}
DOY_USED = np.arange(58, 63)
HOURS = np.arange(0, 24, 6)


def _shape(dims):
  return tuple(len(COORDS[d]) for d in dims)


def make_inputs() -> dict:
  """Seeded float32 fields (and boolean hole / land patterns)."""
  rng = np.random.default_rng(20240229)
  f32 = np.float32
  out = {}
  out['t2'] = rng.normal(280, 10, _shape(D2)).astype(f32)
  out['p2'] = (out['t2'] + rng.normal(0.5, 2, _shape(D2))).astype(f32)
  out['t3'] = rng.normal(5000, 300, _shape(D3)).astype(f32)
  out['p3'] = (out['t3'] + rng.normal(-3, 40, _shape(D3))).astype(f32)
  out['holes2'] = rng.random(_shape(D2)) < 0.07
  out['holes3'] = rng.random(_shape(D3)) < 0.05
  out['land'] = rng.random((NLAT, NLON)) < 0.4
  out['c2_rows'] = rng.normal(
      280, 5, (len(DOY_USED), 4, NLAT, NLON)).astype(f32)
  out['c3_rows'] = rng.normal(
      5000, 100, (len(DOY_USED), 4, len(LEVEL), NLAT, NLON)).astype(f32)
  out['u_t'] = rng.normal(3, 8, _shape(D3)).astype(f32)
  out['v_t'] = rng.normal(-1, 6, _shape(D3)).astype(f32)
  out['u_p'] = (out['u_t'] + rng.normal(0, 2, _shape(D3))).astype(f32)
  out['v_p'] = (out['v_t'] + rng.normal(0, 2, _shape(D3))).astype(f32)
  out['y'] = rng.normal(280, 10, _shape(D_ENS_T)).astype(f32)
  out['x_last'] = (out['y'][..., None] + rng.normal(
      0.3, 3, _shape(D_ENS_LAST))).astype(f32)
  holes = rng.random(_shape(D_ENS_LAST)) < 0.15
  holes[..., :2] = False  # at least two members at every point
  out['member_holes'] = holes
  out['ens_land'] = rng.random((NLAT, NLON)) < 0.4
  out['y_holes'] = rng.random(_shape(D_ENS_T)) < 0.06
  # precipitation-like fields (mm): many exact zeros, values that hit the
  # thresholds exactly (multiples of 0.25) and thresholds that are not float32
  # numbers (0.1): the comparison at the boundary is part of the contract
  wet = rng.random(_shape(D2)) < 0.55
  out['rain_t'] = (np.round(rng.gamma(0.8, 2.0, _shape(D2)) * 4) / 4 * wet
                   ).astype(f32)
  out['rain_p'] = np.maximum(
      0, out['rain_t'] + np.round(rng.normal(0, 1.0, _shape(D2)) * 4) / 4
  ).astype(f32)
  out['rain_p'][rng.random(_shape(D2)) < 0.05] = f32(0.1)
  out['rain_t'][rng.random(_shape(D2)) < 0.05] = f32(0.1)
  out['rain_holes'] = rng.random(_shape(D2)) < 0.06
  out['rain_p_holes'] = rng.random(_shape(D2)) < 0.03
  # SEEPS climatology rows, stored (hour, dayofyear, longitude, latitude) as in
  # the reference's docstring (categorical.py:141-152).  Wet thresholds on the
  # quarter grid of the rain fields (so `x >= wet` is hit exactly), a few equal
  # to the dry threshold (0.25: a value can then be dry AND heavy); dry
  # fractions partly outside [0.1, 0.85] and NaN at some points.
  shape = (4, len(DOY_USED), NLON, NLAT)
  wet_thr = (np.round(rng.uniform(0.5, 3.0, shape) * 4) / 4).astype(f32)
  wet_thr[rng.random(shape) < 0.03] = f32(0.25)
  out['seeps_threshold_rows'] = wet_thr
  dry_frac = rng.uniform(0.0, 1.0, (NLON, NLAT))[None, None] + rng.normal(
      0, 0.02, shape)
  dry_frac[:, :, rng.random((NLON, NLAT)) < 0.04] = np.nan
  out['seeps_dry_fraction_rows'] = dry_frac.astype(f32)
  # an ensemble of targets (e.g. perturbed analyses), member-major
  out['y_ens'] = (out['y'][:, None] + rng.normal(
      0, 1.5, (len(INIT), N_TARGET_MEMBERS, NLAT, NLON))).astype(f32)
  # operational ensemble sizes (member-major), appended LAST so that every
  # array above keeps its values: the sizes the CRPS sort kernel has
  # fixed-size networks for (crps.cu, MFIX = 50 / 51)
  for m in (50, 51):
    out[f'x{m}_major'] = (out['y'][:, None] + rng.normal(
        0.3, 3, (len(INIT), m, NLAT, NLON))).astype(f32)
  holes = np.zeros(out['x51_major'].shape, bool)
  holes[:, 7] = rng.random((len(INIT), NLAT, NLON)) < 0.03   # one NaN member
  holes[:, 50] = rng.random((len(INIT), NLAT, NLON)) < 0.02  # ... or two
  out['member_holes51'] = holes
  return out


SEEPS_DRY_THRESHOLD_MM = 250.0   # 0.25 in the unit of the rain fields
N_TARGET_MEMBERS = 3


RAIN_THRESHOLDS = [0.0, 0.1, 0.5, 2.0, 1e9]
EXCEEDANCE_THRESHOLDS = [0.0, 0.25, 0.1, 3.0]


def full_climatology(rows: np.ndarray) -> np.ndarray:
  """[366, 4, ...] climatology that is NaN outside the stored day-of-year rows
  (a wrong gather index then shows up as NaN instead of passing silently)."""
  full = np.full((366,) + rows.shape[1:], np.nan, np.float32)
  full[DOY_USED - 1] = rows
  return full


def with_nan(values: np.ndarray, holes: np.ndarray) -> np.ndarray:
  out = values.copy()
  out[holes] = np.nan
  return out


def build_cases(ns, inputs):
  """Yields (name, spec, metrics, aggregator, predictions, targets).

  ``spec`` is plain data for consumers that do not go through the class
  surface: {'family', 'reduce_dims', 'weighted', 'masked', 'skipna', 'bins',
  'nan_targets', ...}.
  """
  xr = ns.xr
  agg, binning, weighting = ns.aggregation, ns.binning, ns.weighting
  det, prob, wrappers = ns.deterministic, ns.probabilistic, ns.wrappers

  def da(values, dims, **extra_coords):
    coords = {d: COORDS[d] for d in dims}
    if 'init_time' in dims and 'lead_time' in dims:
      coords['valid_time'] = (('init_time', 'lead_time'),
                              INIT[:, None] + LEAD[None, :])
    coords.update(extra_coords)
    return xr.DataArray(values, dims, coords=coords)

  def area():
    return [weighting.GridAreaWeighting()]

  # -- deterministic ---------------------------------------------------------
  predictions = {'2m_temperature': da(inputs['p2'], D2),
                 'geopotential': da(inputs['p3'], D3)}
  targets = {'2m_temperature': da(inputs['t2'], D2),
             'geopotential': da(inputs['t3'], D3)}
  targets_nan = {
      '2m_temperature': da(with_nan(inputs['t2'], inputs['holes2']), D2,
                           mask=(D2, ~inputs['holes2'])),
      'geopotential': da(with_nan(inputs['t3'], inputs['holes3']), D3,
                         mask=(D3, ~inputs['holes3'])),
  }
  metrics = {'rmse': det.RMSE(), 'mse': det.MSE(), 'mae': det.MAE(),
             'bias': det.Bias()}

  def det_case(name, reduce_dims=None, weighted=True, nan_targets=False,
               bins=None, use_metrics=None, only=None, **flags):
    reduce_dims = reduce_dims or RD
    spec = dict(family='det', reduce_dims=reduce_dims, weighted=weighted,
                masked=flags.get('masked', False),
                skipna=flags.get('skipna', False), bins=bins or [],
                nan_targets=nan_targets)
    aggregator = agg.Aggregator(
        reduce_dims=reduce_dims, weigh_by=area() if weighted else None,
        bin_by=_make_bins(ns, bins, inputs['land']) if bins else None,
        **flags)
    use_targets = targets_nan if nan_targets else targets
    use_predictions = predictions
    if only:  # a coordinate-value binning needs the coordinate on every variable
      use_predictions = {k: predictions[k] for k in only}
      use_targets = {k: use_targets[k] for k in only}
    return (name, spec, use_metrics or metrics, aggregator, use_predictions,
            use_targets)

  yield det_case('det/weighted')
  yield det_case('det/unweighted', weighted=False)
  yield det_case('det/keep_init', reduce_dims=['latitude', 'longitude'])
  yield det_case('det/reduce_all_but_level', reduce_dims=[
      'init_time', 'lead_time', 'latitude', 'longitude'])
  yield det_case('det/nan_default', nan_targets=True)
  yield det_case('det/nan_masked', nan_targets=True, masked=True)
  yield det_case('det/nan_skipna', nan_targets=True, skipna=True)
  yield det_case('det/nan_masked_skipna', nan_targets=True, masked=True,
                 skipna=True)
  yield det_case('det/regions', bins=['regions_land'])
  yield det_case('det/regions_x_landsea', bins=['regions', 'landsea_global'])
  yield det_case('det/regions_nan_masked', bins=['regions_land'],
                 nan_targets=True, masked=True)
  yield det_case('det/regions_nan_default', bins=['regions_land'],
                 nan_targets=True)
  yield det_case('det/lat_lon_bands', bins=['lat30', 'lon90'],
                 use_metrics={'mse': det.MSE()})

  # -- bins over outer dims: time units, value sets (binning.py:394-515,640-704)
  rd_all = ['init_time', 'lead_time', 'latitude', 'longitude']
  yield det_case('det/by_init_hour', bins=['init_hour'])
  yield det_case('det/by_valid_month', bins=['valid_month'], reduce_dims=rd_all)
  yield det_case('det/lead_sets_x_regions', bins=['lead_sets', 'regions_land'],
                 reduce_dims=rd_all)
  yield det_case('det/regions_x_init_hour_global',
                 bins=['regions', 'init_hour_global'])
  yield det_case('det/by_level_sets', bins=['level_sets'],
                 only=['geopotential'])
  yield det_case('det/by_init_hour_nan_default', bins=['init_hour'],
                 nan_targets=True)
  yield det_case('det/by_init_hour_nan_masked', bins=['init_hour_global'],
                 nan_targets=True, masked=True)
  yield det_case('det/by_valid_month_skipna', bins=['valid_month'],
                 reduce_dims=rd_all, nan_targets=True, skipna=True)

  # -- ACC with a (dayofyear, hour) climatology ------------------------------
  clim_coords = {'dayofyear': np.arange(1, 367), 'hour': HOURS}
  c2 = xr.DataArray(
      full_climatology(inputs['c2_rows']),
      ('dayofyear', 'hour', 'latitude', 'longitude'),
      coords=dict(clim_coords, latitude=LAT, longitude=LON))
  c3 = xr.DataArray(
      full_climatology(inputs['c3_rows']),
      ('dayofyear', 'hour', 'level', 'latitude', 'longitude'),
      coords=dict(clim_coords, level=LEVEL, latitude=LAT, longitude=LON))
  climatology = xr.Dataset({'2m_temperature': c2, 'geopotential': c3})
  acc_metrics = {'acc': det.ACC(climatology), 'rmse': det.RMSE()}
  for name, nan_targets, flags in (
      ('acc/weighted', False, {}),
      ('acc/nan_skipna', True, {'skipna': True}),
      ('acc/nan_masked', True, {'masked': True})):
    case = det_case(name, nan_targets=nan_targets, use_metrics=acc_metrics,
                    **flags)
    case[1]['family'] = 'acc'
    yield case

  # -- wind vector -----------------------------------------------------------
  wind_metrics = {'wind_rmse': det.WindVectorRMSE(
      u_name='u_component_of_wind', v_name='v_component_of_wind',
      vector_name='wind_vector')}
  yield ('wind/weighted',
         dict(family='wind', reduce_dims=RD, weighted=True, masked=False,
              skipna=False, bins=[], nan_targets=False),
         wind_metrics, agg.Aggregator(reduce_dims=RD, weigh_by=area()),
         {'u_component_of_wind': da(inputs['u_p'], D3),
          'v_component_of_wind': da(inputs['v_p'], D3)},
         {'u_component_of_wind': da(inputs['u_t'], D3),
          'v_component_of_wind': da(inputs['v_t'], D3)})

  # -- ensembles -------------------------------------------------------------
  y = da(inputs['y'], D_ENS_T)
  x_last = da(inputs['x_last'], D_ENS_LAST)
  x_major_values = np.ascontiguousarray(
      np.moveaxis(inputs['x_last'], -1, 1))
  x_major = da(x_major_values, D_ENS_MAJOR)
  x_nan = da(np.ascontiguousarray(np.moveaxis(
      with_nan(inputs['x_last'], inputs['member_holes']), -1, 1)),
             D_ENS_MAJOR)
  ens_metrics = {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True),
      'crps_unfair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=False),
      'ens_var': prob.EnsembleRootMeanVariance(ensemble_dim=ENS),
      'unbiased_rmse': prob.UnbiasedEnsembleMeanRMSE(ensemble_dim=ENS),
      'unbiased_ssr': prob.UnbiasedSpreadSkillRatio(ensemble_dim=ENS),
  }

  y_nan = da(with_nan(inputs['y'], inputs['y_holes']), D_ENS_T,
             mask=(D_ENS_T, ~inputs['y_holes']))

  def ens_case(name, x, use_metrics, reduce_dims=None, weighted=True,
               bins=None, layout='member_major', member_nan=False,
               family='ens', nan_targets=False, x_key=None, holes_key=None,
               **flags):
    reduce_dims = reduce_dims or RD
    spec = dict(family=family, reduce_dims=reduce_dims, weighted=weighted,
                masked=flags.get('masked', False),
                skipna=flags.get('skipna', False), bins=bins or [],
                layout=layout, member_nan=member_nan, nan_targets=nan_targets)
    if x_key:   # a member-major input other than x_last (50 / 51 members)
      spec.update(x_key=x_key, holes_key=holes_key)
    aggregator = agg.Aggregator(
        reduce_dims=reduce_dims, weigh_by=area() if weighted else None,
        bin_by=_make_bins(ns, bins, inputs['ens_land']) if bins else None,
        **flags)
    return (name, spec, use_metrics, aggregator, {'t2m': x},
            {'t2m': y_nan if nan_targets else y})

  yield ens_case('ens/member_last', x_last, ens_metrics, layout='member_last')
  yield ens_case('ens/member_major', x_major, ens_metrics)
  yield ens_case('ens/use_sort', x_major, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True,
                                     use_sort=True),
      'crps_unfair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=False,
                                       use_sort=True)})
  yield ens_case('ens/unweighted_keep_init', x_major, ens_metrics,
                 reduce_dims=['latitude', 'longitude'], weighted=False)
  yield ens_case('ens/skipna_ensemble', x_nan, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True,
                                     skipna_ensemble=True),
      'ens_var': prob.EnsembleRootMeanVariance(ensemble_dim=ENS,
                                               skipna_ensemble=True),
      'unbiased_rmse': prob.UnbiasedEnsembleMeanRMSE(
          ensemble_dim=ENS, skipna_ensemble=True)}, member_nan=True)
  yield ens_case('ens/nan_members_propagate', x_nan, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True)},
                 reduce_dims=['latitude', 'longitude'], member_nan=True)
  yield ens_case('ens/regions', x_major, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True),
      'unbiased_ssr': prob.UnbiasedSpreadSkillRatio(ensemble_dim=ENS)},
                 bins=['ens_regions_land'])
  # NaN targets with a mask coordinate: only the statistics whose expression
  # touches the targets carry the mask (CRPSSkill, UnbiasedEnsembleMeanSquared
  # Error); CRPSSpread and EnsembleVariance are functions of the predictions.
  yield ens_case('ens/nan_targets_default', x_major, ens_metrics,
                 nan_targets=True)
  yield ens_case('ens/nan_targets_masked', x_major, ens_metrics,
                 nan_targets=True, masked=True)
  yield ens_case('ens/nan_targets_skipna', x_major, ens_metrics,
                 nan_targets=True, skipna=True)
  yield ens_case('ens/regions_nan_targets_masked', x_major, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True),
      'unbiased_ssr': prob.UnbiasedSpreadSkillRatio(ensemble_dim=ENS)},
                 bins=['ens_regions_land'], nan_targets=True, masked=True)
  yield ens_case('ens/ensemble_averaged_rmse', x_major, {
      'rmse_members': prob.EnsembleAveragedMetric(det.RMSE(),
                                                  ensemble_dim=ENS)},
                 family='ens_averaged')
  yield ens_case('ens/ensemble_mean_rmse', x_major, {
      'rmse_mean': wrappers.WrappedMetric(
          det.RMSE(), [wrappers.EnsembleMean('predictions',
                                             ensemble_dim=ENS)])},
                 family='ens_mean')

  # -- operational ensemble sizes: 50 and 51 members --------------------------
  # (the sizes of the public benchmark's probabilistic suite,
  #  run_benchmark_evaluation.py:341-357; in the product these run the
  #  fixed-size sorting networks and the moments-only register kernel)
  x50 = da(inputs['x50_major'], D_ENS_MAJOR, **{ENS: np.arange(50)})
  x51 = da(inputs['x51_major'], D_ENS_MAJOR, **{ENS: np.arange(51)})
  x51_nan = da(with_nan(inputs['x51_major'], inputs['member_holes51']),
               D_ENS_MAJOR, **{ENS: np.arange(51)})
  yield ens_case('ens50/all_metrics', x50, ens_metrics, x_key='x50_major')
  yield ens_case('ens51/use_sort', x51, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True,
                                     use_sort=True)}, x_key='x51_major')
  yield ens_case('ens50/moments_only', x50, {
      'ens_var': prob.EnsembleRootMeanVariance(ensemble_dim=ENS),
      'unbiased_rmse': prob.UnbiasedEnsembleMeanRMSE(ensemble_dim=ENS),
      'unbiased_ssr': prob.UnbiasedSpreadSkillRatio(ensemble_dim=ENS)},
                 x_key='x50_major')
  yield ens_case('ens50/nan_targets_masked', x50, ens_metrics,
                 nan_targets=True, masked=True, x_key='x50_major')
  yield ens_case('ens51/nan_members_propagate', x51_nan, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True)},
                 reduce_dims=['latitude', 'longitude'], member_nan=True,
                 x_key='x51_major', holes_key='member_holes51')
  yield ens_case('ens50/regions', x50, {
      'crps_fair': prob.CRPSEnsemble(ensemble_dim=ENS, fair=True),
      'unbiased_ssr': prob.UnbiasedSpreadSkillRatio(ensemble_dim=ENS)},
                 bins=['ens_regions_land'], x_key='x50_major')

  # -- categorical: thresholded contingency tables, error exceedance ---------
  # (categorical.py:25-101,345-635; wrappers.py:50-88,214-267;
  #  deterministic.py:262-295)
  cat = ns.categorical
  both = [wrappers.ContinuousToBinary('both', RAIN_THRESHOLDS, 'threshold')]
  table_metrics = {
      name: wrappers.WrappedMetric(cls(), both) for name, cls in (
          ('csi', cat.CSI), ('accuracy', cat.Accuracy), ('recall', cat.Recall),
          ('far', cat.FalseAlarmRate), ('precision', cat.Precision),
          ('f1', cat.F1Score), ('frequency_bias', cat.FrequencyBias),
          ('hss', cat.HSS), ('ets', cat.ETS), ('sedi', cat.SEDI))}
  rain_p = {'total_precipitation_6hr': da(inputs['rain_p'], D2)}
  rain_p_nan = {'total_precipitation_6hr': da(
      with_nan(inputs['rain_p'], inputs['rain_p_holes']), D2)}
  rain_t = {'total_precipitation_6hr': da(inputs['rain_t'], D2)}
  rain_t_nan = {'total_precipitation_6hr': da(
      with_nan(inputs['rain_t'], inputs['rain_holes']), D2,
      mask=(D2, ~inputs['rain_holes']))}

  def cat_case(name, use_metrics=None, reduce_dims=None, weighted=True,
               nan_targets=False, nan_predictions=False, bins=None,
               kind='table', **flags):
    reduce_dims = reduce_dims or RD
    spec = dict(family='cat', kind=kind, reduce_dims=reduce_dims,
                weighted=weighted, masked=flags.get('masked', False),
                skipna=flags.get('skipna', False), bins=bins or [],
                nan_targets=nan_targets, nan_predictions=nan_predictions)
    aggregator = agg.Aggregator(
        reduce_dims=reduce_dims, weigh_by=area() if weighted else None,
        bin_by=_make_bins(ns, bins, inputs['land']) if bins else None,
        **flags)
    return (name, spec, use_metrics or table_metrics, aggregator,
            rain_p_nan if nan_predictions else rain_p,
            rain_t_nan if nan_targets else rain_t)

  yield cat_case('cat/table_weighted')
  yield cat_case('cat/table_unweighted_keep_init', weighted=False,
                 reduce_dims=['latitude', 'longitude'])
  yield cat_case('cat/table_nan_default', nan_targets=True,
                 nan_predictions=True)
  yield cat_case('cat/table_nan_masked', nan_targets=True, masked=True)
  yield cat_case('cat/table_nan_masked_nan_predictions', nan_targets=True,
                 nan_predictions=True, masked=True)
  yield cat_case('cat/table_nan_skipna', nan_targets=True,
                 nan_predictions=True, skipna=True)
  yield cat_case('cat/table_by_init_hour', bins=['init_hour_global'],
                 use_metrics={'csi': table_metrics['csi'],
                              'accuracy': table_metrics['accuracy']})
  yield cat_case('cat/table_regions', bins=['regions_land'],
                 use_metrics={'csi': table_metrics['csi'],
                              'accuracy': table_metrics['accuracy']})
  # targets that are binary already: only the predictions are thresholded
  event = {'total_precipitation_6hr': da(
      (inputs['rain_t'] > 0.5).astype(np.float32), D2)}
  name, spec, _, aggregator, p_, _ = cat_case(
      'cat/predictions_thresholded_binary_targets',
      use_metrics={'ets': wrappers.WrappedMetric(
          cat.ETS(), [wrappers.ContinuousToBinary(
              'predictions', [0.5, 2.0], 'threshold')])},
      kind='pred_only')
  yield (name, spec, {'ets': wrappers.WrappedMetric(
      cat.ETS(), [wrappers.ContinuousToBinary(
          'predictions', [0.5, 2.0], 'threshold')])}, aggregator, p_, event)
  exceedance = {'exceedance': det.ErrorExceedance(EXCEEDANCE_THRESHOLDS)}
  yield cat_case('cat/error_exceedance', use_metrics=exceedance,
                 kind='exceedance')
  yield cat_case('cat/error_exceedance_nan_skipna', use_metrics=exceedance,
                 kind='exceedance', nan_targets=True, nan_predictions=True,
                 skipna=True)
  yield cat_case('cat/error_exceedance_nan_default_keep_init',
                 use_metrics=exceedance, kind='exceedance', nan_targets=True,
                 reduce_dims=['latitude', 'longitude'])

  # relative intensity of the spatial means (deterministic.py:28-88); the
  # statistic has no grid dims left, the Aggregator reduces init_time
  intensity = {'relative_intensity': det.RelativeIntensity()}
  yield cat_case('cat/relative_intensity', use_metrics=intensity,
                 kind='relative_intensity', reduce_dims=['init_time'],
                 weighted=False)
  yield cat_case('cat/relative_intensity_masked', use_metrics=intensity,
                 kind='relative_intensity', reduce_dims=['init_time'],
                 weighted=False, nan_targets=True, masked=True)

  # error exceedance averaged over ensemble members (probabilistic.py:836-861)
  ens_exceedance = {'ens_exceedance': prob.EnsembleErrorExceedance(
      [1.0, 2.5, 6.0], ensemble_dim=ENS)}
  yield ens_case('cat/ensemble_error_exceedance', x_major, ens_exceedance,
                 family='ens_exceedance')
  yield ens_case('cat/ensemble_error_exceedance_nan_members', x_nan,
                 ens_exceedance, family='ens_exceedance', member_nan=True)

  # ensemble forecast against an ensemble of targets (probabilistic.py:135-145,
  # 199-204, 691-782); the two ensembles have different sizes
  y_ens = xr.DataArray(
      inputs['y_ens'], D_ENS_MAJOR,
      coords={'init_time': INIT, ENS: np.arange(N_TARGET_MEMBERS),
              'latitude': LAT, 'longitude': LON})
  for name, use_sort in (('ens/distance_to_target_ensemble', False),
                         ('ens/distance_to_target_ensemble_sorted', True)):
    case = ens_case(name, x_major, {
        'crps_distance': prob.CRPSEnsembleDistance(ensemble_dim=ENS,
                                                   use_sort=use_sort)},
                    family='ens_distance')
    yield case[:5] + ({'t2m': y_ens},)

  # -- SEEPS (categorical.py:104-304) -----------------------------------------
  var = 'total_precipitation_6hr'
  seeps_dims = ('hour', 'dayofyear', 'longitude', 'latitude')
  seeps_coords = {'hour': HOURS, 'dayofyear': np.arange(1, 367),
                  'longitude': LON, 'latitude': LAT}

  def seeps_var(rows):
    full = np.full((4, 366, NLON, NLAT), np.nan, np.float32)
    full[:, DOY_USED - 1] = rows
    return xr.DataArray(full, seeps_dims, coords=seeps_coords)

  seeps_climatology = xr.Dataset({
      f'{var}_seeps_threshold': seeps_var(inputs['seeps_threshold_rows']),
      f'{var}_seeps_dry_fraction': seeps_var(
          inputs['seeps_dry_fraction_rows'])})
  seeps_metrics = {'seeps': cat.SEEPS(
      variables=[var], climatology=seeps_climatology,
      dry_threshold_mm=[SEEPS_DRY_THRESHOLD_MM], min_p1=[0.1], max_p1=[0.85])}
  yield cat_case('seeps/masked_weighted', use_metrics=seeps_metrics,
                 kind='seeps', masked=True)
  yield cat_case('seeps/nan_targets_masked', use_metrics=seeps_metrics,
                 kind='seeps', nan_targets=True, masked=True)
  yield cat_case('seeps/nan_both_masked_keep_init', use_metrics=seeps_metrics,
                 kind='seeps', nan_targets=True, nan_predictions=True,
                 masked=True, reduce_dims=['latitude', 'longitude'])
  yield cat_case('seeps/regions_masked', use_metrics=seeps_metrics,
                 kind='seeps', bins=['regions_land'], masked=True)
  yield cat_case('seeps/default_propagates', use_metrics=seeps_metrics,
                 kind='seeps')
  yield cat_case('seeps/skipna_unweighted', use_metrics=seeps_metrics,
                 kind='seeps', nan_targets=True, skipna=True, weighted=False)


def _make_bins(ns, names, land_values):
  xr, binning = ns.xr, ns.binning
  land = xr.DataArray(land_values, ('latitude', 'longitude'),
                      coords={'latitude': LAT, 'longitude': LON})
  out = []
  for name in names:
    if name == 'regions':
      out.append(binning.Regions(REGIONS))
    elif name == 'regions_land':
      out.append(binning.Regions(REGIONS, land_sea_mask=land))
    elif name == 'ens_regions_land':
      out.append(binning.Regions(ENS_REGIONS, land_sea_mask=land))
    elif name == 'landsea_global':
      out.append(binning.LandSea(land.astype(np.float32),
                                 include_global_mask=True))
    elif name == 'init_hour':
      out.append(binning.ByTimeUnit('hour', 'init_time'))
    elif name == 'init_hour_global':
      out.append(binning.ByTimeUnit('hour', 'init_time', add_global_bin=True))
    elif name == 'valid_month':
      out.append(binning.ByTimeUnit('month', 'valid_time'))
    elif name == 'lead_sets':
      out.append(binning.ByTimeUnitSets(LEAD_SETS, 'hour', 'lead_time',
                                        add_global_bin=True))
    elif name == 'level_sets':
      out.append(binning.BySets(LEVEL_SETS, 'level', bin_dim_name='level_set',
                                add_set_complements=True))
    elif name == 'lat30':
      out.append(binning.LatitudeBins(30))
    elif name == 'lon90':
      out.append(binning.LongitudeBins(90))
    else:
      raise KeyError(name)
  return out


def chunked_case(ns, inputs):
  """The det/weighted case evaluated one init_time at a time and combined
  with AggregationState.__add__ (aggregation.py:84-110; the identity
  beam_pipeline_test.py:82-170 checks).  Returns (metrics, state)."""
  base = ns.base
  for name, _, metrics, aggregator, predictions, targets in build_cases(
      ns, inputs):
    if name == 'det/weighted':
      break
  total = ns.aggregation.AggregationState.zero()
  for i in range(len(INIT)):
    chunk_p = {k: v.isel(init_time=[i]) for k, v in predictions.items()}
    chunk_t = {k: v.isel(init_time=[i]) for k, v in targets.items()}
    statistics = base.compute_unique_statistics_for_all_metrics(
        metrics, chunk_p, chunk_t)
    total = total + aggregator.aggregate_statistics(statistics)
  return metrics, total


def namespace(**modules):
  return types.SimpleNamespace(**modules)


def mask_cases(ns):
  """(name, binning instance, statistic) for the coordinate-value binnings on
  a sparse-style statistic: one 'index' dim with non-dimension coordinates."""
  xr, binning = ns.xr, ns.binning
  n = 12
  lead = (np.array([0, 6, 6, 12, 24, 24, 30, 0, 12, 6, 48, 24]) *
          np.timedelta64(1, 'h')).astype('timedelta64[ns]')
  valid = np.datetime64('2020-12-30T00', 'ns') + np.arange(n) * np.timedelta64(
      7, 'h')
  stat = xr.DataArray(
      np.arange(n, dtype=np.float32), ('index',),
      coords={'index': np.arange(n),
              'lead_time': ('index', lead),
              'valid_time': ('index', valid),
              'seconds': ('index', np.arange(n) * 1800 + 10),
              'elevation': ('index', np.linspace(-5.0, 2500.0, n)),
              'station': ('index', np.array(list('abcabcabcabd')))})
  return [
      ('exact_lead', binning.ByExactCoord('lead_time'), stat),
      ('exact_lead_global',
       binning.ByExactCoord('lead_time', add_global_bin=True), stat),
      ('exact_station_global',
       binning.ByExactCoord('station', add_global_bin=True), stat),
      ('lead_day', binning.ByTimeUnit('day', 'lead_time'), stat),
      ('valid_hour_global',
       binning.ByTimeUnit('hour', 'valid_time', add_global_bin=True), stat),
      ('valid_dayofyear', binning.ByTimeUnit('dayofyear', 'valid_time'), stat),
      ('valid_year', binning.ByTimeUnit('year', 'valid_time'), stat),
      ('hour_sets', binning.ByTimeUnitSets(
          {'night': [0, 1, 2, 3, 4, 5, 21, 22, 23], 'noon': 12, 'empty': [9]},
          'hour', 'valid_time', add_global_bin=True), stat),
      ('seconds_minute', binning.ByTimeUnitFromSeconds(
          'minute', 'seconds', bins=[0, 30, 90, 300]), stat),
      ('seconds_hour', binning.ByTimeUnitFromSeconds('hour', 'seconds'), stat),
      ('elevation_bins', binning.ByCoordBins(
          'elevation', np.array([0.0, 500.0, 1000.0, 3000.0])), stat),
      ('elevation_bins_global', binning.ByCoordBins(
          'elevation', np.array([0.0, 500.0, 1000.0, 3000.0]),
          add_global_bin=True), stat),
      ('station_sets', binning.BySets(
          {'ab': ['a', 'b'], 'd': 'd'}, 'station', bin_dim_name='station_set',
          add_set_complements=True, add_global_bin=True), stat),
  ]
