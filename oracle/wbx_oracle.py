"""CPU oracle for the WeatherBench-X statistic + aggregation hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``weatherbenchx_b200/`` does.

It restates, with plain NumPy on (ndarray, dim-names) pairs, the arithmetic the
reference performs through xarray -> NumPy for the path named by
BASELINE.json's north_star.  Every function cites the reference lines it
follows (paths relative to /root/reference/weatherbenchX).

Pinning status
--------------
* PINNED TO OUTPUTS OF THE REFERENCE'S OWN CODE, generated in the build
  container: ``tests/golden/make_reference_golden.py`` imports the unmodified
  modules of /root/reference/weatherbenchX (aggregation, weighting, binning,
  metrics/{base,deterministic,probabilistic,wrappers,categorical}) and stores
  the AggregationState and metric values of 62 cases (all NaN modes, ACC with a
  day-of-year climatology across 29 February, regions / land-sea / band bins,
  both ensemble layouts, pairwise and sorted CRPS, skipna_ensemble, ensemble
  moments, ensemble-averaged and ensemble-mean metrics, thresholded
  contingency tables, error exceedance and SEEPS, chunk combine, five latitude grids) in ``tests/golden/reference_golden.npz``.
  ``tests/test_reference_golden.py`` checks that this oracle reproduces every
  stored array.  Caveat, stated wherever the vectors are used: xarray, jax and
  absl are not installable in the container, so the reference ran on the
  stand-in modules of ``tests/golden/reference_runtime.py`` (this repo's
  labelled-array container under the name ``xarray``; NumPy under ``jax.numpy``)
  -- control flow and arithmetic are the reference's, label alignment /
  broadcasting / ``xr.dot`` (= ``np.einsum``) are the stand-in's.
* Also pinned against every inline known-answer test the reference holds for
  this path (aggregation_test.py:69-246, weighting_test.py:24-46,
  metrics/metrics_test.py:44-98,603-660,983-1006,1199-1308), re-expressed
  without xarray in ``tests/test_oracle_pins.py``, and against
  ``tests/golden/hotpath_golden.npz`` (``tests/golden/make_golden.py``
  evaluates the reference's formulas in float64 by brute force, independently
  of the vectorised code here).
* ``zonal_energy_spectrum``: PARITY UNPINNED.  /root/reference contains no
  implementation, call site or test of an energy spectrum (SURVEY.md finding
  2); the definition restated here is WeatherBench 2's
  ``derived_variables.ZonalEnergySpectrum`` (not vendored, not importable) and
  is validated only against ``numpy.fft`` identities (Parseval, single
  sinusoid, constant field).

Numerics: statistics are evaluated in the input dtype (float32 in, float32
out -- exactly what NumPy ufuncs do for the reference); reductions follow
``np.einsum`` type promotion (float64 as soon as a float64 weight takes part,
which is the GridAreaWeighting case).  ``aggregate(..., exact=True)`` instead
accumulates everything in float64 and is what parity tests compare against.
"""

from __future__ import annotations

import itertools
import warnings
from typing import Mapping, Sequence

import numpy as np

EARTH_RADIUS_M = 1000.0 * (6357.0 + 6378.0) / 2.0  # WB2 derived_variables


# ---------------------------------------------------------------------------
# Per-gridpoint deterministic statistics
# ---------------------------------------------------------------------------


def error(p: np.ndarray, t: np.ndarray) -> np.ndarray:
  """metrics/deterministic.py:94-100 -- ``predictions - targets``."""
  return p - t


def absolute_error(p: np.ndarray, t: np.ndarray) -> np.ndarray:
  """metrics/deterministic.py:106-112 -- ``abs(predictions - targets)``."""
  return np.abs(p - t)


def squared_error(p: np.ndarray, t: np.ndarray) -> np.ndarray:
  """metrics/deterministic.py:118-123 -- ``(predictions - targets) ** 2``."""
  return (p - t) ** 2


def squared_prediction_anomaly(p, t, c) -> np.ndarray:
  """metrics/deterministic.py:225-232 -- ``(p - clim) ** 2``."""
  del t
  return (p - c) ** 2


def squared_target_anomaly(p, t, c) -> np.ndarray:
  """metrics/deterministic.py:238-245 -- ``(t - clim) ** 2``."""
  del p
  return (t - c) ** 2


def anomaly_covariance(p, t, c) -> np.ndarray:
  """metrics/deterministic.py:251-259 -- ``(p - clim) * (t - clim)``."""
  return (p - c) * (t - c)


def wind_vector_squared_error(pu, pv, tu, tv) -> np.ndarray:
  """metrics/deterministic.py:206-219."""
  return (pu - tu) ** 2 + (pv - tv) ** 2


DETERMINISTIC_STATISTICS = {
    'Error': error,
    'AbsoluteError': absolute_error,
    'SquaredError': squared_error,
}
CLIMATOLOGY_STATISTICS = {
    'SquaredPredictionAnomaly': squared_prediction_anomaly,
    'SquaredTargetAnomaly': squared_target_anomaly,
    'AnomalyCovariance': anomaly_covariance,
}


def dayofyear_and_hour(valid_time: np.ndarray):
  """``valid_time.dt.dayofyear`` / ``.dt.hour`` (metrics/base.py:400-402)."""
  vt = np.asarray(valid_time).astype('datetime64[ns]')
  year_start = vt.astype('datetime64[Y]').astype('datetime64[ns]')
  day = vt.astype('datetime64[D]')
  doy = (day - year_start.astype('datetime64[D]')).astype(np.int64) + 1
  hour = ((vt - day.astype('datetime64[ns]')) // np.timedelta64(1, 'h')).astype(
      np.int64)
  return doy, hour


def align_climatology(clim: np.ndarray, clim_dims: Sequence[str],
                      clim_coords: Mapping[str, np.ndarray],
                      init_time: np.ndarray, lead_time: np.ndarray):
  """Vectorised ``climatology.sel(dayofyear=..., hour=...)``.

  metrics/base.py:383-403: valid_time = init_time + lead_time, then a label
  gather along (dayofyear[, hour]).  Returns (array, dims) with the two
  climatology time dims replaced by (init_time, lead_time).
  """
  valid = np.asarray(init_time)[:, None] + np.asarray(lead_time)[None, :]
  doy, hour = dayofyear_and_hour(valid)
  clim_dims = tuple(clim_dims)
  doy_axis = clim_dims.index('dayofyear')
  doy_labels = np.asarray(clim_coords['dayofyear'])
  doy_pos = np.searchsorted(doy_labels, doy)
  if not np.array_equal(doy_labels[doy_pos], doy):
    raise KeyError('dayofyear label missing from climatology')
  moved = np.moveaxis(clim, doy_axis, 0)
  rest = [d for d in clim_dims if d != 'dayofyear']
  if 'hour' in clim_dims:
    hour_axis = rest.index('hour') + 1
    moved = np.moveaxis(moved, hour_axis, 1)
    rest.remove('hour')
    hour_labels = np.asarray(clim_coords['hour'])
    hour_pos = np.searchsorted(hour_labels, hour)
    if not np.array_equal(hour_labels[hour_pos], hour):
      raise KeyError('hour label missing from climatology')
    out = moved[doy_pos, hour_pos]
  else:
    out = moved[doy_pos]
  return out, ('init_time', 'lead_time') + tuple(rest)


# ---------------------------------------------------------------------------
# Weighting
# ---------------------------------------------------------------------------


def latitude_cell_bounds(x: np.ndarray) -> np.ndarray:
  """weighting.py:62-79 -- midpoints, end cells clipped to +-pi/2."""
  x = np.asarray(x)
  if not np.all(np.diff(x) > 0):
    raise AssertionError('Points must be increasing.')
  d = np.diff(x)
  lo = max(x[0] - d[0] / 2, -np.pi / 2)
  hi = min(x[-1] + d[-1] / 2, np.pi / 2)
  return np.concatenate([[lo], (x[:-1] + x[1:]) / 2, [hi]]).astype(x.dtype)


def cell_area_from_latitude(points: np.ndarray) -> np.ndarray:
  """weighting.py:82-88 -- sin(upper) - sin(lower)."""
  b = latitude_cell_bounds(points)
  return np.sin(b[1:]) - np.sin(b[:-1])


def grid_area_weights(latitude_deg: np.ndarray,
                      return_normalized: bool = True) -> np.ndarray:
  """weighting.py:105-130 -- GridAreaWeighting.weights for a latitude coord."""
  lat = np.asarray(latitude_deg)
  diff = np.diff(lat)
  if not (np.all(diff > 0) or np.all(diff < 0)):
    raise AssertionError(f'Points must be strictly monotonic: {lat}')
  reverse = lat[0] > lat[1]
  if reverse:
    lat = lat[::-1]
  w = cell_area_from_latitude(np.deg2rad(lat))
  if reverse:
    w = w[::-1]
  if return_normalized:
    w = w / np.mean(w)
  return w


# ---------------------------------------------------------------------------
# Binning (region masks) -- binning.py:52-89,172-201
# ---------------------------------------------------------------------------


def region_mask(lat: np.ndarray, lon: np.ndarray, lat_lims, lon_lims):
  """binning.py:52-89 -- boolean [lat, lon] rectangle, lon wraps mod 360."""
  if lat_lims[0] >= lat_lims[1]:
    raise ValueError('lat_lims[0] must be smaller than lat_lims[1]')
  lat_m = np.logical_and(lat >= lat_lims[0], lat <= lat_lims[1])
  lon = np.mod(lon, 360)
  l0, l1 = np.mod(lon_lims[0], 360), np.mod(lon_lims[1], 360)
  if l1 > l0:
    lon_m = np.logical_and(lon >= l0, lon <= l1)
  else:
    lon_m = np.logical_or(lon <= l1, lon >= l0)
  return np.logical_and(lat_m[:, None], lon_m[None, :])


def regions_masks(lat, lon, regions: Mapping[str, tuple], land_sea_mask=None):
  """binning.py:172-201 -- stacked [region, lat, lon] masks (+ *_land)."""
  names = list(regions)
  masks = np.stack([region_mask(lat, lon, *regions[n]) for n in names])
  if land_sea_mask is not None:
    land = masks & np.asarray(land_sea_mask, dtype=bool)[None]
    masks = np.concatenate([masks, land])
    names = names + [f'{n}_land' for n in names]
  return masks, names


# ---------------------------------------------------------------------------
# Aggregation
# ---------------------------------------------------------------------------


def _einsum(operands, out_dims, dtype=None):
  """``xr.dot`` restated: einsum over named dims (aggregation.py:334-335)."""
  letters = {}
  subs = []
  arrays = []
  for arr, dims in operands:
    for d in dims:
      if d not in letters:
        letters[d] = chr(ord('a') + len(letters))
    subs.append(''.join(letters[d] for d in dims))
    arrays.append(np.asarray(arr))
  out = ''.join(letters[d] for d in out_dims)
  expr = ','.join(subs) + '->' + out
  if dtype is not None:
    arrays = [a.astype(dtype) for a in arrays]
  return np.einsum(expr, *arrays)


def aggregate(stat: np.ndarray, dims: Sequence[str],
              reduce_dims: Sequence[str], *,
              weights: Sequence[tuple] = (),
              bin_masks: Sequence[tuple] = (),
              mask: np.ndarray | None = None,
              mask_dims: Sequence[str] | None = None,
              masked: bool = False, skipna: bool = False,
              exact: bool = True):
  """Aggregator.aggregate_stat_var + aggregation_fn (aggregation.py:297-366).

  Args:
    stat: per-gridpoint statistic values.
    dims: dim names of ``stat``.
    reduce_dims: dims summed over.
    weights: sequence of (array, dims) broadcastable against stat.
    bin_masks: sequence of (bool array, dims) where dims include the bin dim.
    mask: the 'mask' coordinate (True = valid), dims ``mask_dims``.
    masked / skipna: Aggregator flags.
    exact: accumulate in float64 (parity target).  False follows einsum dtype
      promotion of the reference (float32 when nothing is float64).

  Returns:
    (sum_weighted_statistics, sum_weights, out_dims) or None when the
    aggregation does not apply (aggregation.py:305-309,326-330).
  """
  dims = tuple(dims)
  stat = np.asarray(stat)
  if not set(reduce_dims).issubset(dims):
    return None
  for _, bdims in bin_masks:
    extra = [d for d in bdims if d not in dims]
    if len(extra) != 1:
      # exactly one new (bin) dim is allowed; otherwise not applicable.
      return None
  if masked and mask is not None:
    m = np.broadcast_to(_expand(mask, mask_dims, dims), stat.shape)
    if skipna:
      m = m & ~np.isnan(stat)
    stat = np.where(m, stat, 0).astype(stat.dtype)
    mask_arr = m
  elif skipna:
    mask_arr = ~np.isnan(stat)
    stat = np.where(mask_arr, stat, 0).astype(stat.dtype)
  else:
    mask_arr = np.ones_like(stat)
  kept = [d for d in dims if d not in reduce_dims]
  bin_dims = []
  for _, bdims in bin_masks:
    bin_dims += [d for d in bdims if d not in dims]
  out_dims = tuple(kept + bin_dims)
  ops_tail = [(w, wd) for w, wd in weights] + [(b, bd) for b, bd in bin_masks]
  dtype = np.float64 if exact else None
  sws = _einsum([(stat, dims)] + ops_tail, out_dims, dtype)
  sw = _einsum([(mask_arr.astype(stat.dtype), dims)] + ops_tail, out_dims,
               dtype)
  return sws, sw, out_dims


def _expand(arr, arr_dims, dims):
  """Insert singleton axes so ``arr`` broadcasts against ``dims`` order."""
  arr = np.asarray(arr)
  arr_dims = tuple(arr_dims)
  present = [d for d in dims if d in arr_dims]
  arr = np.transpose(arr, [arr_dims.index(d) for d in present])
  shape = [arr.shape[present.index(d)] if d in present else 1 for d in dims]
  return arr.reshape(shape)


def mean_statistic(sum_ws, sum_w):
  """AggregationState.mean_statistics (aggregation.py:112-120)."""
  with np.errstate(invalid='ignore', divide='ignore'):
    return sum_ws / sum_w


def rmse_from_mean(mean_se):
  """metrics/deterministic.py:319-324."""
  return np.sqrt(mean_se)


def acc_from_means(cov, spa, sta):
  """metrics/deterministic.py:392-400."""
  with np.errstate(invalid='ignore', divide='ignore'):
    return cov / (np.sqrt(spa) * np.sqrt(sta))


def crps_from_means(skill, spread):
  """metrics/probabilistic.py:683-688."""
  return skill - 0.5 * spread


# ---------------------------------------------------------------------------
# CRPS statistics
# ---------------------------------------------------------------------------


def crps_skill(x: np.ndarray, y: np.ndarray, ens_axis: int,
               skipna_ensemble: bool = False) -> np.ndarray:
  """metrics/probabilistic.py:129-145 -- mean_m |x_m - y| (y has no ens dim)."""
  d = np.abs(x - np.expand_dims(y, ens_axis))
  if skipna_ensemble:
    with np.errstate(invalid='ignore'):
      import warnings
      with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        return np.nanmean(d, axis=ens_axis)
  return d.mean(axis=ens_axis)


def rankdata(x: np.ndarray, axis: int) -> np.ndarray:
  """metrics/probabilistic.py:148-158 -- ordinal ranks via double argsort."""
  x = np.swapaxes(np.asarray(x), axis, -1)
  j = np.argsort(x, axis=-1)
  ranks = np.empty(j.shape, dtype=int)
  np.put_along_axis(
      ranks, j,
      np.broadcast_to(np.arange(1, x.shape[-1] + 1, dtype=int), x.shape),
      axis=-1)
  return np.swapaxes(ranks, axis, -1)


def crps_spread(x: np.ndarray, ens_axis: int, fair: bool = True,
                use_sort: bool = False,
                skipna_ensemble: bool = False) -> np.ndarray:
  """metrics/probabilistic.py:194-247 -- E|X - X'| sample estimate."""
  m = x.shape[ens_axis]
  if skipna_ensemble:
    n = np.sum(~np.isnan(x), axis=ens_axis)
  else:
    n = m
    if n < 2:
      raise ValueError('Cannot estimate CRPS spread with n_ensemble < 2.')
  if use_sort:
    if skipna_ensemble:
      raise ValueError('skipna_ensemble is not supported with use_sort=True.')
    rank = rankdata(x, ens_axis)
    return 2 * ((2 * rank - n - 1) * x).mean(axis=ens_axis) / (n - int(fair))
  xm = np.moveaxis(x, ens_axis, -1)
  # Same arithmetic as the reference's [M, M, ...] temporary, but chunked over
  # the leading dims so the oracle stays within memory at M = 50.
  flat = xm.reshape(-1, m)
  out = np.empty(flat.shape[0], dtype=x.dtype)
  step = max(1, (1 << 22) // (m * m))
  for s in range(0, flat.shape[0], step):
    blk = flat[s:s + step]
    d = np.abs(blk[:, :, None] - blk[:, None, :])
    if skipna_ensemble:
      out[s:s + step] = np.nansum(d, axis=(1, 2))
    else:
      out[s:s + step] = d.sum(axis=(1, 2))
  out = out.reshape(xm.shape[:-1])
  with np.errstate(invalid='ignore', divide='ignore'):
    return out / (n * (n - int(fair)))


def ensemble_variance(x: np.ndarray, ens_axis: int,
                      skipna_ensemble: bool = False) -> np.ndarray:
  """EnsembleVariance._compute_per_variable (probabilistic.py:266-273):
  predictions.var(dim=ensemble_dim, ddof=1, skipna=skipna_ensemble)."""
  with np.errstate(all='ignore'), warnings.catch_warnings():
    warnings.simplefilter('ignore')
    fn = np.nanvar if skipna_ensemble else np.var
    return fn(x, axis=ens_axis, ddof=1)


def unbiased_ensemble_mean_squared_error(
    x: np.ndarray, y: np.ndarray, ens_axis: int,
    skipna_ensemble: bool = False) -> np.ndarray:
  """UnbiasedEnsembleMeanSquaredError (probabilistic.py:300-336) for targets
  without an ensemble dim: (mean - y)**2 - var / n, with n the per-point count
  of non-NaN members when skipna_ensemble."""
  with np.errstate(all='ignore'), warnings.catch_warnings():
    warnings.simplefilter('ignore')
    if skipna_ensemble:
      mean = np.nanmean(x, axis=ens_axis)
      n = np.sum(~np.isnan(x), axis=ens_axis)
    else:
      mean = np.mean(x, axis=ens_axis)
      n = x.shape[ens_axis]
    var = ensemble_variance(x, ens_axis, skipna_ensemble)
    return (mean - y) ** 2 - var / n


def ensemble_mean(x: np.ndarray, ens_axis: int,
                  skipna: bool = False) -> np.ndarray:
  """wrappers.EnsembleMean.transform_fn (wrappers.py:145-148):
  da.mean(ensemble_dim, skipna=skipna)."""
  with np.errstate(all='ignore'), warnings.catch_warnings():
    warnings.simplefilter('ignore')
    return (np.nanmean if skipna else np.mean)(x, axis=ens_axis)


def wind_vector_squared_error(pu, tu, pv, tv) -> np.ndarray:
  """WindVectorSquaredError.compute (deterministic.py:211-218)."""
  return (pu - tu) ** 2 + (pv - tv) ** 2


def passthrough(source: np.ndarray, other: np.ndarray,
                copy_nans: bool = False) -> np.ndarray:
  """Prediction/TargetPassthrough (deterministic.py:138-147,162-171):
  source + zeros_like(other), optionally NaN wherever ``other`` is NaN."""
  result = source + np.zeros_like(other)
  if copy_nans:
    result = np.where(~np.isnan(other), result, np.nan).astype(result.dtype)
  return result


def relative_intensity(p: np.ndarray, t: np.ndarray, spatial_axes,
                       mask: np.ndarray | None = None):
  """metrics/deterministic.py:49-86.  Returns (result, result mask or None);
  arithmetic in the input dtype, as NumPy does for the reference."""
  axes = tuple(spatial_axes)
  epsilon = 1e-6
  with np.errstate(all='ignore'):
    if mask is not None:
      mask = mask == 1
      count = mask.sum(axis=axes)
      prediction_mean = np.where(mask, p, 0).sum(axis=axes) / count
      prediction_mean = np.where(count > 0, prediction_mean, 0.0)
      target_mean = np.where(mask, t, 0).sum(axis=axes) / count
      target_mean = np.where(count > 0, target_mean, 0.0)
      ratio = (prediction_mean + epsilon) / (target_mean + epsilon)
      return np.abs(ratio - 1), (count > 0).astype(int)
    ratio = (p.mean(axis=axes) + epsilon) / (t.mean(axis=axes) + epsilon)
    return np.abs(ratio - 1), None


# ---------------------------------------------------------------------------
# Categorical statistics (thresholded contingency tables, error exceedance)
# ---------------------------------------------------------------------------


def binarize_thresholds(x: np.ndarray, thresholds) -> np.ndarray:
  """metrics/wrappers.py:85-88: ``(x > threshold).where(~isnan(x))`` as
  float32 with the thresholds along a NEW TRAILING axis.  The thresholds are
  float64 (a Python list turned into a DataArray), so NumPy promotes the
  float32 field for the comparison; a NaN threshold compares False."""
  t = np.asarray(thresholds, dtype=np.float64)
  with np.errstate(invalid='ignore'):
    out = (x[..., None] > t).astype(np.float32)
  out[np.isnan(x)] = np.nan
  return out


def contingency_table(bp: np.ndarray, bt: np.ndarray) -> dict:
  """metrics/categorical.py:25-101 on binary (0/1/NaN) float inputs:
  ``astype(bool)`` products, NaN where ``predictions * targets`` is NaN."""
  with np.errstate(invalid='ignore'):
    p, t = bp.astype(bool), bt.astype(bool)
    bad = np.isnan(bp * bt)
  table = {'TruePositives': p & t, 'TrueNegatives': ~p & ~t,
           'FalsePositives': p & ~t, 'FalseNegatives': ~p & t}
  out = {}
  for name, v in table.items():
    v = v.astype(np.float32)
    v[bad] = np.nan
    out[name] = v
  return out


def error_exceedance(p: np.ndarray, t: np.ndarray, thresholds) -> np.ndarray:
  """metrics/deterministic.py:285-295: ``abs(p - t) > threshold`` as float,
  NaN where the error or the threshold is NaN; thresholds along a new trailing
  axis."""
  thr = np.asarray(thresholds, dtype=np.float64)
  with np.errstate(invalid='ignore'):
    abs_error = np.abs(p - t)
    out = (abs_error[..., None] > thr).astype(np.float64)
  out[np.isnan(abs_error)] = np.nan
  out[..., np.isnan(thr)] = np.nan
  return out


def categorical_metric(name: str, tp, fp, fn, tn=None):
  """metrics/categorical.py:345-635 on the mean statistics."""
  with np.errstate(all='ignore'):
    if name == 'csi':
      return tp / (tp + fp + fn)
    if name == 'accuracy':
      return (tp + tn) / (tp + fp + fn + tn)
    if name == 'recall':
      return tp / (tp + fn)
    if name == 'far':
      return fp / (tp + fp)
    if name == 'precision':
      return tp / (tp + fp)
    if name == 'f1':
      return 2 * tp / (2 * tp + fp + fn)
    if name == 'frequency_bias':
      return (tp + fp) / (tp + fn)
    if name == 'hss':
      return 2 * (tp * tn - fp * fn) / (
          (tp + fn) * (fn + tn) + (tp + fp) * (fp + tn))
    if name == 'ets':
      tp_random = ((tp + fp) * (tp + fn)) / (tp + fp + fn + tn)
      return (tp - tp_random) / (tp + fp + fn - tp_random)
    if name == 'sedi':
      h = np.clip(tp / (tp + fn), 1e-6, 1 - 1e-6)
      f = np.clip(fp / (fp + tn), 1e-6, 1 - 1e-6)
      num = np.log(f) - np.log(h) + np.log(1 - h) - np.log(1 - f)
      den = np.log(h) + np.log(f) + np.log(1 - h) + np.log(1 - f)
      return num / den
  raise KeyError(name)


def seeps_categories(x: np.ndarray, wet_threshold: np.ndarray,
                     dry_threshold_mm: float) -> np.ndarray:
  """metrics/categorical.py:217-241: [dry, light, heavy] indicators stacked on
  a new leading axis, float64 with NaN where ``x`` is NaN.  ``dry_threshold``
  is a Python float, so NumPy compares the float32 field in float32."""
  dry_threshold = dry_threshold_mm / 1000.0
  with np.errstate(invalid='ignore'):
    dry = x <= dry_threshold
    light = np.logical_and(x > dry_threshold, x < wet_threshold)
    heavy = x >= wet_threshold
  out = np.stack([dry, light, heavy]).astype(np.float64)
  out[:, np.isnan(x)] = np.nan
  return out


def seeps_p1(dry_fraction: np.ndarray, time_axes) -> np.ndarray:
  """metrics/categorical.py:268-272: ``.mean(('hour', 'dayofyear'))`` -- xarray
  skips NaN by default, i.e. ``np.nanmean`` in the dtype of the input."""
  with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    return np.nanmean(dry_fraction, axis=tuple(time_axes))


def seeps(p: np.ndarray, t: np.ndarray, wet_threshold: np.ndarray,
          p1: np.ndarray, dry_threshold_mm: float = 0.25, min_p1: float = 0.1,
          max_p1: float = 0.85):
  """metrics/categorical.py:243-296 for aligned inputs: ``wet_threshold``
  already gathered to the valid times and broadcastable to ``p``; ``p1``
  broadcastable to ``p`` (trailing grid dims).  Returns (score, p1 mask): the
  score is NaN where an input is NaN or p1 lies outside [min_p1, max_p1]."""
  f_cat = seeps_categories(p, wet_threshold, dry_threshold_mm)
  t_cat = seeps_categories(t, wet_threshold, dry_threshold_mm)
  contingency = f_cat[:, None] * t_cat[None, :]          # [forecast, truth, ...]
  with np.errstate(divide='ignore', invalid='ignore'):
    zero = np.zeros_like(p1)
    matrix = [[zero, 1 / (1 - p1), 4 / (1 - p1)],
              [1 / p1, zero, 3 / (1 - p1)],
              [1 / p1 + 3 / (2 + p1), 3 / (2 + p1), zero]]
    matrix = 0.5 * np.stack([np.stack(row) for row in matrix])
    matrix = np.broadcast_to(
        matrix.reshape(matrix.shape[:2] + (1,) * (p.ndim - p1.ndim)
                       + p1.shape), contingency.shape)
    result = np.einsum('ft...,ft...->...', contingency, matrix)
  with np.errstate(invalid='ignore'):
    mask = (p1 >= min_p1) & (p1 <= max_p1)
  result = np.where(np.broadcast_to(mask, result.shape), result, np.nan)
  return result, mask


def crps_spread_brute_force(x: np.ndarray, ens_axis: int, fair: bool):
  """metrics/metrics_test.py:603-608 -- the reference's own brute force."""
  m = x.shape[ens_axis]
  xm = np.moveaxis(np.asarray(x, dtype=np.float64), ens_axis, -1)
  acc = np.zeros(xm.shape[:-1])
  for i, j in itertools.product(range(m), range(m)):
    acc += np.abs(xm[..., i] - xm[..., j])
  return acc / (m * m) * (m / (m - int(fair)))


# ---------------------------------------------------------------------------
# Zonal energy spectrum (PARITY UNPINNED -- see module docstring)
# ---------------------------------------------------------------------------


def zonal_energy_spectrum(f: np.ndarray, latitude_deg: np.ndarray,
                          lat_axis: int = -2, lon_axis: int = -1):
  """WeatherBench 2 ``ZonalEnergySpectrum`` restated.

  Per latitude circle: F = rfft(f, axis=lon, norm='forward');
  S[0] = C |F_0|^2, S[k>0] = 2 C |F_k|^2 with C(lat) = 2 pi R cos(lat).
  Returns an array with the lon axis replaced by zonal_wavenumber 0..N/2.
  """
  f = np.asarray(f)
  n = f.shape[lon_axis]
  spec = np.fft.rfft(f.astype(np.float64), axis=lon_axis, norm='forward')
  power = spec.real ** 2 + spec.imag ** 2
  factor = np.full(n // 2 + 1, 2.0)
  factor[0] = 1.0
  shape = [1] * f.ndim
  shape[lon_axis] = n // 2 + 1
  power = power * factor.reshape(shape)
  circ = 2 * np.pi * EARTH_RADIUS_M * np.cos(np.deg2rad(latitude_deg))
  shape = [1] * f.ndim
  shape[lat_axis] = len(latitude_deg)
  return power * circ.reshape(shape)


# ---------------------------------------------------------------------------
# "Reference-mirroring" CPU path used as the timed CPU baseline.
# Same sequence of full-size temporaries as the reference: statistic ufuncs,
# ones_like + astype, and two einsum contractions per statistic
# (metrics/deterministic.py:118-123, aggregation.py:337-366).  xarray's label
# alignment overhead is NOT included (xarray is not installable here).
# ---------------------------------------------------------------------------


def reference_path_rmse(p: np.ndarray, t: np.ndarray, w_lat: np.ndarray,
                        expr: str = 'tyx,y->'):
  """One variable, one chunk: returns (sum_weighted_se, sum_weights)."""
  se = (p - t) ** 2
  ones = np.ones_like(se)
  sws = np.einsum(expr, se, w_lat)
  sw = np.einsum(expr, ones.astype(se.dtype), w_lat)
  return sws, sw


def reference_path_crps(x: np.ndarray, y: np.ndarray, w_lat: np.ndarray,
                        use_sort: bool = False, fair: bool = True):
  """x: [t, y, x, m], y: [t, y, x].  Returns the four aggregated sums."""
  skill = crps_skill(x, y, ens_axis=-1)
  spread = crps_spread(x, ens_axis=-1, fair=fair, use_sort=use_sort)
  out = []
  for stat in (skill, spread):
    ones = np.ones_like(stat)
    out.append(np.einsum('tyx,y->', stat, w_lat))
    out.append(np.einsum('tyx,y->', ones.astype(stat.dtype), w_lat))
  return tuple(out)
