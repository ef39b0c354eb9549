"""Deterministic statistics and metrics served by the fused CUDA kernel.

Mirrors the hot-path subset of
/root/reference/weatherbenchX/metrics/deterministic.py: Error :91-100,
AbsoluteError :103-112, SquaredError :115-123, SquaredPredictionAnomaly
:222-232, SquaredTargetAnomaly :235-245, AnomalyCovariance :248-259, the
aliases Bias / MAE / MSE :305-307, RMSE :312-324, ACC :374-400 and
PredictionActivity :403-425, plus PredictionPassthrough / TargetPassthrough
:126-171 (aliases PredictionAverage / TargetAverage :308-309),
WindVectorSquaredError :174-219, WindVectorRMSE :327-371, ErrorExceedance
:262-295 and RelativeIntensity :28-88.  unique_name of every statistic follows the reference.
"""

from __future__ import annotations

from typing import Hashable, Mapping, Sequence, Union

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import LazyPassthrough
from weatherbenchx_b200.lazy import LazyStatistic
from weatherbenchx_b200.lazy import LazySumStatistic
from weatherbenchx_b200.metrics import base


class RelativeIntensity(base.PerVariableStatistic):
  """``abs((mean(predictions) + eps) / (mean(targets) + eps) - 1)`` with the
  means taken over ``spatial_dims`` (deterministic.py:28-88; for non-negative
  fields such as precipitation).

  The two spatial means are reductions of full fields, i.e. launches of the
  fused kernel (unweighted sums and counts in float64); the ratio is formed on
  the few numbers that remain.  With a ``mask`` coordinate on the targets the
  means run over ``mask == 1`` only and the result carries ``mask = count > 0``
  (:63-80).
  """

  def __init__(self, spatial_dims: Sequence[str] = ('latitude', 'longitude')):
    self._spatial_dims = spatial_dims

  def _compute_per_variable(self, predictions, targets):
    from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
    predictions = xl.as_data_array(predictions)
    targets = xl.as_data_array(targets)
    spatial_dims = list(self._spatial_dims)
    epsilon = 1e-6
    masked = 'mask' in targets.coords
    if masked:
      mask = targets.coords['mask']
      if not mask.is_device and mask.dtype != np.bool_:
        # `targets.mask == 1` (deterministic.py:68): only the value 1 is valid
        targets = targets.assign_coords(mask=mask._replace(  # pylint: disable=protected-access
            data=mask.to_numpy() == 1))

    def sums(source):
      # source - 0 against a shared zero slab; the coordinates (and so the
      # mask) of the targets ride along
      lazy = LazyPassthrough(source, targets)
      res = engine.aggregate_fused([lazy], spatial_dims, masked=masked)
      if res is None:
        raise ValueError(
            f'spatial dims {spatial_dims} not found in {source.dims}')
      return res['Error']

    prediction_sum, count = sums(predictions)
    target_sum, _ = sums(targets)
    if masked:
      positive = count > 0
      prediction_mean = (prediction_sum / count).where(positive, 0.0)
      target_mean = (target_sum / count).where(positive, 0.0)
    else:
      prediction_mean = prediction_sum / count
      target_mean = target_sum / count
    ratio = (prediction_mean + epsilon) / (target_mean + epsilon)
    result = abs(ratio - 1).astype(np.float32)
    result.name = predictions.name
    if masked:
      result = result.assign_coords(mask=positive.astype(int))
    return result


class _FusedStatistic(base.PerVariableStatistic):
  """predictions/targets statistic evaluated inside the reduction kernel."""

  def _compute_per_variable(self, predictions, targets):
    return LazyStatistic(type(self).__name__, predictions, targets)


class Error(_FusedStatistic):
  """predictions - targets."""


class AbsoluteError(_FusedStatistic):
  """abs(predictions - targets)."""


class SquaredError(_FusedStatistic):
  """(predictions - targets) ** 2."""


def _passthrough(source: xl.DataArray, other: xl.DataArray, copy_nans: bool):
  from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
  if copy_nans or not set(other.dims) <= set(source.dims) or len(source.dims) < 1:
    return engine.passthrough(source, other, copy_nans=copy_nans)
  return LazyPassthrough(source, other)


class PredictionPassthrough(base.PerVariableStatistic):
  """Simply returns predictions (its mean is the PredictionAverage metric)."""

  def __init__(self, copy_nans_from_targets: bool = False):
    self._copy_nans_from_targets = copy_nans_from_targets

  def _compute_per_variable(self, predictions, targets):
    return _passthrough(predictions, targets, self._copy_nans_from_targets)


class TargetPassthrough(base.PerVariableStatistic):
  """Simply returns targets (its mean is the TargetAverage metric)."""

  def __init__(self, copy_nans_from_predictions: bool = False):
    self._copy_nans_from_predictions = copy_nans_from_predictions

  def _compute_per_variable(self, predictions, targets):
    return _passthrough(targets, predictions, self._copy_nans_from_predictions)


class WindVectorSquaredError(base.Statistic):
  """(u_pred - u_target)**2 + (v_pred - v_target)**2 per wind vector."""

  def __init__(self, u_name: Sequence[str], v_name: Sequence[str],
               vector_name: Sequence[str]):
    self._u_name = u_name
    self._v_name = v_name
    self._vector_name = vector_name
    if not len(self._u_name) == len(self._v_name) == len(self._vector_name):
      raise ValueError(
          'u_name, v_name, and vector_name must have the same length')

  @property
  def unique_name(self) -> str:
    return 'WindVectorSquaredError_' + '_'.join(self._vector_name)

  def compute(self, predictions, targets):
    predictions = base._as_mapping(predictions)  # pylint: disable=protected-access
    targets = base._as_mapping(targets)  # pylint: disable=protected-access
    out = {}
    for u, v, vector in zip(self._u_name, self._v_name, self._vector_name):
      parts = [LazyStatistic('SquaredError', predictions[c], targets[c])
               for c in (u, v)]
      for part, component in zip(parts, (u, v)):
        part.var = component  # lets the Aggregator share the per-component pass
      out[vector] = LazySumStatistic('WindVectorSquaredError', parts,
                                     name=vector)
    return out


class _FusedClimatologyStatistic(base.PerVariableStatisticWithClimatology):

  def _compute_per_variable_with_aligned_climatology(
      self, predictions, targets, aligned_climatology):
    return LazyStatistic(type(self).__name__, predictions, targets,
                         aligned_climatology)


class SquaredPredictionAnomaly(_FusedClimatologyStatistic):
  """(predictions - climatology) ** 2."""


class SquaredTargetAnomaly(_FusedClimatologyStatistic):
  """(targets - climatology) ** 2."""


class AnomalyCovariance(_FusedClimatologyStatistic):
  """(predictions - climatology) * (targets - climatology)."""


class ErrorExceedance(base.PerVariableStatistic):
  """``abs(predictions - targets) > threshold`` as 0/1 for every threshold, NaN
  where the error or the threshold is NaN (deterministic.py:262-295).

  ``thresholds``: a sequence of numbers (dim ``error_exceedance_thresholds``),
  a 1-d DataArray or a Dataset of those keyed by variable name.  The
  comparison happens inside the fused reduction; the threshold index is an
  outer dim of the launch.
  """

  def __init__(self, thresholds):
    if isinstance(thresholds, (list, tuple)):
      thresholds = xl.DataArray(
          np.asarray(thresholds, dtype=np.float64),
          dims='error_exceedance_thresholds',
          coords={'error_exceedance_thresholds': thresholds})
    self._thresholds = thresholds

  def _compute_per_variable(self, predictions, targets):
    from weatherbenchx_b200.lazy import LazyCategoricalStatistic  # pylint: disable=g-import-not-at-top
    thresholds = self._thresholds
    if isinstance(thresholds, xl.Dataset):
      thresholds = thresholds[xl.as_data_array(predictions).name]
    thresholds = xl.as_data_array(thresholds)
    if thresholds.ndim != 1:
      raise NotImplementedError(
          'error exceedance thresholds must be one-dimensional on the B200 '
          f'hot path (got dims {thresholds.dims})')
    dim = thresholds.dims[0]
    return LazyCategoricalStatistic(
        'ErrorExceedance', predictions, targets,
        exceedance_thresholds=thresholds.to_numpy(), exceedance_dim=dim,
        exceedance_coord=thresholds.coords.get(dim))


Bias = Error
MAE = AbsoluteError
MSE = SquaredError
PredictionAverage = PredictionPassthrough
TargetAverage = TargetPassthrough


class RMSE(base.PerVariableMetric):
  """Root mean squared error: sqrt of the mean SquaredError."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {'SquaredError': SquaredError()}

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['SquaredError'])


class WindVectorRMSE(base.Metric):
  """Vector RMSE of two wind components; arguments may be names or lists."""

  def __init__(self, u_name: Union[str, list], v_name: Union[str, list],
               vector_name: Union[str, list]):
    self._u_name = [u_name] if isinstance(u_name, str) else u_name
    self._v_name = [v_name] if isinstance(v_name, str) else v_name
    self._vector_name = (
        [vector_name] if isinstance(vector_name, str) else vector_name)
    if not len(self._u_name) == len(self._v_name) == len(self._vector_name):
      raise ValueError(
          'u_name, v_name, and vector_name must have the same length')

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {'WindVectorSquaredError': WindVectorSquaredError(
        self._u_name, self._v_name, self._vector_name)}

  def values_from_mean_statistics(
      self, statistic_values: Mapping[str, Mapping[Hashable, xl.DataArray]]):
    return {k: np.sqrt(v)
            for k, v in statistic_values['WindVectorSquaredError'].items()}


class ACC(base.PerVariableMetric):
  """Anomaly correlation coefficient cov / (sqrt(spa) * sqrt(sta))."""

  def __init__(self, climatology):
    self._climatology = climatology

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'SquaredPredictionAnomaly': SquaredPredictionAnomaly(self._climatology),
        'SquaredTargetAnomaly': SquaredTargetAnomaly(self._climatology),
        'AnomalyCovariance': AnomalyCovariance(self._climatology),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['AnomalyCovariance'] / (
        np.sqrt(statistic_values['SquaredPredictionAnomaly'])
        * np.sqrt(statistic_values['SquaredTargetAnomaly']))


class PredictionActivity(base.PerVariableMetric):
  """Standard deviation of the prediction anomalies."""

  def __init__(self, climatology):
    self._climatology = climatology

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {'SquaredPredictionAnomaly':
            SquaredPredictionAnomaly(self._climatology)}

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['SquaredPredictionAnomaly'])
