"""Deterministic statistics and metrics served by the fused CUDA kernel.

Mirrors the hot-path subset of
/root/reference/weatherbenchX/metrics/deterministic.py: Error :91-100,
AbsoluteError :103-112, SquaredError :115-123, SquaredPredictionAnomaly
:222-232, SquaredTargetAnomaly :235-245, AnomalyCovariance :248-259, the
aliases Bias / MAE / MSE :305-307, RMSE :312-324, ACC :374-400 and
PredictionActivity :403-425.  unique_name of every statistic is the class name,
as in the reference.
"""

from __future__ import annotations

from typing import Mapping

import numpy as np

from weatherbenchx_b200.lazy import LazyStatistic
from weatherbenchx_b200.metrics import base


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


Bias = Error
MAE = AbsoluteError
MSE = SquaredError


class RMSE(base.PerVariableMetric):
  """Root mean squared error: sqrt of the mean SquaredError."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {'SquaredError': SquaredError()}

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['SquaredError'])


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
