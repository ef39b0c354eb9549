"""Categorical statistics and metrics on the fused CUDA reduction.

Mirrors /root/reference/weatherbenchX/metrics/categorical.py: TruePositives
:25-41, TrueNegatives :44-62, FalsePositives :65-81, FalseNegatives :84-101 and
the contingency-table metrics CSI :345-367, Accuracy :370-398, Recall :401-421,
FalseAlarmRate :424-444, Precision :447-467, F1Score :470-500, FrequencyBias
:503-525, HSS :528-555, ETS :558-592, SEDI :595-635 (same unique names and
formulas).

In the reference the inputs of these statistics are binary fields produced by
``wrappers.ContinuousToBinary`` -- one full-size field per threshold and input
-- and every statistic is another full-size field.  Here the statistics are
handles: the Aggregator serves the four entries of a variable's contingency
table, for all thresholds, with ONE launch of the fused reduction kernel that
thresholds the continuous fields as it reads them (8 B per grid point and
threshold; ``wbx_det_desc.xform``).

SEEPS, the ranked-probability / reliability statistics and the tile-based
scores of the reference's module are not part of this path.
"""

from __future__ import annotations

from typing import Mapping

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import LazyCategoricalStatistic
from weatherbenchx_b200.metrics import base


class _ContingencyStatistic(base.PerVariableStatistic):
  """One entry of the 2x2 contingency table of binary predictions/targets:
  0/1 per grid point, NaN where ``predictions * targets`` is NaN."""

  @property
  def unique_name(self) -> str:
    return type(self).__name__

  def _compute_per_variable(self, predictions, targets):
    return LazyCategoricalStatistic(type(self).__name__, predictions, targets)


class TruePositives(_ContingencyStatistic):
  """predictions.astype(bool) * targets.astype(bool)."""


class TrueNegatives(_ContingencyStatistic):
  """~predictions.astype(bool) * ~targets.astype(bool)."""


class FalsePositives(_ContingencyStatistic):
  """predictions.astype(bool) * ~targets.astype(bool)."""


class FalseNegatives(_ContingencyStatistic):
  """~predictions.astype(bool) * targets.astype(bool)."""


def _table(*names):
  classes = {'TruePositives': TruePositives, 'FalsePositives': FalsePositives,
             'FalseNegatives': FalseNegatives, 'TrueNegatives': TrueNegatives}
  return {n: classes[n]() for n in names}


_ALL = ('TruePositives', 'FalsePositives', 'FalseNegatives', 'TrueNegatives')


class CSI(base.PerVariableMetric):
  """Critical Success Index (Threat Score): TP / (TP + FP + FN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalsePositives', 'FalseNegatives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['TruePositives'] / (
        statistic_values['TruePositives']
        + statistic_values['FalsePositives']
        + statistic_values['FalseNegatives'])


class Accuracy(base.PerVariableMetric):
  """(TP + TN) / (TP + FP + FN + TN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table(*_ALL)

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return (
        statistic_values['TruePositives'] + statistic_values['TrueNegatives']
    ) / (
        statistic_values['TruePositives']
        + statistic_values['FalsePositives']
        + statistic_values['FalseNegatives']
        + statistic_values['TrueNegatives'])


class Recall(base.PerVariableMetric):
  """True positive rate / sensitivity: TP / (TP + FN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalseNegatives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['TruePositives'] / (
        statistic_values['TruePositives'] + statistic_values['FalseNegatives'])


class FalseAlarmRate(base.PerVariableMetric):
  """FP / (TP + FP)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalsePositives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['FalsePositives'] / (
        statistic_values['TruePositives'] + statistic_values['FalsePositives'])


class Precision(base.PerVariableMetric):
  """Positive predictive value: TP / (TP + FP)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalsePositives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['TruePositives'] / (
        statistic_values['TruePositives'] + statistic_values['FalsePositives'])


class F1Score(base.PerVariableMetric):
  """2 TP / (2 TP + FP + FN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalsePositives', 'FalseNegatives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return (
        2
        * statistic_values['TruePositives']
        / (
            2 * statistic_values['TruePositives']
            + statistic_values['FalsePositives']
            + statistic_values['FalseNegatives']))


class FrequencyBias(base.PerVariableMetric):
  """(TP + FP) / (TP + FN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table('TruePositives', 'FalsePositives', 'FalseNegatives')

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return (
        statistic_values['TruePositives'] + statistic_values['FalsePositives']
    ) / (statistic_values['TruePositives'] + statistic_values['FalseNegatives'])


class HSS(base.PerVariableMetric):
  """Heidke Skill Score:
  2 (TP TN - FP FN) / ((TP + FN)(FN + TN) + (TP + FP)(FP + TN))."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table(*_ALL)

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    tp = statistic_values['TruePositives']
    tn = statistic_values['TrueNegatives']
    fp = statistic_values['FalsePositives']
    fn = statistic_values['FalseNegatives']
    numerator = 2 * (tp * tn - fp * fn)
    denominator = (tp + fn) * (fn + tn) + (tp + fp) * (fp + tn)
    return numerator / denominator


class ETS(base.PerVariableMetric):
  """Equitable Threat Score (Gilbert Skill Score):
  (TP - TP_random) / (TP + FP + FN - TP_random) with
  TP_random = (TP + FP)(TP + FN) / (TP + FP + FN + TN)."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table(*_ALL)

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    tp = statistic_values['TruePositives']
    tn = statistic_values['TrueNegatives']
    fp = statistic_values['FalsePositives']
    fn = statistic_values['FalseNegatives']
    all_sum = tp + fp + fn + tn
    tp_random = ((tp + fp) * (tp + fn)) / all_sum
    return (tp - tp_random) / (tp + fp + fn - tp_random)


class SEDI(base.PerVariableMetric):
  """Symmetric extremal dependency index (Ferro and Stephenson 2011) from the
  hit rate H = TP / (TP + FN) and the false alarm rate F = FP / (FP + TN),
  both clipped to [1e-6, 1 - 1e-6]."""

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return _table(*_ALL)

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    tp = statistic_values['TruePositives']
    tn = statistic_values['TrueNegatives']
    fp = statistic_values['FalsePositives']
    fn = statistic_values['FalseNegatives']
    h = (tp / (tp + fn)).clip(1e-6, 1 - 1e-6)
    f = (fp / (fp + tn)).clip(1e-6, 1 - 1e-6)
    log_h, log_f = np.log(h), np.log(f)
    log_1_minus_h, log_1_minus_f = np.log(1 - h), np.log(1 - f)
    numerator = log_f - log_h + log_1_minus_h - log_1_minus_f
    denominator = log_h + log_f + log_1_minus_h + log_1_minus_f
    return numerator / denominator
