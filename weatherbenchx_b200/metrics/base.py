"""Metric / Statistic base classes.

Same class names, method names, argument meaning and error behaviour as
/root/reference/weatherbenchX/metrics/base.py (Metric :23-82, Statistic :85-173,
PerVariableStatistic :176-205, PerVariableMetric :208-243, the unique-statistic
drivers :252-335, PerVariableStatisticWithClimatology :338-415), so that
existing Metric subclasses and evaluation scripts keep working.  The difference
is in what flows through: ``Statistic.compute`` may return ``LazyStatistic``
handles, which ``aggregation.Aggregator`` turns into fused CUDA launches.
"""

from __future__ import annotations

import abc
from typing import Hashable, Iterator, Mapping

from weatherbenchx_b200 import engine
from weatherbenchx_b200 import xarray_lite as xl


def _as_mapping(data) -> dict:
  """dict of our DataArrays from a dict / Dataset of (xarray or own) arrays."""
  return {k: xl.as_data_array(v) for k, v in data.items()}


class Metric(abc.ABC):
  """A function of weighted *means* of one or more statistics."""

  @property
  @abc.abstractmethod
  def statistics(self) -> Mapping[str, 'Statistic']:
    """Internal name -> Statistic whose mean value the metric needs."""

  @abc.abstractmethod
  def values_from_mean_statistics(
      self, statistic_values: Mapping[str, Mapping[Hashable, xl.DataArray]],
  ) -> Mapping[Hashable, xl.DataArray]:
    """Metric values from mean statistics keyed by the internal names."""


class Statistic(Metric):
  """A per-chunk function of (predictions, targets), aggregated by a mean.

  A Statistic is itself a Metric whose value is the mean of the statistic.
  ``unique_name`` keys the de-duplication across metrics and must capture every
  parameter that changes the result.
  """

  @property
  def unique_name(self) -> str:
    return type(self).__name__

  @abc.abstractmethod
  def compute(
      self, predictions: Mapping[Hashable, xl.DataArray],
      targets: Mapping[Hashable, xl.DataArray],
  ) -> Mapping[Hashable, xl.DataArray]:
    """Per-variable statistic values (possibly lazy) for one chunk."""

  @property
  def statistics(self) -> Mapping[str, 'Statistic']:
    return {'self': self}

  def values_from_mean_statistics(self, statistic_values):
    return statistic_values['self']


class PerVariableStatistic(Statistic):
  """Statistic evaluated independently for every variable in both inputs."""

  def compute(self, predictions, targets):
    predictions, targets = _as_mapping(predictions), _as_mapping(targets)
    out = {}
    for var, pred in predictions.items():
      if var not in targets:
        continue
      value = self._compute_per_variable(pred, targets[var])
      if value is not None:
        out[var] = value
    return out

  @abc.abstractmethod
  def _compute_per_variable(self, predictions: xl.DataArray,
                            targets: xl.DataArray) -> xl.DataArray | None:
    """Statistic for one variable, or None if it is not defined."""


class PerVariableMetric(Metric):
  """Metric evaluated per variable from per-variable mean statistics."""

  def values_from_mean_statistics(self, statistic_values):
    names = list(self.statistics)
    common = set.intersection(*[set(statistic_values[s]) for s in names])
    return {
        var: self._values_from_mean_statistics_per_variable(
            {s: statistic_values[s][var] for s in names})
        for var in common
    }

  @abc.abstractmethod
  def _values_from_mean_statistics_per_variable(
      self, statistic_values: Mapping[str, xl.DataArray]) -> xl.DataArray:
    """Metric value for a single variable."""


NoOpMetric = lambda statistic: statistic  # deprecated shim kept for parity


def generate_unique_statistics_for_all_metrics(
    metrics: Mapping[str, Metric], predictions, targets,
) -> Iterator[tuple]:
  """Yields (unique_name, per-variable statistic values), one at a time."""
  unique: dict = {}
  for metric in metrics.values():
    for stat in metric.statistics.values():
      unique[stat.unique_name] = stat
  for name, stat in unique.items():
    try:
      yield name, stat.compute(predictions, targets)
    except Exception as e:
      raise ValueError(
          f'Failed to compute statistic {name}={stat} from:\n'
          f'{predictions=}\n{targets=}') from e


def compute_unique_statistics_for_all_metrics(metrics, predictions, targets):
  """{unique_name: {variable: statistic values}} with duplicates removed."""
  return dict(generate_unique_statistics_for_all_metrics(
      metrics, predictions, targets))


def compute_metric_from_statistics(metric: Metric, statistic_values):
  """Re-keys mean statistics from unique to internal names and evaluates."""
  renamed = {internal: statistic_values[stat.unique_name]
             for internal, stat in metric.statistics.items()}
  return metric.values_from_mean_statistics(renamed)


def compute_metrics_from_statistics(metrics, statistic_values):
  return {name: compute_metric_from_statistics(metric, statistic_values)
          for name, metric in metrics.items()}


class PerVariableStatisticWithClimatology(Statistic):
  """Per-variable statistic of (predictions, targets, aligned climatology).

  The climatology is aligned on the predictions' valid time
  (init_time + lead_time, or valid_time) by dayofyear[/hour] or time labels.
  The alignment is carried as index arrays; the gather itself happens inside
  the kernel through per-job addresses.
  """

  def __init__(self, climatology: Mapping[Hashable, xl.DataArray]):
    self._climatology = climatology

  def compute(self, predictions, targets):
    predictions, targets = _as_mapping(predictions), _as_mapping(targets)
    out = {}
    for var, pred in predictions.items():
      clim = xl.as_data_array(self._climatology[var])
      out[var] = self._compute_per_variable(pred, targets[var], clim)
    return out

  def _compute_per_variable(self, predictions, targets, climatology):
    aligned = engine.align_climatology(predictions, climatology)
    return self._compute_per_variable_with_aligned_climatology(
        predictions, targets, aligned)

  @abc.abstractmethod
  def _compute_per_variable_with_aligned_climatology(
      self, predictions, targets, aligned_climatology):
    """Statistic for one variable given the aligned climatology."""
