"""Input transforms and the wrappers that apply them to statistics / metrics.

Mirrors the part of /root/reference/weatherbenchX/metrics/wrappers.py that
feeds the fused deterministic path: InputTransform :92-113, EnsembleMean
:116-148, Rename :745-768, Select :771-808, WrappedStatistic :967-1003,
RenamedStatistic :1006-1022, WrappedMetric :1025-1069,
SubselectVariablesForStatistic :1072-1099, SubselectVariables :1102-1128 (same
constructor arguments and unique_name strings).

``EnsembleMean`` is the compute-carrying transform: the public benchmark's
ensemble-mean RMSE / bias / ACC are ``WrappedMetric(metric,
[EnsembleMean('predictions')])``.  Its mean field comes from the
``wbx_ensemble_mean`` kernel (one HBM pass over the ensemble) and is memoised
per input array, so all wrapped statistics of one variable see the SAME mean
array and the Aggregator fuses them into one launch.  ``ContinuousToBinary`` /
``binarize_thresholds`` (:50-88, :214-267) return lazy handles that the
categorical statistics threshold inside the fused reduction.  The binning /
quantile transforms of the reference are not part of this path.
"""

from __future__ import annotations

import abc
from typing import Any, Hashable, Mapping, Sequence

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200 import xarray_tree
from weatherbenchx_b200.metrics import base


class InputTransform(abc.ABC):
  """Base class for input transformations."""

  def __init__(self, which):
    if which not in ['predictions', 'targets', 'both']:
      raise ValueError(f'Invalid value for `which`: {which}')
    self.which = which

  @property
  @abc.abstractmethod
  def unique_name_suffix(self) -> str:
    """Suffix added to the wrapped statistic's unique name."""

  @abc.abstractmethod
  def transform_fn(self, da: xl.DataArray) -> xl.DataArray:
    """Function applied to predictions and/or targets."""


class EnsembleMean(InputTransform):
  """Ensemble mean over ``ensemble_dim`` (evaluated on the GPU)."""

  def __init__(self, which: str, ensemble_dim='number', skipna=False,
               skip_if_ensemble_dim_missing: bool = False):
    super().__init__(which)
    self._ensemble_dim = ensemble_dim
    self._skipna = skipna
    self._skip_if_ensemble_dim_missing = skip_if_ensemble_dim_missing

  @property
  def unique_name_suffix(self) -> str:
    return f'ensemble_mean_{self._ensemble_dim=}_{self._skipna=}'

  def transform_fn(self, da: xl.DataArray) -> xl.DataArray:
    from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
    da = xl.as_data_array(da)
    if self._ensemble_dim not in da.dims and self._skip_if_ensemble_dim_missing:
      return da
    return engine.ensemble_mean(da, self._ensemble_dim, skipna=self._skipna)


def binarize_thresholds(x: xl.DataArray, thresholds, threshold_dim: str):
  """``(x > threshold).where(~isnan(x)).astype(float32)`` with the thresholds
  along a new trailing dim (wrappers.py:50-88), as a lazy handle.

  ``thresholds``: an iterable of numbers, a 1-d DataArray along
  ``threshold_dim`` or a Dataset of those keyed by variable name.  Thresholds
  that vary over other dims (per grid point / per time) are not part of the
  fused path.
  """
  from weatherbenchx_b200.lazy import LazyBinarized  # pylint: disable=g-import-not-at-top
  x = xl.as_data_array(x)
  if isinstance(thresholds, xl.Dataset):
    assert x.name in thresholds, (
        f'Input DataArray name ({x.name}) not found in thresholds')
    thresholds = thresholds[x.name]
  if isinstance(thresholds, xl.DataArray):
    assert threshold_dim in thresholds.dims, (
        f'threshold_dim ({threshold_dim}) not found in thresholds'
        f' ({thresholds.dims})')
    if thresholds.dims != (threshold_dim,):
      raise NotImplementedError(
          'thresholds that vary along dims other than the threshold dim '
          f'({thresholds.dims}) are outside the B200 hot path')
    labels = thresholds.coords[threshold_dim].to_numpy() if (
        threshold_dim in thresholds.coords) else None
    out = LazyBinarized(x, thresholds.to_numpy(), threshold_dim)
    if labels is not None:  # the dim keeps the labels of the threshold array
      out._coords[threshold_dim] = xl.DataArray(  # pylint: disable=protected-access
          labels, (threshold_dim,), name=threshold_dim)
    else:
      del out._coords[threshold_dim]  # pylint: disable=protected-access
    return out
  return LazyBinarized(x, list(thresholds), threshold_dim)


class ContinuousToBinary(InputTransform):
  """``x > threshold`` for every threshold, along a new dim ``threshold_dim``
  (wrappers.py:214-267).  Returns a handle: the categorical statistics compare
  inside the fused reduction, so no binary field is stored."""

  def __init__(self, which: str, threshold_value, threshold_dim: str,
               unique_name_suffix: str | None = None):
    super().__init__(which)
    import collections.abc  # pylint: disable=g-import-not-at-top
    self._threshold_value = (
        threshold_value
        if isinstance(threshold_value, (collections.abc.Iterable,
                                        xl.DataArray, xl.Dataset))
        else [threshold_value])
    self._threshold_dim = threshold_dim
    if isinstance(self._threshold_value, (xl.DataArray, xl.Dataset)):
      if unique_name_suffix is None:
        raise ValueError(
            'unique_name_suffix must be provided if threshold_value is an'
            ' xarray.DataArray or xarray.Dataset.')
    self._unique_name_suffix = unique_name_suffix
    # one label vector per transform: every call hands the planner the same
    # coordinate payload, so repeated evaluations reuse their plan
    self._labels = None
    if not isinstance(self._threshold_value, (xl.DataArray, xl.Dataset)):
      self._threshold_value = list(self._threshold_value)
      self._labels = np.asarray(self._threshold_value, dtype=np.float64)

  @property
  def unique_name_suffix(self) -> str:
    if self._unique_name_suffix is None:
      suffix = ','.join([str(t) for t in self._threshold_value])
    else:
      suffix = self._unique_name_suffix
    return f'{self._threshold_dim}={suffix}'

  def transform_fn(self, da: xl.DataArray) -> xl.DataArray:
    if self._labels is not None:
      from weatherbenchx_b200.lazy import LazyBinarized  # pylint: disable=g-import-not-at-top
      return LazyBinarized(xl.as_data_array(da), self._labels,
                           self._threshold_dim)
    return binarize_thresholds(da, self._threshold_value, self._threshold_dim)


class Rename(InputTransform):
  """Renames variables, coordinates and dimensions."""

  def __init__(self, which: str, renames: Mapping[Hashable, Hashable]):
    super().__init__(which)
    self._renames = renames

  @property
  def unique_name_suffix(self) -> str:
    return f'rename_{self._renames}'

  def transform_fn(self, da: xl.DataArray) -> xl.DataArray:
    return xl.as_data_array(da).rename(self._renames)


class Select(InputTransform):
  """Selects data with sel and / or isel (views; no data movement)."""

  def __init__(self, which: str, sel: Mapping[Hashable, Any] | None = None,
               isel: Mapping[Hashable, Any] | None = None,
               sel_kwargs: Mapping[Hashable, Any] | None = None,
               isel_kwargs: Mapping[Hashable, Any] | None = None):
    super().__init__(which)
    self._isel = isel
    self._sel = sel
    self._isel_kwargs = isel_kwargs or {}
    self._sel_kwargs = sel_kwargs or {}

  @property
  def unique_name_suffix(self) -> str:
    return (f'select_{self._isel=}_{self._isel_kwargs=}_{self._sel=}_'
            f'{self._sel_kwargs=}')

  def transform_fn(self, da: xl.DataArray) -> xl.DataArray:
    da = xl.as_data_array(da)
    if self._sel is not None:
      da = da.sel(self._sel, **self._sel_kwargs)
    if self._isel is not None:
      da = da.isel(self._isel, **self._isel_kwargs)
    return da


class WrappedStatistic(base.Statistic):
  """A statistic evaluated on transformed inputs; the suffix enters its name."""

  def __init__(self, statistic: base.Statistic, transform: InputTransform):
    self.statistic = statistic
    self.transform = transform

  @property
  def unique_name(self) -> str:
    return (f'{self.statistic.unique_name}_{self.transform.which}_'
            f'{self.transform.unique_name_suffix}')

  def compute(self, predictions, targets):
    if self.transform.which in ('predictions', 'both'):
      predictions = xarray_tree.map_structure(
          self.transform.transform_fn, dict(predictions))
    if self.transform.which in ('targets', 'both'):
      targets = xarray_tree.map_structure(
          self.transform.transform_fn, dict(targets))
    return self.statistic.compute(predictions, targets)


class RenamedStatistic(base.Statistic):
  """A statistic under a new unique name."""

  def __init__(self, statistic: base.Statistic, unique_name: str):
    self._statistic = statistic
    self._unique_name = unique_name

  @property
  def unique_name(self) -> str:
    return self._unique_name

  def compute(self, predictions, targets):
    return self._statistic.compute(predictions, targets)


class WrappedMetric(base.Metric):
  """All statistics of a metric wrapped with input transforms.

  Transforms [f, g, h] turn x into h(g(f(x))).
  """

  def __init__(self, metric: base.Metric, transforms: list,
               unique_name_suffix: str | None = None):
    self.metric = metric
    self.transforms = transforms
    self.unique_name_suffix = unique_name_suffix

  @property
  def statistics(self) -> Mapping[Hashable, base.Statistic]:
    stats = {}
    for name, stat in self.metric.statistics.items():
      original_name = stat.unique_name
      # the outermost wrapper runs first, hence the reverse order
      for wrapper in self.transforms[::-1]:
        stat = WrappedStatistic(stat, wrapper)
      if self.unique_name_suffix is not None:
        stat = RenamedStatistic(
            stat, f'{original_name}_{self.unique_name_suffix}')
      stats[name] = stat
    return stats

  def values_from_mean_statistics(self, statistic_values):
    return self.metric.values_from_mean_statistics(statistic_values)


class SubselectVariablesForStatistic(base.Statistic):
  """A statistic restricted to a subset of variables."""

  def __init__(self, statistic: base.Statistic, variables: Sequence[str]):
    self.statistic = statistic
    self.variables = variables

  @property
  def unique_name(self) -> str:
    return f'{self.statistic.unique_name}_' + '_'.join(self.variables)

  def compute(self, predictions, targets):
    predictions = {k: v for k, v in predictions.items() if k in self.variables}
    targets = {k: v for k, v in targets.items() if k in self.variables}
    return self.statistic.compute(predictions, targets)


class SubselectVariables(base.Metric):
  """A metric restricted to a subset of variables."""

  def __init__(self, metric: base.Metric, variables: Sequence[str]):
    self.metric = metric
    self.variables = variables

  @property
  def statistics(self) -> Mapping[Hashable, base.Statistic]:
    return {name: SubselectVariablesForStatistic(stat, self.variables)
            for name, stat in self.metric.statistics.items()}

  def values_from_mean_statistics(self, statistic_values):
    return self.metric.values_from_mean_statistics(statistic_values)


# Deprecated no-op aliases kept by the reference (wrappers.py:1131-1135).
IntersectPredictionAndTargetVariablesForStatistic = lambda statistic: statistic  # pylint: disable=invalid-name
IntersectPredictionAndTargetVariables = lambda metric: metric  # pylint: disable=invalid-name
