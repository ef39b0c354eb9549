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

``SEEPS`` (:104-304), the precipitation score of the public benchmark's
deterministic suite (run_benchmark_evaluation.py:331-340), depends on two
per-point parameters (climatological wet threshold of the valid time, dry
fraction of the location); it is evaluated as a field by an elementwise kernel
and aggregated by the fused masked reduction (region bins included).

The ranked-probability / reliability statistics and the tile-based scores of
the reference's module are not part of this path.
"""

from __future__ import annotations

import collections
import threading
import warnings
import weakref
from typing import Hashable, Mapping, Sequence, Union

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import LazyCategoricalStatistic
from weatherbenchx_b200.lazy import LazyPassthrough
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


# p1 (mean dry fraction per grid point) of a climatology variable: a host
# reduction over (hour, dayofyear) that the reference repeats for every chunk
# (categorical.py:268-272); here once per climatology array.
_P1_CACHE: 'collections.OrderedDict' = collections.OrderedDict()
_P1_LOCK = threading.Lock()


def _dry_fraction_mean(dry_fraction: xl.DataArray) -> xl.DataArray:
  key = (id(dry_fraction), id(dry_fraction.data))
  with _P1_LOCK:
    hit = _P1_CACHE.get(key)
    if hit is not None and hit[1]() is dry_fraction.data:
      return hit[0]
  time_dims = [d for d in ('hour', 'dayofyear') if d in dry_fraction.dims]
  if len(time_dims) != 2:
    raise ValueError(
        'seeps_dry_fraction needs the dims `hour` and `dayofyear` '
        f'(got {dry_fraction.dims})')
  values = dry_fraction.to_numpy()
  axes = tuple(dry_fraction.dims.index(d) for d in time_dims)
  with warnings.catch_warnings():
    warnings.simplefilter('ignore')  # all-NaN grid points stay NaN
    # xarray's .mean() skips NaN by default and keeps the input dtype
    mean = np.nanmean(values, axis=axes)
  rest = tuple(d for d in dry_fraction.dims if d not in time_dims)
  p1 = xl.DataArray(mean, rest, coords={
      d: dry_fraction.coords[d] for d in rest if d in dry_fraction.coords})
  try:
    ref = weakref.ref(dry_fraction.data)
  except TypeError:
    return p1
  with _P1_LOCK:
    _P1_CACHE[key] = (p1, ref)
    for k in [k for k, v in _P1_CACHE.items() if v[1]() is None]:
      del _P1_CACHE[k]
    while len(_P1_CACHE) > 8:
      _P1_CACHE.popitem(last=False)
  return p1


class SEEPS(base.Statistic):
  """Stable Equitable Error in Probability Space (Rodwell et al. 2010;
  categorical.py:104-304).

  ``climatology`` holds ``{variable}_seeps_dry_fraction`` and
  ``{variable}_seeps_threshold`` with dims (hour, dayofyear, <grid dims>).
  The result carries a ``mask`` coordinate -- p1 within [min_p1, max_p1],
  combined with the mask of the predictions or targets if there is one -- and
  is NaN outside it: use ``Aggregator(masked=True)``.
  """

  def __init__(self, variables: Sequence[str], climatology,
               dry_threshold_mm: Union[float, Sequence[float]] = 0.25,
               min_p1: Union[float, Sequence[float]] = 0.1,
               max_p1: Union[float, Sequence[float]] = 0.85):
    as_list = lambda v: (list(v) if isinstance(v, Sequence)  # noqa: E731
                         else [v] * len(variables))
    self._variables = variables
    self._climatology = climatology
    self._dry_threshold_mm = as_list(dry_threshold_mm)
    self._min_p1 = as_list(min_p1)
    self._max_p1 = as_list(max_p1)
    assert (len(self._variables) == len(self._dry_threshold_mm)
            == len(self._min_p1) == len(self._max_p1)
            ), 'All arguments must have the same length.'

  @property
  def unique_name(self) -> str:
    suffix = (
        '_'.join(self._variables)
        + '_dry_threshold_mm_'
        + '_'.join([str(s) for s in self._dry_threshold_mm])
        + '_min_p1_'
        + '_'.join([str(s) for s in self._min_p1])
        + '_max_p1_'
        + '_'.join([str(s) for s in self._max_p1]))
    return f'SEEPS_{suffix}'

  def compute(self, predictions: Mapping[Hashable, xl.DataArray],
              targets: Mapping[Hashable, xl.DataArray]
              ) -> Mapping[Hashable, xl.DataArray]:
    out = {}
    for variable, dry_threshold_mm, min_p1, max_p1 in zip(
        self._variables, self._dry_threshold_mm, self._min_p1, self._max_p1):
      out[variable] = self._compute_seeps_per_variable(
          xl.as_data_array(predictions[variable]),
          xl.as_data_array(targets[variable]), variable, dry_threshold_mm,
          min_p1, max_p1)
    return out

  def _compute_seeps_per_variable(self, predictions, targets, variable,
                                  dry_threshold_mm, min_p1, max_p1):
    from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
    wet_threshold = xl.as_data_array(
        self._climatology[f'{variable}_seeps_threshold'])
    aligned = engine.align_climatology(predictions, wet_threshold)
    p1 = _dry_fraction_mean(xl.as_data_array(
        self._climatology[f'{variable}_seeps_dry_fraction']))
    # Python-float bounds compare in the dtype of p1 (NumPy weak scalars)
    values = p1.to_numpy()
    with np.errstate(invalid='ignore'):
      in_range = (values >= min_p1) & (values <= max_p1)
    p1_masked = p1._replace(data=np.where(  # pylint: disable=protected-access
        in_range, values, np.nan).astype(np.float32))
    # metres; the comparison `da <= dry_threshold` is a float32 one
    dry_threshold = float(np.float32(dry_threshold_mm / 1000.0))
    field = engine.seeps_field(predictions, targets, aligned, p1_masked,
                               dry_threshold)
    mask = p1._replace(data=in_range)  # pylint: disable=protected-access
    if 'mask' in predictions.coords:
      if 'mask' in targets.coords:
        raise ValueError(
            'Both predictions and targets have masks. This should not happen.')
      mask = engine.and_masks(mask, predictions.coords['mask'])
    elif 'mask' in targets.coords:
      mask = engine.and_masks(mask, targets.coords['mask'])
    field = field.assign_coords(mask=mask)
    # handle: `field - 0` in the fused masked reduction (region bins included)
    return LazyPassthrough(field, field)
