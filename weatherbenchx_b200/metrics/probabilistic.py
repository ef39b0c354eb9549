"""Ensemble CRPS statistics and metric served by the CUDA CRPS kernel.

Mirrors the hot-path subset of
/root/reference/weatherbenchX/metrics/probabilistic.py: CRPSSkill :116-145,
CRPSSpread :165-247, CRPSEnsemble :606-688, and the ensemble-moment family
EnsembleVariance :250-273, UnbiasedEnsembleMeanSquaredError :276-336 with their
metrics :864-1005 (same constructor arguments, same unique_name strings, same
errors).  All four statistics of one (predictions, targets) pair come out of
ONE kernel launch that reads the ensemble once.  As in the reference ``use_sort`` does not
enter the statistic's unique_name -- both estimators compute the same statistic.
``use_sort=False`` (default) runs the tiled O(M^2) member-pair kernel;
``use_sort=True`` runs the sort / probability-weighted-moment estimator in a
register sorting network (ensembles of up to 64 members; larger ones fall back
to the pair sum).
"""

from __future__ import annotations

from typing import Mapping

import numpy as np

from weatherbenchx_b200.lazy import LazyEnsembleStatistic
from weatherbenchx_b200.metrics import base
from weatherbenchx_b200.metrics import deterministic

ENSEMBLE_DIM = 'number'


class EnsembleAveragedStatistic(base.Statistic):
  """A statistic averaged over the ensemble dimension (per-member scores).

  Reference: probabilistic.py:35-69.  Lazy statistics stay lazy: the
  Aggregator folds the ensemble average into the fused reduction.
  """

  def __init__(self, wrapped_statistic: base.Statistic, *, ensemble_dim: str,
               skipna_ensemble: bool):
    self._wrapped_statistic = wrapped_statistic
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def unique_name(self) -> str:
    return self._wrapped_statistic.unique_name + '_each_' + self._ensemble_dim

  def compute(self, predictions, targets):
    from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
    from weatherbenchx_b200 import xarray_lite as xl  # pylint: disable=g-import-not-at-top
    from weatherbenchx_b200.lazy import LazyEnsembleAveraged  # pylint: disable=g-import-not-at-top
    from weatherbenchx_b200.lazy import LazyStatistic  # pylint: disable=g-import-not-at-top
    statistics = self._wrapped_statistic.compute(predictions, targets)
    out = {}
    for var, da in statistics.items():
      da = xl.as_data_array(da)
      if self._ensemble_dim not in da.dims:
        raise ValueError(
            f'Dimension {self._ensemble_dim} not found in {da.dims}')
      if (isinstance(da, LazyStatistic) and da.is_lazy and
          type(da) is LazyStatistic):
        out[var] = LazyEnsembleAveraged(da, self._ensemble_dim,
                                        self._skipna_ensemble)
      else:
        out[var] = engine.ensemble_mean(da, self._ensemble_dim,
                                        skipna=self._skipna_ensemble)
    return out


class EnsembleErrorExceedance(deterministic.ErrorExceedance):
  """Error exceedance averaged over the ensemble members
  (probabilistic.py:836-861: ``ErrorExceedance`` of every member, then
  ``.mean(dim=ensemble_dim)`` -- xarray's NaN-skipping mean).

  The member mean followed by the weighted sums equals the fused reduction over
  ``reduce_dims + [ensemble_dim]`` divided by the member count as long as no
  NaN takes part; the Aggregator evaluates it that way (one launch reading the
  ensemble once per threshold) and recomputes through the per-point mean field
  only if the result shows a NaN.
  """

  def __init__(self, thresholds, ensemble_dim: str = ENSEMBLE_DIM):
    super().__init__(thresholds=thresholds)
    self._ensemble_dim = ensemble_dim

  def _compute_per_variable(self, predictions, targets):
    from weatherbenchx_b200.lazy import LazyEnsembleAveraged  # pylint: disable=g-import-not-at-top
    inner = super()._compute_per_variable(predictions, targets)
    return LazyEnsembleAveraged(inner, self._ensemble_dim,
                                skipna_ensemble=True, optimistic=True)


class EnsembleAveragedMetric(base.Metric):
  """Wraps a metric so that its statistics are averaged over the ensemble
  dimension, i.e. deterministic scores of the individual members
  (probabilistic.py:72-113)."""

  def __init__(self, wrapped_metric: base.Metric, *,
               ensemble_dim: str = ENSEMBLE_DIM, skipna_ensemble: bool = False):
    self._wrapped_metric = wrapped_metric
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        name: EnsembleAveragedStatistic(
            wrapped_statistic=stat, ensemble_dim=self._ensemble_dim,
            skipna_ensemble=self._skipna_ensemble)
        for name, stat in self._wrapped_metric.statistics.items()}

  def values_from_mean_statistics(self, statistic_values):
    return self._wrapped_metric.values_from_mean_statistics(statistic_values)


class CRPSSkill(base.PerVariableStatistic):
  """The skill term of CRPS, E|X - Y|."""

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def unique_name(self) -> str:
    return f'CRPSSkill_{self._ensemble_dim}'

  def _compute_per_variable(self, predictions, targets):
    from weatherbenchx_b200 import xarray_lite as xl  # pylint: disable=g-import-not-at-top
    targets = xl.as_data_array(targets)
    if self._ensemble_dim in targets.dims:
      # An ensemble of targets (probabilistic.py:135-145): the mean of
      # |x_m - y_k| over both member dims is the mean over the target members
      # of the usual skill.  Every target member is a strided view of the
      # target array; the launches of the members are merged into one.
      if self._skipna_ensemble:
        raise NotImplementedError(
            'skipna_ensemble with an ensemble of targets: the NaN-skipping '
            'mean over member pairs is not a mean of per-member means')
      from weatherbenchx_b200.lazy import LazySumStatistic  # pylint: disable=g-import-not-at-top
      n_target = targets.sizes[self._ensemble_dim]
      parts = [
          LazyEnsembleStatistic(
              'CRPSSkill', predictions,
              targets.isel({self._ensemble_dim: k}), self._ensemble_dim,
              fair=True, skipna_ensemble=False)
          for k in range(n_target)]
      return LazySumStatistic('CRPSSkill', parts,
                              name=xl.as_data_array(predictions).name,
                              scale=1.0 / n_target)
    return LazyEnsembleStatistic(
        'CRPSSkill', predictions, targets, self._ensemble_dim, fair=True,
        skipna_ensemble=self._skipna_ensemble)


class CRPSSpread(base.PerVariableStatistic):
  """Sample estimate of the spread term of CRPS, E|X - X'|."""

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM, use_sort: bool = False,
               fair: bool = True, which: str = 'predictions',
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._use_sort = use_sort
    self._which = which
    self._fair = fair
    self._skipna_ensemble = skipna_ensemble

  @property
  def unique_name(self) -> str:
    fair_str = 'fair' if self._fair else 'unfair'
    return f'CRPSSpread_{self._ensemble_dim}_{fair_str}_{self._which}'

  def _compute_per_variable(self, predictions, targets):
    from weatherbenchx_b200 import xarray_lite as xl  # pylint: disable=g-import-not-at-top
    predictions = xl.as_data_array(predictions)
    targets = xl.as_data_array(targets)
    if self._which not in ('predictions', 'targets'):
      raise ValueError(f'Unhandled {self._which=}')
    ens = self._ensemble_dim

    def one_member(da):
      # the spread is a function of one input alone (probabilistic.py:
      # 199-204); the kernel still wants a member-free companion slab
      return da.isel({ens: 0}) if ens in da.dims else da

    if self._which == 'targets':
      predictions, targets = targets, one_member(predictions)
    else:
      targets = one_member(targets)
    if self._use_sort and self._skipna_ensemble:
      raise ValueError('skipna_ensemble is not supported with use_sort=True.')
    if (not self._skipna_ensemble and
        predictions.sizes.get(self._ensemble_dim, 2) < 2):
      raise ValueError('Cannot estimate CRPS spread with n_ensemble < 2.')
    return LazyEnsembleStatistic(
        'CRPSSpread', predictions, targets, self._ensemble_dim,
        fair=self._fair, skipna_ensemble=self._skipna_ensemble,
        use_sort=self._use_sort)


class EnsembleVariance(base.PerVariableStatistic):
  """Variance over the ensemble dimension, standard unbiased (ddof=1) estimator.

  Reference: probabilistic.py:250-273.  Served by the same launch as the CRPS
  statistics (slot 2 of wbx_crps_plan_run).
  """

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def unique_name(self) -> str:
    return (f'EnsembleVariance_{self._ensemble_dim}_skipna_ensemble_'
            f'{self._skipna_ensemble}')

  def _compute_per_variable(self, predictions, targets):
    return LazyEnsembleStatistic(
        'EnsembleVariance', predictions, targets, self._ensemble_dim,
        fair=True, skipna_ensemble=self._skipna_ensemble)


class UnbiasedEnsembleMeanSquaredError(base.PerVariableStatistic):
  """(ensemble mean - target)^2 minus the finite-ensemble bias variance / n.

  Reference: probabilistic.py:276-336 (the usual case of deterministic
  targets; an ensemble of targets is outside the hot path).  Slot 3 of
  wbx_crps_plan_run.
  """

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def unique_name(self) -> str:
    return (f'UnbiasedEnsembleMeanSquaredError_{self._ensemble_dim}_'
            f'skipna_ensemble_{self._skipna_ensemble}')

  def _compute_per_variable(self, predictions, targets):
    return LazyEnsembleStatistic(
        'UnbiasedEnsembleMeanSquaredError', predictions, targets,
        self._ensemble_dim, fair=True, skipna_ensemble=self._skipna_ensemble)


class CRPSEnsemble(base.PerVariableMetric):
  """CRPS = E|X - Y| - 0.5 E|X - X'| for an ensemble prediction.

  ``fair=True`` gives the unbiased (eFAIR) estimate of Zamo & Naveau (2018);
  ``skipna_ensemble=True`` treats NaN members as missing, with a per-point
  ensemble size.
  """

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM, use_sort: bool = False,
               fair: bool = True, skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._use_sort = use_sort
    self._fair = fair
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'CRPSSkill': CRPSSkill(ensemble_dim=self._ensemble_dim,
                               skipna_ensemble=self._skipna_ensemble),
        'CRPSSpread': CRPSSpread(ensemble_dim=self._ensemble_dim,
                                 use_sort=self._use_sort, fair=self._fair,
                                 skipna_ensemble=self._skipna_ensemble),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return statistic_values['CRPSSkill'] - 0.5 * statistic_values['CRPSSpread']


class CRPSEnsembleDistance(base.PerVariableMetric):
  """Unbiased CRPS distance between an ensemble forecast and an ensemble of
  targets: E|X - Y| - 0.5 E|X - X'| - 0.5 E|Y - Y'| (probabilistic.py:691-782).
  Both inputs carry ``ensemble_dim`` (sizes may differ)."""

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM, use_sort: bool = False,
               fair: bool = True, skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._use_sort = use_sort
    self._fair = fair
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'CRPSSkill': CRPSSkill(ensemble_dim=self._ensemble_dim),
        'CRPSSpread': CRPSSpread(
            ensemble_dim=self._ensemble_dim, use_sort=self._use_sort,
            fair=self._fair, skipna_ensemble=self._skipna_ensemble),
        'CRPSTargetSpread': CRPSSpread(
            ensemble_dim=self._ensemble_dim, use_sort=self._use_sort,
            fair=self._fair, which='targets'),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return (statistic_values['CRPSSkill']
            - 0.5 * statistic_values['CRPSSpread']
            - 0.5 * statistic_values['CRPSTargetSpread'])


class UnbiasedEnsembleMeanRMSE(base.PerVariableMetric):
  """Square root of the unbiased ensemble mean MSE (probabilistic.py:864-894)."""

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'UnbiasedEnsembleMeanSquaredError': UnbiasedEnsembleMeanSquaredError(
            ensemble_dim=self._ensemble_dim,
            skipna_ensemble=self._skipna_ensemble),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['UnbiasedEnsembleMeanSquaredError'])


def SpreadSkillRatio(**unused_kwargs):  # pylint: disable=invalid-name
  """Refuses like the reference (probabilistic.py:897-902)."""
  raise ValueError(
      'SpreadSkillRatio is not supported (the reference withdrew it as '
      'incorrectly implemented); use UnbiasedSpreadSkillRatio instead.')


class UnbiasedSpreadSkillRatio(base.PerVariableMetric):
  """sqrt(mean ensemble variance / unbiased ensemble mean MSE).

  Reference: probabilistic.py:905-967.
  """

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'EnsembleVariance': EnsembleVariance(
            ensemble_dim=self._ensemble_dim,
            skipna_ensemble=self._skipna_ensemble),
        'UnbiasedEnsembleMeanSquaredError': UnbiasedEnsembleMeanSquaredError(
            ensemble_dim=self._ensemble_dim,
            skipna_ensemble=self._skipna_ensemble),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['EnsembleVariance'] /
                     statistic_values['UnbiasedEnsembleMeanSquaredError'])


class EnsembleRootMeanVariance(base.PerVariableMetric):
  """Square root of the mean ensemble variance (probabilistic.py:970-1005)."""

  def __init__(self, ensemble_dim: str = ENSEMBLE_DIM,
               skipna_ensemble: bool = False):
    self._ensemble_dim = ensemble_dim
    self._skipna_ensemble = skipna_ensemble

  @property
  def statistics(self) -> Mapping[str, base.Statistic]:
    return {
        'EnsembleVariance': EnsembleVariance(
            ensemble_dim=self._ensemble_dim,
            skipna_ensemble=self._skipna_ensemble),
    }

  def _values_from_mean_statistics_per_variable(self, statistic_values):
    return np.sqrt(statistic_values['EnsembleVariance'])
