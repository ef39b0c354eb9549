"""Deferred per-gridpoint statistics.

``Statistic.compute`` in the reference returns fully materialised DataArrays
(metrics/base.py:135-158), which the Aggregator then reads again
(aggregation.py:337-366).  Here ``compute`` returns a ``LazyStatistic`` -- a
DataArray-shaped handle that records *which* statistic of *which* operands is
meant.  ``Aggregator`` recognises the handle and runs the fused CUDA kernel
(statistic + weights + reduction in one pass over HBM); any other consumer that
touches ``.data`` / ``.values`` gets the materialised field, evaluated on the
GPU by the elementwise kernel.
"""

from __future__ import annotations

import dataclasses
import weakref
from typing import Sequence

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl


# (id(pred coords), id(target coords)) -> weakrefs + mutation stamps of pairs
# whose index coordinates were already compared
_CHECKED_PAIRS: dict = {}


@dataclasses.dataclass
class AlignedClimatology:
  """Climatology + the (dayofyear[, hour]) gather that aligns it.

  Restates metrics/base.py:383-403 as index arrays instead of a materialised
  ``climatology.sel(...)``: ``time_dims`` are the prediction dims the valid
  time depends on, ``positions[clim_dim]`` the integer position along
  ``clim_dim`` for every valid time.
  """
  climatology: xl.DataArray          # dims e.g. (dayofyear, hour, level, lat, lon)
  time_dims: tuple                   # ('init_time', 'lead_time') or ('valid_time',)
  positions: dict                    # clim time dim -> int64 array over time_dims

  @property
  def clim_time_dims(self) -> tuple:
    return tuple(self.positions)

  @property
  def dims(self) -> tuple:
    rest = tuple(d for d in self.climatology.dims
                 if d not in self.positions)
    return self.time_dims + rest

  @property
  def sizes(self) -> dict:
    first = next(iter(self.positions.values()))
    out = dict(zip(self.time_dims, first.shape))
    for d, n in self.climatology.sizes.items():
      if d not in self.positions:
        out[d] = n
    return out


# Statistics whose defining expression touches only one of the two inputs.  In
# the reference the result of e.g. ``(predictions - climatology) ** 2`` carries
# the coordinates of the predictions and the climatology only -- in particular
# NOT the 'mask' coordinate that ``add_nan_mask_to_data`` puts on the targets
# (deterministic.py:225-232, probabilistic.py:241-247,266-273), so
# ``Aggregator(masked=True)`` leaves such a statistic unmasked
# (aggregation.py:339: ``hasattr(stat, 'mask')``).
_KIND_OPERANDS = {
    'SquaredPredictionAnomaly': ('predictions',),
    'SquaredTargetAnomaly': ('targets',),
    'CRPSSpread': ('predictions',),
    'EnsembleVariance': ('predictions',),
}


def _reference_mask(kind: str, predictions: xl.DataArray,
                    targets: xl.DataArray, dims, drop_dim=None):
  """The 'mask' coordinate the reference's result would carry, or None.

  The mask of the operand(s) the expression touches; when both carry one they
  must agree, otherwise xarray's coordinate merge drops it.
  """
  operands = {'predictions': predictions, 'targets': targets}
  found = []
  for which in _KIND_OPERANDS.get(kind, ('predictions', 'targets')):
    mask = operands[which]._coords.get('mask')  # pylint: disable=protected-access
    if mask is None or not set(mask.dims) <= set(dims):
      continue
    if drop_dim is not None and drop_dim in mask.dims:
      continue
    found.append(mask)
  if not found:
    return None
  if len(found) == 2:
    a, b = found
    if a is not b and a._data is not b._data and (  # pylint: disable=protected-access
        a.dims != b.dims or not np.array_equal(a.to_numpy(), b.to_numpy())):
      return None
  return found[0]


def _with_reference_mask(coords: dict, mask) -> dict:
  coords = dict(coords)
  coords.pop('mask', None)
  if mask is not None:
    coords['mask'] = mask
  return coords


def mask_identity(stat) -> int | None:
  """Groups statistics that can share a masked launch (same mask payload)."""
  mask = stat._coords.get('mask')  # pylint: disable=protected-access
  return None if mask is None else id(mask._data)  # pylint: disable=protected-access


class LazyStatistic(xl.DataArray):
  """A per-gridpoint statistic that is evaluated on demand."""

  # True when the field is `kind` applied elementwise to (predictions, targets
  # [, climatology]); composite handles (sums, member means) say False.
  elementwise_of_operands = True

  def __init__(self, kind: str, predictions: xl.DataArray,
               targets: xl.DataArray,
               climatology: AlignedClimatology | None = None):
    same_grid = (predictions.dims == targets.dims and
                 predictions.shape == targets.shape and
                 xl._same_coords(predictions._coords, targets._coords))  # pylint: disable=protected-access
    if not same_grid:
      # every statistic of a variable pairs the same two arrays: the label
      # comparison is made once per pair of coordinate sets
      pair = (id(predictions._coords), id(targets._coords))  # pylint: disable=protected-access
      hit = _CHECKED_PAIRS.get(pair)
      if not (hit is not None and hit[0]() is predictions and
              hit[1]() is targets and hit[2] == (
                  getattr(predictions, '_version', 0),
                  getattr(targets, '_version', 0))):
        xl._check_index_coords(predictions, targets)  # pylint: disable=protected-access
        try:
          if len(_CHECKED_PAIRS) > 256:
            _CHECKED_PAIRS.clear()
          _CHECKED_PAIRS[pair] = (
              weakref.ref(predictions), weakref.ref(targets),
              (getattr(predictions, '_version', 0),
               getattr(targets, '_version', 0)))
        except TypeError:
          pass
    dims = predictions.dims + tuple(
        d for d in targets.dims if d not in predictions.dims)
    sizes = dict(targets.sizes, **predictions.sizes)
    if climatology is not None:
      for d, n in climatology.sizes.items():
        if d in sizes and sizes[d] != n:
          raise ValueError(
              f'climatology size {n} != data size {sizes[d]} along {d!r}')
        if d not in sizes:
          dims = dims + (d,)
          sizes[d] = n
    self.kind = kind
    self.predictions = predictions
    self.targets = targets
    self.climatology = climatology
    self.dims = dims
    self._sizes = {d: sizes[d] for d in dims}
    self.name = predictions.name
    self.attrs = {}
    coords = (dict(predictions._coords) if same_grid else  # pylint: disable=protected-access
              xl._merge_coords(predictions, targets, dims))  # pylint: disable=protected-access
    if 'mask' in predictions._coords or 'mask' in targets._coords:  # pylint: disable=protected-access
      coords = _with_reference_mask(
          coords, _reference_mask(kind, predictions, targets, dims))
    self._coords = coords
    self._materialized = None

  # -- metadata without materialising ---------------------------------------

  @property
  def shape(self):
    return tuple(self._sizes[d] for d in self.dims)

  @property
  def sizes(self):
    return dict(self._sizes)

  @property
  def dtype(self):
    return np.dtype(np.float32)

  @property
  def is_device(self) -> bool:
    return True

  @property
  def is_lazy(self) -> bool:
    return self._materialized is None

  def group_key(self):
    """Statistics with the same key can share one pass over the operands."""
    clim = self.climatology
    return (id(self.predictions), id(self.targets),
            id(clim.climatology) if clim is not None else None)

  def __repr__(self):
    return f'<LazyStatistic {self.kind} {self.name!r} {self.sizes}>'

  # -- materialisation --------------------------------------------------------

  @property
  def _data(self):
    if self._materialized is None:
      from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
      self._materialized = engine.materialize(self)
    return self._materialized

  @_data.setter
  def _data(self, value):
    self._materialized = value

  def _replace(self, data=None, dims=None, coords=None, name='__keep__'):
    out = xl.DataArray.__new__(xl.DataArray)
    out._data = self._data if data is None else xl._as_payload(data)  # pylint: disable=protected-access
    out.dims = self.dims if dims is None else tuple(dims)
    out.name = self.name if name == '__keep__' else name
    out.attrs = dict(self.attrs)
    out._coords = dict(self._coords if coords is None else coords)
    return out


class LazyPassthrough(LazyStatistic):
  """``source + zeros_like(other)`` (Prediction/TargetPassthrough,
  deterministic.py:138-147,162-171) as the fused 'Error' statistic of
  ``source`` against a shared all-zero slab: ``source - 0`` is exact, NaN
  propagates, and the kernel reads ``source`` from HBM once while the zero slab
  stays in L2.  Coordinates follow the reference: those of both inputs.
  """

  def __init__(self, source: xl.DataArray, other: xl.DataArray):
    from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
    if not set(other.dims) <= set(source.dims):
      raise ValueError('passthrough: the other input has extra dims')
    inner = source.dims[-2:]
    zeros = engine.zero_slab(
        inner, tuple(source.sizes[d] for d in inner),
        source.data.device if source.is_device else None)
    super().__init__('Error', source, zeros)
    xl._check_index_coords(source, other)  # pylint: disable=protected-access
    self._coords = xl._merge_coords(source, other, self.dims)  # pylint: disable=protected-access


class LazyEnsembleAveraged(LazyStatistic):
  """Mean of a lazy statistic over the ensemble dimension
  (EnsembleAveragedStatistic.compute, probabilistic.py:56-69).

  Without NaN skipping the mean over members followed by the weighted spatial
  sums equals the fused reduction of the inner statistic over
  ``reduce_dims + [ensemble_dim]`` divided by the member count; that is how the
  Aggregator evaluates it.  ``.values`` gives the per-point mean field.
  """

  elementwise_of_operands = False

  def __init__(self, inner: LazyStatistic, ensemble_dim, skipna_ensemble: bool,
               optimistic: bool = False):
    if ensemble_dim not in inner.dims:
      raise ValueError(f'Dimension {ensemble_dim} not found in {inner.dims}')
    # optimistic: a NaN-skipping member mean (xarray's default ``.mean``) equals
    # the plain one whenever no NaN takes part; the Aggregator then tries the
    # fused reduction first and only falls back to the per-point mean field when
    # the result shows that a NaN was met.
    self.optimistic = bool(optimistic)
    self.kind = inner.kind
    self.inner = inner
    self.ensemble_dim = ensemble_dim
    self.skipna_ensemble = bool(skipna_ensemble)
    self.predictions = inner.predictions
    self.targets = inner.targets
    self.climatology = inner.climatology
    self.dims = tuple(d for d in inner.dims if d != ensemble_dim)
    self._sizes = {d: inner.sizes[d] for d in self.dims}
    self.name = inner.name
    self.attrs = {}
    self._coords = {k: v for k, v in inner.coords.items()
                    if ensemble_dim not in v.dims}
    self._materialized = None

  @property
  def n_members(self) -> int:
    return self.inner.sizes[self.ensemble_dim]

  def group_key(self):
    return ('ensemble-mean', self.ensemble_dim) + tuple(self.inner.group_key())

  @property
  def _data(self):
    if self._materialized is None:
      from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
      field = xl.DataArray(self.inner.data, self.inner.dims,
                           coords=self.inner.coords, name=self.name)
      self._materialized = engine.ensemble_mean(
          field, self.ensemble_dim, skipna=self.skipna_ensemble).data
    return self._materialized

  @_data.setter
  def _data(self, value):
    self._materialized = value


class LazySumStatistic(LazyStatistic):
  """Sum of lazy statistics on a common grid (WindVectorSquaredError =
  SquaredError(u) + SquaredError(v), deterministic.py:206-219).

  The Aggregator evaluates the parts in their own fused launches -- shared with
  any other statistic of the same operands, e.g. the per-component RMSE -- and
  adds the states; ``.values`` gives the summed field.
  """

  elementwise_of_operands = False

  def __init__(self, kind: str, parts: Sequence[LazyStatistic], name=None,
               scale: float = 1.0):
    first = parts[0]
    # the field is scale * (part0 + part1 + ...): 1 for sums, 1 / n for the
    # mean over an ensemble of targets (CRPSSkill, probabilistic.py:135-145)
    self.scale = float(scale)
    for p in parts[1:]:
      if p.dims != first.dims or p.sizes != first.sizes:
        raise ValueError(
            f'cannot add statistics on different grids: {first.sizes} vs '
            f'{p.sizes}')
    self.kind = kind
    self.parts = tuple(parts)
    self.predictions = first.predictions
    self.targets = first.targets
    self.climatology = None
    self.dims = first.dims
    self._sizes = dict(first.sizes)
    self.name = name
    self.attrs = {}
    self._coords = dict(first.coords)
    self._materialized = None

  def group_key(self):
    return ('sum',) + tuple(p.group_key() for p in self.parts)


class LazyEnsembleStatistic(LazyStatistic):
  """CRPSSkill / CRPSSpread of an ensemble prediction (deferred).

  ``predictions`` carries ``ensemble_dim``; the statistic does not.
  """

  def __init__(self, kind: str, predictions: xl.DataArray,
               targets: xl.DataArray, ensemble_dim: str, fair: bool,
               skipna_ensemble: bool, use_sort: bool = False):
    if ensemble_dim not in predictions.dims:
      raise ValueError(
          f'Dimension {ensemble_dim} not found in {predictions.dims}')
    if ensemble_dim in targets.dims:
      # CRPSSkill / CRPSSpread split an ensemble of targets into member views
      # before they get here (metrics/probabilistic.py)
      raise NotImplementedError(
          f'{kind} with an ensemble of targets is outside the B200 hot path')
    xl._check_index_coords(predictions, targets)  # pylint: disable=protected-access
    pdims = tuple(d for d in predictions.dims if d != ensemble_dim)
    dims = pdims + tuple(d for d in targets.dims if d not in pdims)
    sizes = dict(targets.sizes, **predictions.sizes)
    self.kind = kind
    self.predictions = predictions
    self.targets = targets
    self.climatology = None
    self.ensemble_dim = ensemble_dim
    self.fair = bool(fair)
    self.skipna_ensemble = bool(skipna_ensemble)
    self.use_sort = bool(use_sort)
    self.dims = dims
    self._sizes = {d: sizes[d] for d in dims}
    self.name = predictions.name
    self.attrs = {}
    coords = xl._merge_coords(predictions, targets, dims)  # pylint: disable=protected-access
    coords = {k: v for k, v in coords.items() if ensemble_dim not in v.dims}
    if 'mask' in predictions._coords or 'mask' in targets._coords:  # pylint: disable=protected-access
      coords = _with_reference_mask(coords, _reference_mask(
          kind, predictions, targets, dims, drop_dim=ensemble_dim))
    self._coords = coords
    self._materialized = None

  @property
  def n_members(self) -> int:
    return self.predictions.sizes[self.ensemble_dim]


def threshold_f32(thresholds) -> np.ndarray:
  """float32 thresholds t' with ``x > t'`` == ``x > t`` for every float32 x.

  The reference compares float32 fields with float64 thresholds (a Python list
  turned into a DataArray, wrappers.py:85-88, deterministic.py:272-277); NumPy
  promotes the field, so a threshold that is not a float32 number (0.1, 0.25
  ...) must be rounded DOWN, not to nearest, for the float32 comparison in the
  kernel to give the same answer at the boundary.  NaN stays NaN.
  """
  t64 = np.asarray(thresholds, dtype=np.float64)
  with np.errstate(over='ignore', invalid='ignore'):
    t32 = t64.astype(np.float32)
    above = t32.astype(np.float64) > t64
    t32 = np.where(above, np.nextafter(t32, np.float32(-np.inf)), t32)
  return np.ascontiguousarray(t32, dtype=np.float32)


class LazyBinarized(xl.DataArray):
  """``binarize_thresholds(x, thresholds, threshold_dim)`` (wrappers.py:50-88)
  as a handle: ``(x > threshold).where(~isnan(x)).astype(float32)`` with the
  threshold dim appended.  The categorical statistics read ``source`` and the
  thresholds and compare inside the fused reduction; nothing is written to
  HBM unless a caller touches ``.data`` (then the elementwise kernel
  materialises the [..., threshold] field).
  """

  def __init__(self, source: xl.DataArray, thresholds, threshold_dim):
    source = xl.as_data_array(source)
    # a float64 vector is kept as the object it is: the planner recognises a
    # repeated request by the identity of every payload involved, and the
    # threshold labels become a coordinate of the statistic
    labels = np.asarray(thresholds, dtype=np.float64)
    if labels.ndim != 1:
      labels = labels.reshape(-1)
    if threshold_dim in source.dims:
      raise ValueError(
          f'{threshold_dim!r} is already a dimension of the input')
    self.source = source
    self.threshold_dim = threshold_dim
    self.threshold_labels = labels
    self.thresholds = threshold_f32(labels)
    self.dims = tuple(source.dims) + (threshold_dim,)
    self._sizes = dict(source.sizes)
    self._sizes[threshold_dim] = len(labels)
    self.name = source.name
    self.attrs = {}
    coords = dict(source._coords)  # pylint: disable=protected-access
    coords[threshold_dim] = xl.DataArray(labels, (threshold_dim,),
                                         name=threshold_dim)
    self._coords = coords
    self._materialized = None

  @property
  def shape(self):
    return tuple(self._sizes[d] for d in self.dims)

  @property
  def sizes(self):
    return dict(self._sizes)

  @property
  def dtype(self):
    return np.dtype(np.float32)

  @property
  def is_device(self) -> bool:
    return True

  @property
  def is_lazy(self) -> bool:
    return self._materialized is None

  def __repr__(self):
    return f'<LazyBinarized {self.name!r} {self.sizes}>'

  @property
  def _data(self):
    if self._materialized is None:
      from weatherbenchx_b200 import engine  # pylint: disable=g-import-not-at-top
      self._materialized = engine.materialize_binarized(self)
    return self._materialized

  @_data.setter
  def _data(self, value):
    self._materialized = value

  _replace = LazyStatistic._replace


CONTINGENCY_KINDS = ('TruePositives', 'FalsePositives', 'FalseNegatives',
                     'TrueNegatives')


class LazyCategoricalStatistic(LazyStatistic):
  """TruePositives / FalsePositives / FalseNegatives / TrueNegatives
  (categorical.py:25-101) of inputs that are thresholded on the fly, and
  ErrorExceedance (deterministic.py:262-295).

  The operands the kernels read are the *continuous* fields: an input that came
  through ``ContinuousToBinary`` (a LazyBinarized handle) contributes its
  source and its thresholds, any other input is taken as binary already
  (``.astype(bool)``: non-zero is True).  One launch evaluates the whole
  contingency table of a variable for all thresholds; the threshold index is a
  kept outer dim of the job table (``plan_dims`` puts it first so that the
  grid dims stay the contiguous slab).
  """

  elementwise_of_operands = False

  def __init__(self, kind: str, predictions: xl.DataArray,
               targets: xl.DataArray, exceedance_thresholds=None,
               exceedance_dim=None, exceedance_coord=None):
    from weatherbenchx_b200 import _cabi  # pylint: disable=g-import-not-at-top
    p_bin = predictions if (isinstance(predictions, LazyBinarized)
                            and predictions.is_lazy) else None
    t_bin = targets if (isinstance(targets, LazyBinarized)
                        and targets.is_lazy) else None
    p_src = p_bin.source if p_bin is not None else predictions
    t_src = t_bin.source if t_bin is not None else targets
    super().__init__(kind, p_src, t_src)
    self.thr_pred = self.thr_target = None
    thr_dim = labels = thr_coord = None
    if kind == 'ErrorExceedance':
      self.xform = _cabi.XF_ERROR_EXCEEDANCE
      if p_bin is not None or t_bin is not None:
        raise NotImplementedError('ErrorExceedance of thresholded inputs')
      thr_dim = exceedance_dim
      labels = np.asarray(exceedance_thresholds, dtype=np.float64).reshape(-1)
      thr_coord = exceedance_coord
      self.thr_pred = threshold_f32(labels)
    else:
      if kind not in CONTINGENCY_KINDS:
        raise ValueError(f'unknown categorical statistic {kind!r}')
      self.xform = _cabi.XF_CONTINGENCY
      if p_bin is not None and t_bin is not None and (
          p_bin.threshold_dim != t_bin.threshold_dim or
          not np.array_equal(p_bin.threshold_labels, t_bin.threshold_labels,
                             equal_nan=True)):
        raise NotImplementedError(
            'predictions and targets thresholded along different dims / '
            'with different values')
      for which, b in (('pred', p_bin), ('target', t_bin)):
        if b is None:
          self.xform |= (_cabi.XF_PRED_NONZERO if which == 'pred'
                         else _cabi.XF_TARGET_NONZERO)
        else:
          thr_dim, labels = b.threshold_dim, b.threshold_labels
          thr_coord = b._coords.get(thr_dim)  # pylint: disable=protected-access
          setattr(self, f'thr_{which}', b.thresholds)
    self.threshold_dim = thr_dim
    base_dims = self.dims
    if thr_dim is not None:
      if thr_dim in base_dims:
        raise ValueError(f'{thr_dim!r} is already a dimension of the inputs')
      # dims as xarray broadcasting orders them (wrappers.py:88,
      # categorical.py:37-41, deterministic.py:290)
      if kind != 'ErrorExceedance' and p_bin is not None:
        n_p = len(p_src.dims)
        self.dims = base_dims[:n_p] + (thr_dim,) + base_dims[n_p:]
      else:
        self.dims = base_dims + (thr_dim,)
      self._sizes[thr_dim] = len(labels)
      self._sizes = {d: self._sizes[d] for d in self.dims}
      self._coords = dict(self._coords)
      if thr_coord is not None:
        self._coords[thr_dim] = xl.DataArray(
            thr_coord.to_numpy(), (thr_dim,), name=thr_dim)
      self.plan_dims = (thr_dim,) + base_dims
    else:
      self.plan_dims = base_dims

  def group_key(self):
    def key(a):
      return None if a is None else a.tobytes()
    return (('xf', self.xform, self.threshold_dim, key(self.thr_pred),
             key(self.thr_target), id(self.predictions)),
            id(self.targets), None)

  def __repr__(self):
    return f'<LazyCategoricalStatistic {self.kind} {self.name!r} {self.sizes}>'


def statistic_names(stats: Sequence[LazyStatistic]) -> list:
  return [s.kind for s in stats]
