"""A small labelled-array type for the host side of the statistic/aggregation path.

The reference's plug-in surface is typed with ``xarray.DataArray``
(/root/reference/weatherbenchX/metrics/base.py:135-158,
aggregation.py:297-366).  xarray is not available on the build or GPU boxes,
so this module provides the subset of its behaviour that the path relies on:
named dims, coordinate variables (including non-index coordinates such as the
``mask`` coordinate produced by ``add_nan_mask_to_data``,
data_loaders/base.py:25-57), broadcast-by-dim-name arithmetic, ``where`` /
``isnull`` / reductions on *small host arrays* (aggregated sums and metric
values), ``dot`` and zero-filled outer ``align``.

It is not a general xarray replacement.  Label alignment is exact-match only:
operands whose shared index coordinates differ raise ``ValueError`` instead of
being inner-joined.  Arrays may be backed by a NumPy array (host) or by a
``torch.Tensor`` on a CUDA device (field data handed to the kernels); only
metadata operations are defined for device-backed arrays -- arithmetic on full
fields is the job of the CUDA kernels, never of this module.

When xarray *is* importable, ``from_xarray`` / ``to_xarray`` convert at the
boundary without copying the payload.
"""

from __future__ import annotations

import collections.abc
import numbers
from typing import Any, Hashable, Iterable, Mapping, Sequence

import numpy as np


def _is_device(data) -> bool:
  return type(data).__module__.startswith('torch') and hasattr(data, 'device')


def _as_payload(data):
  if _is_device(data):
    return data
  return np.asarray(data)


class _Coords(collections.abc.MutableMapping):
  """``da.coords``: a live view; assignment validates like the constructor.

  The reference assigns label arrays this way (``masks.coords[bin_dim] =
  np.array(labels)``, binning.py:135,181,199).
  """

  __slots__ = ('_owner',)

  def __init__(self, owner: 'DataArray'):
    self._owner = owner

  def __getitem__(self, key):
    return self._owner._coord_view(key)  # pylint: disable=protected-access

  def __setitem__(self, key, value):
    self._owner._set_coord(key, value)  # pylint: disable=protected-access

  def __delitem__(self, key):
    del self._owner._coords[key]  # pylint: disable=protected-access
    self._owner._version = getattr(self._owner, '_version', 0) + 1  # pylint: disable=protected-access

  def __iter__(self):
    return iter(self._owner._coords)  # pylint: disable=protected-access

  def __len__(self):
    return len(self._owner._coords)  # pylint: disable=protected-access

  def __contains__(self, key):
    return key in self._owner._coords  # pylint: disable=protected-access

  # Bulk iteration hands out the stored coordinate variables themselves (no
  # per-item view): the planner walks these on every aggregation.
  def values(self):
    return self._owner._coords.values()  # pylint: disable=protected-access

  def items(self):
    return self._owner._coords.items()  # pylint: disable=protected-access

  def keys(self):
    return self._owner._coords.keys()  # pylint: disable=protected-access

  def get(self, key, default=None):
    if key in self._owner._coords:  # pylint: disable=protected-access
      return self[key]
    return default

  def __repr__(self):
    return f'Coords({list(self)})'


class DataArray:
  """N-d array with named dims and coordinate variables."""

  __array_priority__ = 60
  _version = 0   # mutation stamp of the coordinates (bumped per instance)

  def __init__(self, data, dims: Sequence[Hashable] | str | None = None,
               coords: Mapping[Hashable, Any] | None = None,
               name: Hashable | None = None,
               attrs: Mapping | None = None):
    payload = _as_payload(data)
    if dims is None:
      if payload.ndim != 0 and not (coords and len(coords) == payload.ndim):
        dims = tuple(f'dim_{i}' for i in range(payload.ndim))
      elif payload.ndim != 0:
        dims = tuple(coords.keys())
      else:
        dims = ()
    if isinstance(dims, str):
      dims = (dims,)
    dims = tuple(dims)
    if len(dims) != payload.ndim:
      raise ValueError(f'{len(dims)} dims {dims} for {payload.ndim}-d data')
    if len(set(dims)) != len(dims):
      raise ValueError(f'duplicate dims {dims}')
    self._data = payload
    self.dims = dims
    self.name = name
    self.attrs = dict(attrs or {})
    self._coords: dict[Hashable, DataArray] = {}
    if coords:
      sizes = dict(zip(dims, payload.shape))
      for key, value in coords.items():
        self._set_coord(key, value, sizes)

  # -- construction helpers -------------------------------------------------

  def _set_coord(self, key, value, sizes=None):
    if isinstance(value, DataArray):
      cv = DataArray(value._data, value.dims, name=key, attrs=value.attrs)
    elif isinstance(value, tuple) and len(value) == 2 and (
        isinstance(value[0], str) or (
            isinstance(value[0], (tuple, list)) and
            all(isinstance(d, str) for d in value[0]))):
      cdims, cdata = value                    # (dims, data), dims a name or names
      cv = DataArray(cdata, cdims, name=key)
    else:
      arr = np.asarray(value)
      if arr.ndim == 0:
        cv = DataArray(arr, (), name=key)
      elif arr.ndim == 1:
        cv = DataArray(arr, (key,), name=key)
      else:
        raise ValueError(f'coordinate {key!r} needs explicit dims')
    if sizes is None:
      sizes = self.sizes
    for d, n in zip(cv.dims, cv.shape):
      if d not in sizes:
        raise ValueError(f'coordinate {key!r} has dim {d!r} not on the array')
      if sizes[d] != n:
        raise ValueError(
            f'coordinate {key!r} size {n} != array size {sizes[d]} on {d!r}')
    self._coords[key] = cv
    # mutation stamp: lets identity-keyed memos (fastpath.py) notice that the
    # coordinates of a live array were edited in place
    self._version = getattr(self, '_version', 0) + 1

  @classmethod
  def _fast(cls, payload, dims: tuple, coords: dict, name=None) -> 'DataArray':
    """Unchecked construction from parts that are known to be consistent
    (``coords`` maps names to coordinate DataArrays and is taken as is)."""
    out = cls.__new__(cls)
    out._data = payload
    out.dims = dims
    out.name = name
    out.attrs = {}
    out._coords = coords
    return out

  def _coord_view(self, key) -> 'DataArray':
    """A coordinate as an array that carries the coordinates on its own dims
    (``stat.latitude.latitude`` works, as binning.py:190-196 expects)."""
    cv = self._coords[key]
    if not cv.dims:
      return cv
    own = set(cv.dims)
    return cv._replace(coords={k: v for k, v in self._coords.items()
                               if v.dims and set(v.dims) <= own})

  def _replace(self, data=None, dims=None, coords=None, name='__keep__'):
    out = DataArray.__new__(DataArray)
    out._data = self._data if data is None else _as_payload(data)
    out.dims = self.dims if dims is None else tuple(dims)
    out.name = self.name if name == '__keep__' else name
    out.attrs = dict(self.attrs)
    out._coords = dict(self._coords if coords is None else coords)
    return out

  # -- basic properties -----------------------------------------------------

  @property
  def data(self):
    return self._data

  @property
  def values(self) -> np.ndarray:
    return self.to_numpy()

  def to_numpy(self) -> np.ndarray:
    if _is_device(self._data):
      return self._data.detach().cpu().numpy()
    return self._data

  @property
  def is_device(self) -> bool:
    return _is_device(self._data)

  @property
  def shape(self):
    return tuple(self._data.shape)

  @property
  def ndim(self):
    return len(self.dims)

  @property
  def size(self):
    return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

  @property
  def dtype(self):
    return self._data.dtype

  @property
  def sizes(self) -> dict:
    return dict(zip(self.dims, self.shape))

  @property
  def coords(self) -> '_Coords':
    return _Coords(self)

  def __len__(self):
    return self.shape[0]

  def __getitem__(self, key):
    if isinstance(key, Mapping):
      return self.isel(key)
    if key in self._coords:
      return self._coord_view(key)
    raise KeyError(key)

  def __getattr__(self, item):
    # Attribute-style coordinate access (``stat.mask``, ``da.latitude``), which
    # the reference uses together with hasattr() (aggregation.py:339).
    if item.startswith('_'):
      raise AttributeError(item)
    coords = self.__dict__.get('_coords', {})
    if item in coords:
      return self._coord_view(item)
    raise AttributeError(item)

  def __repr__(self):
    loc = 'device' if self.is_device else 'host'
    return (f'<wbx DataArray {self.name!r} {self.sizes} {self.dtype} {loc} '
            f'coords={list(self._coords)}>')

  def item(self):
    return self.to_numpy().item()

  def __float__(self):
    return float(self.to_numpy())

  def __bool__(self):
    return bool(self.to_numpy())

  def __array__(self, dtype=None, copy=None):
    arr = self.to_numpy()
    return arr.astype(dtype) if dtype is not None else arr

  # -- metadata operations (valid for host and device payloads) -------------

  def rename(self, new_name_or_dims=None, **dim_kwargs):
    if isinstance(new_name_or_dims, Mapping) or dim_kwargs:
      mapping = dict(new_name_or_dims or {}, **dim_kwargs)
      dims = tuple(mapping.get(d, d) for d in self.dims)
      coords = {}
      for k, cv in self._coords.items():
        nk = mapping.get(k, k)
        coords[nk] = cv._replace(
            dims=tuple(mapping.get(d, d) for d in cv.dims), name=nk)
      return self._replace(dims=dims, coords=coords)
    return self._replace(name=new_name_or_dims)

  def copy(self, deep: bool = True, data=None):
    if data is not None:
      return self._replace(data=data)
    if deep:
      payload = (self._data.clone() if self.is_device else self._data.copy())
      return self._replace(data=payload)
    return self._replace()

  def transpose(self, *dims):
    if not dims:
      dims = self.dims[::-1]
    if set(dims) != set(self.dims):
      raise ValueError(f'{dims} is not a permutation of {self.dims}')
    order = [self.dims.index(d) for d in dims]
    payload = (self._data.permute(*order) if self.is_device
               else np.transpose(self._data, order))
    return self._replace(data=payload, dims=dims)

  def isel(self, indexers: Mapping | None = None, drop: bool = False,
           **kwargs):
    indexers = dict(indexers or {}, **kwargs)
    for d in indexers:
      if d not in self.dims:
        raise ValueError(f'dim {d!r} not in {self.dims}')

    def apply(arr: 'DataArray'):
      key = tuple(indexers.get(d, slice(None)) for d in arr.dims)
      key = tuple(np.asarray(k) if isinstance(k, (list, tuple)) else k
                  for k in key)
      new_dims = tuple(d for d, k in zip(arr.dims, key)
                       if not isinstance(k, numbers.Integral))
      if _is_device(arr._data) and any(isinstance(k, np.ndarray) for k in key):
        import torch  # pylint: disable=g-import-not-at-top
        payload = arr._data
        for axis, k in enumerate(key):   # one index_select per array indexer
          if isinstance(k, np.ndarray):
            payload = payload.index_select(
                axis, torch.as_tensor(k.astype(np.int64), device=payload.device))
        rest = tuple(slice(None) if isinstance(k, np.ndarray) else k
                     for k in key)
        return payload[rest], new_dims
      return arr._data[key], new_dims

    payload, new_dims = apply(self)
    coords = {}
    for k, cv in self._coords.items():
      cdata, cdims = apply(cv)
      if drop and not cdims and k in indexers:
        continue
      coords[k] = cv._replace(data=cdata, dims=cdims, coords={})
    return self._replace(data=payload, dims=new_dims, coords=coords)

  def sel(self, indexers: Mapping | None = None, drop: bool = False, **kwargs):
    """Exact-label selection (scalars, lists or slices of labels)."""
    indexers = dict(indexers or {}, **kwargs)
    positional = {}
    for d, label in indexers.items():
      index = self._coords[d].to_numpy()
      if isinstance(label, slice):
        lo = -np.inf if label.start is None else label.start
        hi = np.inf if label.stop is None else label.stop
        positional[d] = np.nonzero((index >= lo) & (index <= hi))[0]
      elif np.ndim(label) == 0:
        hits = np.nonzero(index == label)[0]
        if not len(hits):
          raise KeyError(f'{label!r} not found along {d!r}')
        positional[d] = int(hits[0])
      else:
        lookup = {v: i for i, v in enumerate(index.tolist())}
        positional[d] = np.array([lookup[v] for v in np.asarray(label).tolist()])
    return self.isel(positional, drop=drop)

  def expand_dims(self, dim=None, axis: int = 0, **dim_kwargs):
    """Adds leading dims.  ``dim`` may be a name, or {name: size|labels}."""
    spec = {}
    if isinstance(dim, Mapping):
      spec.update(dim)
    elif dim is not None:
      spec[dim] = 1
    spec.update(dim_kwargs)
    out = self
    for pos, (name, value) in enumerate(spec.items()):
      labels = None
      if isinstance(value, numbers.Integral):
        n = int(value)
      else:
        labels = np.asarray(value)
        n = len(labels)
      arr = out.to_numpy() if not out.is_device else out._data
      if out.is_device:
        payload = arr.unsqueeze(axis + pos).expand(
            *arr.shape[:axis + pos], n, *arr.shape[axis + pos:])
      else:
        payload = np.broadcast_to(
            np.expand_dims(arr, axis + pos),
            arr.shape[:axis + pos] + (n,) + arr.shape[axis + pos:])
      dims = out.dims[:axis + pos] + (name,) + out.dims[axis + pos:]
      out = out._replace(data=payload, dims=dims)
      if labels is not None:
        out._coords[name] = DataArray(labels, (name,), name=name)
    return out

  def squeeze(self, dim=None, drop: bool = False):
    dims = [dim] if isinstance(dim, str) else (
        dim or [d for d, n in self.sizes.items() if n == 1])
    return self.isel({d: 0 for d in dims}, drop=drop)

  def drop_vars(self, names, errors: str = 'raise'):
    names = [names] if isinstance(names, str) else list(names)
    coords = {k: v for k, v in self._coords.items() if k not in names}
    return self._replace(coords=coords)

  drop = drop_vars

  def isin(self, test_elements) -> 'DataArray':
    self._require_host('isin')
    return self._replace(data=np.isin(self._data, np.asarray(test_elements)))

  @property
  def dt(self):
    """Datetime / timedelta field access (``valid_time.dt.dayofyear``,
    base.py:398-401; ``lead_time.dt.total_seconds()``, binning.py:376-389)."""
    kind = self.dtype.kind
    if kind == 'M':
      return DatetimeAccessor(self)
    if kind == 'm':
      return TimedeltaAccessor(self)
    raise AttributeError(
        f"'.dt' needs datetime64 or timedelta64 values, got {self.dtype}")

  def assign_coords(self, coords: Mapping | None = None, **kwargs):
    out = self._replace()
    for k, v in dict(coords or {}, **kwargs).items():
      out._set_coord(k, v)
    return out

  def broadcast_like(self, other: 'DataArray'):
    dims = tuple(d for d in other.dims if d not in self.dims) + self.dims
    sizes = dict(other.sizes, **self.sizes)
    arr = _expand(self.to_numpy(), self.dims, dims)
    arr = np.broadcast_to(arr, tuple(sizes[d] for d in dims))
    ordered = tuple(d for d in other.dims) + tuple(
        d for d in self.dims if d not in other.dims)
    out = self._replace(data=arr, dims=dims).transpose(*ordered)
    for k, cv in other._coords.items():
      if k not in out._coords and set(cv.dims) <= set(out.dims):
        out._coords[k] = cv
    return out

  def to_host(self) -> 'DataArray':
    if not self.is_device:
      return self
    return self._replace(data=self.to_numpy())

  def astype(self, dtype):
    self._require_host('astype')
    return self._replace(data=self._data.astype(dtype))

  # -- host arithmetic ------------------------------------------------------

  def _require_host(self, what: str):
    if self.is_device:
      raise TypeError(
          f'{what} on a device-resident DataArray is not defined on the host '
          'side; per-gridpoint arithmetic belongs to the CUDA kernels. Call '
          '.to_host() explicitly if a host copy is really wanted.')

  def _binary(self, other, op, reflexive=False):
    self._require_host('arithmetic')
    if isinstance(other, DataArray):
      other._require_host('arithmetic')
      if (self.dims == other.dims and self._data.shape == other._data.shape
          and _same_coords(self._coords, other._coords)):
        # same grid, same coordinate payloads (e.g. the two halves of an
        # AggregationState): no alignment work to do
        res = (op(other._data, self._data) if reflexive
               else op(self._data, other._data))
        return self._replace(
            data=res, name=self.name if self.name == other.name else None)
      dims, a, b = _broadcast_pair(self, other)
      coords = _merge_coords(self, other, dims)
      res = op(b, a) if reflexive else op(a, b)
      return DataArray(res, dims, coords=coords, name=self.name
                       if self.name == other.name else None)
    if hasattr(other, 'sum_along_dims') or isinstance(other, (dict, list)):
      return NotImplemented
    res = op(other, self._data) if reflexive else op(self._data, other)
    return self._replace(data=res)

  def __add__(self, o): return self._binary(o, np.add)
  def __radd__(self, o): return self._binary(o, np.add, True)
  def __sub__(self, o): return self._binary(o, np.subtract)
  def __rsub__(self, o): return self._binary(o, np.subtract, True)
  def __mul__(self, o): return self._binary(o, np.multiply)
  def __rmul__(self, o): return self._binary(o, np.multiply, True)
  def __truediv__(self, o): return self._binary(o, _divide)
  def __rtruediv__(self, o): return self._binary(o, _divide, True)
  def __pow__(self, o): return self._binary(o, np.power)
  def __floordiv__(self, o): return self._binary(o, np.floor_divide)
  def __mod__(self, o): return self._binary(o, np.mod)
  def __and__(self, o): return self._binary(o, np.logical_and)
  def __or__(self, o): return self._binary(o, np.logical_or)
  def __lt__(self, o): return self._binary(o, np.less)
  def __le__(self, o): return self._binary(o, np.less_equal)
  def __gt__(self, o): return self._binary(o, np.greater)
  def __ge__(self, o): return self._binary(o, np.greater_equal)
  def __eq__(self, o): return self._binary(o, np.equal)  # type: ignore
  def __ne__(self, o): return self._binary(o, np.not_equal)  # type: ignore
  __hash__ = None  # type: ignore

  def __neg__(self):
    self._require_host('negation')
    return self._replace(data=-self._data)

  def __invert__(self):
    self._require_host('invert')
    return self._replace(data=~self._data)

  def __abs__(self):
    self._require_host('abs')
    return self._replace(data=np.abs(self._data))

  def clip(self, min=None, max=None):  # pylint: disable=redefined-builtin
    self._require_host('clip')
    return self._replace(data=np.clip(self._data, min, max))

  def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
    if method != '__call__' or kwargs.get('out') is not None:
      return NotImplemented
    arrays = [x for x in inputs if isinstance(x, DataArray)]
    if len(inputs) == 1:
      self._require_host(ufunc.__name__)
      return self._replace(data=ufunc(self._data, **kwargs))
    if len(inputs) == 2:
      a, b = inputs
      if isinstance(a, DataArray):
        return a._binary(b, lambda x, y: ufunc(x, y, **kwargs))
      return b._binary(a, lambda x, y: ufunc(x, y, **kwargs), True)
    del arrays
    return NotImplemented

  def isnull(self):
    self._require_host('isnull')
    data = self._data
    if data.dtype.kind in 'fc':
      return self._replace(data=np.isnan(data))
    return self._replace(data=np.zeros(data.shape, dtype=bool))

  def notnull(self):
    return ~self.isnull()

  def where(self, cond, other=np.nan):
    self._require_host('where')
    if isinstance(cond, DataArray):
      dims, a, c = _broadcast_pair(self, cond)
      coords = _merge_coords(self, cond, dims)
    else:
      dims, a, c, coords = self.dims, self._data, np.asarray(cond), self._coords
    if isinstance(other, DataArray):
      o = _expand(other.to_numpy(), other.dims, dims)
    else:
      o = other
    if a.dtype.kind not in 'fc' and isinstance(o, float) and np.isnan(o):
      a = a.astype(np.float64)
    res = np.where(c, a, o)
    if a.dtype.kind == 'f' and not isinstance(other, DataArray):
      res = res.astype(a.dtype, copy=False)
    return DataArray(res, dims, coords=coords, name=self.name,
                     attrs=self.attrs)

  def fillna(self, value):
    return self.where(self.notnull(), value)

  def _reduce(self, func, nanfunc, dim, skipna, **kwargs):
    self._require_host('reduction')
    if dim is None:
      dims = self.dims
    elif isinstance(dim, str):
      dims = (dim,)
    else:
      dims = tuple(dim)
    for d in dims:
      if d not in self.dims:
        raise ValueError(f'{d!r} not found in array dimensions {self.dims}')
    axes = tuple(self.dims.index(d) for d in dims)
    use_nan = skipna if skipna is not None else self.dtype.kind in 'fc'
    fn = nanfunc if (use_nan and self.dtype.kind in 'fc') else func
    import warnings
    with warnings.catch_warnings(), np.errstate(invalid='ignore',
                                                divide='ignore'):
      warnings.simplefilter('ignore', RuntimeWarning)
      res = fn(self._data, axis=axes, **kwargs)
    new_dims = tuple(d for d in self.dims if d not in dims)
    coords = {k: v for k, v in self._coords.items()
              if set(v.dims) <= set(new_dims)}
    return DataArray(res, new_dims, coords=coords, name=self.name,
                     attrs=self.attrs)

  def sum(self, dim=None, skipna=None):
    return self._reduce(np.sum, np.nansum, dim, skipna)

  def mean(self, dim=None, skipna=None):
    return self._reduce(np.mean, np.nanmean, dim, skipna)

  def max(self, dim=None, skipna=None):
    return self._reduce(np.max, np.nanmax, dim, skipna)

  def min(self, dim=None, skipna=None):
    return self._reduce(np.min, np.nanmin, dim, skipna)

  def var(self, dim=None, skipna=None, ddof=0):
    return self._reduce(np.var, np.nanvar, dim, skipna, ddof=ddof)

  def count(self, dim=None):
    return self.notnull()._reduce(np.sum, np.sum, dim, False)

  def all(self, dim=None):
    return self._reduce(np.all, np.all, dim, False)

  def any(self, dim=None):
    return self._reduce(np.any, np.any, dim, False)

  def equals(self, other: 'DataArray') -> bool:
    if not isinstance(other, DataArray) or self.dims != other.dims:
      return False
    a, b = self.to_numpy(), other.to_numpy()
    return a.shape == b.shape and bool(np.array_equal(a, b, equal_nan=(
        a.dtype.kind in 'fc')))


def _divide(a, b):
  with np.errstate(invalid='ignore', divide='ignore'):
    return np.true_divide(a, b)


def _expand(arr: np.ndarray, arr_dims: Sequence, dims: Sequence) -> np.ndarray:
  """View of ``arr`` with singleton axes inserted to follow ``dims`` order."""
  arr_dims = tuple(arr_dims)
  present = [d for d in dims if d in arr_dims]
  if len(present) != len(arr_dims):
    raise ValueError(f'dims {arr_dims} not all contained in {tuple(dims)}')
  arr = np.transpose(arr, [arr_dims.index(d) for d in present])
  shape = [arr.shape[present.index(d)] if d in present else 1 for d in dims]
  return arr.reshape(shape)


def _same_coords(a: dict, b: dict) -> bool:
  """True if both coordinate dicts hold the very same payload objects."""
  if a is b:
    return True
  if len(a) != len(b):
    return False
  for k, va in a.items():
    vb = b.get(k)
    if vb is None or va._data is not vb._data or va.dims != vb.dims:  # pylint: disable=protected-access
      return False
  return True


def _check_index_coords(a: DataArray, b: DataArray):
  sa, sb = a.sizes, b.sizes
  ca, cb = a._coords, b._coords  # pylint: disable=protected-access
  for d in a.dims:
    if d in sb:
      if sa[d] != sb[d]:
        raise ValueError(
            f'size mismatch along {d!r}: {sa[d]} vs {sb[d]} '
            '(xarray_lite only supports exact alignment)')
      if d in ca and d in cb:
        ia, ib = ca[d]._data, cb[d]._data  # pylint: disable=protected-access
        if ia is ib:
          continue
        ia, ib = ca[d].to_numpy(), cb[d].to_numpy()
        if ia is not ib and not np.array_equal(ia, ib):
          raise ValueError(
              f'index coordinate {d!r} differs between operands; xarray_lite '
              'only supports exact alignment')


def reorder_like(da: DataArray, ref: DataArray, what: str = 'operand'
                 ) -> DataArray:
  """``da`` with its index coordinates in the label order of ``ref``.

  xarray aligns the operands of arithmetic and of ``xr.dot`` by coordinate
  LABEL, never by position.  Equal labels: ``da`` is returned as is; the same
  labels in another order (e.g. a climatology or a land-sea mask stored with
  descending latitude): a re-ordered copy; different label sets: ValueError
  (only exact alignment is provided).
  """
  indexers = {}
  for d in da.dims:
    if d not in ref.dims:
      continue
    a, b = da._coords.get(d), ref._coords.get(d)  # pylint: disable=protected-access
    if a is None or b is None or a._data is b._data:  # pylint: disable=protected-access
      continue
    la, lb = a.to_numpy(), b.to_numpy()
    if la.shape != lb.shape:
      raise ValueError(
          f'{what}: size mismatch along {d!r}: {la.shape[0]} vs {lb.shape[0]} '
          '(only exact alignment is supported)')
    if np.array_equal(la, lb):
      continue
    order = np.argsort(la, kind='stable')
    pos = np.clip(np.searchsorted(la[order], lb), 0, len(la) - 1)
    found = order[pos]
    if (not np.array_equal(la[found], lb) or
        len(np.unique(found)) != len(found)):
      raise ValueError(
          f'{what}: index coordinate {d!r} holds different labels than the '
          'data (only exact alignment is supported)')
    indexers[d] = found
  if not indexers:
    return da
  return da.isel(indexers)


def _broadcast_pair(a: DataArray, b: DataArray):
  _check_index_coords(a, b)
  dims = a.dims + tuple(d for d in b.dims if d not in a.dims)
  return dims, _expand(a.to_numpy(), a.dims, dims), _expand(
      b.to_numpy(), b.dims, dims)


def _merge_coords(a: DataArray, b: DataArray, dims) -> dict:
  """Union of coordinates; conflicting non-index coordinates are dropped."""
  coords = dict(a.coords)
  for k, cv in b.coords.items():
    if k not in coords:
      coords[k] = cv
    elif k not in dims:
      mine = coords[k]
      if mine.dims != cv.dims or not np.array_equal(
          mine.to_numpy(), cv.to_numpy(), equal_nan=False):
        del coords[k]
  return {k: v for k, v in coords.items() if set(v.dims) <= set(dims)}


# ---------------------------------------------------------------------------
# Module-level functions mirroring the xarray API used on the path
# ---------------------------------------------------------------------------


def ones_like(da: DataArray, dtype=None) -> DataArray:
  da._require_host('ones_like')
  return da._replace(data=np.ones_like(da.data, dtype=dtype))


def zeros_like(da: DataArray, dtype=None) -> DataArray:
  da._require_host('zeros_like')
  return da._replace(data=np.zeros_like(da.data, dtype=dtype))


def full_like(da: DataArray, fill_value, dtype=None) -> DataArray:
  da._require_host('full_like')
  return da._replace(data=np.full_like(da.data, fill_value, dtype=dtype))


def dot(*arrays: DataArray, dim=None) -> DataArray:
  """``xr.dot``: product of all arrays summed over ``dim`` (einsum)."""
  if dim is None:
    raise ValueError('dim must be given')
  reduce = {dim} if isinstance(dim, str) else set(dim)
  letters: dict = {}
  subs, payloads = [], []
  all_dims: list = []
  for arr in arrays:
    arr._require_host('dot')
    for d in arr.dims:
      if d not in letters:
        letters[d] = chr(ord('a') + len(letters))
        all_dims.append(d)
    subs.append(''.join(letters[d] for d in arr.dims))
    payloads.append(arr.to_numpy())
  for first, second in zip(arrays[:-1], arrays[1:]):
    _check_index_coords(first, second)
  out_dims = tuple(d for d in all_dims if d not in reduce)
  expr = ','.join(subs) + '->' + ''.join(letters[d] for d in out_dims)
  res = np.einsum(expr, *payloads)
  coords = {}
  for arr in arrays:
    for k, cv in arr.coords.items():
      if k not in coords and set(cv.dims) <= set(out_dims):
        coords[k] = cv
  return DataArray(res, out_dims, coords=coords, name=arrays[0].name,
                   attrs=arrays[0].attrs)


def align(*arrays: DataArray, join: str = 'outer', fill_value=0):
  """Outer join on index coordinates with ``fill_value`` (aggregation.py:58)."""
  if join != 'outer':
    raise NotImplementedError('only join="outer" is provided')
  union: dict = {}
  for arr in arrays:
    for d in arr.dims:
      if d in arr.coords:
        labels = arr.coords[d].to_numpy()
        if d not in union:
          union[d] = labels
        elif not np.array_equal(union[d], labels):
          union[d] = np.union1d(union[d], labels)
  out = []
  for arr in arrays:
    need = [d for d in arr.dims if d in union and d in arr.coords and
            not np.array_equal(union[d], arr.coords[d].to_numpy())]
    if not need:
      out.append(arr)
      continue
    host = arr.to_numpy()
    shape = tuple(len(union[d]) if d in need else n
                  for d, n in zip(arr.dims, arr.shape))
    filled = np.full(shape, fill_value, dtype=host.dtype)
    index = []
    for d in arr.dims:
      if d in need:
        index.append(np.searchsorted(union[d], arr.coords[d].to_numpy()))
      else:
        index.append(np.arange(arr.sizes[d]))
    filled[np.ix_(*index)] = host
    coords = {k: v for k, v in arr.coords.items()
              if not (set(v.dims) & set(need))}
    for d in need:
      coords[d] = DataArray(union[d], (d,), name=d)
    out.append(DataArray(filled, arr.dims, coords=coords, name=arr.name,
                         attrs=arr.attrs))
  return tuple(out)


def concat(arrays: Sequence[DataArray], dim: str) -> DataArray:
  first = arrays[0]
  if dim in first.dims:
    axis = first.dims.index(dim)
    payload = np.concatenate([a.to_numpy() for a in arrays], axis=axis)
    dims = first.dims
    labels = (np.concatenate([a.coords[dim].to_numpy() for a in arrays])
              if all(dim in a.coords for a in arrays) else None)
  else:
    payload = np.stack([a.to_numpy() for a in arrays], axis=0)
    dims = (dim,) + first.dims
    labels = None
  coords = {k: v for k, v in first.coords.items() if dim not in v.dims}
  out = DataArray(payload, dims, coords=coords, name=first.name)
  if labels is not None:
    out._coords[dim] = DataArray(labels, (dim,), name=dim)
  return out


class DatetimeAccessor:
  """The numeric calendar fields of a datetime64 array."""

  FIELDS = ('year', 'month', 'day', 'hour', 'minute', 'second', 'dayofyear',
            'dayofweek', 'weekday', 'quarter')

  def __init__(self, arr: DataArray):
    self._arr = arr

  def __getattr__(self, unit):
    if unit.startswith('_') or unit not in self.FIELDS:
      raise AttributeError(unit)
    t = self._arr.to_numpy().astype('datetime64[ns]')
    years = t.astype('datetime64[Y]')
    months = t.astype('datetime64[M]')
    days = t.astype('datetime64[D]')

    def since(coarse, fine_unit):
      fine = t.astype(f'datetime64[{fine_unit}]')
      return (fine - coarse.astype(f'datetime64[{fine_unit}]')).astype(np.int64)

    month = months.astype(np.int64) % 12 + 1
    values = {
        'year': lambda: years.astype(np.int64) + 1970,
        'month': lambda: month,
        'day': lambda: since(months, 'D') + 1,
        'hour': lambda: since(days, 'h'),
        'minute': lambda: since(t.astype('datetime64[h]'), 'm'),
        'second': lambda: since(t.astype('datetime64[m]'), 's'),
        'dayofyear': lambda: since(years, 'D') + 1,
        'dayofweek': lambda: (days.astype(np.int64) + 3) % 7,  # Monday = 0
        'weekday': lambda: (days.astype(np.int64) + 3) % 7,
        'quarter': lambda: (month - 1) // 3 + 1,
    }[unit]()
    return self._arr._replace(data=values)  # pylint: disable=protected-access


class TimedeltaAccessor:
  """total_seconds / days / seconds of a timedelta64 array."""

  def __init__(self, arr: DataArray):
    self._arr = arr

  def _ns(self):
    return self._arr.to_numpy().astype('timedelta64[ns]').astype(np.int64)

  def total_seconds(self) -> DataArray:
    return self._arr._replace(data=self._ns() / 1e9)  # pylint: disable=protected-access

  @property
  def days(self) -> DataArray:
    return self._arr._replace(data=self._ns() // (86400 * 10**9))  # pylint: disable=protected-access

  @property
  def seconds(self) -> DataArray:
    return self._arr._replace(  # pylint: disable=protected-access
        data=(self._ns() // 10**9) % 86400)


class Dataset(dict):
  """Mapping of variable name -> DataArray (what ``metric_values`` returns)."""

  @property
  def data_vars(self):
    return self

  @property
  def dims(self):
    dims: dict = {}
    for da in self.values():
      dims.update(da.sizes)
    return dims

  def __getattr__(self, item):
    for da in self.values():
      if item in da.coords:
        return da.coords[item]
    raise AttributeError(item)

  def rename(self, mapping=None, **kwargs):
    mapping = dict(mapping or {}, **kwargs)
    return Dataset({k: v.rename(
        {a: b for a, b in mapping.items() if a in v.dims or a in v.coords})
                    for k, v in self.items()})

  def map(self, fn):
    return Dataset({k: fn(v) for k, v in self.items()})


class DataTree:
  """The subset of ``xr.DataTree`` that AggregationState serialisation uses
  (aggregation.py:203-265 of the reference): a node holds a Dataset and named
  children; ``to_dict`` / ``from_dict`` map between a tree and
  ``{'/path/to/node': Dataset}``."""

  def __init__(self, dataset=None, children: Mapping | None = None,
               name: str | None = None):
    self.dataset = Dataset(dataset or {})
    self.name = name
    self.children: dict = {}
    for key, child in (children or {}).items():
      if not isinstance(child, DataTree):
        raise TypeError(f'child {key!r} is not a DataTree')
      if '/' in str(key):
        raise ValueError(f"node names cannot contain '/': {key!r}")
      child.name = str(key)
      self.children[str(key)] = child

  def __getitem__(self, path: str):
    node = self
    for part in [p for p in str(path).split('/') if p]:
      if part in node.children:
        node = node.children[part]
      elif part in node.dataset:
        return node.dataset[part]
      else:
        raise KeyError(path)
    return node

  @property
  def subtree(self):
    """(path, node) pairs, depth first, the root ('/') first."""
    stack = [('/', self)]
    while stack:
      path, node = stack.pop(0)
      yield path, node
      base = path.rstrip('/')
      stack = [(f'{base}/{k}', c) for k, c in node.children.items()] + stack

  def to_dict(self) -> dict:
    return {path: Dataset(node.dataset) for path, node in self.subtree}

  @classmethod
  def from_dict(cls, d: Mapping, name: str | None = None) -> 'DataTree':
    root = cls(name=name)
    for path, dataset in d.items():
      node = root
      for part in [p for p in str(path).split('/') if p]:
        if part not in node.children:
          node.children[part] = cls(name=part)
        node = node.children[part]
      node.dataset = Dataset(dataset or {})
    return root

  def __repr__(self):
    return (f'<wbx DataTree {self.name!r} vars={list(self.dataset)} '
            f'children={list(self.children)}>')


# ---------------------------------------------------------------------------
# xarray interop (only exercised when xarray is importable)
# ---------------------------------------------------------------------------


def from_xarray(obj):
  """xr.DataArray -> DataArray (payload shared, not copied)."""
  coords = {}
  for k, cv in obj.coords.items():
    coords[k] = DataArray(np.asarray(cv.data), tuple(cv.dims), name=k)
  return DataArray(obj.data, tuple(obj.dims), coords=coords, name=obj.name,
                   attrs=dict(obj.attrs))


def to_xarray(da: DataArray):
  import xarray as xr  # pylint: disable=g-import-not-at-top
  coords = {k: (cv.dims, cv.to_numpy()) for k, cv in da.coords.items()}
  return xr.DataArray(da.to_numpy(), dims=da.dims, coords=coords,
                      name=da.name, attrs=da.attrs)


def as_data_array(obj) -> DataArray:
  if isinstance(obj, DataArray):
    return obj
  if type(obj).__module__.startswith('xarray'):
    return from_xarray(obj)
  raise TypeError(f'expected a DataArray, got {type(obj)}')


class testing:  # pylint: disable=invalid-name
  """``xr.testing`` look-alike used by the ported reference tests."""

  @staticmethod
  def assert_allclose(a, b, rtol=1e-5, atol=1e-8, check_dim_order=False):
    if isinstance(a, Mapping):
      assert set(a) == set(b), (set(a), set(b))
      for k in a:
        testing.assert_allclose(a[k], b[k], rtol=rtol, atol=atol,
                                check_dim_order=check_dim_order)
      return
    if check_dim_order:
      assert a.dims == b.dims, (a.dims, b.dims)
    assert set(a.dims) == set(b.dims), (a.dims, b.dims)
    b = b.transpose(*a.dims)
    np.testing.assert_allclose(a.to_numpy(), b.to_numpy(), rtol=rtol,
                               atol=atol, equal_nan=True)
    for d in a.dims:
      if d in a.coords and d in b.coords:
        np.testing.assert_array_equal(a.coords[d].to_numpy(),
                                      b.coords[d].to_numpy())

  @staticmethod
  def assert_equal(a, b):
    assert a.dims == b.dims, (a.dims, b.dims)
    np.testing.assert_array_equal(a.to_numpy(), b.to_numpy())


def iter_leaves(tree) -> Iterable[DataArray]:
  if isinstance(tree, Mapping):
    for v in tree.values():
      yield from iter_leaves(v)
  elif tree is not None:
    yield tree
