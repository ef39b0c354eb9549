"""Runs the reference's OWN Python code in the build container.

TEST INFRASTRUCTURE (golden-vector generation only; never imported by the
product, by ``-m gpu`` tests, ``smoke()`` or ``bench.py`` -- /root/reference
does not exist on the GPU box).

``weatherbenchX`` is pure Python on top of xarray -> NumPy.  xarray, jax and
absl are not installed here and cannot be (no network), so ``import
weatherbenchX.aggregation`` fails out of the box.  ``install()`` registers
stand-ins under those module names *in this process only*:

* ``xarray`` / ``xarray.ufuncs``: the labelled-array container of
  ``weatherbenchx_b200.xarray_lite`` (dims, coords, broadcast-by-name,
  ``xr.dot`` = ``np.einsum`` exactly as xarray falls back to without
  opt_einsum, ``where`` / ``mean`` / ``var`` / ``sum`` = the NumPy functions
  xarray dispatches to), plus the few extra calls the reference's hot path
  makes (vectorised ``.sel`` with labelled indexers, ``.compute()``,
  ``xr.concat`` along a labelled new dim, ``xr.dot(dims=)``,
  ``xr.set_options``, list indexing of a Dataset, ``xr.core.accessor_dt``);
* ``jax`` / ``jax.numpy``: NumPy (only referenced by RMSE/ACC value functions
  for the autodiff tracing hook, metrics/deterministic.py:18-20);
* ``absl.logging``: the stdlib logger.

What this does and does not prove: every line of control flow and arithmetic
of ``weatherbenchX/{aggregation,weighting,binning}.py`` and
``weatherbenchX/metrics/{base,deterministic,probabilistic,wrappers}.py`` that
the golden cases touch is the reference's, executed unmodified from
/root/reference; the container semantics underneath (label alignment,
broadcasting, reductions) are the stand-in's restatement of xarray on top of
the same NumPy calls.  The vectors are therefore "reference code on stand-in
xarray", which is stated wherever they are used.
"""

from __future__ import annotations

import contextlib
import logging
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))))


def available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'weatherbenchX'))


def _patch_data_array(xl):
  """Adds the extra DataArray calls of the reference's hot path."""
  data_array = xl.DataArray
  plain_sel = data_array.sel

  def sel(self, indexers=None, drop=False, **kwargs):
    indexers = dict(indexers or {}, **kwargs)
    # a dimension without a coordinate is indexed by position, as xarray's
    # default RangeIndex does (metrics_test.py:983-1006 builds its climatology
    # with expand_dims(dayofyear=366, hour=4))
    missing = {d: np.arange(self.sizes[d]) for d in indexers
               if d in self.dims and d not in self.coords}
    if missing:
      self = self.assign_coords(missing)
    labelled = {d: v for d, v in indexers.items()
                if isinstance(v, data_array) and v.ndim > 0}
    if not labelled:
      return plain_sel(self, indexers, drop=drop)
    rest = {d: v for d, v in indexers.items() if d not in labelled}
    arr = plain_sel(self, rest, drop=drop) if rest else self
    # Vectorised (pointwise) selection: all labelled indexers share dims; the
    # new dims replace the first indexed dim (xarray's outer placement).
    new_dims = None
    positions = {}
    for d, label in labelled.items():
      if new_dims is None:
        new_dims = label.dims
      elif label.dims != new_dims:
        label = label.transpose(*new_dims)
      index = arr.coords[d].to_numpy()
      lookup = {v: i for i, v in enumerate(index.tolist())}
      flat = [lookup[v] for v in label.to_numpy().ravel().tolist()]
      positions[d] = np.asarray(flat).reshape(label.shape)
    first = min(arr.dims.index(d) for d in labelled)
    key = tuple(positions[d] if d in positions else slice(None)
                for d in arr.dims)
    payload = arr.to_numpy()[key]
    # NumPy puts the broadcast index dims first unless the advanced indices
    # are adjacent; normalise to "new dims at the first indexed position".
    indexed_axes = [arr.dims.index(d) for d in labelled]
    adjacent = indexed_axes == list(range(min(indexed_axes),
                                          max(indexed_axes) + 1))
    kept = [d for d in arr.dims if d not in labelled]
    if adjacent:
      dims = tuple(arr.dims[:first]) + tuple(new_dims) + tuple(
          d for d in arr.dims[first:] if d not in labelled)
    else:
      dims = tuple(new_dims) + tuple(kept)
    coords = {}
    for k, cv in arr.coords.items():
      if not set(cv.dims) & set(labelled):
        coords[k] = cv
    some = next(iter(labelled.values()))
    for k, cv in some.coords.items():
      if set(cv.dims) <= set(new_dims):
        coords.setdefault(k, cv)
    return data_array(payload, dims, coords=coords, name=arr.name,
                      attrs=arr.attrs)

  def setitem(self, key, value):
    """``da[{dim: index}] = value`` on a host payload (metrics_test.py:1225)."""
    if not isinstance(key, dict):
      raise TypeError('only dict indexers are supported')
    index = tuple(key.get(d, slice(None)) for d in self.dims)
    self.to_numpy()[index] = value

  data_array.__setitem__ = setitem
  data_array.sel = sel
  data_array.compute = lambda self: self
  data_array.load = lambda self: self


def install():
  """Registers the stand-in modules and returns the stand-in ``xarray``."""
  if 'weatherbenchX' in sys.modules:
    return sys.modules['xarray']
  if not available():
    raise RuntimeError(f'{REFERENCE_ROOT} is not present on this machine')
  sys.dont_write_bytecode = True  # /root/reference is read-only
  if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
  from weatherbenchx_b200 import xarray_lite as xl  # pylint: disable=g-import-not-at-top
  _patch_data_array(xl)

  xr = types.ModuleType('xarray')
  for key in dir(xl):
    if not key.startswith('_'):
      setattr(xr, key, getattr(xl, key))

  class Dataset(xl.Dataset):
    """Mapping of DataArrays with the Dataset calls the reference makes:
    ``ds[[names]]`` subsets (base.py:370) and -- for running the reference's
    own unit tests on this stand-in (tests/test_reference_own_tests.py) --
    construction from ``(dims, values)`` pairs with shared coordinates,
    ``expand_dims / rename / isel / sel / where / copy / mean``, coordinate
    access, and variable-wise arithmetic."""

    def __init__(self, data_vars=None, coords=None):
      dict.__init__(self)
      coords = dict(coords or {})
      for name, value in dict(data_vars or {}).items():
        if isinstance(value, tuple) and len(value) == 2:
          dims, values = value
          dims = (dims,) if isinstance(dims, str) else tuple(dims)
          value = xl.DataArray(
              np.asarray(values), dims,
              coords={d: np.asarray(coords[d]) for d in dims if d in coords},
              name=name)
        elif coords:
          value = xl.as_data_array(value)
          extra = {k: np.asarray(v) for k, v in coords.items()
                   if k in value.dims and k not in value.coords}
          if extra:
            value = value.assign_coords(extra)
        dict.__setitem__(self, name, value)

    def _map(self, fn, only_with=None):
      return Dataset({k: (fn(v) if only_with is None or
                          set(only_with) & set(v.dims) else v)
                      for k, v in dict.items(self)})

    def __getitem__(self, key):
      if isinstance(key, list):
        return Dataset({k: dict.__getitem__(self, k) for k in key})
      if key not in self.keys():
        for da in self.values():   # a coordinate shared by the variables
          if key in da.coords:
            return da.coords[key]
      return dict.__getitem__(self, key)

    @property
    def sizes(self):
      return self.dims

    @property
    def coords(self):
      out = {}
      for da in self.values():
        for k, v in da.coords.items():
          out.setdefault(k, v)
      return out

    def copy(self, deep=True):
      return self._map(lambda v: v.copy(deep=deep))

    def rename(self, mapping=None, **kwargs):
      mapping = dict(mapping or {}, **kwargs)
      out = Dataset()
      for k, v in dict.items(self):
        out[mapping.get(k, k)] = v.rename(
            {a: b for a, b in mapping.items() if a in v.dims or a in v.coords})
      return out

    def expand_dims(self, dim=None, **kwargs):
      spec = dict(dim or {}, **kwargs) if not isinstance(dim, str) else {
          dim: 1}
      return self._map(lambda v: v.expand_dims(spec))

    def isel(self, indexers=None, drop=False, **kwargs):
      indexers = dict(indexers or {}, **kwargs)
      return self._map(lambda v: v.isel(
          {d: i for d, i in indexers.items() if d in v.dims}, drop=drop))

    def sel(self, indexers=None, drop=False, **kwargs):
      indexers = dict(indexers or {}, **kwargs)
      return self._map(lambda v: v.sel(
          {d: i for d, i in indexers.items() if d in v.dims}, drop=drop))

    def where(self, cond, other=np.nan):
      return self._map(lambda v: v.where(cond, other))

    def mean(self, dim=None, skipna=None):
      dims = (dim,) if isinstance(dim, str) else tuple(dim or ())
      return self._map(lambda v: v.mean(
          [d for d in dims if d in v.dims] if dims else None, skipna=skipna))

    def isnull(self):
      return self._map(lambda v: v.isnull())

    def _binary(self, other, op):
      if isinstance(other, dict):
        return Dataset({k: op(v, other[k]) for k, v in dict.items(self)
                        if k in other})
      return self._map(lambda v: op(v, other))

    def __sub__(self, o): return self._binary(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._binary(o, lambda a, b: b - a)
    def __add__(self, o): return self._binary(o, lambda a, b: a + b)
    def __radd__(self, o): return self._binary(o, lambda a, b: b + a)
    def __mul__(self, o): return self._binary(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._binary(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._binary(o, lambda a, b: a / b)
    def __abs__(self): return self._map(abs)

  DataTree = xl.DataTree

  @contextlib.contextmanager
  def set_options(**unused_kwargs):
    yield

  def concat(arrays, dim):
    """xr.concat; ``dim`` may be a DataArray that names and labels the new
    dimension (categorical.py:223-226,283-285)."""
    if isinstance(dim, xl.DataArray):
      name = dim.dims[0]
      out = xl.concat(list(arrays), name)
      out._coords[name] = xl.DataArray(  # pylint: disable=protected-access
          dim.to_numpy(), (name,), name=name)
      return out
    return xl.concat(list(arrays), dim)

  def dot(*arrays, dim=None, dims=None):
    """xr.dot with the older ``dims=`` spelling (categorical.py:290)."""
    return xl.dot(*arrays, dim=dims if dim is None else dim)

  def like(fn):
    def wrapped(obj, dtype=None):
      if isinstance(obj, dict):
        return Dataset({k: fn(v, dtype) for k, v in obj.items()})
      return fn(obj, dtype)
    return wrapped

  xr.zeros_like = like(xl.zeros_like)
  xr.ones_like = like(xl.ones_like)
  if not hasattr(xl.DataArray, 'drop'):
    xl.DataArray.drop = xl.DataArray.drop_vars   # the older spelling
  xr.concat = concat
  xr.dot = dot
  xr.Dataset = Dataset
  xr.DataTree = DataTree
  # binning.py:376-391 dispatches on the accessor type
  xr.core = types.SimpleNamespace(accessor_dt=types.SimpleNamespace(
      TimedeltaAccessor=xl.TimedeltaAccessor,
      DatetimeAccessor=xl.DatetimeAccessor))
  xr.set_options = set_options
  xu = types.ModuleType('xarray.ufuncs')
  for name in ('sqrt', 'isnan', 'log', 'minimum', 'maximum', 'logical_and',
               'abs', 'square', 'exp'):
    setattr(xu, name, getattr(np, name))
  xr.ufuncs = xu
  sys.modules['xarray'] = xr
  sys.modules['xarray.ufuncs'] = xu

  jax = types.ModuleType('jax')
  jnp = types.ModuleType('jax.numpy')
  for key in dir(np):
    if not key.startswith('_'):
      setattr(jnp, key, getattr(np, key))
  jax.numpy = jnp
  jax.Array = np.ndarray
  jax.jit = lambda fn, **unused: fn
  jax.vmap = None
  sys.modules['jax'] = jax
  sys.modules['jax.numpy'] = jnp

  try:   # the real absl when the image has it (the reference's unit tests
    # need absl.testing); a logging-only stand-in otherwise
    import absl.logging  # noqa: F401  pylint: disable=g-import-not-at-top,unused-import
  except ImportError:
    absl = types.ModuleType('absl')
    absl.logging = logging
    sys.modules['absl'] = absl
    sys.modules['absl.logging'] = logging

  sys.path.insert(0, REFERENCE_ROOT)
  import weatherbenchX  # noqa: F401  pylint: disable=g-import-not-at-top,unused-import
  return xr
