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
    """Mapping of DataArrays; ``ds[[names]]`` subsets (base.py:370)."""

    def __getitem__(self, key):
      if isinstance(key, list):
        return Dataset({k: dict.__getitem__(self, k) for k in key})
      return dict.__getitem__(self, key)

  class DataTree:  # only named in annotations
    pass

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

  absl = types.ModuleType('absl')
  absl.logging = logging
  sys.modules['absl'] = absl
  sys.modules['absl.logging'] = logging

  sys.path.insert(0, REFERENCE_ROOT)
  import weatherbenchX  # noqa: F401  pylint: disable=g-import-not-at-top,unused-import
  return xr
