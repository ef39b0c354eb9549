"""NetCDF-3 output / input of metric Datasets and aggregation states.

The reference writes its results with ``Dataset.to_netcdf()`` behind an atomic
rename (beam_pipeline.py:405-448, beam_utils.atomic_write).  Here the files are
classic NetCDF-3 written through ``scipy.io.netcdf_file`` (the only NetCDF
writer in this image); xarray opens them with its scipy engine.  Times are
stored CF-style as float64 ``seconds since 1970-01-01`` / ``seconds``; a
DataArray name that is not a legal NetCDF name is kept in the ``wbx_name``
attribute of a sanitised variable.
"""

from __future__ import annotations

import os
import re
import tempfile
from typing import Mapping

import numpy as np
from scipy.io import netcdf_file

from weatherbenchx_b200 import xarray_lite as xl

_EPOCH_UNITS = 'seconds since 1970-01-01 00:00:00'


def _legal(name: str) -> str:
  out = re.sub(r'[^A-Za-z0-9_.@+\-]', '_', str(name))
  return out if re.match(r'[A-Za-z_]', out) else '_' + out


def _encode(values: np.ndarray):
  """(array NetCDF-3 can hold, attributes)."""
  values = np.asarray(values)
  if values.dtype.kind == 'M':
    ns = values.astype('datetime64[ns]').astype(np.int64)
    return ns / 1e9, {'units': _EPOCH_UNITS, 'wbx_kind': 'datetime64'}
  if values.dtype.kind == 'm':
    ns = values.astype('timedelta64[ns]').astype(np.int64)
    return ns / 1e9, {'units': 'seconds', 'wbx_kind': 'timedelta64'}
  if values.dtype.kind == 'b':
    return values.astype(np.int8), {'wbx_kind': 'bool'}
  if values.dtype.kind in 'iu' and values.dtype.itemsize == 8:
    if values.size and (values.min() < -2**31 or values.max() >= 2**31):
      return values.astype(np.float64), {'wbx_kind': 'int64'}
    return values.astype(np.int32), {'wbx_kind': 'int64'}
  if values.dtype.kind in 'US' or values.dtype.kind == 'O':
    text = np.asarray(values, dtype=str)
    width = max(1, max((len(s) for s in text.ravel()), default=1))
    chars = np.array([list(s.ljust(width)) for s in text.ravel()], dtype='S1')
    return chars.reshape(text.shape + (width,)), {'wbx_kind': 'str'}
  return values, {}


def _decode(values: np.ndarray, attrs: Mapping):
  kind = attrs.get('wbx_kind')
  if isinstance(kind, bytes):
    kind = kind.decode()
  if kind == 'datetime64':
    return (np.round(np.asarray(values) * 1e9).astype(np.int64)
            ).astype('datetime64[ns]')
  if kind == 'timedelta64':
    return (np.round(np.asarray(values) * 1e9).astype(np.int64)
            ).astype('timedelta64[ns]')
  if kind == 'bool':
    return np.asarray(values).astype(bool)
  if kind == 'int64':
    return np.asarray(values).astype(np.int64)
  if kind == 'str':
    chars = np.asarray(values)
    flat = chars.reshape(-1, chars.shape[-1])
    text = np.array([b''.join(row).decode().rstrip() for row in flat])
    return text.reshape(chars.shape[:-1])
  values = np.array(values)
  # NetCDF-3 is big-endian on disk: hand back native byte order
  return values.astype(values.dtype.newbyteorder('='))


def to_netcdf(dataset: Mapping[str, xl.DataArray], path: str) -> None:
  """Writes {name: DataArray} to ``path`` atomically (write + rename)."""
  directory = os.path.dirname(os.path.abspath(path)) or '.'
  os.makedirs(directory, exist_ok=True)
  fd, tmp = tempfile.mkstemp(dir=directory, suffix='.tmp')
  os.close(fd)
  try:
    f = netcdf_file(tmp, 'w', version=2)
    dim_sizes: dict = {}

    def need_dim(name, size):
      name = _legal(name)
      if name in dim_sizes:
        if dim_sizes[name] != size:
          raise ValueError(f'dimension {name!r} has sizes {dim_sizes[name]} '
                           f'and {size} in one file')
      else:
        dim_sizes[name] = size
        f.createDimension(name, size)
      return name

    written: dict = {}   # legal NetCDF name -> (original name, host values)

    def write_var(name, dims, values, extra, is_coord=False):
      """Writes one variable and returns its NetCDF name.  A coordinate that
      several variables share is written once -- but only if it IS the same
      coordinate: equal name and different values raise instead of silently
      keeping the first.  Data variables whose names collide after replacing
      the characters NetCDF-3 forbids get a numeric suffix (the original name
      is kept in the `wbx_name` attribute and restored on reading)."""
      nc_name = _legal(name)
      raw_values = np.asarray(values)
      if nc_name in written:
        prev_name, prev_values = written[nc_name]
        if is_coord and prev_name == str(name):
          same = (prev_values.shape == raw_values.shape and (
              np.array_equal(prev_values, raw_values) or (
                  prev_values.dtype.kind == 'f' and raw_values.dtype.kind == 'f'
                  and np.array_equal(prev_values, raw_values, equal_nan=True))))
          if same:
            return nc_name
          raise ValueError(
              f'coordinate {name!r} has different values on two variables of '
              'one file; rename one of them before writing')
        k = 2
        while f'{nc_name}_{k}' in written:
          k += 1
        nc_name = f'{nc_name}_{k}'
      values, attrs = _encode(values)
      nc_dims = [need_dim(d, n) for d, n in zip(dims, values.shape)]
      if values.ndim > len(dims):  # char arrays
        nc_dims.append(need_dim(f'string{values.shape[-1]}', values.shape[-1]))
      if values.dtype == np.float16:
        values = values.astype(np.float32)
      var = f.createVariable(nc_name, values.dtype, tuple(nc_dims))
      if values.ndim:
        var[:] = values
      else:
        var.data[...] = values  # assignValue() indexes 0-d data with [:]
      for k, v in dict(attrs, **extra).items():
        setattr(var, k, v)
      if nc_name != str(name):
        var.wbx_name = str(name)
      written[nc_name] = (str(name), raw_values)
      return nc_name

    for name, da in dataset.items():
      da = xl.as_data_array(da)
      coord_names = []
      for cname, cv in da.coords.items():
        if cname == 'mask':
          continue
        nc_cname = write_var(cname, cv.dims, cv.to_numpy(), {}, is_coord=True)
        if cv.dims != (cname,):
          coord_names.append(nc_cname)
      extra = {'coordinates': ' '.join(coord_names)} if coord_names else {}
      write_var(name, da.dims, da.to_numpy(), extra)
    f.close()
    os.replace(tmp, path)
  except BaseException:
    if os.path.exists(tmp):
      os.remove(tmp)
    raise


def open_dataset(path: str) -> xl.Dataset:
  """Reads a file written by ``to_netcdf`` back into a Dataset."""
  with netcdf_file(path, 'r', mmap=False) as f:
    raw = {}
    for nc_name, var in f.variables.items():
      attrs = dict(var._attributes)  # pylint: disable=protected-access
      name = attrs.get('wbx_name', nc_name)
      if isinstance(name, bytes):
        name = name.decode()
      values = _decode(var.data if var.shape else var.getValue(), attrs)
      dims = tuple(var.dimensions)[:np.ndim(values)]
      raw[name] = (dims, values, attrs, nc_name)
  dim_names = {d for dims, _, _, _ in raw.values() for d in dims}
  coord_vars = {n for n in raw if n in dim_names}
  for _, _, attrs, _ in raw.values():
    listed = attrs.get('coordinates', b'')
    if isinstance(listed, bytes):
      listed = listed.decode()
    legal_to_name = {v[3]: k for k, v in raw.items()}
    coord_vars.update(legal_to_name.get(c, c) for c in listed.split())
  out = xl.Dataset()
  for name, (dims, values, attrs, _) in raw.items():
    if name in coord_vars:
      continue
    coords = {}
    for c in coord_vars:
      cdims, cvalues, _, _ = raw[c]
      if set(cdims) <= set(dims):
        coords[c] = xl.DataArray(cvalues, cdims, name=c)
    out[name] = xl.DataArray(values, dims, coords=coords, name=name)
  return out
