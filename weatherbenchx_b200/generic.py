"""Host side of the generic strided aggregation (wbx_reduce_generic).

Used by ``aggregation.Aggregator`` for everything the fused slab kernel cannot
express: bin masks (aggregation.py:320-335), N-d weights, statistics whose
reduced dims are not the trailing contiguous ones, broadcasting inside the
reduced dims, and statistics that are not lazy (already materialised).  The
statistic (when lazy), the mask / skipna logic, the weights, the bin masks and
the reduction all happen in one kernel; there is no host arithmetic on fields.
"""

from __future__ import annotations

import ctypes
from ctypes import c_int64
from typing import Hashable, Sequence

import numpy as np

from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.lazy import LazyStatistic

MAX_DIMS, MAX_FACTORS = _cabi.MAX_DIMS, _cabi.MAX_FACTORS
DTYPE_F64, DTYPE_F32, DTYPE_U8 = _cabi.DTYPE_F64, _cabi.DTYPE_F32, _cabi.DTYPE_U8
GenericDesc = _cabi.GenericDesc


def _device_tensor(da: xl.DataArray, kind: str, device=None):
  """(tensor, dims) on the GPU with a dtype the kernel understands."""
  torch = engine._torch()  # pylint: disable=protected-access
  payload = da.data
  if not xl._is_device(payload):  # pylint: disable=protected-access
    arr = np.asarray(payload)
    if kind == 'field':
      arr = arr.astype(np.float32, copy=False)
    elif arr.dtype == np.bool_:
      arr = arr.view(np.uint8)
    elif arr.dtype.kind in 'iu':
      arr = arr.astype(np.float64)
    elif arr.dtype not in (np.float32, np.float64):
      arr = arr.astype(np.float64)
    dev = torch.device('cuda', torch.cuda.current_device()
                       if device is None else device)
    payload = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
  else:
    if kind == 'field' and payload.dtype != torch.float32:
      payload = payload.to(torch.float32)
    elif payload.dtype == torch.bool:
      payload = payload.view(torch.uint8)
    elif payload.dtype not in (torch.float32, torch.float64, torch.uint8):
      payload = payload.to(torch.float64)
  return payload, da.dims


def _dtype_code(tensor) -> int:
  torch = engine._torch()  # pylint: disable=protected-access
  return {torch.float64: DTYPE_F64, torch.float32: DTYPE_F32,
          torch.uint8: DTYPE_U8}[tensor.dtype]


def _strides(tensor, tdims, dims):
  out = (c_int64 * MAX_DIMS)()
  st = dict(zip(tdims, tensor.stride()))
  sz = dict(zip(tdims, tensor.shape))
  for i, d in enumerate(dims):
    out[i] = st[d] if d in st and sz[d] != 1 else 0
  return out


def aggregate(stat: xl.DataArray, factors: Sequence[xl.DataArray],
              reduce_dims, *, mask: xl.DataArray | None = None,
              skipna: bool = False, extra_dims: Sequence[Hashable] = (),
              device: int | None = None):
  """(sum_weighted_statistics, sum_weights) of ``stat`` over ``reduce_dims``.

  ``factors`` are weights / bin masks whose dims are stat dims or one of
  ``extra_dims`` (the bin dims, kept in the output after the stat dims).
  """
  reduce_set = set(reduce_dims)
  dims = tuple(stat.dims) + tuple(extra_dims)
  if len(dims) > MAX_DIMS:
    raise NotImplementedError(f'more than {MAX_DIMS} dims: {dims}')
  if len(factors) > MAX_FACTORS:
    raise NotImplementedError(f'more than {MAX_FACTORS} weight/bin operands')
  sizes = dict(stat.sizes)
  for f in factors:
    for d, n in f.sizes.items():
      if d in sizes and sizes[d] != n and n != 1:
        raise ValueError(f'operand size {n} != {sizes[d]} along {d!r}')
      sizes.setdefault(d, n)
  for d in extra_dims:
    if d not in sizes:
      raise ValueError(f'bin dim {d!r} not found on any operand')

  keep = []  # tensors must outlive the launch
  desc = GenericDesc()
  desc.ndim = len(dims)
  desc.flags = _cabi.FLAG_SKIPNA if skipna else 0
  for i, d in enumerate(dims):
    desc.size[i] = sizes[d]
    desc.reduced[i] = 1 if d in reduce_set else 0

  if (isinstance(stat, LazyStatistic) and stat.is_lazy and
      (stat.kind not in _cabi.STAT_SLOT or not stat.elementwise_of_operands)):
    # CRPS fields, sums of statistics, member means: not an elementwise
    # function of (predictions, targets[, climatology]) -- materialise
    stat = stat._replace()  # pylint: disable=protected-access
  if isinstance(stat, LazyStatistic) and stat.is_lazy:
    desc.op = _cabi.STAT_SLOT[stat.kind]
    ta, da_ = _device_tensor(stat.predictions, 'field', device)
    tb, db_ = _device_tensor(stat.targets, 'field', device)
    keep += [ta, tb]
    desc.a, desc.a_stride = ta.data_ptr(), _strides(ta, da_, dims)
    desc.b, desc.b_stride = tb.data_ptr(), _strides(tb, db_, dims)
    if stat.climatology is not None:
      tc, dc_ = _aligned_climatology_tensor(stat, device)
      keep.append(tc)
      desc.c, desc.c_stride = tc.data_ptr(), _strides(tc, dc_, dims)
  else:
    desc.op = -1
    ta, da_ = _device_tensor(stat, 'field', device)
    keep.append(ta)
    desc.a, desc.a_stride = ta.data_ptr(), _strides(ta, da_, dims)
  if mask is not None:
    tm, dm_ = _device_tensor(mask, 'mask', device)
    torch = engine._torch()  # pylint: disable=protected-access
    if tm.dtype != torch.uint8:
      tm = (tm != 0).view(torch.uint8)
    keep.append(tm)
    desc.mask, desc.mask_stride = tm.data_ptr(), _strides(tm, dm_, dims)
  desc.n_factors = len(factors)
  for k, f in enumerate(factors):
    tf, df_ = _device_tensor(f, 'factor', device)
    keep.append(tf)
    desc.factor[k] = tf.data_ptr()
    desc.factor_dtype[k] = _dtype_code(tf)
    desc.factor_stride[k] = _strides(tf, df_, dims)

  kept = [d for d in dims if d not in reduce_set]
  kept_shape = [sizes[d] for d in kept]
  n_cells = int(np.prod(kept_shape, dtype=np.int64)) if kept else 1
  ws = np.empty(n_cells, np.float64)
  w = np.empty(n_cells, np.float64)
  ctx = _cabi.get_context(device)
  ctx.use_torch_stream()
  _cabi.check(ctx.lib.wbx_reduce_generic(
      ctx.handle, ctypes.byref(desc), ws.ctypes.data, w.ctypes.data,
      _cabi.SPACE_HOST))
  del keep
  coords = {}
  for src in [stat] + list(factors):
    for name, cv in src.coords.items():
      if name not in coords and name != 'mask' and set(cv.dims) <= set(kept):
        coords[name] = cv
  name = stat.name
  return (xl.DataArray(ws.reshape(kept_shape), kept, coords=coords, name=name),
          xl.DataArray(w.reshape(kept_shape), kept, coords=coords, name=name))


def _aligned_climatology_tensor(stat: LazyStatistic, device=None):
  """climatology.sel(dayofyear=..., hour=...) as a device gather (torch)."""
  torch = engine._torch()  # pylint: disable=protected-access
  ac = stat.climatology
  tc, cdims = _device_tensor(ac.climatology, 'field', device)
  front = list(ac.clim_time_dims)
  rest = [d for d in cdims if d not in front]
  tc = tc.permute(*[cdims.index(d) for d in front + rest])
  index = tuple(torch.as_tensor(ac.positions[d], device=tc.device)
                for d in front)
  return tc[index], tuple(ac.time_dims) + tuple(rest)
