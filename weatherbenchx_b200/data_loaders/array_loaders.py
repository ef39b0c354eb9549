"""Loaders over in-memory (or memory-mapped) gridded arrays.

The counterpart of the reference's xarray/Zarr loaders
(/root/reference/weatherbenchX/data_loaders/xarray_loaders.py:160-263) for the
on-box driver: a dataset is a mapping variable -> DataArray (NumPy, np.memmap
or CUDA payload).  Selection semantics are the reference's:

* ``PredictionsFromArrays``: dims ``init_time`` and ``lead_time``;
  ``load_chunk`` selects the exact init times and the exact lead times (or the
  inclusive lead-time interval of a slice) -- xarray_loaders.py:191-206.
* ``TargetsFromArrays``: dim ``valid_time``; the chunk is gathered at
  ``valid_time = init_time + lead_time`` and gets dims (init_time, lead_time,
  ...) with a 2-d ``valid_time`` coordinate; without lead times the init times
  are taken as valid times -- xarray_loaders.py:242-263.

Contiguous selections are returned as views (no copy), so an init-time chunk of
a pinned or memory-mapped array goes to the GPU straight from its source.
"""

from __future__ import annotations

import threading
from typing import Hashable, Iterable, Mapping, Optional

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.data_loaders import base


def _positions(index: np.ndarray, labels: np.ndarray, dim: str) -> np.ndarray:
  """Integer positions of ``labels`` in the coordinate ``index`` (exact)."""
  if index.dtype.kind in 'mM':
    unit = 'datetime64[ns]' if index.dtype.kind == 'M' else 'timedelta64[ns]'
    index, labels = index.astype(unit), np.asarray(labels).astype(unit)
  labels = np.asarray(labels)
  order = np.argsort(index, kind='stable')
  pos = np.searchsorted(index, labels.ravel(), sorter=order)
  pos = np.clip(pos, 0, len(index) - 1)
  found = order[pos]
  if not np.array_equal(index[found], labels.ravel()):
    missing = labels.ravel()[index[found] != labels.ravel()]
    raise KeyError(f'not all values found in index {dim!r}: {missing[:3]}')
  return found.reshape(labels.shape)


def _take(da: xl.DataArray, dim: str, pos: np.ndarray) -> xl.DataArray:
  """da.isel({dim: pos}) for 1-d positions; a view when they are a range."""
  if len(pos) and np.array_equal(pos, np.arange(pos[0], pos[0] + len(pos))):
    return da.isel({dim: slice(int(pos[0]), int(pos[0]) + len(pos))})
  return da.isel({dim: pos})


def _window_view(payload, axis: int, pos: np.ndarray):
  """Zero-copy (init_time, lead_time) view of ``payload`` along ``axis`` when
  the valid-time positions form a regular lattice pos[i, j] = a + i*di + j*dj
  (regularly spaced init and lead times on a regularly spaced analysis): the
  rows of neighbouring init times overlap in memory, which a strided view
  expresses and the job tables of the kernels consume as is.  None otherwise.
  """
  n_i, n_j = pos.shape
  a = int(pos[0, 0])
  di = int(pos[1, 0]) - a if n_i > 1 else 0
  dj = int(pos[0, 1]) - a if n_j > 1 else 0
  lattice = a + di * np.arange(n_i)[:, None] + dj * np.arange(n_j)[None, :]
  if di < 0 or dj < 0 or not np.array_equal(lattice, pos):
    return None
  if xl._is_device(payload):  # pylint: disable=protected-access
    import torch  # pylint: disable=g-import-not-at-top
    stride = list(payload.stride())
    size = list(payload.shape)
    offset = payload.storage_offset() + a * stride[axis]
    return torch.as_strided(
        payload, size[:axis] + [n_i, n_j] + size[axis + 1:],
        stride[:axis] + [di * stride[axis], dj * stride[axis]] +
        stride[axis + 1:], offset)
  strides = list(payload.strides)
  shape = list(payload.shape)
  start = [slice(None)] * payload.ndim
  start[axis] = slice(a, None)
  return np.lib.stride_tricks.as_strided(
      payload[tuple(start)],
      shape[:axis] + [n_i, n_j] + shape[axis + 1:],
      strides[:axis] + [di * strides[axis], dj * strides[axis]] +
      strides[axis + 1:], writeable=False)


class _ArrayLoader(base.DataLoader):

  def __init__(self, ds: Mapping[Hashable, xl.DataArray],
               variables: Optional[Iterable[str]] = None,
               sel_kwargs: Optional[Mapping] = None,
               rename_variables: Optional[Mapping[str, str]] = None, **kwargs):
    super().__init__(**kwargs)
    ds = {k: xl.as_data_array(v) for k, v in dict(ds).items()}
    if rename_variables:
      ds = {rename_variables.get(k, k): v.rename(rename_variables.get(k, k))
            for k, v in ds.items()}
    if variables is not None:
      ds = {k: ds[k] for k in variables}
    if sel_kwargs:
      ds = {k: v.sel({d: s for d, s in sel_kwargs.items() if d in v.dims})
            for k, v in ds.items()}
    self._ds = ds


class PredictionsFromArrays(_ArrayLoader):
  """Forecasts with dims (init_time, lead_time, ...)."""

  def _load_chunk_from_source(self, init_times, lead_times=None):
    out = {}
    for var, da in self._ds.items():
      chunk = _take(da, 'init_time', _positions(
          da.coords['init_time'].to_numpy(), init_times, 'init_time'))
      if isinstance(lead_times, slice):
        lead = da.coords['lead_time'].to_numpy()
        keep = np.nonzero((lead >= lead_times.start) &
                          (lead <= lead_times.stop))[0]
        chunk = _take(chunk, 'lead_time', keep)
      elif lead_times is not None:
        chunk = _take(chunk, 'lead_time', _positions(
            da.coords['lead_time'].to_numpy(), lead_times, 'lead_time'))
      out[var] = chunk
    return out


class _DeviceRows:
  """Rows of one analysis array along ``valid_time`` kept on the GPU.

  Consecutive (init, lead) chunks of an evaluation share most of their target
  rows (valid_time = init_time + lead_time: with 12-hourly init times and
  6-hourly lead times 10 of the 12 rows of an (init=1, lead=12) chunk were
  already needed by the chunk before).  The reference re-reads them from its
  Zarr store for every chunk (xarray_loaders.py:242-263); a B200 has 180 GB of
  HBM, so each row is uploaded ONCE, the first time a chunk needs it, and
  stays.  The device array is allocated in full (address space only; pages
  are touched row by row) and rows are filled lazily.
  """

  def __init__(self, da: xl.DataArray, device):
    import torch  # pylint: disable=g-import-not-at-top
    self._host = da.data
    self._lock = threading.Lock()
    self._present = np.zeros(da.shape[0], bool)
    self.device = torch.device('cuda', device)
    self.rows = torch.empty(tuple(da.shape), dtype=torch.float32,
                            device=self.device)
    self._stream = torch.cuda.Stream(self.device)
    self._last_event = None
    self.uploaded_bytes = 0

  def ensure(self, positions: np.ndarray, wait: bool = True):
    """Uploads the rows at ``positions`` that are not on the device yet.

    ``wait=True`` returns after they have arrived.  ``wait=False`` only issues
    the copies and returns the event that follows the last upload of this
    cache (all uploads share one stream, so it covers every row a chunk can
    need); whoever consumes the chunk waits for it.  The loader thread of the
    chunk driver must not wait itself: its small copy queues behind the
    100 MB forecast copies of the evaluation lanes on the same PCIe link, and
    a loader that blocks for each chunk starves the lanes."""
    import torch  # pylint: disable=g-import-not-at-top
    with self._lock:
      need = np.unique(positions)
      need = need[~self._present[need]]
      if not len(need):
        return self._last_event
      runs = np.split(need, np.nonzero(np.diff(need) != 1)[0] + 1)
      with torch.cuda.stream(self._stream):
        for run in runs:
          lo, hi = int(run[0]), int(run[-1]) + 1
          src = self._host[lo:hi]
          if src.dtype != np.float32:
            src = src.astype(np.float32)
          self.rows[lo:hi].copy_(torch.from_numpy(np.ascontiguousarray(src)),
                                 non_blocking=True)
          self.uploaded_bytes += (hi - lo) * self.rows[0].numel() * 4
        event = torch.cuda.Event()
        event.record(self._stream)
      self._last_event = event
      self._present[need] = True
    if wait:
      event.synchronize()
    return event


class TargetsFromArrays(_ArrayLoader):
  """Analyses / observations on a grid with dim valid_time.

  ``device_cache=True`` keeps every target row a chunk has needed on the GPU
  (see ``_DeviceRows``): chunks then come back as device arrays -- strided
  views of the resident rows -- and only the predictions of a chunk cross
  PCIe.  ``device_cache_bytes`` bounds the HBM one loader may claim (default:
  half of what is free at first use); variables that do not fit, whose
  ``valid_time`` is not the leading dim, or that already live on the device
  are served from where they are.
  """

  def __init__(self, ds, *args, device_cache: bool = False,
               device_cache_bytes: Optional[int] = None,
               device: Optional[int] = None, **kwargs):
    super().__init__(ds, *args, **kwargs)
    self._device_cache = bool(device_cache)
    self._device_cache_bytes = device_cache_bytes
    self._device = device
    self._resident: dict = {}
    self._claimed = 0
    self._resident_lock = threading.Lock()
    # chunk driver: do not wait for the uploads inside load_chunk, hand the
    # event to the consumer (take_ready_events)
    self.async_uploads = False
    self._chunk_events: list = []

  def take_ready_events(self) -> list:
    """Events the arrays of the last load_chunk call of this thread's loader
    have to wait for before they are read (async_uploads only)."""
    events, self._chunk_events = self._chunk_events, []
    return events

  def __getstate__(self):
    state = dict(self.__dict__)
    state['_resident'] = {}      # device memory is per process
    state['_claimed'] = 0
    state['_chunk_events'] = []
    state.pop('_resident_lock', None)
    return state

  def __setstate__(self, state):
    self.__dict__.update(state)
    self._resident_lock = threading.Lock()

  @property
  def uploaded_bytes(self) -> int:
    """Bytes sent to the device by the cache so far (each row once)."""
    return sum(r.uploaded_bytes for r in self._resident.values()
               if r is not None)

  def _device_rows(self, var, da):
    """The resident rows of ``var``, or None when it is not cached."""
    if not self._device_cache or da.is_device or not da.dims or (
        da.dims[0] != 'valid_time'):
      return None
    with self._resident_lock:
      if var in self._resident:
        return self._resident[var]
      import torch  # pylint: disable=g-import-not-at-top
      if not torch.cuda.is_available():
        raise RuntimeError('device_cache=True needs a CUDA device')
      device = (torch.cuda.current_device() if self._device is None
                else self._device)
      if self._device_cache_bytes is None:
        free, _ = torch.cuda.mem_get_info(device)
        self._device_cache_bytes = free // 2
      nbytes = int(np.prod(da.shape, dtype=np.int64)) * 4
      rows = None
      if self._claimed + nbytes <= self._device_cache_bytes:
        rows = _DeviceRows(da, device)
        self._claimed += nbytes
      self._resident[var] = rows
      return rows

  def _load_chunk_from_source(self, init_times, lead_times=None):
    if isinstance(lead_times, slice):
      raise ValueError('Lead time slice not supported for target data loaders.')
    init_times = np.asarray(init_times).astype('datetime64[ns]')
    out = {}
    for var, da in self._ds.items():
      index = da.coords['valid_time'].to_numpy()
      if lead_times is None:
        chunk = _take(da, 'valid_time',
                      _positions(index, init_times, 'valid_time'))
        out[var] = chunk.rename({'valid_time': 'init_time'})
        continue
      lead = np.asarray(lead_times).astype('timedelta64[ns]')
      valid = init_times[:, None] + lead[None, :]
      pos = _positions(index, valid, 'valid_time')
      axis = da.dims.index('valid_time')
      source = da.data
      resident = self._device_rows(var, da)
      if resident is not None:
        event = resident.ensure(pos.ravel(), wait=not self.async_uploads)
        if self.async_uploads and event is not None:
          self._chunk_events.append(event)
        source = resident.rows
      payload = _window_view(source, axis, pos)
      if payload is None:
        if xl._is_device(source):  # pylint: disable=protected-access
          import torch  # pylint: disable=g-import-not-at-top
          flat = torch.as_tensor(pos.ravel(), device=source.device)
          payload = source.index_select(axis, flat)
        else:
          payload = np.take(source, pos.ravel(), axis=axis)
        shape = list(payload.shape)
        shape[axis:axis + 1] = list(pos.shape)
        payload = payload.reshape(shape)
      dims = da.dims[:axis] + ('init_time', 'lead_time') + da.dims[axis + 1:]
      coords = {k: v for k, v in da.coords.items()
                if 'valid_time' not in v.dims}
      coords['init_time'] = xl.DataArray(init_times, ('init_time',))
      coords['lead_time'] = xl.DataArray(lead, ('lead_time',))
      coords['valid_time'] = xl.DataArray(valid, ('init_time', 'lead_time'))
      out[var] = xl.DataArray(payload, dims, coords=coords, name=da.name,
                              attrs=da.attrs)
    return out
