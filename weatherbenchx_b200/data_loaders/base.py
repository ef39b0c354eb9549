"""Data loader interface of the chunk driver.

Same contract as /root/reference/weatherbenchX/data_loaders/base.py:25-175:
``load_chunk(init_times, lead_times, reference)`` returns a mapping of
variable -> DataArray for one time chunk, after the optional
``process_chunk_fn``, NaN-mask and values-as-coords steps.  Interpolation to
sparse targets is outside the gridded hot path and is refused.
"""

from __future__ import annotations

import abc
from typing import Callable, Collection, Hashable, Mapping, Optional, Union

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl


def add_nan_mask_to_data(data: Mapping[Hashable, xl.DataArray],
                         variable_subset: Collection[str] | None = None
                         ) -> dict:
  """Adds a boolean 'mask' coordinate (False at NaN) to every variable; the
  Aggregator uses it when ``masked=True`` (data_loaders/base.py:25-57)."""
  out = {}
  for var, da in data.items():
    da = xl.as_data_array(da)
    if variable_subset is None or var in variable_subset:
      if da.is_device:  # device-resident rows: the mask is made there too
        mask = xl.DataArray(~da.data.isnan(), da.dims)
      else:
        mask = xl.DataArray(~np.isnan(da.to_numpy()), da.dims)
      da = da.assign_coords(mask=mask)
    out[var] = da
  return out


class DataLoader(abc.ABC):
  """Returns chunks of data that broadcast against the other loader's."""

  def __init__(self, interpolation=None, compute: bool = True,
               add_nan_mask: bool = False,
               process_chunk_fn: Optional[Callable[[Mapping], Mapping]] = None,
               add_values_to_coords: bool = False):
    if interpolation is not None:
      raise NotImplementedError(
          'interpolation to sparse targets is outside the B200 gridded path')
    self._compute = compute
    self._add_nan_mask = add_nan_mask
    self._process_chunk_fn = process_chunk_fn
    self._add_values_to_coords = add_values_to_coords

  @abc.abstractmethod
  def _load_chunk_from_source(
      self, init_times: np.ndarray,
      lead_times: Optional[Union[np.ndarray, slice]] = None,
  ) -> Mapping[Hashable, xl.DataArray]:
    """Raw chunk for the given times."""

  def load_chunk(self, init_times: np.ndarray,
                 lead_times: Optional[Union[np.ndarray, slice]] = None,
                 reference: Optional[Mapping] = None) -> Mapping:
    del reference  # only used by interpolating loaders
    chunk = dict(self._load_chunk_from_source(init_times, lead_times))
    if self._process_chunk_fn is not None:
      chunk = dict(self._process_chunk_fn(chunk))
    chunk = {k: xl.as_data_array(v) for k, v in chunk.items()}
    if self._add_nan_mask:
      chunk = add_nan_mask_to_data(chunk)
    if self._add_values_to_coords:
      chunk = {k: v.assign_coords(values_as_coord=v) for k, v in chunk.items()}
    return chunk
