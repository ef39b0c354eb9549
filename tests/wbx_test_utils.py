"""Synthetic fixtures shaped like the reference's test_utils.py:27-90.

Same grid (10 degrees -> 19 x 36), levels (500, 700, 850), daily times, zeros or
``default_rng(seed).random`` float64 values, optional trailing ``realization``
dim, and the same dim order as the reference produces for 3-d variables
(prediction_timedelta, time, latitude, longitude, level[, realization]).  The
reference derives the 2-d dim order from a ``set`` (test_utils.py:67), i.e. it
is arbitrary; here it is (time, latitude, longitude[, realization]).
"""

from __future__ import annotations

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl


def _times(start: str, stop: str, step='1D'):
  step = np.timedelta64(int(step[:-1]), step[-1])
  return np.arange(np.datetime64(start, 'ns'), np.datetime64(stop, 'ns'), step)


def mock_target_data(*, variables_3d=('geopotential',),
                     variables_2d=('2m_temperature',),
                     levels=(500, 700, 850),
                     spatial_resolution_in_degrees=10.0,
                     time_start='2020-01-01', time_stop='2021-01-01',
                     dtype=np.float32, ensemble_size=None, random=False,
                     seed=None, time_name='time') -> dict:
  rng = np.random.default_rng(seed)

  def values(shape):
    return rng.random(size=shape) if random else np.zeros(shape, dtype=dtype)

  nlat = round(180 / spatial_resolution_in_degrees) + 1
  nlon = round(360 / spatial_resolution_in_degrees)
  coords = {
      time_name: _times(time_start, time_stop),
      'latitude': np.linspace(-90, 90, nlat),
      'longitude': np.linspace(0, 360, nlon, endpoint=False),
      'level': np.array(levels),
  }
  if ensemble_size is not None:
    coords['realization'] = np.arange(ensemble_size)
  out = {}
  dims3 = tuple(coords)
  for name in variables_3d:
    out[name] = xl.DataArray(
        values(tuple(len(coords[d]) for d in dims3)), dims3,
        coords={d: coords[d] for d in dims3}, name=name)
  dims2 = tuple(d for d in coords if d != 'level')
  for name in variables_2d:
    out[name] = xl.DataArray(
        values(tuple(len(coords[d]) for d in dims2)), dims2,
        coords={d: coords[d] for d in dims2}, name=name)
  return out


def mock_prediction_data(*, lead_start=0, lead_stop=10, lead_name=
                         'prediction_timedelta', **kwargs) -> dict:
  """Adds a leading lead-time dim of whole days [lead_start, lead_stop]."""
  lead = (np.arange(lead_start, lead_stop + 1) *
          np.timedelta64(1, 'D')).astype('timedelta64[ns]')
  data = mock_target_data(**kwargs)
  return {k: v.expand_dims({lead_name: lead}) for k, v in data.items()}


def rename_all(data: dict, **mapping) -> dict:
  return {k: v.rename({a: b for a, b in mapping.items() if a in v.dims})
          for k, v in data.items()}


def add_nan_mask_to_data(data: dict) -> dict:
  """data_loaders/base.py:25-57 -- boolean 'mask' coordinate, False at NaN."""
  out = {}
  for k, v in data.items():
    mask = xl.DataArray(~np.isnan(v.to_numpy()), v.dims)
    out[k] = v.assign_coords(mask=mask)
  return out


def to_f32(data: dict) -> dict:
  return {k: v.astype(np.float32) for k, v in data.items()}
