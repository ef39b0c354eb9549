"""Binning classes whose masks enter the aggregation as extra operands.

Mirrors the part of /root/reference/weatherbenchX/binning.py used by the
evaluation scripts on the gridded path: Binning :22-49, the lat/lon rectangle
helpers :52-89, Regions :147-201, LandSea :92-144, LatitudeBins :204-243 and
LongitudeBins :246-298.  Masks are small boolean
host arrays ([bins, latitude, longitude]); the kernels read them as uint8.
"""

from __future__ import annotations

import abc
from typing import Hashable, Mapping, Optional, Tuple

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl


class Binning(abc.ABC):
  """Creates a boolean mask with a new ``bin_dim_name`` dimension."""

  def __init__(self, bin_dim_name: str):
    self.bin_dim_name = bin_dim_name

  @abc.abstractmethod
  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    """Boolean mask that broadcasts against ``statistic``."""


def _lat_mask(lat: np.ndarray, lims) -> np.ndarray:
  if lims[0] >= lims[1]:
    raise ValueError(
        f'`lat_lims[0]` must be smaller than `lat_lims[1]`, got {lims}`')
  return (lat >= lims[0]) & (lat <= lims[1])


def _lon_mask(lon: np.ndarray, lims) -> np.ndarray:
  lon = np.mod(lon, 360)
  west, east = np.mod(lims[0], 360), np.mod(lims[1], 360)
  if east > west:
    return (lon >= west) & (lon <= east)
  return (lon <= east) | (lon >= west)  # region wraps across 0 degrees


class Regions(Binning):
  """Rectangular lat/lon regions {name: ((lat_lo, lat_hi), (lon_lo, lon_hi))}.

  With ``land_sea_mask`` (True = land) every region gets a second
  '<name>_land' bin restricted to land points.
  """

  def __init__(self, regions: Mapping[Hashable, Tuple[Tuple[float, float],
                                                      Tuple[float, float]]],
               bin_dim_name: str = 'region',
               land_sea_mask: Optional[xl.DataArray] = None):
    super().__init__(bin_dim_name)
    self._regions = regions
    self._land_sea_mask = land_sea_mask

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    lat = statistic.coords['latitude'].to_numpy()
    lon = statistic.coords['longitude'].to_numpy()
    # The mask only depends on the grid: reuse it while the coordinate arrays
    # are the same objects (every chunk of an evaluation shares them).
    cached = getattr(self, '_cache', None)
    if cached is not None and cached[0] is lat and cached[1] is lon:
      return cached[2]
    out = self._create_bin_mask(lat, lon)
    self._cache = (lat, lon, out)
    return out

  def __getstate__(self):
    state = dict(self.__dict__)
    state.pop('_cache', None)
    return state

  def _create_bin_mask(self, lat, lon) -> xl.DataArray:
    names = list(self._regions)
    masks = np.stack([
        _lat_mask(lat, lat_lims)[:, None] & _lon_mask(lon, lon_lims)[None, :]
        for lat_lims, lon_lims in self._regions.values()])
    if self._land_sea_mask is not None:
      lsm = xl.as_data_array(self._land_sea_mask)
      assert (np.array_equal(np.sort(lat),
                             np.sort(lsm.coords['latitude'].to_numpy()))
              and np.array_equal(lon, lsm.coords['longitude'].to_numpy())), (
                  'Land/sea mask coordinates do not match.')
      land = lsm.transpose('latitude', 'longitude').to_numpy().astype(bool)
      if not np.array_equal(lat, lsm.coords['latitude'].to_numpy()):
        order = np.argsort(lsm.coords['latitude'].to_numpy())
        land = land[order][np.argsort(np.argsort(lat))]
      masks = np.concatenate([masks, masks & land[None]])
      names = names + [f'{n}_land' for n in names]
    return xl.DataArray(
        masks, (self.bin_dim_name, 'latitude', 'longitude'),
        coords={self.bin_dim_name: np.array(names), 'latitude': lat,
                'longitude': lon})


class _BandBins(Binning):
  """Bins that depend on one grid coordinate; the mask is [bins, coordinate]
  (it broadcasts against the statistic like the reference's full-shape one)."""

  _coord = ''

  def _band_mask(self, values: np.ndarray, start: float) -> np.ndarray:
    raise NotImplementedError

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    values = statistic.coords[self._coord].to_numpy()
    cached = getattr(self, '_cache', None)
    if cached is not None and cached[0] is values:
      return cached[1]
    starts = self._starts
    masks = np.stack([self._band_mask(values, s) for s in starts])
    out = xl.DataArray(
        masks, (self.bin_dim_name, self._coord),
        coords={self.bin_dim_name: self._labels(starts), self._coord: values})
    self._cache = (values, out)
    return out

  def _labels(self, starts: np.ndarray) -> np.ndarray:
    return np.asarray(starts)

  def __getstate__(self):
    state = dict(self.__dict__)
    state.pop('_cache', None)
    return state


class LatitudeBins(_BandBins):
  """Latitude bands of ``degrees`` width (binning.py:204-243); both band
  edges are inclusive, as in the reference."""

  _coord = 'latitude'

  def __init__(self, degrees: float, lat_range: Tuple[int, int] = (-90, 90),
               bin_dim_name: str = 'latitude_bins'):
    super().__init__(bin_dim_name)
    self._degrees = degrees
    self._starts = np.arange(lat_range[0], lat_range[1] + degrees,
                             degrees)[:-1]

  def _band_mask(self, values, start):
    return _lat_mask(values, (start, start + self._degrees))


class LongitudeBins(_BandBins):
  """Longitude bands of ``degrees`` width (binning.py:246-298), wrapping at
  360 degrees; labelled by the band start modulo 360."""

  _coord = 'longitude'

  def __init__(self, degrees: float, lon_range: Tuple[int, int] = (0, 360),
               bin_dim_name: str = 'longitude_bins'):
    super().__init__(bin_dim_name)
    self._degrees = degrees
    lon_end = lon_range[1]
    if lon_range[0] >= lon_range[1]:
      lon_end += 360
    self._starts = np.arange(lon_range[0], lon_end + degrees, degrees)[:-1]

  def _band_mask(self, values, start):
    return _lon_mask(values, (start, start + self._degrees))

  def _labels(self, starts):
    return np.mod(starts, 360)


class LandSea(Binning):
  """['land', 'sea'(, 'global')] bins from a land fraction field."""

  def __init__(self, land_sea_fraction: xl.DataArray,
               land_sea_threshold: float = 0.5,
               bin_dim_name: str = 'land_sea',
               include_global_mask: bool = False):
    super().__init__(bin_dim_name)
    frac = xl.as_data_array(land_sea_fraction)
    self._land = frac._replace(data=frac.to_numpy() >= land_sea_threshold)  # pylint: disable=protected-access
    self._include_global_mask = include_global_mask

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    land = self._land.to_numpy()
    masks, labels = [land, ~land], ['land', 'sea']
    if self._include_global_mask:
      masks.append(np.ones_like(land))
      labels.append('global')
    coords = {k: v for k, v in self._land.coords.items()
              if k in self._land.dims}
    coords[self.bin_dim_name] = np.array(labels)
    return xl.DataArray(np.stack(masks), (self.bin_dim_name,) + self._land.dims,
                        coords=coords)
