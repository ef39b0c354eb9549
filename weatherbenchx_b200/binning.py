"""Binning classes whose masks enter the aggregation as extra operands.

Mirrors /root/reference/weatherbenchX/binning.py: Binning :22-49, the lat/lon
rectangle helpers :52-89, LandSea :92-144, Regions :147-201, LatitudeBins
:204-243, LongitudeBins :246-298, vectorized_coord_mask :301-332, ByExactCoord
:335-355, the time-unit bins :358-567 (ByTimeUnit, ByTimeUnitSets,
ByTimeUnitFromSeconds), ByCoordBins :570-637 and BySets :640-704.  Masks are
small boolean host arrays.  Masks over the grid ([bins, latitude, longitude])
are folded into the class map of the fused kernel; masks over outer dims
([bins, init_time], [bins, init_time, lead_time], ...) are folded into the
job -> cell table of the launch (engine.OuterClasses), so neither kind costs an
extra pass over the fields.
"""

from __future__ import annotations

import abc
from typing import Any, Hashable, Mapping, Optional, Sequence, Tuple

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


# ---------------------------------------------------------------------------
# Bins defined by the values of a coordinate (binning.py:301-704)
# ---------------------------------------------------------------------------


def vectorized_coord_mask(coord: xl.DataArray, coord_name: str,
                          bin_dim_name: str,
                          add_global_bin: bool = False) -> xl.DataArray:
  """One bin per unique value of ``coord`` (binning.py:301-332).

  The mask has dims (bin_dim_name, *coord.dims).  With ``add_global_bin`` an
  all-True 'global' bin comes FIRST; if the labels are not strings all labels
  are then cast to str, as in the reference.
  """
  coord = xl.as_data_array(coord)
  values = coord.to_numpy()
  unique = np.unique(values)
  masks = np.equal(values, unique.reshape((-1,) + (1,) * values.ndim))
  labels = unique
  if add_global_bin:
    masks = np.concatenate([np.ones((1,) + values.shape, bool), masks])
    global_label = np.array(['global'])
    if global_label.dtype != unique.dtype:
      labels = unique.astype('str')
    labels = np.concatenate([global_label, labels])
  coords = {d: coord.coords[d] for d in coord.dims if d in coord.coords
            and d != bin_dim_name}
  coords[bin_dim_name] = labels
  return xl.DataArray(masks, (bin_dim_name,) + tuple(coord.dims),
                      coords=coords)


class ByExactCoord(Binning):
  """A bin for each unique value of a non-dimension coordinate, e.g. the lead
  time of sparse forecasts (binning.py:335-355).  The bin dim is named like
  the coordinate."""

  def __init__(self, coord: str, add_global_bin: bool = False):
    super().__init__(coord)
    self.coord = coord
    self.add_global_bin = add_global_bin

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    assert self.coord not in statistic.dims, (
        'For dimensions, specify reduce_dims in aggregation.')
    return vectorized_coord_mask(statistic.coords[self.coord], self.coord,
                                 self.coord, self.add_global_bin)


def _extract_time_unit(time_coord: xl.DataArray, unit: str) -> xl.DataArray:
  """binning.py:358-391: ``.dt.<unit>`` of datetimes; whole units of the
  total seconds of timedeltas (second, minute, hour, day, week, year)."""
  time_coord = xl.as_data_array(time_coord)
  if time_coord.dtype.kind == 'm':
    seconds = time_coord.dt.total_seconds()
    per_unit = {'second': None, 'minute': 60, 'hour': 60 * 60,
                'day': 60 * 60 * 24, 'week': 60 * 60 * 24 * 7,
                'year': 60 * 60 * 24 * 365}
    if unit not in per_unit:
      raise ValueError(f'Unsupported unit for timedelta: {unit}')
    if per_unit[unit] is None:
      return seconds
    return seconds._replace(data=seconds.to_numpy() // per_unit[unit])  # pylint: disable=protected-access
  assert time_coord.dtype.kind == 'M', time_coord.dtype
  return getattr(time_coord.dt, unit)


class ByTimeUnit(Binning):
  """Bins by a time unit of a datetime64 / timedelta64 coordinate, e.g. the
  hour of ``init_time`` or the month of ``valid_time`` (binning.py:394-442).
  The bin dim is ``f'{time_dim}_{unit}'``."""

  def __init__(self, unit: str, time_dim: str, add_global_bin: bool = False):
    super().__init__(f'{time_dim}_{unit}')
    self.unit = unit
    self.time_dim = time_dim
    self.add_global_bin = add_global_bin

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    coord = _extract_time_unit(statistic.coords[self.time_dim], self.unit)
    return vectorized_coord_mask(coord, self.time_dim,
                                 f'{self.time_dim}_{self.unit}',
                                 self.add_global_bin)


def _as_value_array(values) -> np.ndarray:
  if isinstance(values, Sequence) and not isinstance(values, str):
    return np.array(list(values))
  return np.array([values])


def _stack_named_masks(masks, labels, template: xl.DataArray,
                       bin_dim_name: str) -> xl.DataArray:
  coords = {d: template.coords[d] for d in template.dims
            if d in template.coords and d != bin_dim_name}
  coords[bin_dim_name] = np.array(labels)
  shape = (len(masks),) + tuple(template.shape)
  data = np.stack(masks) if masks else np.zeros(shape, bool)
  return xl.DataArray(data, (bin_dim_name,) + tuple(template.dims),
                      coords=coords)


class ByTimeUnitSets(Binning):
  """Bins by named sets of time unit values, e.g. {'00/12': [0, 12], '06/18':
  [6, 18]} of the hour of ``init_time`` (binning.py:445-515).  Sets may
  overlap; 'global' (optional) comes last."""

  def __init__(self, sets: Mapping[str, Any], unit: str, dim: str,
               bin_dim_name: Optional[str] = None,
               add_global_bin: bool = False):
    super().__init__(bin_dim_name if bin_dim_name is not None
                     else f'{dim}_{unit}_sets')
    self.sets = sets
    self.unit = unit
    self.dim = dim
    self.add_global_bin = add_global_bin

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    values = _extract_time_unit(statistic.coords[self.dim], self.unit)
    masks = [np.isin(values.to_numpy(), _as_value_array(s))
             for s in self.sets.values()]
    labels = list(self.sets)
    if self.add_global_bin:
      masks.append(np.ones(values.shape, bool))
      labels.append('global')
    return _stack_named_masks(masks, labels, values, self.bin_dim_name)


class ByTimeUnitFromSeconds(Binning):
  """As ByTimeUnit for a coordinate that holds plain seconds
  (binning.py:518-567): bins default to 0..59 (second, minute) / 0..23 (hour)
  and a value outside every bin falls in none."""

  def __init__(self, unit: str, time_dim: str,
               bins: Optional[Sequence[int]] = None):
    super().__init__(f'{time_dim}_{unit}')
    self.unit = unit
    self.time_dim = time_dim
    self.bins = bins

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    coord = xl.as_data_array(statistic.coords[self.time_dim])
    values = coord.to_numpy()
    bins = self.bins
    if self.unit == 'second':
      bins = bins if bins is not None else np.arange(0, 60)
    elif self.unit == 'minute':
      values = values // 60
      bins = bins if bins is not None else np.arange(0, 60)
    elif self.unit == 'hour':
      values = values // (60 * 60)
      bins = bins if bins is not None else np.arange(0, 24)
    else:
      raise ValueError(f'Unsupported unit: {self.unit}')
    bins = np.asarray(bins)
    # the reference broadcasts ``coord == bins``: coordinate dims first
    masks = np.equal(values[..., None], bins)
    coords = {d: coord.coords[d] for d in coord.dims if d in coord.coords}
    coords[self.bin_dim_name] = bins
    return xl.DataArray(masks, tuple(coord.dims) + (self.bin_dim_name,),
                        coords=coords)


class ByCoordBins(Binning):
  """Half-open bins [start, stop) over a non-dimension coordinate
  (binning.py:570-637).  The bin dim is named like the coordinate and labelled
  by the left edges (as strings when a 'global' bin is added)."""

  def __init__(self, dim_name: str, bin_edges: np.ndarray,
               add_global_bin: bool = False):
    super().__init__(dim_name)
    self.dim_name = dim_name
    self.bin_edges = bin_edges
    self.add_global_bin = add_global_bin

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    coord = xl.as_data_array(statistic.coords[self.dim_name])
    if self.dim_name in coord.dims:
      raise ValueError(
          f'{self.dim_name!r} is a dimension of the statistic; ByCoordBins '
          'bins over a non-dimension coordinate')
    values = coord.to_numpy()
    masks, labels = [], []
    for start, stop in zip(self.bin_edges[:-1], self.bin_edges[1:]):
      masks.append(np.logical_and(values >= start, values < stop))
      labels.append(str(start) if self.add_global_bin else start)
    if self.add_global_bin:
      masks.append(np.ones(values.shape, bool))
      labels.append('global')
    if not labels:
      labels = np.array([], dtype=coord.dtype)
    return _stack_named_masks(masks, labels, coord, self.bin_dim_name)


class BySets(Binning):
  """Bins by named sets of values of a coordinate, e.g. sets of station names
  or of levels (binning.py:640-704)."""

  def __init__(self, sets: Mapping[str, Any], coord_name: str,
               bin_dim_name: Optional[str] = None,
               add_set_complements: bool = False,
               add_global_bin: bool = False):
    if bin_dim_name is None or bin_dim_name == coord_name:
      raise ValueError(
          'bin_dim_name must be defined and be different from coord_name.')
    super().__init__(bin_dim_name)
    self.sets = sets
    self.coord_name = coord_name
    self.add_set_complements = add_set_complements
    self.add_global_bin = add_global_bin

  def create_bin_mask(self, statistic: xl.DataArray) -> xl.DataArray:
    coord = xl.as_data_array(statistic.coords[self.coord_name])
    values = coord.to_numpy()
    masks, labels = [], []
    for name, s in self.sets.items():
      mask = np.isin(values, _as_value_array(s))
      masks.append(mask)
      labels.append(name)
      if self.add_set_complements:
        masks.append(~mask)
        labels.append(f'not_in_{name}')
    if self.add_global_bin:
      masks.append(np.ones(values.shape, bool))
      labels.append('global')
    return _stack_named_masks(masks, labels, coord, self.bin_dim_name)
