"""Weighting classes (host side; the weight vectors are tiny).

Mirrors /root/reference/weatherbenchX/weighting.py: Weighting :24-42,
latitude_cell_bounds :62-79, cell_area_from_latitude :82-88, GridAreaWeighting
:91-130.  The weights are computed on the host in float64 and handed to the
kernel as the per-row vector w_y (or w_x for longitude-major layouts).
StationDensityWeighting (sparse observations) is outside the hot path.
"""

from __future__ import annotations

import abc
import dataclasses
import threading
import weakref

import numpy as np

from weatherbenchx_b200 import xarray_lite as xl


class Weighting(abc.ABC):
  """Returns weights that broadcast against a statistic's dims."""

  @abc.abstractmethod
  def weights(self, statistic: xl.DataArray) -> xl.DataArray:
    """Weights for ``statistic`` (only its dims / coordinates are used)."""


def latitude_cell_bounds(x: np.ndarray) -> np.ndarray:
  """Cell edges (radians) for increasing cell centres: midpoints, with the two
  end cells extended by half a spacing and clipped at the poles."""
  x = np.asarray(x)
  spacing = np.diff(x)
  assert np.all(spacing > 0), 'Points must be increasing.'
  south = max(x[0] - spacing[0] / 2, -np.pi / 2)
  north = min(x[-1] + spacing[-1] / 2, np.pi / 2)
  edges = np.empty(len(x) + 1, dtype=x.dtype)
  edges[0], edges[-1] = south, north
  edges[1:-1] = (x[:-1] + x[1:]) / 2
  return edges


def cell_area_from_latitude(points: np.ndarray) -> np.ndarray:
  """Integral of cos(lat) over every latitude cell."""
  edges = latitude_cell_bounds(points)
  return np.sin(edges[1:]) - np.sin(edges[:-1])


_WEIGHT_CACHE: dict = {}
_WEIGHT_LOCK = threading.Lock()


@dataclasses.dataclass
class GridAreaWeighting(Weighting):
  """Weights proportional to the area of rectangular lat/lon grid boxes."""

  latitude_name: str = 'latitude'
  return_normalized: bool = True

  def weights(self, statistic: xl.DataArray) -> xl.DataArray:
    if self.latitude_name not in statistic.dims:
      return xl.DataArray(1)
    lat = statistic.coords[self.latitude_name].to_numpy()
    # memoised on the identity of the latitude payload: every statistic of a
    # chunk asks for the same vector, and a stable result object lets the
    # engine recognise a repeated aggregation plan.
    key = (id(lat), self.latitude_name, self.return_normalized)
    hit = _WEIGHT_CACHE.get(key)
    if hit is not None and hit[0]() is lat:
      return hit[1]
    out = self._weights(lat)
    try:
      ref = weakref.ref(lat)
    except TypeError:
      return out
    with _WEIGHT_LOCK:
      _WEIGHT_CACHE[key] = (ref, out)
      while len(_WEIGHT_CACHE) > 16:
        _WEIGHT_CACHE.pop(next(iter(_WEIGHT_CACHE)))
    return out

  def _weights(self, lat: np.ndarray) -> xl.DataArray:
    steps = np.diff(lat)
    assert np.all(steps > 0) or np.all(steps < 0), (
        f'Points must be strictly monotonic: {lat}')
    descending = lat[0] > lat[1]
    ordered = lat[::-1] if descending else lat
    area = cell_area_from_latitude(np.deg2rad(ordered))
    if descending:
      area = area[::-1]
    if self.return_normalized:
      area = area / np.mean(area)
    return xl.DataArray(area, (self.latitude_name,),
                        coords={self.latitude_name: lat},
                        name=self.latitude_name)
