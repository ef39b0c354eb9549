"""Zonal energy spectrum (north_star item "EnergySpectrum").

PARITY UNPINNED: /root/reference/weatherbenchX contains no energy-spectrum
implementation, call site or test (SURVEY.md finding 2, section 8 row a16).  The
definition implemented here is WeatherBench 2's
``derived_variables.ZonalEnergySpectrum`` (not vendored, restated from its
published formula) and is validated against ``numpy.fft`` only:

    F = rfft(f, axis=longitude, norm='forward')
    S[0] = C |F_0|^2,   S[k > 0] = 2 C |F_k|^2,   C(lat) = 2 pi R cos(lat)

It is exposed through the WeatherBench-X plug-in surface as a
``PerVariableStatistic`` whose values carry a new ``zonal_wavenumber`` dim in
place of ``longitude``; averaging over time / latitude bands is then an ordinary
``Aggregator`` reduction.  The FFT is a shared-memory mixed-radix kernel
(csrc/spectrum.cu); nothing is computed on the host.
"""

from __future__ import annotations

import ctypes

import numpy as np

from weatherbenchx_b200 import _cabi
from weatherbenchx_b200 import engine
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base

EARTH_RADIUS_M = 1000.0 * (6357.0 + 6378.0) / 2.0


_COORD_CACHE: dict = {}


def _spectral_coords(lat: np.ndarray, lon: np.ndarray, nx: int,
                     latitude_name: str) -> dict:
  """zonal_wavenumber / frequency / wavelength coordinates of a spectrum;
  they depend on the grid only and are built once per grid (the wavelength
  table is [latitude, wavenumber])."""
  key = (lat.tobytes(), float(lon[0]), float(lon[1]) if nx > 1 else 0.0, nx,
         latitude_name)
  hit = _COORD_CACHE.get(key)
  if hit is None:
    scale = 2 * np.pi * EARTH_RADIUS_M * np.cos(np.deg2rad(lat))
    k = np.arange(nx // 2 + 1)
    spacing_deg = float(lon[1] - lon[0]) if nx > 1 else 360.0
    with np.errstate(divide='ignore'):
      freq = k / (nx * spacing_deg)             # cycles per degree longitude
      hit = {
          'zonal_wavenumber': xl.DataArray(k, ('zonal_wavenumber',)),
          'frequency': xl.DataArray(freq, ('zonal_wavenumber',)),
          'wavelength': xl.DataArray(
              scale[:, None] / np.where(k > 0, k, np.nan)[None, :],
              (latitude_name, 'zonal_wavenumber'))}
    if len(_COORD_CACHE) > 8:
      _COORD_CACHE.clear()
    _COORD_CACHE[key] = hit
  return hit


def zonal_energy_spectrum(field: xl.DataArray, latitude_name: str = 'latitude',
                          longitude_name: str = 'longitude',
                          device: int | None = None) -> xl.DataArray:
  """Spectrum of ``field`` along longitude; returns a device DataArray with
  ``longitude`` replaced by ``zonal_wavenumber`` (0 .. N/2) and ``wavelength`` /
  ``frequency`` coordinates as WeatherBench 2 provides them."""
  torch = engine._torch()  # pylint: disable=protected-access
  field = xl.as_data_array(field)
  if longitude_name not in field.dims or latitude_name not in field.dims:
    raise ValueError(f'{latitude_name!r} and {longitude_name!r} dims required')
  outer = [d for d in field.dims if d not in (latitude_name, longitude_name)]
  order = tuple(outer) + (latitude_name, longitude_name)
  canon = field if field.dims == order else field.transpose(*order)
  canon = engine.to_device(engine._normalise(canon, 'field'), device)  # pylint: disable=protected-access
  t = canon.data.contiguous()
  ny, nx = t.shape[-2], t.shape[-1]
  n_jobs = int(np.prod(t.shape[:-2], dtype=np.int64)) if outer else 1
  lat = canon.coords[latitude_name].to_numpy().astype(np.float64)
  scale = np.ascontiguousarray(
      2 * np.pi * EARTH_RADIUS_M * np.cos(np.deg2rad(lat)))
  out = torch.empty(t.shape[:-1] + (nx // 2 + 1,), dtype=torch.float32,
                    device=t.device)
  addr = (np.uint64(t.data_ptr()) +
          np.arange(n_jobs, dtype=np.uint64) * np.uint64(ny * nx * 4))
  addr = np.ascontiguousarray(addr, np.uint64)
  desc = _cabi.SpectrumDesc(
      n_jobs=n_jobs, ny=ny, nx=nx,
      field=addr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
      row_scale=scale.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
      spectrum=out.data_ptr())
  ctx = _cabi.get_context(t.device.index)
  ctx.use_torch_stream()
  _cabi.check(ctx.lib.wbx_zonal_spectrum(ctx.handle, ctypes.byref(desc)))
  dims = tuple(outer) + (latitude_name, 'zonal_wavenumber')
  coords = {k: v for k, v in canon.coords.items()
            if longitude_name not in v.dims and k != 'mask'}
  lon = canon.coords[longitude_name].to_numpy() if (
      longitude_name in canon.coords) else np.arange(nx) * 360.0 / nx
  coords.update(_spectral_coords(lat, lon, nx, latitude_name))
  return xl.DataArray(out, dims, coords=coords, name=field.name)


class ZonalEnergySpectrum(base.PerVariableStatistic):
  """Zonal energy spectrum of the predictions (or the targets).

  Args:
    which: 'predictions' or 'targets'.
  """

  def __init__(self, which: str = 'predictions',
               latitude_name: str = 'latitude',
               longitude_name: str = 'longitude'):
    if which not in ('predictions', 'targets'):
      raise ValueError(f'Unhandled {which=}')
    self._which = which
    self._lat, self._lon = latitude_name, longitude_name

  @property
  def unique_name(self) -> str:
    return f'ZonalEnergySpectrum_{self._which}'

  def _compute_per_variable(self, predictions, targets):
    da = predictions if self._which == 'predictions' else targets
    if self._lon not in da.dims or self._lat not in da.dims:
      return None
    return zonal_energy_spectrum(da, self._lat, self._lon)
