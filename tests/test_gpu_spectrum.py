"""GPU tests of the zonal energy spectrum kernel.

PARITY UNPINNED (no reference implementation exists, SURVEY.md finding 2):
the only checker is the numpy.fft float64 oracle and FFT identities.  f32 FFT
error is absolute w.r.t. the largest bin of a row, so the tolerance is
|S - S_ref| <= 2e-5 * max_k S_ref + 1e-4 * S_ref.
"""

import numpy as np
import pytest

import wbx_oracle as oracle
from weatherbenchx_b200 import aggregation
from weatherbenchx_b200 import weighting
from weatherbenchx_b200 import xarray_lite as xl
from weatherbenchx_b200.metrics import base as metrics_base
from weatherbenchx_b200.metrics import spectral

pytestmark = pytest.mark.gpu


def _close(got, ref):
  tol = 2e-5 * ref.max(axis=-1, keepdims=True) + 1e-4 * ref
  assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) / tol))


def _field(shape, nlat, nlon, seed=0, red=True):
  rng = np.random.default_rng(seed)
  f = rng.normal(size=shape + (nlat, nlon))
  if red:  # red-ish spectrum like atmospheric fields
    spec = np.fft.rfft(f, axis=-1)
    spec *= 1.0 / (1.0 + np.arange(spec.shape[-1])) ** 1.5
    f = np.fft.irfft(spec, n=nlon, axis=-1) + 3.0
  return f.astype(np.float32)


@pytest.mark.parametrize('nlon', [8, 36, 64, 90, 240, 256, 360, 720, 1440, 2880])
def test_matches_numpy_rfft(nlon):
  nlat = 5
  lat = np.linspace(-60, 60, nlat)
  f = _field((2, 3), nlat, nlon, seed=nlon)
  da = xl.DataArray(f, ('time', 'level', 'latitude', 'longitude'),
                    coords={'latitude': lat,
                            'longitude': np.arange(nlon) * 360.0 / nlon},
                    name='u')
  s = spectral.zonal_energy_spectrum(da)
  assert s.dims == ('time', 'level', 'latitude', 'zonal_wavenumber')
  assert s.shape == (2, 3, nlat, nlon // 2 + 1)
  _close(s.values.astype(np.float64), oracle.zonal_energy_spectrum(f, lat))


@pytest.mark.parametrize('kernel', ['default', 'fixed2', 'generic'])
@pytest.mark.parametrize('nlat,nlon', [(721, 1440), (361, 720)])
def test_fixed_shape_kernels_on_many_rows(nlat, nlon, kernel, monkeypatch):
  """The operational grids run compile-time-shaped kernels (spectrum.cu: the
  two-pass 24 x 30 kernel for N = 1440, the three-pass (5, 6, 12) one for
  N = 720; WBX_SPECTRUM_KERNEL selects the three-pass kernel for N = 1440 too,
  or the runtime-shaped generic kernel).  More rows than one
  sweep of the persistent grid (148 SMs x 2 CTAs x 8 rows), several slabs, every
  row against numpy.fft."""
  if kernel == 'default':
    monkeypatch.delenv('WBX_SPECTRUM_KERNEL', raising=False)
  else:
    monkeypatch.setenv('WBX_SPECTRUM_KERNEL', kernel)
  n_fields = 4 if nlat == 721 else 7
  assert n_fields * nlat > 148 * 2 * 8
  lat = np.linspace(-90, 90, nlat)
  f = _field((n_fields,), nlat, nlon, seed=nlat + len(kernel))
  da = xl.DataArray(f, ('level', 'latitude', 'longitude'),
                    coords={'latitude': lat,
                            'longitude': np.arange(nlon) * 360.0 / nlon},
                    name='u')
  s = spectral.zonal_energy_spectrum(da)
  assert s.shape == (n_fields, nlat, nlon // 2 + 1)
  _close(s.values.astype(np.float64), oracle.zonal_energy_spectrum(f, lat))


def test_unsupported_lengths_raise():
  da = xl.DataArray(np.zeros((3, 14), np.float32), ('latitude', 'longitude'),
                    coords={'latitude': [-10.0, 0.0, 10.0]})
  with pytest.raises(Exception):
    spectral.zonal_energy_spectrum(da)          # 7 is not 2^a 3^b 5^c
  odd = xl.DataArray(np.zeros((3, 15), np.float32), ('latitude', 'longitude'),
                     coords={'latitude': [-10.0, 0.0, 10.0]})
  with pytest.raises(Exception):
    spectral.zonal_energy_spectrum(odd)


def test_identities_at_full_size():
  """0.25 degree rows (N = 1440): constant, single sinusoid, Parseval,
  scaling -- the self-made known answers of SURVEY.md section 8(c)."""
  nlat, n = 721, 1440
  lat = np.linspace(-90, 90, nlat)
  circ = 2 * np.pi * oracle.EARTH_RADIUS_M * np.cos(np.deg2rad(lat))
  coords = {'latitude': lat, 'longitude': np.arange(n) * 0.25}
  dims = ('latitude', 'longitude')
  const = spectral.zonal_energy_spectrum(
      xl.DataArray(np.full((nlat, n), 3.0, np.float32), dims, coords=coords)
  ).values.astype(np.float64)
  np.testing.assert_allclose(const[:, 0], circ * 9.0, rtol=1e-5,
                             atol=1e-5 * circ.max() * 9)
  assert np.all(const[:, 1:] <= 1e-9 * (circ[:, None] * 9.0 + 1))
  k0 = 37
  wave = np.cos(2 * np.pi * k0 * np.arange(n) / n).astype(np.float32)
  s = spectral.zonal_energy_spectrum(
      xl.DataArray(np.broadcast_to(wave, (nlat, n)).copy(), dims, coords=coords)
  ).values.astype(np.float64)
  mid = slice(1, nlat - 1)   # cos(lat) ~ 0 at the poles
  np.testing.assert_allclose(s[mid, k0], circ[mid] / 2, rtol=2e-5)
  others = np.delete(s[mid], k0, axis=1)
  assert others.max() <= 1e-9 * circ.max()
  rng = np.random.default_rng(1)
  f = rng.normal(size=(nlat, n)).astype(np.float32)
  s = spectral.zonal_energy_spectrum(
      xl.DataArray(f, dims, coords=coords)).values.astype(np.float64)
  total = s.sum(-1) - s[:, -1] / 2
  parseval = circ / n * (f.astype(np.float64) ** 2).sum(-1)
  np.testing.assert_allclose(total[mid], parseval[mid], rtol=2e-5)
  s2 = spectral.zonal_energy_spectrum(
      xl.DataArray(2 * f, dims, coords=coords)).values.astype(np.float64)
  np.testing.assert_array_equal(s2, 4 * s)   # exact: power-of-two scaling


def test_statistic_and_aggregation():
  """Spectrum as a Statistic, lat-weighted mean over (time, latitude)."""
  nlat, nlon = 19, 36
  lat = np.linspace(-90, 90, nlat)
  f = _field((4,), nlat, nlon, seed=5)
  da = xl.DataArray(f, ('time', 'latitude', 'longitude'),
                    coords={'time': np.arange(4), 'latitude': lat,
                            'longitude': np.arange(nlon) * 10.0}, name='z')
  metrics = {'spectrum': spectral.ZonalEnergySpectrum()}
  stats = metrics_base.compute_unique_statistics_for_all_metrics(
      metrics, {'z': da}, {'z': da})
  assert list(stats) == ['ZonalEnergySpectrum_predictions']
  agg = aggregation.Aggregator(reduce_dims=['time', 'latitude'],
                               weigh_by=[weighting.GridAreaWeighting()])
  values = agg.aggregate_statistics(stats).metric_values(metrics)
  got = values['spectrum.z']
  assert got.dims == ('zonal_wavenumber',)
  ref = oracle.zonal_energy_spectrum(f, lat)
  w = oracle.grid_area_weights(lat)
  expected = (ref * w[None, :, None]).sum((0, 1)) / (4 * w.sum())
  np.testing.assert_allclose(got.values, expected, rtol=1e-4,
                             atol=2e-5 * expected.max())
