"""GPU tests of the SEEPS path (categorical.SEEPS -> wbx_seeps_elementwise ->
fused masked reduction) and of probabilistic.EnsembleErrorExceedance (XF
reduction over the members, per-point fallback when a member is NaN).

Added after the main GPU session of round 1 and confirmed on a B200 in a
separate short run (profiles/gpu_tests_seeps_late_cases_r1.log: 23 passed).
Also verified on the CPU: the per-point device function, compiled for the host
from the same header, equals the oracle bit for bit (tests/test_seeps_host.py);
the oracle reproduces the reference's own results on the ten cases used here;
the class surface with interpreted plans reproduces them too
(tests/test_reference_golden.py).
"""

import numpy as np
import pytest

import test_reference_golden as ref
import wbx_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def golden():
  with np.load(ref.GOLDEN) as data:
    return {k: data[k] for k in data.files}


@pytest.fixture(scope='module')
def inputs(golden):
  return {k[3:]: v for k, v in golden.items() if k.startswith('in/')}


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('case', ref.LATE_CASES)
def test_cuda_path_reproduces_reference_late_cases(golden, inputs, case, space):
  """State and values of the reference's own SEEPS / EnsembleErrorExceedance,
  from the CUDA path."""
  ref._run_product_case(golden, inputs, case, space)  # pylint: disable=protected-access


@pytest.mark.parametrize('space', ['host', 'device'])
@pytest.mark.parametrize('case', ref.UNCONFIRMED_CASES)
def test_cuda_path_reproduces_reference_ensemble_of_targets(golden, inputs,
                                                            case, space):
  """Ensemble-of-targets CRPS (a composition of the CRPS launches); passing
  on the B200 since the round-1 driver run."""
  ref._run_product_case(golden, inputs, case, space)  # pylint: disable=protected-access


def _random_inputs(seed, n):
  rng = np.random.default_rng(seed)
  quarter = lambda lo, hi: (np.round(rng.uniform(lo, hi, n) * 4) / 4  # noqa: E731
                            ).astype(np.float32)
  p = (quarter(0, 4) * (rng.random(n) < 0.7)).astype(np.float32)
  t = (quarter(0, 4) * (rng.random(n) < 0.7)).astype(np.float32)
  wet = quarter(0.5, 3)
  wet[rng.random(n) < 0.05] = np.float32(0.25)
  p[rng.random(n) < 0.03] = np.nan
  t[rng.random(n) < 0.03] = np.nan
  return p, t, wet


@pytest.mark.parametrize('slab', [24 * 40, 19 * 37])   # float4 / scalar path
def test_kernel_equals_oracle_bit_for_bit(slab):
  import torch
  from weatherbenchx_b200 import _cabi
  n_rep = 7
  n = slab * n_rep
  p, t, wet = _random_inputs(3, n)
  rng = np.random.default_rng(4)
  p1 = rng.uniform(0.02, 0.98, slab).astype(np.float32)
  p1[rng.random(slab) < 0.05] = np.nan
  dp, dt, dw, dq = (torch.from_numpy(a).cuda() for a in (p, t, wet, p1))
  out = torch.empty(n, dtype=torch.float32, device='cuda')
  ctx = _cabi.get_context()
  ctx.use_torch_stream()
  _cabi.seeps_elementwise(ctx, dp.data_ptr(), dt.data_ptr(), dw.data_ptr(),
                          dq.data_ptr(), slab, 0.25, n, out.data_ptr())
  got = out.cpu().numpy()
  want, _ = oracle.seeps(p, t, wet, np.tile(p1, n_rep),
                         dry_threshold_mm=250.0, min_p1=-1.0, max_p1=2.0)
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  ok = ~np.isnan(want)
  np.testing.assert_array_equal(got[ok], want[ok].astype(np.float32))
  with pytest.raises(_cabi.WbxError):   # n must be a multiple of p1_len
    _cabi.seeps_elementwise(ctx, dp.data_ptr(), dt.data_ptr(), dw.data_ptr(),
                            dq.data_ptr(), slab - 1, 0.25, n, out.data_ptr())


def test_known_answers_through_the_class_surface():
  """metrics/metrics_test.py:546-602: perfect forecast -> 0; forecast light /
  observation dry -> 0.5 / p1 = 1.25; list and scalar parameters agree."""
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.metrics import categorical
  lat, lon = np.linspace(-90, 90, 19), np.linspace(0, 360, 36, endpoint=False)
  init = np.datetime64('2020-01-01T00', 'ns') + np.arange(2) * np.timedelta64(
      1, 'D')
  lead = (np.arange(3) * np.timedelta64(6, 'h')).astype('timedelta64[ns]')
  dims = ('init_time', 'lead_time', 'latitude', 'longitude')
  coords = {'init_time': init, 'lead_time': lead, 'latitude': lat,
            'longitude': lon}
  variables = ['total_precipitation_6hr', 'total_precipitation_24hr']
  zeros = np.zeros((2, 3, 19, 36), np.float32)
  target = {v: xl.DataArray(zeros, dims, coords=coords, name=v)
            for v in variables}
  cdims = ('dayofyear', 'hour', 'latitude', 'longitude')
  ccoords = {'dayofyear': np.arange(1, 367), 'hour': np.arange(0, 24, 6),
             'latitude': lat, 'longitude': lon}
  clim = {}
  for v in variables:
    clim[f'{v}_seeps_dry_fraction'] = xl.DataArray(
        np.full((366, 4, 19, 36), 0.4, np.float32), cdims, coords=ccoords)
    clim[f'{v}_seeps_threshold'] = xl.DataArray(
        np.ones((366, 4, 19, 36), np.float32), cdims, coords=ccoords)
  clim = xl.Dataset(clim)
  seeps = categorical.SEEPS(climatology=clim, variables=variables)
  statistic = seeps.compute(target, target)
  for v in variables:
    np.testing.assert_allclose(statistic[v].values, 0, atol=1e-4)
    assert statistic[v].coords['mask'].values.all()
  prediction = {v: xl.DataArray(zeros + np.float32(0.5), dims, coords=coords,
                                name=v) for v in variables}
  statistic = seeps.compute(prediction, target)
  for v in variables:
    np.testing.assert_allclose(statistic[v].values, 1.25, atol=1e-4)
  seeps2 = categorical.SEEPS(
      climatology=clim, variables=variables, dry_threshold_mm=[0.25, 0.25],
      min_p1=[0.1, 0.1], max_p1=[0.85, 0.85])
  statistic2 = seeps2.compute(prediction, target)
  for v in variables:
    np.testing.assert_array_equal(statistic[v].values, statistic2[v].values)
  assert seeps.unique_name == (
      'SEEPS_total_precipitation_6hr_total_precipitation_24hr_'
      'dry_threshold_mm_0.25_0.25_min_p1_0.1_0.1_max_p1_0.85_0.85')
