"""The SEEPS per-point function of the CUDA kernel, checked on the CPU.

``weatherbenchx_b200/csrc/seeps_point.h`` holds the arithmetic of one grid
point as a host/device inline function; ``seeps.cu`` calls it from the
elementwise kernel.  Here the SAME header is compiled for the host with g++
(into a temporary shared object, test infrastructure only) and compared bit for
bit with the oracle's restatement of metrics/categorical.py:217-296 -- which
tests/test_reference_golden.py pins to the reference's own code -- on inputs
that hit every comparison boundary.  This validates the arithmetic of the
device function without a GPU; indexing / launch of the kernel itself are
covered by the ``-m gpu`` tests.
"""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import wbx_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER_DIR = os.path.join(ROOT, 'weatherbenchx_b200', 'csrc')

SHIM = r'''
#include "seeps_point.h"
extern "C" void seeps_points(const float* p, const float* t, const float* wet,
                             const float* p1, float dry, long n, float* out) {
  for (long i = 0; i < n; ++i)
    out[i] = wbx_seeps_point(p[i], t[i], wet[i], p1[i], dry);
}
'''


@pytest.fixture(scope='module')
def host_copy(tmp_path_factory):
  gxx = shutil.which('g++')
  if gxx is None:
    pytest.skip('g++ not available')
  work = tmp_path_factory.mktemp('seeps_host')
  src = work / 'shim.cc'
  src.write_text(SHIM)
  lib = work / 'libseeps_host.so'
  subprocess.run([gxx, '-O2', '-ffp-contract=off', '-shared', '-fPIC',
                  '-I', HEADER_DIR, '-o', str(lib), str(src)], check=True)
  dll = ctypes.CDLL(str(lib))
  fptr = ctypes.POINTER(ctypes.c_float)
  dll.seeps_points.argtypes = [fptr, fptr, fptr, fptr, ctypes.c_float,
                               ctypes.c_long, fptr]
  dll.seeps_points.restype = None

  def run(p, t, wet, p1, dry):
    arrays = [np.ascontiguousarray(a, np.float32) for a in (p, t, wet, p1)]
    out = np.empty(arrays[0].shape, np.float32)
    dll.seeps_points(*[a.ctypes.data_as(fptr) for a in arrays],
                     ctypes.c_float(dry), out.size, out.ctypes.data_as(fptr))
    return out

  return run


def _inputs(seed, n=20000):
  rng = np.random.default_rng(seed)
  quarter = lambda lo, hi: (np.round(rng.uniform(lo, hi, n) * 4) / 4  # noqa: E731
                            ).astype(np.float32)
  p = quarter(0, 4) * (rng.random(n) < 0.7)
  t = quarter(0, 4) * (rng.random(n) < 0.7)
  wet = quarter(0.5, 3)
  wet[rng.random(n) < 0.05] = np.float32(0.25)     # wet == dry threshold
  wet[rng.random(n) < 0.02] = np.float32(0.0)      # wet < dry threshold
  wet[rng.random(n) < 0.02] = np.nan
  p1 = rng.uniform(0.02, 0.98, n).astype(np.float32)
  p1[rng.random(n) < 0.02] = 0.0
  p1[rng.random(n) < 0.02] = 1.0
  p1[rng.random(n) < 0.03] = np.nan
  p[rng.random(n) < 0.03] = np.nan
  t[rng.random(n) < 0.03] = np.nan
  return p.astype(np.float32), t.astype(np.float32), wet, p1


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_device_function_equals_oracle_bit_for_bit(host_copy, seed):
  p, t, wet, p1 = _inputs(seed)
  dry_mm = 250.0                      # 0.25 in field units: hit exactly
  got = host_copy(p, t, wet, p1, np.float32(dry_mm / 1000.0))
  want, _ = oracle.seeps(p, t, wet, p1, dry_threshold_mm=dry_mm, min_p1=-1.0,
                         max_p1=2.0)  # no range mask: NaN p1 -> NaN only
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  ok = ~np.isnan(want)
  # the oracle's float64 sum of float32 matrix entries; a single non-zero term
  # is exact in float32, sums of several terms round once
  np.testing.assert_array_equal(got[ok], want[ok].astype(np.float32))
  # p1 in {0, 1}: the matrix holds 1/0 = inf, and 0 * inf = NaN poisons the
  # sum over the 3 x 3 table -- in NumPy's einsum and here alike
  edge = (p1 == 0) | (p1 == 1)
  assert edge.sum() > 100 and np.isnan(got[edge]).all()
  assert (got[ok] == 0).sum() > 100           # diagonal of the scoring matrix


def test_known_answers(host_copy):
  """metrics/metrics_test.py:546-584: perfect forecast -> 0; forecast light,
  observation dry -> 0.5 / p1."""
  n = 16
  zeros = np.zeros(n, np.float32)
  wet = np.ones(n, np.float32)
  p1 = np.full(n, 0.4, np.float32)
  dry = np.float32(0.25 / 1000.0)
  np.testing.assert_array_equal(host_copy(zeros, zeros, wet, p1, dry), 0)
  got = host_copy(zeros + np.float32(0.5), zeros, wet, p1, dry)
  np.testing.assert_array_equal(got, np.float32(0.5) * (np.float32(1) / p1))
  np.testing.assert_allclose(got, 1.25, atol=1e-4)
  # heavy forecast, dry observation: 0.5 * (1/p1 + 3/(2+p1))
  got = host_copy(zeros + 2, zeros, wet, p1, dry)
  want = np.float32(0.5) * (np.float32(1) / p1 + np.float32(3) /
                            (np.float32(2) + p1))
  np.testing.assert_array_equal(got, want)
  # NaN anywhere -> NaN
  nan = np.full(n, np.nan, np.float32)
  for args in ((nan, zeros, wet, p1), (zeros, nan, wet, p1),
               (zeros, zeros, wet, nan)):
    assert np.isnan(host_copy(*args, dry)).all()
  # a NaN wet threshold compares false: the point is dry or in no category
  np.testing.assert_array_equal(host_copy(zeros, zeros, nan, p1, dry), 0)
  np.testing.assert_array_equal(host_copy(zeros + 1, zeros + 1, nan, p1, dry), 0)


def test_emulator_contract_matches_device_function(host_copy):
  """The NumPy stand-in the CPU class-surface tests use for
  ``engine.seeps_field`` computes the documented per-point contract."""
  import wbx_emulator
  from weatherbenchx_b200 import xarray_lite as xl
  from weatherbenchx_b200.lazy import AlignedClimatology
  p, t, wet, p1 = _inputs(7, n=6 * 5 * 4)
  dims = ('valid_time', 'latitude', 'longitude')
  P = xl.DataArray(p.reshape(6, 5, 4), dims, name='rain')
  T = xl.DataArray(t.reshape(6, 5, 4), dims, name='rain')
  clim = xl.DataArray(wet.reshape(6, 5, 4), ('time',) + dims[1:])
  aligned = AlignedClimatology(clim, ('valid_time',),
                               {'time': np.arange(6)})
  q = xl.DataArray(p1.reshape(6, 5, 4)[0], dims[1:])
  got = wbx_emulator.seeps_field(P, T, aligned, q, 0.25).to_numpy()
  want = host_copy(p, t, wet, np.tile(p1[:20], 6), np.float32(0.25))
  np.testing.assert_array_equal(got.reshape(-1), want)
