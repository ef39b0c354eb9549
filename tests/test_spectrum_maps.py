"""CPU tests of the index arithmetic behind the fixed-shape spectrum kernels
(csrc/spectrum.cu) -- the constants are read from the CUDA source and replayed
in NumPy; the kernels themselves only ever run in the `-m gpu` tests.

  * every ``dft_pfa<N1, N2, A, C>`` instantiation: A, C are the Chinese-
    remainder constants of the Good-Thomas maps and the two-stage transform
    equals the DFT;
  * the two-pass kernel's lane / register layout (pass 1 scatter with an odd
    pitch, pass 2 gather + [k][t] twiddles, lane k holding Z[k + B1 t], the
    shuffle that fetches Z[H - m], the paired real-FFT split) produces the
    WeatherBench 2 zonal power spectrum of a real row;
  * the padded inter-pass layout of the three-pass kernel is conflict free.

PARITY UNPINNED: the reference holds no spectrum (SURVEY.md finding 2); the
checker is numpy.fft."""

import os
import re

import numpy as np
import pytest

SRC = open(os.path.join(os.path.dirname(__file__), '..', 'weatherbenchx_b200',
                        'csrc', 'spectrum.cu')).read()


def _pfa_instantiations():
  found = re.findall(
      r'dft<(\d+)>\(float2\* v\) \{\s*dft_pfa<(\d+), (\d+), (\d+), (\d+)>\(v\);',
      SRC)
  assert len(found) >= 4, found
  return [tuple(int(x) for x in f) for f in found]


def _dft_pfa(v, n1, n2, a, c):
  """The two loops of dft_pfa, with numpy.fft for the inner transforms."""
  n = n1 * n2
  x = np.array([[v[(n2 * i1 + n1 * i2) % n] for i2 in range(n2)]
                for i1 in range(n1)])
  y = np.fft.fft(np.fft.fft(x, axis=0), axis=1)
  out = np.zeros(n, complex)
  for k1 in range(n1):
    for k2 in range(n2):
      out[(a * k1 + c * k2) % n] = y[k1, k2]
  return out


@pytest.mark.parametrize('n,n1,n2,a,c', _pfa_instantiations())
def test_prime_factor_butterflies(n, n1, n2, a, c):
  assert n == n1 * n2 and np.gcd(n1, n2) == 1
  assert a % n1 == 1 % n1 and a % n2 == 0      # A = N2 (N2^-1 mod N1)
  assert c % n2 == 1 % n2 and c % n1 == 0      # C = N1 (N1^-1 mod N2)
  rng = np.random.default_rng(n)
  v = rng.normal(size=n) + 1j * rng.normal(size=n)
  np.testing.assert_allclose(_dft_pfa(v, n1, n2, a, c), np.fft.fft(v),
                             atol=1e-12)


def _two_pass_shapes():
  m = re.search(r'constexpr int kR0 = (\d+), kR1 = (\d+);\s*(?:.*\n)*?.*'
                r'zonal_spectrum_2pass_kernel<(\d+), kR0, kR1>;', SRC)
  assert m, 'two-pass dispatch not found'
  return [(int(m.group(3)), int(m.group(1)), int(m.group(2)))]


@pytest.mark.parametrize('h,r0,r1', _two_pass_shapes())
def test_two_pass_layout(h, r0, r1):
  assert r0 * r1 == h and r0 <= 32 and r1 <= 32
  n = 2 * h
  b0, b1, pitch, half = r1, r0, r0 + 1, r1 // 2
  assert pitch % 2 == 1
  rng = np.random.default_rng(h)
  x = rng.normal(size=n)
  z = x[0::2] + 1j * x[1::2]
  a = np.zeros(pitch * b0, complex)
  for lane in range(b0):                       # pass 1
    v = np.fft.fft(np.array([z[lane + b0 * t] for t in range(r0)]))
    a[pitch * lane:pitch * lane + r0] = v
  reg = np.zeros((b1, r1), complex)            # pass 2: [lane k][register t]
  for k in range(b1):
    v = np.array([a[pitch * t + k] for t in range(r1)])
    v = v * np.exp(-2j * np.pi * np.arange(r1) * k / h)
    reg[k] = np.fft.fft(v)
  full = np.fft.fft(z)
  for k in range(b1):
    for t in range(r1):
      assert abs(reg[k, t] - full[k + b1 * t]) < 1e-9
  s = np.full(h + 1, np.nan)
  w = np.exp(-2j * np.pi * np.arange(h + 1) / n)

  def emit(m, zk, zc):
    e, o = zk + np.conj(zc), zk - np.conj(zc)
    p = w[m] * o
    s[m] = (1 if m == 0 else 2) * 0.25 * (
        (e.real + p.imag) ** 2 + (e.imag - p.real) ** 2) / n ** 2
    s[h - m] = 2 * 0.25 * ((e.real - p.imag) ** 2 + (e.imag + p.real) ** 2
                           ) / n ** 2

  for lane in range(b1):
    src = (2 * b1 - lane) % b1
    for t in range(half):
      # what lane `src` hands over in shuffle step t
      give = reg[src, (r1 - t) % r1] if src == 0 else reg[src, r1 - 1 - t]
      m = lane + b1 * t
      assert abs(give - full[(h - m) % h]) < 1e-9
      emit(m, reg[lane, t], give)
  emit(h // 2, reg[0, half], reg[0, half])     # lane 0, its own partner
  ref = np.abs(np.fft.rfft(x) / n) ** 2 * np.r_[1.0, 2.0 * np.ones(h)]
  assert not np.isnan(s).any()                 # every bin written
  np.testing.assert_allclose(s, ref, rtol=1e-10, atol=1e-16)


@pytest.mark.parametrize('r0,r1,r2', [(5, 12, 12), (5, 6, 12)])
def test_three_pass_padding_is_conflict_free(r0, r1, r2):
  """Second pass of the three-pass kernel: lane j = q r0 + k stores to
  q kQ1 + k + t r0; with kQ1 = r0 (mod 16) the 16 lanes of a half-warp hit 16
  different 8-byte banks (unpadded, R0 R1 = 60 or 30, they collide)."""
  h = r0 * r1 * r2
  assert f'zonal_spectrum_fixed2_kernel<{h}, {r0}, {r1}, {r2}>' in SRC
  kq1 = r0 * r1 + ((r0 - r0 * r1) % 16 + 16) % 16
  assert kq1 % 16 == r0 % 16 and kq1 >= r0 * r1
  b1 = h // r1
  for j0 in range(0, b1, 32):
    for half_warp in (0, 16):
      lanes = [j for j in range(j0 + half_warp, min(j0 + half_warp + 16, b1))]
      banks = [((j // r0) * kq1 + j % r0) % 16 for j in lanes]
      assert len(set(banks)) == len(banks)
      unpadded = [((j // r0) * r0 * r1 + j % r0) % 16 for j in lanes]
      if len(lanes) == 16:
        assert len(set(unpadded)) < 16
  # the third pass reads element j + t B2 of the unpadded order at j + t kQ1
  b2 = h // r2
  assert b2 == r0 * r1
