// Zonal energy spectrum: real FFT along longitude in shared memory.
//
// north_star row a16.  /root/reference contains NO implementation, call site or
// test of an energy spectrum (SURVEY.md finding 2) -- PARITY UNPINNED.  The
// definition restated here is WeatherBench 2's ZonalEnergySpectrum:
//   F = rfft(f, axis=longitude, norm='forward')
//   S[0] = C |F_0|^2,  S[k>0] = 2 C |F_k|^2,   C(lat) = 2 pi R cos(lat)
// (C arrives as the per-row scale vector; the kernel itself is a generic
// row-wise power spectrum).  Validation: oracle = numpy.fft (float64).
//
// Algorithm: a length-N real row is read as H = N/2 complex numbers
// z[n] = x[2n] + i x[2n+1] (the row bytes ARE that array), transformed by a
// mixed-radix Stockham autosort FFT ping-ponging between two shared-memory
// buffers, and split into the N/2+1 real-FFT bins
//   X[k] = (Z[k] + conj Z[H-k])/2 - i w_N^k (Z[k] - conj Z[H-k])/2 .
// Three kernels share this file:
//  * zonal_spectrum_kernel: any N = 2 * 2^a 3^b 5^c up to 4096.  Radices 2, 3,
//    4, 5 and the in-register composites 8, 9, 10, 16; shared-memory indices
//    padded by one complex per 16 (e + e/16) so that the strided Stockham
//    scatter of the early passes is bank-conflict free; (j / Ns, j % Ns) by
//    host-computed magic multipliers; one group of `gsize` threads per row.
//  * zonal_spectrum_fixed2_kernel<H, 5, R1, 12>: compile-time shapes, one warp
//    per row, three passes with prime-factor 6- / 12-point butterflies, first
//    pass straight from global memory (what N = 720 runs).
//  * zonal_spectrum_2pass_kernel<720, 24, 30>: two passes, one round trip
//    through shared memory, partner bins by warp shuffle (what N = 1440 runs).
// Twiddles come from shared-memory tables computed once per CTA in double
// precision.  Several rows per CTA, persistent grid; HBM traffic is
// 4 B/point in + 4 (N/2+1)/N B/point out.
#include <algorithm>
#include <string.h>

#include "common.cuh"

namespace wbx {

constexpr int kSpecMaxPasses = 12;
constexpr int kSpecMaxH = 2048;

struct SpecParams {
  const uint64_t* field;    // [n_jobs] slab addresses
  const double* row_scale;  // [ny] or NULL
  float* out;               // [n_jobs, ny, H + 1]
  long long n_rows;         // n_jobs * ny
  int ny, nx, H;
  int gsize;                // threads per row (multiple of 32)
  int pad;                  // 1: one pad element per 16 (power-of-two sizes)
  int rows;                 // rows in flight per CTA
  int n_passes;
  int radix[kSpecMaxPasses];
  unsigned magic[kSpecMaxPasses];  // ceil(2^32 / Ns) of every pass
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  return make_float2(a.x - b.x, a.y - b.y);
}
__device__ __forceinline__ float2 mul_mi(float2 a) {  // a * (-i)
  return make_float2(a.y, -a.x);
}

// exp(-2 pi i m / N) for the in-register composite radices.
__constant__ float2 kW8[8] = {
    {1.f, 0.f}, {0.70710678118654752440f, -0.70710678118654752440f},
    {0.f, -1.f}, {-0.70710678118654752440f, -0.70710678118654752440f},
    {-1.f, 0.f}, {-0.70710678118654752440f, 0.70710678118654752440f},
    {0.f, 1.f}, {0.70710678118654752440f, 0.70710678118654752440f}};
__constant__ float2 kW9[9] = {
    {1.f, 0.f},
    {0.76604444311897803520f, -0.64278760968653932632f},
    {0.17364817766693034885f, -0.98480775301220805937f},
    {-0.5f, -0.86602540378443864676f},
    {-0.93969262078590838405f, -0.34202014332566873304f},
    {-0.93969262078590838405f, 0.34202014332566873304f},
    {-0.5f, 0.86602540378443864676f},
    {0.17364817766693034885f, 0.98480775301220805937f},
    {0.76604444311897803520f, 0.64278760968653932632f}};
__constant__ float2 kW10[10] = {
    {1.f, 0.f},
    {0.80901699437494742410f, -0.58778525229247312917f},
    {0.30901699437494742410f, -0.95105651629515357212f},
    {-0.30901699437494742410f, -0.95105651629515357212f},
    {-0.80901699437494742410f, -0.58778525229247312917f},
    {-1.f, 0.f},
    {-0.80901699437494742410f, 0.58778525229247312917f},
    {-0.30901699437494742410f, 0.95105651629515357212f},
    {0.30901699437494742410f, 0.95105651629515357212f},
    {0.80901699437494742410f, 0.58778525229247312917f}};
__constant__ float2 kW16[16] = {
    {1.f, 0.f},
    {0.92387953251128675613f, -0.38268343236508977173f},
    {0.70710678118654752440f, -0.70710678118654752440f},
    {0.38268343236508977173f, -0.92387953251128675613f},
    {0.f, -1.f},
    {-0.38268343236508977173f, -0.92387953251128675613f},
    {-0.70710678118654752440f, -0.70710678118654752440f},
    {-0.92387953251128675613f, -0.38268343236508977173f},
    {-1.f, 0.f},
    {-0.92387953251128675613f, 0.38268343236508977173f},
    {-0.70710678118654752440f, 0.70710678118654752440f},
    {-0.38268343236508977173f, 0.92387953251128675613f},
    {0.f, 1.f},
    {0.38268343236508977173f, 0.92387953251128675613f},
    {0.70710678118654752440f, 0.70710678118654752440f},
    {0.92387953251128675613f, 0.38268343236508977173f}};

template <int R>
__device__ __forceinline__ void dft(float2* v);

template <>
__device__ __forceinline__ void dft<2>(float2* v) {
  const float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void dft<4>(float2* v) {
  const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
  const float2 c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}

template <>
__device__ __forceinline__ void dft<3>(float2* v) {
  const float s = 0.86602540378443864676f;
  const float2 t1 = cadd(v[1], v[2]);
  const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
  const float2 d = csub(v[1], v[2]);
  const float2 t3 = make_float2(s * d.y, -s * d.x);  // -i s d
  v[0] = cadd(v[0], t1);
  v[1] = cadd(t2, t3);
  v[2] = csub(t2, t3);
}

template <>
__device__ __forceinline__ void dft<5>(float2* v) {
  const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
  const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
  const float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
  const float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
  const float2 x0 = v[0];
  v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
  const float2 p1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x,
                                x0.y + c1 * a1.y + c2 * a2.y);
  const float2 p2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x,
                                x0.y + c2 * a1.y + c1 * a2.y);
  const float2 q1 = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
  const float2 q2 = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
  v[1] = cadd(p1, q1);
  v[4] = csub(p1, q1);
  v[2] = cadd(p2, q2);
  v[3] = csub(p2, q2);
}

// In-register Cooley-Tukey for N = R1 * R2 with compile-time indices:
//   X[k1 + R1 k2] = sum_n2 W_N^{n2 k1} ( sum_n1 x[R2 n1 + n2] W_R1^{n1 k1} )
//                   W_R2^{n2 k2}
template <int R1, int R2>
__device__ __forceinline__ void dft_composite(float2* v, const float2* wN) {
  float2 y[R1 * R2];
#pragma unroll
  for (int n2 = 0; n2 < R2; ++n2) {
    float2 t[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) t[n1] = v[R2 * n1 + n2];
    dft<R1>(t);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1)
      y[k1 * R2 + n2] = (n2 * k1 == 0) ? t[k1] : cmul(t[k1], wN[n2 * k1]);
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    float2 t[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) t[n2] = y[k1 * R2 + n2];
    dft<R2>(t);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
  }
}

template <>
__device__ __forceinline__ void dft<8>(float2* v) {
  dft_composite<4, 2>(v, kW8);
}
template <>
__device__ __forceinline__ void dft<9>(float2* v) {
  dft_composite<3, 3>(v, kW9);
}
template <>
__device__ __forceinline__ void dft<10>(float2* v) {
  dft_composite<5, 2>(v, kW10);
}
template <>
__device__ __forceinline__ void dft<16>(float2* v) {
  dft_composite<4, 4>(v, kW16);
}

// Good-Thomas prime-factor DFT for coprime N1 * N2 (compile-time indices): with
//   n = (N2 n1 + N1 n2) mod N,   k = (A k1 + C k2) mod N,
//   A = N2 (N2^-1 mod N1),       C = N1 (N1^-1 mod N2)
// the transform is N2 DFTs of length N1 followed by N1 of length N2 with NO
// twiddles in between: 12 points cost 4 x dft<3> + 3 x dft<4>, fewer FP32
// instructions per point and level than the radix-9 / radix-10 composites.
template <int N1, int N2, int A, int C>
__device__ __forceinline__ void dft_pfa(float2* v) {
  constexpr int N = N1 * N2;
  float2 y[N];
#pragma unroll
  for (int n2 = 0; n2 < N2; ++n2) {
    float2 t[N1];
#pragma unroll
    for (int n1 = 0; n1 < N1; ++n1) t[n1] = v[(N2 * n1 + N1 * n2) % N];
    dft<N1>(t);
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) y[k1 * N2 + n2] = t[k1];
  }
#pragma unroll
  for (int k1 = 0; k1 < N1; ++k1) {
    float2 t[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) t[n2] = y[k1 * N2 + n2];
    dft<N2>(t);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) v[(A * k1 + C * k2) % N] = t[k2];
  }
}
template <>
__device__ __forceinline__ void dft<6>(float2* v) {
  dft_pfa<2, 3, 3, 4>(v);
}
template <>
__device__ __forceinline__ void dft<12>(float2* v) {
  dft_pfa<3, 4, 4, 9>(v);
}
template <>
__device__ __forceinline__ void dft<24>(float2* v) {
  dft_pfa<3, 8, 16, 9>(v);
}
template <>
__device__ __forceinline__ void dft<30>(float2* v) {
  dft_pfa<5, 6, 6, 25>(v);
}

// One Stockham pass of radix R over a length-H sequence held in shared memory.
// q = j / Ns via the magic multiplier; `pm` is the padding mask (see kernel).
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in,
                                              float2* __restrict__ out,
                                              const float2* __restrict__ tw,
                                              const int H, const int Ns,
                                              const unsigned magic, const int pm,
                                              const int lane, const int gsize) {
  const int B = H / R;
  const int tstep = H / (Ns * R);  // twiddle table stride for this pass
  for (int j = lane; j < B; j += gsize) {
    const int q = (Ns == 1) ? j : static_cast<int>(__umulhi(j, magic));
    const int k = j - q * Ns;
    float2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      const int e = j + t * B;
      v[t] = in[e + ((e >> 4) & pm)];
    }
    if (Ns > 1) {
      const int kt = k * tstep;
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul(v[t], tw[t * kt]);
    }
    dft<R>(v);
    const int j0 = q * Ns * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) {
      const int e = j0 + t * Ns;
      out[e + ((e >> 4) & pm)] = v[t];
    }
  }
}

// Barrier over the `gsize` threads that share one row (rows are independent).
__device__ __forceinline__ void group_sync(const int group, const int gsize) {
  if (gsize == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(gsize) : "memory");
  }
}

__global__ void __launch_bounds__(512)
    zonal_spectrum_kernel(const SpecParams P) {
  extern __shared__ __align__(16) unsigned char spec_smem[];
  const int H = P.H;
  const int pm = P.pad ? -1 : 0;
  const int Hp = (P.pad ? H + (H >> 4) : H) + 2;  // padded buffer length (even)
  float2* tw = reinterpret_cast<float2*>(spec_smem);  // exp(-2 pi i q / H)
  float2* twn = tw + H;                               // exp(-2 pi i k / N), k <= H
  float2* bufs = twn + (H + 2);                       // [rows][2][Hp]
  const int group = threadIdx.x / P.gsize;
  const int lane = threadIdx.x - group * P.gsize;
  for (int q = threadIdx.x; q < H; q += blockDim.x) {
    double s, c;
    sincospi(2.0 * q / H, &s, &c);
    tw[q] = make_float2(static_cast<float>(c), static_cast<float>(-s));
  }
  for (int q = threadIdx.x; q <= H; q += blockDim.x) {
    double s, c;
    sincospi(2.0 * q / P.nx, &s, &c);
    twn[q] = make_float2(static_cast<float>(c), static_cast<float>(-s));
  }
  __syncthreads();  // tables ready; from here on the row groups run decoupled
  float2* a = bufs + static_cast<size_t>(group) * 2 * Hp;
  float2* b = a + Hp;
  const float inv_n2 =
      1.0f / (static_cast<float>(P.nx) * static_cast<float>(P.nx));
  const long long rows_per_iter = static_cast<long long>(gridDim.x) * P.rows;
  // Register prefetch of the NEXT row (issued before the FFT passes of the
  // current one) hides the DRAM latency that a one-row-per-warp schedule would
  // otherwise expose.
  constexpr int kPre = 12;
  const int nvec = H >> 1;
  const bool vec_ok = !P.pad && (P.nx & 3) == 0 && nvec <= kPre * P.gsize;
  float4 pre[kPre];
  bool pre_valid = false;
  auto row_pointer = [&](long long r, int* y_out) {
    const long long job = r / P.ny;
    const int y = static_cast<int>(r - job * P.ny);
    *y_out = y;
    return reinterpret_cast<const float*>(__ldg(P.field + job)) +
           static_cast<long long>(y) * P.nx;
  };
  auto prefetch = [&](long long r) {
    int yy;
    const float* rp = row_pointer(r, &yy);
    pre_valid = vec_ok && (reinterpret_cast<uintptr_t>(rp) & 15) == 0;
    if (pre_valid) {
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        const int n = lane + i * P.gsize;
        if (n < nvec) pre[i] = ldg_stream_f4(rp + 4 * n);
      }
    }
  };
  long long row = static_cast<long long>(blockIdx.x) * P.rows + group;
  if (row < P.n_rows) prefetch(row);
  for (; row < P.n_rows; row += rows_per_iter) {
    int y;
    const float* rowp = row_pointer(row, &y);
    group_sync(group, P.gsize);  // the previous row's readers are done
    if (pre_valid) {
      float4* a4 = reinterpret_cast<float4*>(a);
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        const int n = lane + i * P.gsize;
        if (n < nvec) a4[n] = pre[i];
      }
    } else {
      const float2* s2 = reinterpret_cast<const float2*>(rowp);
      for (int n = lane; n < H; n += P.gsize)
        a[n + ((n >> 4) & pm)] = __ldg(s2 + n);
    }
    if (row + rows_per_iter < P.n_rows) prefetch(row + rows_per_iter);
    float2* in = a;
    float2* out = b;
    int Ns = 1;
    for (int p = 0; p < P.n_passes; ++p) {
      group_sync(group, P.gsize);
      const unsigned mg = P.magic[p];
      switch (P.radix[p]) {
        case 2: stockham_pass<2>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 3: stockham_pass<3>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 4: stockham_pass<4>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 5: stockham_pass<5>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 8: stockham_pass<8>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 9: stockham_pass<9>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        case 10: stockham_pass<10>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
        default: stockham_pass<16>(in, out, tw, H, Ns, mg, pm, lane, P.gsize); break;
      }
      Ns *= P.radix[p];
      float2* tmp = in;
      in = out;
      out = tmp;
    }
    group_sync(group, P.gsize);
    // `in` now holds Z[0..H-1]; produce S[k] for k = 0..H.
    const float scale =
        (P.row_scale ? static_cast<float>(__ldg(P.row_scale + y)) : 1.0f) *
        inv_n2;
    float* dst = P.out + row * static_cast<long long>(H + 1);
    for (int k = lane; k <= H; k += P.gsize) {
      const int ek = (k == H) ? 0 : k;
      const int ec = (k == 0 || k == H) ? 0 : H - k;
      const float2 zk = in[ek + ((ek >> 4) & pm)];
      const float2 zc = in[ec + ((ec >> 4) & pm)];
      const float2 zr = make_float2(zc.x, -zc.y);  // conj Z[H-k]
      const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y + zr.y));
      const float2 o = make_float2(0.5f * (zk.x - zr.x), 0.5f * (zk.y - zr.y));
      const float2 wo = cmul(twn[k], o);
      const float xr = e.x + wo.y;  // X = e - i * wo
      const float xi = e.y - wo.x;
      const float factor = (k == 0) ? 1.0f : 2.0f;
      dst[k] = factor * scale * (xr * xr + xi * xi);
    }
  }
}

// ---------------------------------------------------------------------------
// Fixed-shape variants: H and the radices are compile-time constants, one warp
// per row, no padding.  Every index of the Stockham passes (q = j / Ns,
// k = j % Ns, the twiddle stride, the scatter base) folds into constants or
// shift/multiply sequences and the butterfly loops unroll; about two thirds of
// the generic kernel's instructions were this address arithmetic.  They serve
// the operational grids (0.25 deg: N = 1440, 0.5 deg: N = 720).  The first
// such kernel (round 1: radices (9, 10, 8), row staged through shared memory,
// 0.178 ms per C4 step) was replaced by the two below and is gone.
// ---------------------------------------------------------------------------
// IN_T: distance between the R operands of a butterfly in `in` (H / R unless
// the producer padded its output); OUT_Q: distance between consecutive output
// blocks q in `out` (NS * R unless padded for the consumer).
template <int R, int H, int NS, int IN_T = H / R, int OUT_Q = NS * R>
__device__ __forceinline__ void stockham_pass_fixed(
    const float2* __restrict__ in, float2* __restrict__ out,
    const float2* __restrict__ tw, const int lane) {
  constexpr int B = H / R;
#pragma unroll
  for (int j0 = 0; j0 < B; j0 += 32) {
    const int j = j0 + lane;
    if (j0 + 32 <= B || j < B) {
      const int q = j / NS;   // compile-time divisor
      const int k = j - q * NS;
      float2 v[R];
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = in[j + t * IN_T];
      if constexpr (NS > 1) {
        // pass table laid out [k][t]: a lane reads R-1 consecutive twiddles and
        // neighbouring lanes are an odd number of float2 apart (no bank
        // conflicts; the shared exp(-2 pi i q / H) table gave up to 5-way ones)
        const float2* tk = tw + k * (R - 1) - 1;
#pragma unroll
        for (int t = 1; t < R; ++t) v[t] = cmul(v[t], tk[t]);
      }
      dft<R>(v);
      const int base = q * OUT_Q + k;
#pragma unroll
      for (int t = 0; t < R; ++t) out[base + t * NS] = v[t];
    }
  }
}

// ---------------------------------------------------------------------------
// Three-pass fixed-shape kernel: radices (5, 6, 12) for H = 360 (what N = 720
// runs) and (5, 12, 12) for H = 720 (kept selectable, WBX_SPECTRUM_KERNEL=fixed2,
// as the cross-check of the two-pass kernel in the tests).
//  * the butterfly counts 144 / 60 / 60 fill 5 / 2 / 2 warp iterations to
//    90-94 % (80 / 72 / 90 of the (9, 10, 8) split filled 3 each to 75-94 %),
//    and the prime-factor 12- and 6-point butterflies need no inner twiddles;
//  * the first pass takes its operands straight from global memory in the
//    order it needs them (lane j reads z[j + t B0]: 8-byte loads, 256
//    contiguous bytes per warp instruction), prefetched one row ahead in
//    registers -- no staging copy of the row through shared memory;
//  * the real-FFT split handles the bins k and H - k together: they share
//    E = Z[k] + conj Z[H-k], O = Z[k] - conj Z[H-k] and w^(H-k) = -conj w^k,
//        X[k]   = (E.x + P.y,  E.y - P.x) / 2,     P = w^k O
//        X[H-k] = (E.x - P.y, -E.y - P.x) / 2
//    (k = 0 pairs with the Nyquist bin through Z[H] = Z[0]; k = H/2 with
//    itself), which halves the shared-memory reads and the twiddle table of
//    that stage and removes a third of its arithmetic.
// ---------------------------------------------------------------------------
template <int H, int R0, int R1, int R2>
__global__ void __launch_bounds__(256, 2)
    zonal_spectrum_fixed2_kernel(const SpecParams P) {
  static_assert(R0 * R1 * R2 == H, "radices must factor H");
  static_assert(R0 % 2 == 1, "the first pass scatters with stride R0");
  extern __shared__ __align__(16) unsigned char spec_smem[];
  constexpr int N = 2 * H;
  constexpr int B0 = H / R0;
  constexpr int kIt0 = (B0 + 31) / 32;
  float2* tw1 = reinterpret_cast<float2*>(spec_smem);  // exp(-2 pi i t k / (R0 R1))
  float2* tw2 = tw1 + R0 * (R1 - 1);                   // exp(-2 pi i t k / H)
  float2* twn = tw1 + H;                               // exp(-2 pi i k / N), k <= H/2
  float2* bufs = twn + (H / 2 + 2);                    // [rows][H + kPadB]
  // The second pass scatters blocks of R0 consecutive elements R0 R1 apart:
  // with that distance padded to = R0 (mod 16) in 8-byte units the 32 lanes of
  // a store hit 32 different banks (unpadded: 3.9 M two-way conflicts per C4
  // step, 15 % of all shared-memory wavefronts).
  constexpr int kQ1 = R0 * R1 + ((R0 - R0 * R1) % 16 + 16) % 16;
  constexpr int kPadB = R2 * kQ1;                      // length of buffer b
  const int group = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int q = threadIdx.x; q < R0 * (R1 - 1); q += blockDim.x) {
    const int k = q / (R1 - 1), t = q - k * (R1 - 1) + 1;
    double sn, cs;
    sincospi(2.0 * (t * k) / (R0 * R1), &sn, &cs);
    tw1[q] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
  }
  for (int q = threadIdx.x; q < R0 * R1 * (R2 - 1); q += blockDim.x) {
    const int k = q / (R2 - 1), t = q - k * (R2 - 1) + 1;
    double sn, cs;
    sincospi(2.0 * (t * k) / H, &sn, &cs);
    tw2[q] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
  }
  for (int q = threadIdx.x; q <= H / 2; q += blockDim.x) {
    double sn, cs;
    sincospi(2.0 * q / N, &sn, &cs);
    twn[q] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
  }
  __syncthreads();
  float2* a = bufs + static_cast<size_t>(group) * (H + kPadB);
  float2* b = a + H;
  // |X|^2 / N^2 with X = (E -+ ...) / 2
  constexpr float inv_4n2 =
      0.25f / (static_cast<float>(N) * static_cast<float>(N));
  const long long rows_per_iter = static_cast<long long>(gridDim.x) * P.rows;
  float2 pre[kIt0 * R0];
  auto row_pointer = [&](long long r, int* y_out) {
    const long long job = r / P.ny;
    const int y = static_cast<int>(r - job * P.ny);
    *y_out = y;
    return reinterpret_cast<const float2*>(
               reinterpret_cast<const float*>(__ldg(P.field + job)) +
               static_cast<long long>(y) * N);
  };
  auto prefetch = [&](long long r) {
    int yy;
    const float2* rp = row_pointer(r, &yy) + lane;
#pragma unroll
    for (int it = 0; it < kIt0; ++it) {
      if (it * 32 + 32 <= B0 || it * 32 + lane < B0) {
#pragma unroll
        for (int t = 0; t < R0; ++t)
          pre[it * R0 + t] = ldg_stream_f2(rp + it * 32 + t * B0);
      }
    }
  };
  long long row = static_cast<long long>(blockIdx.x) * P.rows + group;
  if (row < P.n_rows) prefetch(row);
  for (; row < P.n_rows; row += rows_per_iter) {
    const int y = static_cast<int>(row % P.ny);
    __syncwarp();  // the previous row's readers of `a` are done
    // pass 1 (no twiddles): registers -> a[j R0 + t]
#pragma unroll
    for (int it = 0; it < kIt0; ++it) {
      const int j = it * 32 + lane;
      if (it * 32 + 32 <= B0 || j < B0) {
        float2 v[R0];
#pragma unroll
        for (int t = 0; t < R0; ++t) v[t] = pre[it * R0 + t];
        dft<R0>(v);
#pragma unroll
        for (int t = 0; t < R0; ++t) a[j * R0 + t] = v[t];
      }
    }
    if (row + rows_per_iter < P.n_rows) prefetch(row + rows_per_iter);
    __syncwarp();
    stockham_pass_fixed<R1, H, R0, H / R1, kQ1>(a, b, tw1, lane);
    __syncwarp();
    stockham_pass_fixed<R2, H, R0 * R1, kQ1>(b, a, tw2, lane);
    __syncwarp();
    const float2* in = a;  // Z[0..H-1]
    const float scale =
        (P.row_scale ? static_cast<float>(__ldg(P.row_scale + y)) : 1.0f) *
        inv_4n2;
    float* dst = P.out + row * static_cast<long long>(H + 1);
#pragma unroll 4
    for (int k = lane; k <= H / 2; k += 32) {
      const float2 zk = in[k];
      const float2 zc = in[k == 0 ? 0 : H - k];
      const float2 e = make_float2(zk.x + zc.x, zk.y - zc.y);  // Z[k] + conj Z[H-k]
      const float2 o = make_float2(zk.x - zc.x, zk.y + zc.y);  // Z[k] - conj Z[H-k]
      const float2 pw = cmul(twn[k], o);
      const float ar = e.x + pw.y, ai = e.y - pw.x;            // 2 X[k]
      const float br = e.x - pw.y, bi = e.y + pw.x;            // 2 X[H-k] (conj)
      const float two = 2.0f * scale;
      dst[k] = (k == 0 ? scale : two) * (ar * ar + ai * ai);
      dst[H - k] = two * (br * br + bi * bi);
    }
  }
}

// ---------------------------------------------------------------------------
// Two-pass variant for H = 720 = 24 * 30: ONE round trip through shared memory
// per row instead of two (the three-pass kernels are bound by shared-memory
// wavefronts: 450 per row against 383 ideal, ncu_spectrum_r2_fixed2_a.txt).
//   pass 1  lanes 0..29: 24 operands z[j + 30 t] straight from global memory
//           (prefetched one row ahead in registers), prime-factor DFT-24
//           (3 x 8, no inner twiddles), stored to a[25 j + t] (odd stride);
//   pass 2  lanes 0..23: operands a[25 t + k], twiddles exp(-2 pi i t k / H)
//           from a [k][t] table, prime-factor DFT-30 (5 x 6); lane k then holds
//           Z[k + 24 t] in register t;
//   split   the partner Z[H - (k + 24 t)] of register t is register 29 - t of
//           lane 24 - k (register (30 - t) % 30 of lane 0 for k = 0): one
//           shuffle per component, no second trip through shared memory; the
//           bins m = k + 24 t < H/2 and H - m are formed together as in the
//           three-pass kernel (lane 0 also owns m = H/2).
// ---------------------------------------------------------------------------
template <int H, int R0, int R1>
__global__ void __launch_bounds__(256, 2)
    zonal_spectrum_2pass_kernel(const SpecParams P) {
  static_assert(R0 * R1 == H, "radices must factor H");
  static_assert(R0 <= 32 && R1 <= 32 && R0 % 2 == 0 && R1 % 2 == 0,
                "one butterfly per lane in both passes");
  extern __shared__ __align__(16) unsigned char spec_smem[];
  constexpr int N = 2 * H;
  constexpr int B0 = R1;          // butterflies (lanes) of pass 1
  constexpr int B1 = R0;          // butterflies (lanes) of pass 2
  constexpr int kPitch = R0 + 1;  // odd: conflict-free scatter of pass 1
  constexpr int kHalfT = R1 / 2;  // registers t < kHalfT hold the bins < H/2
  float2* tw = reinterpret_cast<float2*>(spec_smem);   // [B1][R1 - 1]
  float2* twn = tw + B1 * (R1 - 1);                    // exp(-2 pi i k / N), k <= H/2
  float2* bufs = twn + (H / 2 + 2);                    // [rows][B0 * kPitch]
  const int group = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int q = threadIdx.x; q < B1 * (R1 - 1); q += blockDim.x) {
    const int k = q / (R1 - 1), t = q - k * (R1 - 1) + 1;
    double sn, cs;
    sincospi(2.0 * (t * k) / H, &sn, &cs);
    tw[q] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
  }
  for (int q = threadIdx.x; q <= H / 2; q += blockDim.x) {
    double sn, cs;
    sincospi(2.0 * q / N, &sn, &cs);
    twn[q] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
  }
  __syncthreads();
  float2* a = bufs + static_cast<size_t>(group) * (B0 * kPitch);
  constexpr float inv_4n2 =
      0.25f / (static_cast<float>(N) * static_cast<float>(N));
  const long long rows_per_iter = static_cast<long long>(gridDim.x) * P.rows;
  float2 pre[R0];
  auto prefetch = [&](long long r) {
    const long long job = r / P.ny;
    const int y = static_cast<int>(r - job * P.ny);
    const float2* rp = reinterpret_cast<const float2*>(
                           reinterpret_cast<const float*>(__ldg(P.field + job)) +
                           static_cast<long long>(y) * N) + lane;
    if (lane < B0) {
#pragma unroll
      for (int t = 0; t < R0; ++t) pre[t] = ldg_stream_f2(rp + t * B0);
    }
  };
  const int src = (2 * B1 - lane) % B1;  // partner lane of the split
  long long row = static_cast<long long>(blockIdx.x) * P.rows + group;
  if (row < P.n_rows) prefetch(row);
  for (; row < P.n_rows; row += rows_per_iter) {
    const int y = static_cast<int>(row % P.ny);
    __syncwarp();  // the previous row's readers of `a` are done
    if (lane < B0) {
      dft<R0>(pre);
#pragma unroll
      for (int t = 0; t < R0; ++t) a[lane * kPitch + t] = pre[t];
    }
    __syncwarp();
    float2 z[R1];
    if (lane < B1) {
#pragma unroll
      for (int t = 0; t < R1; ++t) z[t] = a[t * kPitch + lane];
      const float2* tk = tw + lane * (R1 - 1) - 1;
#pragma unroll
      for (int t = 1; t < R1; ++t) z[t] = cmul(z[t], tk[t]);
      dft<R1>(z);
    } else {
#pragma unroll
      for (int t = 0; t < R1; ++t) z[t] = make_float2(0.f, 0.f);
    }
    // the next row's operands travel while this one is split and stored
    if (row + rows_per_iter < P.n_rows) prefetch(row + rows_per_iter);
    const float scale =
        (P.row_scale ? static_cast<float>(__ldg(P.row_scale + y)) : 1.0f) *
        inv_4n2;
    const float two = 2.0f * scale;
    float* dst = P.out + row * static_cast<long long>(H + 1);
    // idle lanes (>= B1) run the arithmetic on lane 0's indices and store
    // nothing: predicated stores instead of a divergent region per register
    const bool owner = lane < B1;
    const int lk = owner ? lane : 0;
    auto emit = [&](const int m, const float2 zk, const float2 zc,
                    const bool store) {
      const float2 e = make_float2(zk.x + zc.x, zk.y - zc.y);  // Z[m] + conj Z[H-m]
      const float2 o = make_float2(zk.x - zc.x, zk.y + zc.y);  // Z[m] - conj Z[H-m]
      const float2 pw = cmul(twn[m], o);
      const float ar = e.x + pw.y, ai = e.y - pw.x;            // 2 X[m]
      const float br = e.x - pw.y, bi = e.y + pw.x;            // 2 conj X[H-m]
      const float lo = (m == 0 ? scale : two) * (ar * ar + ai * ai);
      const float hi = two * (br * br + bi * bi);
      if (store) {
        dst[m] = lo;
        dst[H - m] = hi;
      }
    };
#pragma unroll
    for (int t = 0; t < kHalfT; ++t) {
      const float2 give = (lane == 0) ? z[(R1 - t) % R1] : z[R1 - 1 - t];
      float2 zc;
      zc.x = __shfl_sync(0xffffffffu, give.x, src);
      zc.y = __shfl_sync(0xffffffffu, give.y, src);
      emit(lk + B1 * t, z[t], zc, owner);
    }
    emit(H / 2, z[kHalfT], z[kHalfT], lane == 0);  // its own partner
  }
}

// H = 2^a 3^b 5^c -> radix list preferring few, large passes.
static bool factorise(int H, int* radix, int* n_passes) {
  int a = 0, b = 0, c = 0, rem = H;
  while (rem % 2 == 0) { rem /= 2; ++a; }
  while (rem % 3 == 0) { rem /= 3; ++b; }
  while (rem % 5 == 0) { rem /= 5; ++c; }
  if (rem != 1) return false;
  int n = 0;
  auto push = [&](int r) {
    if (n < kSpecMaxPasses) radix[n] = r;
    ++n;
  };
  while (a >= 1 && c >= 1) { push(10); --a; --c; }
  while (b >= 2) { push(9); b -= 2; }
  while (a >= 4) { push(16); a -= 4; }
  if (a == 3) { push(8); a = 0; }
  if (a == 2) { push(4); a = 0; }
  if (a == 1) { push(2); a = 0; }
  while (b >= 1) { push(3); --b; }
  while (c >= 1) { push(5); --c; }
  if (n > kSpecMaxPasses) return false;
  // Order: an odd radix first if there is one -- the first pass scatters with
  // stride R, which is bank-conflict free for odd R (and needs no twiddles);
  // then descending.
  std::sort(radix, radix + n, [](int x, int y) { return x > y; });
  for (int i = 0; i < n; ++i) {
    if (radix[i] % 2 == 1) {
      std::rotate(radix, radix + i, radix + i + 1);
      break;
    }
  }
  *n_passes = n;
  return true;
}

}  // namespace wbx

extern "C" int wbx_zonal_spectrum(wbx_ctx* ctx, const wbx_spectrum_desc* d) {
  using namespace wbx;
  WBX_REQUIRE(ctx && d, "wbx_zonal_spectrum: NULL argument");
  WBX_REQUIRE(d->n_jobs >= 1 && d->ny >= 1, "spectrum: bad shape");
  WBX_REQUIRE(d->field && d->spectrum, "spectrum: NULL field table / output");
  WBX_REQUIRE(d->nx >= 4 && d->nx % 2 == 0,
              "spectrum: longitude count must be even and >= 4 (got %lld)",
              (long long)d->nx);
  const int H = static_cast<int>(d->nx / 2);
  if (H > kSpecMaxH) {
    set_error("spectrum: nx = %lld exceeds the shared-memory FFT limit %d",
              (long long)d->nx, 2 * kSpecMaxH);
    return WBX_ERR_UNSUPPORTED;
  }
  SpecParams P;
  memset(&P, 0, sizeof(P));
  if (!factorise(H, P.radix, &P.n_passes)) {
    set_error("spectrum: nx/2 = %d has a prime factor other than 2, 3, 5", H);
    return WBX_ERR_UNSUPPORTED;
  }
  int max_b = 1, Ns = 1;
  for (int p = 0; p < P.n_passes; ++p) {
    max_b = std::max(max_b, H / P.radix[p]);
    P.magic[p] = Ns == 1 ? 0u
                         : static_cast<unsigned>(((1ull << 32) + Ns - 1) / Ns);
    Ns *= P.radix[p];
  }
  for (int64_t j = 0; j < d->n_jobs; ++j)
    WBX_REQUIRE(d->field[j] != 0 && d->field[j] % 8 == 0,
                "spectrum: slab %lld is NULL or not 8-byte aligned",
                (long long)j);
  WBX_CUDA(cudaSetDevice(ctx->device));
  // One warp per row while a pass has at most 128 butterflies (ILP instead of
  // block-wide barriers); wider groups only for the longest rows.
  P.gsize = max_b <= 128 ? 32 : std::min(256, (max_b / 2 + 31) / 32 * 32);
  P.rows = std::min(15, std::max(1, 256 / P.gsize));
  P.pad = (P.radix[0] % 2 == 0) ? 1 : 0;
  const int threads = P.gsize * P.rows;
  const int Hp = (P.pad ? H + (H >> 4) : H) + 2;
  const size_t smem = (static_cast<size_t>(H) + (H + 2) +
                       static_cast<size_t>(P.rows) * 2 * Hp) * sizeof(float2);
  WBX_REQUIRE(smem <= std::min<size_t>(ctx->smem_optin, 227 * 1024),
              "spectrum: shared memory budget exceeded");
  // upload the job table and the row scale
  const size_t tbytes = static_cast<size_t>(d->n_jobs) * 8;
  const size_t sbytes = d->row_scale ? static_cast<size_t>(d->ny) * 8 : 0;
  // The tables go into the next slot of a small ring; a slot is rewritten only
  // after the kernel that last read it has finished (normally long ago), so
  // the call itself never waits for the GPU.
  const int slot = ctx->ring_next;
  ctx->ring_next = (slot + 1) % wbx_ctx::kTableRing;
  WBX_CUDA(cudaEventSynchronize(ctx->ring_ev[slot]));
  int rc = ctx->ring_tables[slot].reserve(tbytes + sbytes + 16);
  if (rc != WBX_OK) return rc;
  unsigned char* base = ctx->ring_tables[slot].as<unsigned char>();
  WBX_CUDA(cudaMemcpyAsync(base, d->field, tbytes, cudaMemcpyHostToDevice,
                           ctx->stream));
  if (sbytes)
    WBX_CUDA(cudaMemcpyAsync(base + tbytes, d->row_scale, sbytes,
                             cudaMemcpyHostToDevice, ctx->stream));
  P.field = reinterpret_cast<const uint64_t*>(base);
  P.row_scale = sbytes ? reinterpret_cast<const double*>(base + tbytes) : nullptr;
  P.out = d->spectrum;
  P.n_rows = d->n_jobs * d->ny;
  P.ny = static_cast<int>(d->ny);
  P.nx = static_cast<int>(d->nx);
  P.H = H;
  WBX_CUDA(cudaFuncSetAttribute(zonal_spectrum_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const size_t per_sm = std::min<size_t>(ctx->smem_optin, 227 * 1024);
  long long ctas_per_sm = std::max<size_t>(1, per_sm / (smem + 1024));
  ctas_per_sm = std::min<long long>(ctas_per_sm, std::max(1, 2048 / threads));
  const long long want = (P.n_rows + P.rows - 1) / P.rows;
  const int grid = static_cast<int>(
      std::max(1ll, std::min<long long>(want, ctx->sm_count * ctas_per_sm)));
  int prc = ctx->prof_begin();
  if (prc != WBX_OK) return prc;
  // The operational grids: radices chosen for the warp-per-row kernel, not by
  // factorise() (whose "few large passes" rule serves the generic kernel).
  const bool fixed2_ok = H == 720 || H == 360;
  // WBX_SPECTRUM_KERNEL (debugging / tests): "fixed2" = the three-pass kernel
  // for N = 1440 too, "generic" = the runtime-shaped kernel for every N
  const char* which = getenv("WBX_SPECTRUM_KERNEL");
  const bool want_generic = which && !strcmp(which, "generic");
  const bool want_3pass = which && !strcmp(which, "fixed2");
  if (H == 720 && !want_generic && !want_3pass) {
    constexpr int kR0 = 24, kR1 = 30;
    const int rows2 = 8, threads2 = rows2 * 32;
    const size_t smem2 = (static_cast<size_t>(kR0) * (kR1 - 1) + (H / 2 + 2) +
                          static_cast<size_t>(rows2) * kR1 * (kR0 + 1)) *
                         sizeof(float2);
    P.rows = rows2;
    P.gsize = 32;
    const long long want2 = (P.n_rows + rows2 - 1) / rows2;
    const int grid2 = static_cast<int>(std::max(
        1ll, std::min<long long>(want2, ctx->sm_count * 2ll)));
    auto kern = zonal_spectrum_2pass_kernel<720, kR0, kR1>;
    WBX_CUDA(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(smem2)));
    kern<<<grid2, threads2, smem2, ctx->stream>>>(P);
  } else if (fixed2_ok && !want_generic) {
    const int rows2 = 8, threads2 = rows2 * 32;
    const int r0 = 5, r1 = H == 720 ? 12 : 6, r2 = 12;
    const int q1 = r0 * r1 + (((r0 - r0 * r1) % 16) + 16) % 16;
    const size_t smem2 = (static_cast<size_t>(H) + (H / 2 + 2) +
                          static_cast<size_t>(rows2) * (H + r2 * q1)) *
                         sizeof(float2);
    P.rows = rows2;
    P.gsize = 32;
    const long long want2 = (P.n_rows + rows2 - 1) / rows2;
    const int grid2 = static_cast<int>(std::max(
        1ll, std::min<long long>(want2, ctx->sm_count * 2ll)));
    if (H == 720) {
      auto kern = zonal_spectrum_fixed2_kernel<720, 5, 12, 12>;
      WBX_CUDA(cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
          static_cast<int>(smem2)));
      kern<<<grid2, threads2, smem2, ctx->stream>>>(P);
    } else {
      auto kern = zonal_spectrum_fixed2_kernel<360, 5, 6, 12>;
      WBX_CUDA(cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
          static_cast<int>(smem2)));
      kern<<<grid2, threads2, smem2, ctx->stream>>>(P);
    }
  } else {
    zonal_spectrum_kernel<<<grid, threads, smem, ctx->stream>>>(P);
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  prc = ctx->prof_end();
  if (prc != WBX_OK) return prc;
  WBX_CUDA(cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
  return WBX_OK;
}
