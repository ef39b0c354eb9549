// Ensemble CRPS statistics on the GPU: CRPSSkill and CRPSSpread
// (metrics/probabilistic.py:116-145, 165-247) fused with the weighted
// aggregation of Aggregator.aggregate_stat_var (aggregation.py:337-366).
//
//   skill(point)  = mean_m |x_m - y|
//   spread(point) = sum_{i,j} |x_i - x_j| / (M (M - fair))
//                 = 2 sum_{i<j} |x_i - x_j| / (M (M - fair))
// with skipna_ensemble: NaN members are dropped per point and M becomes the
// per-point count (probabilistic.py:206-209); otherwise a NaN member makes the
// point NaN, exactly as the NumPy reductions of the reference do.
//
// The member-pair sum is O(M^2): 2 FP32 instructions per pair, ~2.55 kFLOP and
// 4 (M + 1) bytes per grid point at M = 50, i.e. at the FP32-issue / HBM ridge
// (DESIGN.md).  One thread owns one grid point.  A CTA stages a tile of G
// points x M members in shared memory (member-major, so the per-thread reads
// are conflict-free whatever the global layout is), then every thread walks
// the pair triangle in 8 x 8 register tiles: 16 LDS feed 128 FP32 instructions.
// The weighted reduction reuses the record scheme of the deterministic kernel
// (per-(CTA, cell, warp) partials, fixed-order second pass).
#include <algorithm>
#include <new>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace wbx {

constexpr int kCrpsThreads = 128;          // = grid points per tile
constexpr int kCrpsWarps = kCrpsThreads / 32;
constexpr int kCrpsPitch = kCrpsThreads + 1;  // smem row pitch (floats)
constexpr int kBlk = 8;                    // register tile edge

struct CrpsParams {
  const uint64_t* ens;      // [n_jobs] address of (member 0, point 0)
  const uint64_t* target;   // [n_jobs]
  const uint64_t* mask;     // [n_jobs] or NULL
  const int32_t* cell;
  const double* w_outer;
  const double* w_y;
  const double* w_x;
  long long n_jobs;
  long long total_tiles;
  long long member_stride;  // elements
  long long point_stride;   // elements
  int cell_base;
  int ny, nx, slab;
  int n_members;
  int tiles_per_slab;
  int fair;
  int skipna_stat;          // Aggregator(skipna=True)
  double* records;
  float* fields[4];         // optional per-point outputs [n_jobs][slab] or NULL
};

// skill / spread of one grid point whose members sit in a shared-memory column
// xs[m * pitch].  ENS_SKIPNA drops NaN members.
// Ensemble moments of one grid point (probabilistic.py:250-336):
//   variance  = sum (x - mean)^2 / (n - 1)             EnsembleVariance
//   umse      = (mean - y)^2 - variance / n            UnbiasedEnsembleMeanSquaredError
template <bool ENS_SKIPNA>
__device__ __forceinline__ void ensemble_moments(const float* __restrict__ xs,
                                                 const int pitch, const int M,
                                                 const float y, float* variance,
                                                 float* umse) {
  float sum = 0.f;
  int n = 0;
  for (int m = 0; m < M; ++m) {
    const float a = xs[m * pitch];
    if (!ENS_SKIPNA || a == a) {
      sum += a;
      ++n;
    }
  }
  const float fn = static_cast<float>(n);
  const float mean = __fdiv_rn(sum, fn);
  float ss = 0.f;
  for (int m = 0; m < M; ++m) {
    const float a = xs[m * pitch];
    if (!ENS_SKIPNA || a == a) {
      const float d = a - mean;
      ss = __fadd_rn(ss, __fmul_rn(d, d));  // NumPy order, no FMA contraction
    }
  }
  const float var = __fdiv_rn(ss, fn - 1.f);
  const float e = mean - y;
  *variance = var;
  *umse = __fsub_rn(__fmul_rn(e, e), __fdiv_rn(var, fn));
}

template <bool ENS_SKIPNA>
__device__ __forceinline__ void crps_point(const float* __restrict__ xs,
                                           const int pitch, const int M,
                                           const float y, const int fair,
                                           float* skill, float* spread) {
  float sk = 0.f, sp = 0.f;
  int n = 0;
  const int Mb = M & ~(kBlk - 1);
  auto pair = [&](float a, float b) {
    const float d = fabsf(a - b);
    if constexpr (ENS_SKIPNA) {
      sp += (d == d) ? d : 0.f;
    } else {
      sp += d;
    }
  };
  auto single = [&](float a) {
    const float d = fabsf(a - y);
    if constexpr (ENS_SKIPNA) {
      const bool ok = (a == a);
      // NaN target still poisons the skill (mean over members of NaN == NaN).
      sk += ok ? d : 0.f;
      n += ok ? 1 : 0;
    } else {
      sk += d;
    }
  };
  for (int ib = 0; ib < Mb; ib += kBlk) {
    float xi[kBlk];
#pragma unroll
    for (int i = 0; i < kBlk; ++i) xi[i] = xs[(ib + i) * pitch];
#pragma unroll
    for (int i = 0; i < kBlk; ++i) single(xi[i]);
#pragma unroll
    for (int i = 0; i < kBlk; ++i)
#pragma unroll
      for (int j = i + 1; j < kBlk; ++j) pair(xi[i], xi[j]);
    for (int jb = ib + kBlk; jb < Mb; jb += kBlk) {
      float xj[kBlk];
#pragma unroll
      for (int j = 0; j < kBlk; ++j) xj[j] = xs[(jb + j) * pitch];
#pragma unroll
      for (int i = 0; i < kBlk; ++i)
#pragma unroll
        for (int j = 0; j < kBlk; ++j) pair(xi[i], xj[j]);
    }
  }
  // tail members (M % 8): pair each with every earlier member.
  for (int t = Mb; t < M; ++t) {
    const float xt = xs[t * pitch];
    single(xt);
    for (int m = 0; m < t; ++m) pair(xt, xs[m * pitch]);
  }
  if constexpr (ENS_SKIPNA) {
    const float fn = static_cast<float>(n);
    *skill = __fdiv_rn(sk, fn);  // n == 0 -> 0/0 = NaN, as nanmean does
    *spread = __fdiv_rn(2.f * sp, fn * (fn - static_cast<float>(fair)));
  } else {
    *skill = __fdiv_rn(sk, static_cast<float>(M));
    *spread = __fdiv_rn(2.f * sp, static_cast<float>(M * (M - fair)));
  }
}

constexpr int kCrpsStats = 4;  // skill, spread, variance, unbiased MSE
// what a launch evaluates (template parameter WHAT): slots 0-1 and / or 2-3.
constexpr int kWantCrps = 1, kWantMoments = 2;
constexpr int kCrpsAcc = 2 * kCrpsStats;  // + one weight sum per statistic

__device__ __forceinline__ void crps_accumulate(const float (&v)[kCrpsStats],
                                                const bool base,
                                                const int skipna_stat,
                                                const double w, double* acc) {
#pragma unroll
  for (int k = 0; k < kCrpsStats; ++k) {
    const bool ok = base && (!skipna_stat || v[k] == v[k]);
    acc[k] += (ok ? static_cast<double>(v[k]) : 0.0) * w;
    acc[kCrpsStats + k] += (ok ? 1.0 : 0.0) * w;
  }
}

// Per-point values of the requested statistics, for callers that bin or
// otherwise post-process the field (wbx_crps_plan_run_fields).
__device__ __forceinline__ void crps_store_fields(const CrpsParams& P,
                                                  const long long job,
                                                  const unsigned e,
                                                  const float (&v)[kCrpsStats]) {
#pragma unroll
  for (int k = 0; k < kCrpsStats; ++k)
    if (P.fields[k]) __stcs(P.fields[k] + job * P.slab + e, v[k]);
}

template <bool ENS_SKIPNA, bool MASK, int WHAT>
__global__ void __launch_bounds__(kCrpsThreads)
    crps_reduce_kernel(const CrpsParams P) {
  extern __shared__ float smem[];
  float* xs = smem;                                  // [M][pitch]
  float* ys = xs + static_cast<size_t>(P.n_members) * kCrpsPitch;  // [G]
  unsigned char* ms = reinterpret_cast<unsigned char*>(ys + kCrpsThreads);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = P.n_members;
  const long long t_begin =
      (static_cast<long long>(blockIdx.x) * P.total_tiles) / gridDim.x;
  const long long t_end =
      (static_cast<long long>(blockIdx.x + 1) * P.total_tiles) / gridDim.x;
  double acc[kCrpsAcc] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int cur_cell = -1;
  long long job = t_begin / P.tiles_per_slab;
  int k = static_cast<int>(t_begin - job * P.tiles_per_slab);
  for (long long g = t_begin; g < t_end; ++g) {
    const float* ea = reinterpret_cast<const float*>(__ldg(P.ens + job));
    const float* ta = reinterpret_cast<const float*>(__ldg(P.target + job));
    const unsigned char* ma = nullptr;
    if constexpr (MASK)
      ma = reinterpret_cast<const unsigned char*>(__ldg(P.mask + job));
    const int cell = __ldg(P.cell + job);
    const double wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
    if (cell != cur_cell) {
      if (cur_cell >= 0) {
        double* rec = P.records +
                      ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                           kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
        for (int a = 0; a < kCrpsAcc; ++a) {
          const double v = warp_sum(acc[a]);
          if (lane == 0) rec[a] = v;
          acc[a] = 0.0;
        }
      }
      cur_cell = cell;
    }
    const int e0 = k * kCrpsThreads;
    const int len = min(kCrpsThreads, P.slab - e0);
    __syncthreads();  // everyone is done with the previous tile
    // ---- stage the tile: coalesce along whichever axis is contiguous ------
    if (P.point_stride == 1) {
      for (int m = warp; m < M; m += kCrpsWarps) {
        const float* src = ea + static_cast<long long>(m) * P.member_stride + e0;
        for (int q = lane; q < len; q += 32)
          xs[m * kCrpsPitch + q] = ldg_stream_f1(src + q);
      }
    } else {
      // member axis fastest (member_stride == 1 typically): walk the tile in
      // memory order and transpose into the member-major layout.
      const long long total = static_cast<long long>(len) * M;
      for (long long q = tid; q < total; q += kCrpsThreads) {
        const int pt = static_cast<int>(q / M);
        const int m = static_cast<int>(q - static_cast<long long>(pt) * M);
        xs[m * kCrpsPitch + pt] = ldg_stream_f1(
            ea + static_cast<long long>(e0 + pt) * P.point_stride +
            static_cast<long long>(m) * P.member_stride);
      }
    }
    if (tid < len) {
      ys[tid] = ldg_stream_f1(ta + e0 + tid);
      if constexpr (MASK) ms[tid] = __ldg(ma + e0 + tid);
    }
    __syncthreads();
    if (tid < len) {
      float v[kCrpsStats] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (WHAT & kWantCrps)
        crps_point<ENS_SKIPNA>(xs + tid, kCrpsPitch, M, ys[tid], P.fair, &v[0],
                               &v[1]);
      if constexpr (WHAT & kWantMoments)
        ensemble_moments<ENS_SKIPNA>(xs + tid, kCrpsPitch, M, ys[tid], &v[2],
                                     &v[3]);
      const unsigned e = static_cast<unsigned>(e0 + tid);
      crps_store_fields(P, job, e, v);
      const unsigned yy = e / static_cast<unsigned>(P.nx);
      const unsigned xx = e - yy * static_cast<unsigned>(P.nx);
      double w = wo;
      if (P.w_y) w *= __ldg(P.w_y + yy);
      if (P.w_x) w *= __ldg(P.w_x + xx);
      bool base = true;
      if constexpr (MASK) base = ms[tid] != 0;
      crps_accumulate(v, base, P.skipna_stat, w, acc);
    }
    if (++k == P.tiles_per_slab) {
      k = 0;
      ++job;
    }
  }
  if (cur_cell >= 0) {
    double* rec = P.records +
                  ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                       kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
    for (int a = 0; a < kCrpsAcc; ++a) {
      const double v = warp_sum(acc[a]);
      if (lane == 0) rec[a] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// TMA-staged variant of the pair kernel for member-major ensembles
// (point_stride == 1, 16-byte aligned rows): a producer warp streams the tile
// of the NEXT 128 points -- one 512-byte cp.async.bulk per member row, plus the
// target row -- into the other half of a two-stage shared-memory ring while the
// four compute warps walk the pair triangle of the current tile.  No LDG/STS or
// address arithmetic in the compute warps, no block-wide barrier per tile.
// ---------------------------------------------------------------------------
constexpr int kCrpsTmaThreads = kCrpsThreads + 32;
constexpr int kCrpsStages = 2;

struct CrpsStageMeta {
  int cell;
  int len;
  int e0;
  int job;
  double wo;
};

template <bool ENS_SKIPNA, bool MASK, int WHAT>
__global__ void __launch_bounds__(kCrpsTmaThreads)
    crps_reduce_tma_kernel(const CrpsParams P) {
  extern __shared__ __align__(128) unsigned char crps_smem[];
  const int M = P.n_members;
  // stage layout: xs [M][128] f32 | ys [128] f32 | ms [128] u8
  const size_t stage_bytes =
      (static_cast<size_t>(M) * kCrpsThreads + kCrpsThreads) * 4 + kCrpsThreads;
  const size_t stage_stride = (stage_bytes + 127) / 128 * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(crps_smem + kCrpsStages * stage_stride);
  uint64_t* empty = full + kCrpsStages;
  CrpsStageMeta* meta = reinterpret_cast<CrpsStageMeta*>(empty + kCrpsStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_begin =
      (static_cast<long long>(blockIdx.x) * P.total_tiles) / gridDim.x;
  const long long t_end =
      (static_cast<long long>(blockIdx.x + 1) * P.total_tiles) / gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < kCrpsStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kCrpsWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kCrpsWarps) {
    if (lane == 0) {
      const uint64_t policy = l2_evict_first_policy();
      long long job = t_begin / P.tiles_per_slab;
      int k = static_cast<int>(t_begin - job * P.tiles_per_slab);
      int s = 0;
      uint32_t ph = 0;
      for (long long g = t_begin; g < t_end; ++g) {
        const float* ea = reinterpret_cast<const float*>(__ldg(P.ens + job));
        const float* ta = reinterpret_cast<const float*>(__ldg(P.target + job));
        const int e0 = k * kCrpsThreads;
        const int len = min(kCrpsThreads, P.slab - e0);
        mbar_wait(&empty[s], ph ^ 1u);
        CrpsStageMeta mt;
        mt.cell = __ldg(P.cell + job);
        mt.len = len;
        mt.e0 = e0;
        mt.job = static_cast<int>(job);
        mt.wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
        meta[s] = mt;
        unsigned char* st = crps_smem + s * stage_stride;
        float* xs = reinterpret_cast<float*>(st);
        const uint32_t row_bytes = static_cast<uint32_t>(len) * 4u;
        mbar_expect_tx(&full[s], row_bytes * static_cast<uint32_t>(M + 1) +
                                     (MASK ? static_cast<uint32_t>(len) : 0u));
        for (int m = 0; m < M; ++m)
          bulk_g2s(xs + m * kCrpsThreads,
                   ea + static_cast<long long>(m) * P.member_stride + e0,
                   row_bytes, &full[s], policy);
        bulk_g2s(xs + M * kCrpsThreads, ta + e0, row_bytes, &full[s], policy);
        if constexpr (MASK) {
          const unsigned char* ma =
              reinterpret_cast<const unsigned char*>(__ldg(P.mask + job));
          bulk_g2s(st + (static_cast<size_t>(M) + 1) * kCrpsThreads * 4, ma + e0,
                   static_cast<uint32_t>(len), &full[s], policy);
        }
        if (++k == P.tiles_per_slab) {
          k = 0;
          ++job;
        }
        if (++s == kCrpsStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  double acc[kCrpsAcc] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int cur_cell = -1;
  int s = 0;
  uint32_t ph = 0;
  for (long long g = t_begin; g < t_end; ++g) {
    mbar_wait(&full[s], ph);
    const CrpsStageMeta mt = meta[s];
    if (mt.cell != cur_cell) {
      if (cur_cell >= 0) {
        double* rec = P.records +
                      ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                           kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
        for (int a = 0; a < kCrpsAcc; ++a) {
          const double v = warp_sum(acc[a]);
          if (lane == 0) rec[a] = v;
          acc[a] = 0.0;
        }
      }
      cur_cell = mt.cell;
    }
    const unsigned char* st = crps_smem + s * stage_stride;
    const float* xs = reinterpret_cast<const float*>(st);
    if (tid < mt.len) {
      float v[kCrpsStats] = {0.f, 0.f, 0.f, 0.f};
      const float yv = xs[M * kCrpsThreads + tid];
      if constexpr (WHAT & kWantCrps)
        crps_point<ENS_SKIPNA>(xs + tid, kCrpsThreads, M, yv, P.fair, &v[0],
                               &v[1]);
      if constexpr (WHAT & kWantMoments)
        ensemble_moments<ENS_SKIPNA>(xs + tid, kCrpsThreads, M, yv, &v[2],
                                     &v[3]);
      const unsigned e = static_cast<unsigned>(mt.e0 + tid);
      crps_store_fields(P, mt.job, e, v);
      const unsigned yy = e / static_cast<unsigned>(P.nx);
      const unsigned xx = e - yy * static_cast<unsigned>(P.nx);
      double w = mt.wo;
      if (P.w_y) w *= __ldg(P.w_y + yy);
      if (P.w_x) w *= __ldg(P.w_x + xx);
      bool base = true;
      if constexpr (MASK)
        base = st[(static_cast<size_t>(M) + 1) * kCrpsThreads * 4 + tid] != 0;
      crps_accumulate(v, base, P.skipna_stat, w, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == kCrpsStages) {
      s = 0;
      ph ^= 1u;
    }
  }
  if (cur_cell >= 0) {
    double* rec = P.records +
                  ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                       kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
    for (int a = 0; a < kCrpsAcc; ++a) {
      const double v = warp_sum(acc[a]);
      if (lane == 0) rec[a] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Sort / probability-weighted-moment estimator (probabilistic.py:214-240):
//   sum_{i,j} |x_i - x_j| = 2 sum_k (2k - n - 1) x_(k)      (x_(k) sorted)
// The n <= MAXM members of a point live in registers and are sorted by a fully
// unrolled Batcher odd-even merge network (543 compare-exchanges at MAXM = 64
// against 2450 FADDs of the pair sum at M = 50); padding lanes and, with
// skipna_ensemble, NaN members hold +inf and sort to the end.  The moment sum
// is taken about the minimum (the coefficients sum to zero), which removes the
// cancellation the reference avoids by evaluating this branch in float64.
// Members are read straight from global memory (coalesced across threads for
// member-major ensembles, L1-served 4 (M) B runs for member-last ones).
// ---------------------------------------------------------------------------
// Batcher's odd-even merge sort as compile-time recursion, so that every
// compare-exchange has constant register indices (a loop-nest formulation made
// ptxas spill the array to local memory).
template <int MAXM>
__device__ __forceinline__ void cmp_exchange(float (&x)[MAXM], const int a,
                                             const int b) {
  const float lo = fminf(x[a], x[b]);
  const float hi = fmaxf(x[a], x[b]);
  x[a] = lo;
  x[b] = hi;
}

// merges the two sorted halves of x[LO .. LO+N) taken with stride R
template <int MAXM, int LO, int N, int R>
struct OddEvenMerge {
  static __device__ __forceinline__ void run(float (&x)[MAXM]) {
    constexpr int kStep = R * 2;
    if constexpr (kStep < N) {
      OddEvenMerge<MAXM, LO, N, kStep>::run(x);
      OddEvenMerge<MAXM, LO + R, N, kStep>::run(x);
#pragma unroll
      for (int i = LO + R; i + R < LO + N; i += kStep)
        cmp_exchange<MAXM>(x, i, i + R);
    } else {
      cmp_exchange<MAXM>(x, LO, LO + R);
    }
  }
};

template <int MAXM, int LO, int N>
struct OddEvenSort {
  static __device__ __forceinline__ void run(float (&x)[MAXM]) {
    if constexpr (N > 1) {
      OddEvenSort<MAXM, LO, N / 2>::run(x);
      OddEvenSort<MAXM, LO + N / 2, N / 2>::run(x);
      OddEvenMerge<MAXM, LO, N, 1>::run(x);
    }
  }
};

template <int MAXM>
__device__ __forceinline__ void sort_network(float (&x)[MAXM]) {
  OddEvenSort<MAXM, 0, MAXM>::run(x);
}

// Fixed-size networks for the operational ensemble sizes (sort_networks.inc,
// generated by gen_sort_networks.py): odd-even merge sort on runs of arbitrary
// length, 395 (n = 50) / 408 (n = 51) compare-exchanges against the 421 / 433
// ptxas leaves of the 64-wire network with +inf pads.  The sorted order sits
// in a fixed permutation of the wires (WBX_SORT<n>_RANKS).
#include "sort_networks.inc"

template <int MFIX>
struct FixedSort {
  static constexpr bool kAvailable = false;
};

// The pipe that bounds the sorting network is the ALU pipe (FMNMX: one warp
// instruction per two cycles per scheduler); the FMA pipe next to it is mostly
// idle.  Two rewrites move work across (exact algebra, float rounding only):
//  (1) the moment is formed WITHOUT the network's last layer: a pair of
//      adjacent ranks (q, q + 1) that sits unsorted on wires (a, b) contributes
//      (2q + 2 - n) (a + b) + |a - b| -- FMA-pipe instructions instead of
//      two FMNMX + two FFMA.  The signed part is written with differences
//      only (sort_networks.inc: _HEAD / _TAIL / _MOMENT), so identical
//      members give a spread of exactly zero, as the reference's float64 sum;
//  (2) MIXPCT per cent of the remaining compare-exchanges compute their
//      maximum as (a + b) - min(a, b): one FMNMX + two FADD.  The members are
//      shifted by the first one beforehand, so the rounding of a + b is
//      relative to the ensemble's range, not to the field's magnitude or the
//      forecast error (measured: no change of the spread above 1e-8 relative
//      at 100 %).
constexpr int kSortMixPct = 45;     // measured: 30 / 35 / 40 / 45 / 50 within 1 %
constexpr int kSortFixedCtas = 4;   // resident CTAs per SM asked of ptxas (128 reg.)
__host__ __device__ constexpr bool sort_ce_mixed(const int index,
                                                 const int mixpct) {
  return (index * 61) % 100 < mixpct;
}

template <int MAXM, bool MIXED>
__device__ __forceinline__ void cmp_exchange_v(float (&x)[MAXM], const int a,
                                               const int b) {
  const float lo = fminf(x[a], x[b]);
  float hi;
  if constexpr (MIXED) hi = __fsub_rn(__fadd_rn(x[a], x[b]), lo);
  else hi = fmaxf(x[a], x[b]);
  x[a] = lo;
  x[b] = hi;
}

#define WBX_CE(a, b) cmp_exchange<MAXM>(x, a, b);
#define WBX_CEV(i, a, b) \
  cmp_exchange_v<MAXM, sort_ce_mixed(i, MIXPCT)>(x, a, b);
#define WBX_RANK_TERM(q, i) \
  if ((q) > 0 && (q) < n) sp += static_cast<float>(2 * (q) + 1 - n) * (x[i] - c);
#define WBX_TAIL_ABS(a, b) sd += fabsf(x[a] - x[b]);
#define WBX_MOM_D2(m, a1, b1, a2, b2) \
  sp += static_cast<float>(m) * ((x[a2] + x[b2]) - (x[a1] + x[b1]));
#define WBX_MOM_D1(c, w1, w2) sp += static_cast<float>(c) * (x[w2] - x[w1]);
#define WBX_MOM_A2(m, a, b, w0) \
  sp += static_cast<float>(m) * fmaf(-2.f, x[w0], x[a] + x[b]);
#define WBX_MOM_A1(c, w, w0) sp += static_cast<float>(c) * (x[w] - x[w0]);
#define WBX_FIXED_SORT(NN)                                                     \
  template <>                                                                  \
  struct FixedSort<NN> {                                                       \
    static constexpr bool kAvailable = true;                                   \
    static constexpr int N = NN;                                               \
    template <int MAXM>                                                        \
    static __device__ __forceinline__ void sort(float (&x)[MAXM]) {            \
      WBX_SORT##NN(WBX_CE)                                                     \
    }                                                                          \
    /* sum_q (2q + 1 - n) (x_(q) - x_(0)) over the first n ranks */            \
    template <int MAXM>                                                        \
    static __device__ __forceinline__ float moment(const float (&x)[MAXM],     \
                                                   const int n) {              \
      const float c = x[0]; /* rank 0 is wire 0 in both networks */            \
      float sp = 0.f;                                                          \
      WBX_SORT##NN##_RANKS(WBX_RANK_TERM)                                      \
      return sp;                                                               \
    }                                                                          \
    /* sum_q (2q + 1 - N) x_(q) of N CENTRED members: the network up to its   \
       last layer, then the pair form of the moment */                         \
    template <int MAXM, int MIXPCT>                                            \
    static __device__ __forceinline__ float sorted_moment(float (&x)[MAXM]) {  \
      WBX_SORT##NN##_HEAD(WBX_CEV)                                             \
      float sp = 0.f, sd = 0.f;                                                \
      WBX_SORT##NN##_MOMENT(WBX_MOM_D2, WBX_MOM_D1, WBX_MOM_A2, WBX_MOM_A1)    \
      WBX_SORT##NN##_TAIL(WBX_TAIL_ABS)                                        \
      return sp + sd;                                                          \
    }                                                                          \
  };
WBX_FIXED_SORT(50)
WBX_FIXED_SORT(51)
#undef WBX_FIXED_SORT
#undef WBX_MOM_A1
#undef WBX_MOM_A2
#undef WBX_MOM_D1
#undef WBX_MOM_D2
#undef WBX_TAIL_ABS
#undef WBX_RANK_TERM
#undef WBX_CEV
#undef WBX_CE

// Skill, spread (sort / PWM estimator) and the optional moments of one grid
// point whose members are in x[0 .. M) (x is sorted in place).
template <int MAXM, int MFIX, bool ENS_SKIPNA, bool MOMENTS, int MIXPCT,
          bool CRPS = true>
__device__ __forceinline__ void sort_point(float (&x)[MAXM], const float y,
                                           const int M, const int fair,
                                           float (&v)[kCrpsStats]) {
  const float inf = __int_as_float(0x7f800000);
  // skill, moments + NaN bookkeeping (order-independent sums)
  float sk = 0.f, msum = 0.f;
  int n_nan = 0;
  // The fixed-size networks without skipna_ensemble need no per-member NaN
  // test (three min/max-pipe instructions per member, the pipe that bounds
  // this kernel): sum_m |x_m| is NaN iff some member is (no cancellation:
  // every term is >= 0, and +-inf members stay inf), and a NaN member only
  // has to turn skill and spread into NaN at the end.
  constexpr bool kCheapNan = !ENS_SKIPNA && FixedSort<MFIX>::kAvailable;
  // kSplit: the members are shifted by the first one before they are sorted
  // (the spread does not change under a shift, and the shift keeps the
  // rounding of the FMA-pipe compare-exchanges relative to the ensemble's own
  // range); the skill sum doubles as the NaN detector.
  constexpr bool kSplit = kCheapNan && MIXPCT >= 0;
  float sabs = 0.f;
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    if (m < M) {
      if constexpr (kSplit) {
        if constexpr (MOMENTS) msum += x[m];
      } else if constexpr (kCheapNan) {
        sk += fabsf(x[m] - y);
        sabs += fabsf(x[m]);
        if constexpr (MOMENTS) msum += x[m];
      } else {
        const bool isn = !(x[m] == x[m]);
        n_nan += isn ? 1 : 0;
        const float d = fabsf(x[m] - y);
        sk += (ENS_SKIPNA && isn) ? 0.f : d;
        if constexpr (MOMENTS) msum += (ENS_SKIPNA && isn) ? 0.f : x[m];
        if constexpr (!MOMENTS) {
          if (isn) x[m] = inf;
        }
      }
    }
  }
  if constexpr (kCheapNan && !kSplit) n_nan = (sabs == sabs) ? 0 : 1;
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if constexpr (MOMENTS) {
    const float fnm = static_cast<float>(ENS_SKIPNA ? (M - n_nan) : M);
    const float mean = __fdiv_rn(msum, fnm);
    float ss = 0.f;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        const float d = x[m] - mean;
        if constexpr (kCheapNan) {
          ss = __fadd_rn(ss, __fmul_rn(d, d));  // NaN members propagate
        } else {
          const bool isn = !(x[m] == x[m]);
          ss = __fadd_rn(ss, (ENS_SKIPNA && isn) ? 0.f : __fmul_rn(d, d));
          if (isn) x[m] = inf;
        }
      }
    }
    v[2] = __fdiv_rn(ss, fnm - 1.f);
    v[3] = __fsub_rn(__fmul_rn(mean - y, mean - y), __fdiv_rn(v[2], fnm));
  }
  if constexpr (!CRPS) return;  // moments only: nothing to sort
  const int n = ENS_SKIPNA ? (M - n_nan) : M;
  float sp = 0.f;
  if constexpr (kSplit) {
    // n == M == MFIX here.  The network runs on x_m - x_0: every operand of
    // its FMA-pipe arithmetic is then bounded by the ensemble's own range
    // whatever the field's magnitude or the forecast error is (any shift
    // serves: the coefficients sum to zero).  The skill is summed from the
    // raw members, as everywhere else; for a finite target it is NaN iff a
    // member is (every term is >= 0), which saves a separate NaN sum.
    const float c = x[0];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < MFIX) {
        sk += fabsf(x[m] - y);
        x[m] -= c;
      }
    }
    if (fabsf(y) < inf) {
      n_nan = (sk == sk) ? 0 : 1;
    } else {
      // masked analysis cells: the skill is inf / NaN by itself
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < MFIX) s += x[m];
      }
      n_nan = (s == s) ? 0 : 1;
    }
    sp = FixedSort<MFIX>::template sorted_moment<MAXM, MIXPCT>(x);
  } else if constexpr (FixedSort<MFIX>::kAvailable) {
    FixedSort<MFIX>::template sort<MAXM>(x);
    sp = FixedSort<MFIX>::template moment<MAXM>(x, n);
  } else {
    sort_network<MAXM>(x);
    const float c = x[0];
#pragma unroll
    for (int q = 1; q < MAXM; ++q) {
      if (q < n) sp += static_cast<float>(2 * q + 1 - n) * (x[q] - c);
    }
  }
  const float fn = static_cast<float>(n);
  float skill = __fdiv_rn(sk, fn);
  float spread = __fdiv_rn(2.f * sp,
                           fn * (fn - static_cast<float>(fair)));
  if (!ENS_SKIPNA && n_nan > 0) {
    skill = __int_as_float(0x7fc00000);
    spread = skill;
  }
  v[0] = skill;
  v[1] = spread;
}

// MFIX > 0 fixes the member count at compile time: the +inf padding lanes
// become constants, ptxas folds every compare-exchange that touches them and
// the network shrinks to the size of the real ensemble (M = 50: 64 -> 50 wires).
// CRPS = false: the moments alone (variance, unbiased MSE) from the same
// register-resident members, no network -- an HBM-bound stream of 51 loads and
// ~250 FP32 instructions per point.
template <int MAXM, int MFIX, bool ENS_SKIPNA, bool MASK, bool MOMENTS,
          int MINB = 1, int MIXPCT = -1, bool CRPS = true>
__global__ void __launch_bounds__(kCrpsThreads, MINB)
    crps_sort_kernel(const CrpsParams P) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = MFIX > 0 ? MFIX : P.n_members;
  const long long t_begin =
      (static_cast<long long>(blockIdx.x) * P.total_tiles) / gridDim.x;
  const long long t_end =
      (static_cast<long long>(blockIdx.x + 1) * P.total_tiles) / gridDim.x;
  double acc[kCrpsAcc] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int cur_cell = -1;
  long long job = t_begin / P.tiles_per_slab;
  int k = static_cast<int>(t_begin - job * P.tiles_per_slab);
  const float inf = __int_as_float(0x7f800000);
  for (long long g = t_begin; g < t_end; ++g) {
    const float* ea = reinterpret_cast<const float*>(__ldg(P.ens + job));
    const float* ta = reinterpret_cast<const float*>(__ldg(P.target + job));
    const int cell = __ldg(P.cell + job);
    const double wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
    if (cell != cur_cell) {
      if (cur_cell >= 0) {
        double* rec = P.records +
                      ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                           kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
        for (int a = 0; a < kCrpsAcc; ++a) {
          const double v = warp_sum(acc[a]);
          if (lane == 0) rec[a] = v;
          acc[a] = 0.0;
        }
      }
      cur_cell = cell;
    }
    const int e0 = k * kCrpsThreads;
    const int len = min(kCrpsThreads, P.slab - e0);
    if (tid < len) {
      const unsigned e = static_cast<unsigned>(e0 + tid);
      const float* src = ea + static_cast<long long>(e) * P.point_stride;
      // (one mad.wide.u32 per member address instead of nvcc's chains of
      // 64-bit adds was measured SLOWER: 1.044 -> 1.095 ms, GPU call 31)
      float x[MAXM];
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        x[m] = (m < M) ? ldg_stream_f1(src + static_cast<long long>(m) *
                                                  P.member_stride)
                       : inf;
      const float y = ldg_stream_f1(ta + e);
      float v[kCrpsStats];
      sort_point<MAXM, MFIX, ENS_SKIPNA, MOMENTS, MIXPCT, CRPS>(x, y, M, P.fair,
                                                                v);
      crps_store_fields(P, job, e, v);
      const unsigned yy = e / static_cast<unsigned>(P.nx);
      const unsigned xx = e - yy * static_cast<unsigned>(P.nx);
      double w = wo;
      if (P.w_y) w *= __ldg(P.w_y + yy);
      if (P.w_x) w *= __ldg(P.w_x + xx);
      bool base = true;
      if constexpr (MASK) {
        const unsigned char* ma =
            reinterpret_cast<const unsigned char*>(__ldg(P.mask + job));
        base = __ldg(ma + e) != 0;
      }
      crps_accumulate(v, base, P.skipna_stat, w, acc);
    }
    if (++k == P.tiles_per_slab) {
      k = 0;
      ++job;
    }
  }
  if (cur_cell >= 0) {
    double* rec = P.records +
                  ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                       kCrpsWarps + warp) * kCrpsAcc;
#pragma unroll
    for (int a = 0; a < kCrpsAcc; ++a) {
      const double v = warp_sum(acc[a]);
      if (lane == 0) rec[a] = v;
    }
  }
}


// out[c*4 + s] (statistics) and out_w[c*4 + s] (weights), one CTA per cell.
struct CrpsFinalizeParams {
  const double* records;
  const int32_t* cell_first_job;
  double* out_ws;
  double* out_w;
  long long total_tiles;
  int n_cells, grid_main, tiles_per_slab, accumulate;
};

// One CTA per cell: thread t adds accumulator t % 8 of the records t / 8,
// t / 8 + 128, ... (a warp reads 256 consecutive bytes), then the 128 partial
// sums of every accumulator are combined in a fixed order (shuffles over the
// 4 records of a warp, the 32 warps through shared memory).
constexpr int kCrpsFinalizeThreads = 1024;
static_assert(kCrpsAcc == 8, "crps_finalize_kernel: lane -> accumulator map");

__global__ void __launch_bounds__(kCrpsFinalizeThreads) crps_finalize_kernel(
    const CrpsFinalizeParams F) {
  __shared__ double part[kCrpsFinalizeThreads / 32][kCrpsAcc];
  const int c = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int a = t & (kCrpsAcc - 1);
  const long long ft =
      static_cast<long long>(F.cell_first_job[c]) * F.tiles_per_slab;
  const long long lt =
      static_cast<long long>(F.cell_first_job[c + 1]) * F.tiles_per_slab - 1;
  const long long G = F.grid_main;
  const int b_lo = static_cast<int>(((ft + 1) * G - 1) / F.total_tiles);
  const int b_hi = static_cast<int>(((lt + 1) * G - 1) / F.total_tiles);
  const int n = (b_hi - b_lo + 1) * kCrpsWarps;
  const double* rec =
      F.records + (static_cast<size_t>(b_lo) + c) * kCrpsWarps * kCrpsAcc + a;
  double sum = 0.0;
  for (int i = t >> 3; i < n; i += kCrpsFinalizeThreads / kCrpsAcc)
    sum += rec[static_cast<size_t>(i) * kCrpsAcc];
  sum += __shfl_xor_sync(0xffffffffu, sum, 8);
  sum += __shfl_xor_sync(0xffffffffu, sum, 16);
  if (lane < kCrpsAcc) part[warp][lane] = sum;
  __syncthreads();
  if (t < kCrpsAcc) {
    double total = 0.0;
#pragma unroll 8
    for (int w = 0; w < kCrpsFinalizeThreads / 32; ++w) total += part[w][t];
    double* dst = t < kCrpsStats
                      ? F.out_ws + (size_t)c * kCrpsStats + t
                      : F.out_w + (size_t)c * kCrpsStats + (t - kCrpsStats);
    *dst = F.accumulate ? (*dst + total) : total;
  }
}

// ---------------------------------------------------------------------------
// Per-point CRPS statistics for arbitrary layouts (what CRPSSkill / CRPSSpread
// .compute return when the full field is wanted).
// ---------------------------------------------------------------------------
struct CrpsPointParams {
  const float* ens;
  const float* target;
  long long size[WBX_MAX_DIMS];
  long long e_stride[WBX_MAX_DIMS];
  long long t_stride[WBX_MAX_DIMS];
  long long member_stride;
  long long n_points;
  int ndim, n_members, fair;
  float* skill;
  float* spread;
  float* variance;
  float* umse;
};

template <bool ENS_SKIPNA>
__global__ void __launch_bounds__(kCrpsThreads)
    crps_pointwise_kernel(const CrpsPointParams P) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const long long pt = blockIdx.x * static_cast<long long>(kCrpsThreads) + tid;
  if (pt >= P.n_points) return;
  long long rem = pt, eo = 0, to = 0;
  for (int d = P.ndim - 1; d >= 0; --d) {
    const long long i = rem % P.size[d];
    rem /= P.size[d];
    eo += i * P.e_stride[d];
    to += i * P.t_stride[d];
  }
  float* col = smem + tid;
  for (int m = 0; m < P.n_members; ++m)
    col[m * kCrpsPitch] = P.ens[eo + m * P.member_stride];
  float skill = 0.f, spread = 0.f, variance = 0.f, umse = 0.f;
  const float yv = P.target[to];
  if (P.skill || P.spread)
    crps_point<ENS_SKIPNA>(col, kCrpsPitch, P.n_members, yv, P.fair, &skill,
                           &spread);
  if (P.variance || P.umse)
    ensemble_moments<ENS_SKIPNA>(col, kCrpsPitch, P.n_members, yv, &variance,
                                 &umse);
  if (P.skill) P.skill[pt] = skill;
  if (P.spread) P.spread[pt] = spread;
  if (P.variance) P.variance[pt] = variance;
  if (P.umse) P.umse[pt] = umse;
}

// ---------------------------------------------------------------------------
// Ensemble mean field (wrappers.py:145-148).  One thread per VEC consecutive
// points of the innermost dim; every member row it touches is a coalesced
// 4*VEC-byte-per-lane load, 8 members in flight per thread.  HBM-bound:
// 4*(M+1) B per point.
// ---------------------------------------------------------------------------
template <bool ENS_SKIPNA, int VEC>
__global__ void __launch_bounds__(256)
    ensemble_mean_kernel(const CrpsPointParams P, float* __restrict__ out) {
  const long long g = blockIdx.x * 256ll + threadIdx.x;
  const long long pt = g * VEC;
  if (pt >= P.n_points) return;
  long long rem = pt, eo = 0;
  for (int d = P.ndim - 1; d >= 0; --d) {
    const long long i = rem % P.size[d];
    rem /= P.size[d];
    eo += i * P.e_stride[d];
  }
  const float* __restrict__ src = P.ens + eo;
  float sum[VEC], cnt[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) sum[v] = cnt[v] = 0.f;
  auto add = [&](const float (&x)[VEC]) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (!ENS_SKIPNA || x[v] == x[v]) {
        sum[v] += x[v];
        cnt[v] += 1.f;
      }
    }
  };
  constexpr int kInFlight = 8;
  const int M = P.n_members;
  int m = 0;
  for (; m + kInFlight <= M; m += kInFlight) {
    float x[kInFlight][VEC];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) {
      const float* q = src + (m + k) * P.member_stride;
      if constexpr (VEC == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(q));
        x[k][0] = t.x; x[k][1] = t.y; x[k][2] = t.z; x[k][3] = t.w;
      } else {
        x[k][0] = __ldcs(q);
      }
    }
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) add(x[k]);
  }
  for (; m < M; ++m) {
    float x[VEC];
    const float* q = src + m * P.member_stride;
    if constexpr (VEC == 4) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(q));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
      x[0] = __ldcs(q);
    }
    add(x);
  }
  if constexpr (VEC == 4) {
    float4 r;
    r.x = __fdiv_rn(sum[0], cnt[0]);
    r.y = __fdiv_rn(sum[1], cnt[1]);
    r.z = __fdiv_rn(sum[2], cnt[2]);
    r.w = __fdiv_rn(sum[3], cnt[3]);
    *reinterpret_cast<float4*>(out + pt) = r;
  } else {
    out[pt] = __fdiv_rn(sum[0], cnt[0]);
  }
}

}  // namespace wbx

struct wbx_crps_plan {
  int32_t space = 0, flags = 0;
  int64_t n_jobs = 0, ny = 0, nx = 0, n_members = 0, n_cells = 0;
  int64_t member_stride = 0, point_stride = 0;
  bool has_mask = false, has_wo = false, has_wy = false, has_wx = false;
  std::vector<uint64_t> ens, target, mask;
  std::vector<int32_t> cell;
  std::vector<double> wo, wy, wx;
  int tiles_per_slab = 0;
  size_t smem_bytes = 0;
  bool use_sort = false;  // register sorting network (n_members <= 64)
  int what = 3;           // kWantCrps | kWantMoments
  float* fields[4] = {nullptr, nullptr, nullptr, nullptr};  // run_fields only
  bool tma_ok = false;    // TMA-staged pair kernel allowed (alignment, size)
  size_t smem_plain = 0;  // shared memory of the non-TMA pair kernel
  wbx::DevBuf tables, weights;
  wbx::CrpsParams params{};
  const int32_t* d_cell_first_job = nullptr;
  const double* d_wy = nullptr;
  const double* d_wx = nullptr;
  int grid = 0;
  std::vector<unsigned char> chunk_host[2];
};

namespace wbx {

static inline size_t rup(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int crps_launch(wbx_ctx* ctx, const wbx_crps_plan* plan,
                       const CrpsParams& P, int grid) {
  const bool ens_skipna = (plan->flags & WBX_CRPS_SKIPNA_ENSEMBLE) != 0;
  int prc = ctx->prof_begin();
  if (prc != WBX_OK) return prc;
  const int what = plan->what;
  if (plan->use_sort) {
    // Without skipna_ensemble the fixed-size networks run with the FMA-pipe
    // rewrites (kSortMixPct, sort_ce_mixed) at 4 resident CTAs per SM.
#define WBX_SORT_LAUNCH2(MAXM, MFIX, MOM)                                      \
  do {                                                                         \
    constexpr int kMinB = (MFIX) > 0 ? kSortFixedCtas : 1;                     \
    constexpr int kMix = (MFIX) > 0 ? kSortMixPct : -1;                        \
    if (ens_skipna && plan->has_mask)                                          \
      crps_sort_kernel<MAXM, MFIX, true, true, MOM>                            \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else if (ens_skipna)                                                       \
      crps_sort_kernel<MAXM, MFIX, true, false, MOM>                           \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else if (plan->has_mask)                                                   \
      crps_sort_kernel<MAXM, MFIX, false, true, MOM, kMinB, kMix>              \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else                                                                       \
      crps_sort_kernel<MAXM, MFIX, false, false, MOM, kMinB, kMix>             \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
  } while (0)
#define WBX_MOMENTS_LAUNCH(MAXM, MFIX)                                         \
  do {                                                                         \
    if (ens_skipna && plan->has_mask)                                          \
      crps_sort_kernel<MAXM, MFIX, true, true, true, 1, -1, false>             \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else if (ens_skipna)                                                       \
      crps_sort_kernel<MAXM, MFIX, true, false, true, 1, -1, false>            \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else if (plan->has_mask)                                                   \
      crps_sort_kernel<MAXM, MFIX, false, true, true, 1, -1, false>            \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
    else                                                                       \
      crps_sort_kernel<MAXM, MFIX, false, false, true, 1, -1, false>           \
          <<<grid, kCrpsThreads, 0, ctx->stream>>>(P);                         \
  } while (0)
#define WBX_SORT_LAUNCH(MAXM, MFIX)                                            \
  do {                                                                         \
    if (!(what & kWantCrps)) WBX_MOMENTS_LAUNCH(MAXM, MFIX);                   \
    else if (what & kWantMoments) WBX_SORT_LAUNCH2(MAXM, MFIX, true);          \
    else WBX_SORT_LAUNCH2(MAXM, MFIX, false);                                  \
  } while (0)
    // the common operational ensemble sizes get a pruned network
    if (plan->n_members == 50) WBX_SORT_LAUNCH(64, 50);
    else if (plan->n_members == 51) WBX_SORT_LAUNCH(64, 51);
    else if (plan->n_members <= 8) WBX_SORT_LAUNCH(8, 0);
    else if (plan->n_members <= 16) WBX_SORT_LAUNCH(16, 0);
    else if (plan->n_members <= 32) WBX_SORT_LAUNCH(32, 0);
    else WBX_SORT_LAUNCH(64, 0);
#undef WBX_SORT_LAUNCH
#undef WBX_MOMENTS_LAUNCH
#undef WBX_SORT_LAUNCH2
    WBX_CUDA(cudaGetLastError());
    ctx->launches++;
    return ctx->prof_end();
  }
#define WBX_CRPS_LAUNCH3(A, B, W)                                              \
  do {                                                                         \
    if (use_tma) {                                                             \
      auto kern = crps_reduce_tma_kernel<A, B, W>;                             \
      WBX_CUDA(cudaFuncSetAttribute(                                           \
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
          static_cast<int>(plan->smem_bytes)));                                \
      kern<<<grid, kCrpsTmaThreads, plan->smem_bytes, ctx->stream>>>(P);       \
    } else {                                                                   \
      auto kern = crps_reduce_kernel<A, B, W>;                                 \
      WBX_CUDA(cudaFuncSetAttribute(                                           \
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
          static_cast<int>(plan->smem_bytes)));                                \
      kern<<<grid, kCrpsThreads, plan->smem_bytes, ctx->stream>>>(P);          \
    }                                                                          \
  } while (0)
#define WBX_CRPS_LAUNCH(A, B)                                                  \
  do {                                                                         \
    if (what == kWantCrps) WBX_CRPS_LAUNCH3(A, B, 1);                          \
    else if (what == kWantMoments) WBX_CRPS_LAUNCH3(A, B, 2);                  \
    else WBX_CRPS_LAUNCH3(A, B, 3);                                            \
  } while (0)
  // host-space chunks are staged member-major and aligned, so they take the
  // TMA path whenever the plan allows it.
  const bool use_tma = plan->tma_ok && P.point_stride == 1 &&
                       (P.member_stride % 4) == 0;
  if (ens_skipna && plan->has_mask) WBX_CRPS_LAUNCH(true, true);
  else if (ens_skipna) WBX_CRPS_LAUNCH(true, false);
  else if (plan->has_mask) WBX_CRPS_LAUNCH(false, true);
  else WBX_CRPS_LAUNCH(false, false);
#undef WBX_CRPS_LAUNCH
#undef WBX_CRPS_LAUNCH3
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return ctx->prof_end();
}

static int crps_grid(const wbx_ctx* ctx, const wbx_crps_plan* plan,
                     long long total_tiles) {
  // several CTAs per SM so that staging of one overlaps the pair loop of others
  const size_t per_sm = std::min<size_t>(ctx->smem_optin, 227 * 1024);
  long long ctas_per_sm = std::max<size_t>(1, per_sm / (plan->smem_bytes + 1024));
  ctas_per_sm = std::min<long long>(ctas_per_sm, 8);
  // register-limited, no shared memory; every CTA owns an equal run of tiles,
  // and many short runs balance better than one per resident CTA (measured at
  // M = 50, 4 resident: 8 / 12 / 16 / 24 / 48 per SM -> 1.10 / 1.08 / 1.07 /
  // 1.06 / 1.05 ms; profiles/exp_crps_r2_mix.log)
  if (plan->use_sort) ctas_per_sm = 24;
  const long long g = ctx->sm_count * ctas_per_sm;
  return static_cast<int>(std::max(1ll, std::min(g, total_tiles)));
}

static void crps_cell_first(const wbx_crps_plan* plan, int64_t j0, int64_t j1,
                            std::vector<int32_t>* first) {
  const int c0 = plan->cell[j0];
  const int n = plan->cell[j1 - 1] - c0 + 1;
  first->assign(n + 1, 0);
  for (int64_t j = j0; j < j1; ++j)
    (*first)[plan->cell[j] - c0 + 1] = static_cast<int32_t>(j - j0 + 1);
}

// Packs the job tables of [j0, j1) (with the given operand addresses) into one
// host blob, uploads it and fills P.  Returns the device pointer of the
// cell_first_job table through `d_first`.
static int crps_upload_tables(wbx_ctx* ctx, const wbx_crps_plan* plan,
                              int64_t j0, int64_t j1, const uint64_t* ens,
                              const uint64_t* target, const uint64_t* mask,
                              std::vector<unsigned char>* host, DevBuf* dev,
                              cudaStream_t stream, CrpsParams* P,
                              const int32_t** d_first, int* n_cells) {
  const size_t nj = static_cast<size_t>(j1 - j0);
  std::vector<int32_t> first;
  crps_cell_first(plan, j0, j1, &first);
  *n_cells = static_cast<int>(first.size()) - 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += rup(bytes, 16);
    return o;
  };
  const size_t o_ens = take(nj * 8), o_tgt = take(nj * 8);
  const size_t o_mask = plan->has_mask ? take(nj * 8) : 0;
  const size_t o_wo = plan->has_wo ? take(nj * 8) : 0;
  const size_t o_cell = take(nj * 4);
  const size_t o_first = take(first.size() * 4);
  host->assign(off, 0);
  memcpy(host->data() + o_ens, ens, nj * 8);
  memcpy(host->data() + o_tgt, target, nj * 8);
  if (plan->has_mask) memcpy(host->data() + o_mask, mask, nj * 8);
  if (plan->has_wo) memcpy(host->data() + o_wo, plan->wo.data() + j0, nj * 8);
  memcpy(host->data() + o_cell, plan->cell.data() + j0, nj * 4);
  memcpy(host->data() + o_first, first.data(), first.size() * 4);
  int rc = dev->reserve(off);
  if (rc != WBX_OK) return rc;
  unsigned char* base = dev->as<unsigned char>();
  WBX_CUDA(cudaMemcpyAsync(base, host->data(), off, cudaMemcpyHostToDevice,
                           stream));
  P->ens = reinterpret_cast<const uint64_t*>(base + o_ens);
  P->target = reinterpret_cast<const uint64_t*>(base + o_tgt);
  P->mask = plan->has_mask ? reinterpret_cast<const uint64_t*>(base + o_mask)
                           : nullptr;
  P->w_outer =
      plan->has_wo ? reinterpret_cast<const double*>(base + o_wo) : nullptr;
  P->cell = reinterpret_cast<const int32_t*>(base + o_cell);
  P->w_y = plan->d_wy;
  P->w_x = plan->d_wx;
  P->n_jobs = static_cast<long long>(nj);
  P->total_tiles = static_cast<long long>(nj) * plan->tiles_per_slab;
  P->cell_base = plan->cell[j0];
  P->ny = static_cast<int>(plan->ny);
  P->nx = static_cast<int>(plan->nx);
  P->slab = static_cast<int>(plan->ny * plan->nx);
  P->n_members = static_cast<int>(plan->n_members);
  P->tiles_per_slab = plan->tiles_per_slab;
  P->fair = (plan->flags & WBX_CRPS_FAIR) ? 1 : 0;
  P->skipna_stat = (plan->flags & WBX_FLAG_SKIPNA) ? 1 : 0;
  *d_first = reinterpret_cast<const int32_t*>(base + o_first);
  return WBX_OK;
}

static int crps_finalize(wbx_ctx* ctx, const wbx_crps_plan* plan,
                         const double* records, const int32_t* d_first,
                         int n_cells, int grid_main, long long total_tiles,
                         double* out_ws, double* out_w, int accumulate) {
  CrpsFinalizeParams F;
  F.records = records;
  F.cell_first_job = d_first;
  F.out_ws = out_ws;
  F.out_w = out_w;
  F.total_tiles = total_tiles;
  F.n_cells = n_cells;
  F.grid_main = grid_main;
  F.tiles_per_slab = plan->tiles_per_slab;
  F.accumulate = accumulate;
  crps_finalize_kernel<<<n_cells, kCrpsFinalizeThreads, 0, ctx->stream>>>(F);
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

}  // namespace wbx

extern "C" {

int wbx_crps_plan_create(wbx_ctx* ctx, const wbx_crps_desc* d,
                         wbx_crps_plan** out) {
  WBX_REQUIRE(ctx && d && out, "wbx_crps_plan_create: NULL argument");
  *out = nullptr;
  WBX_REQUIRE(d->space == WBX_SPACE_DEVICE || d->space == WBX_SPACE_HOST,
              "crps: bad space");
  WBX_REQUIRE(d->n_jobs >= 1 && d->n_jobs < (1ll << 31), "crps: bad n_jobs");
  WBX_REQUIRE(d->ny >= 1 && d->nx >= 1 && d->ny * d->nx < (1ll << 30),
              "crps: bad slab shape");
  const bool ens_skipna = (d->flags & WBX_CRPS_SKIPNA_ENSEMBLE) != 0;
  const int stat_mask = d->stat_mask ? d->stat_mask : 15;
  WBX_REQUIRE(stat_mask > 0 && stat_mask < 16, "crps: bad stat_mask");
  if (!ens_skipna && d->n_members < 2 && (stat_mask & 2)) {
    wbx::set_error("Cannot estimate CRPS spread with n_ensemble < 2.");
    return WBX_ERR_INVALID;  // probabilistic.py:210-212
  }
  WBX_REQUIRE(d->n_members >= 1 && d->n_members <= 1024,
              "crps: n_members %lld out of range [1, 1024]",
              (long long)d->n_members);
  WBX_REQUIRE(d->ens && d->target && d->cell, "crps: missing tables");
  WBX_REQUIRE(d->member_stride >= 1 && d->point_stride >= 1,
              "crps: strides must be positive");
  WBX_REQUIRE(d->n_cells >= 1 && d->n_cells <= d->n_jobs, "crps: bad n_cells");
  const bool masked = (d->flags & WBX_FLAG_MASKED) != 0;
  WBX_REQUIRE(masked == (d->mask != nullptr),
              "crps: WBX_FLAG_MASKED and a mask table must be given together");
  WBX_REQUIRE(d->cell[0] == 0, "crps: cell[0] must be 0");
  for (int64_t j = 1; j < d->n_jobs; ++j) {
    const int step = d->cell[j] - d->cell[j - 1];
    WBX_REQUIRE(step == 0 || step == 1, "crps: cell[] must be non-decreasing");
  }
  WBX_REQUIRE(d->cell[d->n_jobs - 1] == d->n_cells - 1,
              "crps: cell[] must end at n_cells - 1");
  for (int64_t j = 0; j < d->n_jobs; ++j)
    WBX_REQUIRE(d->ens[j] && d->target[j] && (!d->mask || d->mask[j]),
                "crps: NULL slab address (job %lld)", (long long)j);
  if (d->space == WBX_SPACE_HOST) {
    const int64_t slab = d->ny * d->nx;
    const bool a = d->point_stride == 1 && d->member_stride >= slab;
    const bool b = d->member_stride == 1 && d->point_stride == d->n_members;
    if (!a && !b) {
      wbx::set_error("crps: host-space plans need member-major slabs "
                     "(point_stride 1) or member-last points (member_stride 1)");
      return WBX_ERR_UNSUPPORTED;
    }
  }
  wbx_crps_plan* p = new (std::nothrow) wbx_crps_plan();
  if (!p) {
    wbx::set_error("crps: out of host memory");
    return WBX_ERR_NOMEM;
  }
  p->space = d->space;
  p->flags = d->flags;
  p->n_jobs = d->n_jobs;
  p->ny = d->ny;
  p->nx = d->nx;
  p->n_members = d->n_members;
  p->n_cells = d->n_cells;
  p->member_stride = d->member_stride;
  p->point_stride = d->point_stride;
  p->has_mask = d->mask != nullptr;
  p->has_wo = d->w_outer != nullptr;
  p->has_wy = d->w_y != nullptr;
  p->has_wx = d->w_x != nullptr;
  p->ens.assign(d->ens, d->ens + d->n_jobs);
  p->target.assign(d->target, d->target + d->n_jobs);
  if (p->has_mask) p->mask.assign(d->mask, d->mask + d->n_jobs);
  p->cell.assign(d->cell, d->cell + d->n_jobs);
  if (p->has_wo) p->wo.assign(d->w_outer, d->w_outer + d->n_jobs);
  if (p->has_wy) p->wy.assign(d->w_y, d->w_y + d->ny);
  if (p->has_wx) p->wx.assign(d->w_x, d->w_x + d->nx);
  const int64_t slab = d->ny * d->nx;
  p->tiles_per_slab =
      static_cast<int>((slab + wbx::kCrpsThreads - 1) / wbx::kCrpsThreads);
  p->smem_bytes = static_cast<size_t>(d->n_members) * wbx::kCrpsPitch * 4 +
                  wbx::kCrpsThreads * 4 + wbx::kCrpsThreads + 64;
  p->smem_plain = p->smem_bytes;
  p->what = ((stat_mask & 3) ? wbx::kWantCrps : 0) |
            ((stat_mask & 12) ? wbx::kWantMoments : 0);
  // register-resident members (n_members <= 64): the sorting network for the
  // CRPS statistics, the same skeleton without the network for moments alone
  // (0.734 ms per 20.8 M points at M = 50 = 0.88 of HBM peak; the pair-kernel
  // skeleton that served them before: 0.828 ms, profiles/exp_crps_r2_call34.log)
  p->use_sort = (d->flags & WBX_CRPS_USE_SORT) != 0 && d->n_members <= 64;
  {
    // TMA variant: member-major rows, everything 16-byte aligned.
    const size_t stage =
        ((static_cast<size_t>(d->n_members) * wbx::kCrpsThreads +
          wbx::kCrpsThreads) * 4 + wbx::kCrpsThreads + 127) / 128 * 128;
    const size_t need = wbx::kCrpsStages * stage + 256;
    bool ok = !(d->flags & WBX_FLAG_FORCE_LDG) && (slab % 4) == 0 &&
              (!p->has_mask || (slab % 16) == 0) &&
              need <= std::min<size_t>(ctx->smem_optin, 227 * 1024);
    if (d->space == WBX_SPACE_DEVICE) {
      ok = ok && d->point_stride == 1 && (d->member_stride % 4) == 0;
      for (int64_t j = 0; j < d->n_jobs && ok; ++j)
        ok = (d->ens[j] % 16) == 0 && (d->target[j] % 16) == 0 &&
             (!p->has_mask || (d->mask[j] % 16) == 0);
    } else {
      ok = ok && d->point_stride == 1;  // host chunks are re-staged member-major
    }
    p->tma_ok = ok && !p->use_sort;
    if (p->tma_ok) p->smem_bytes = need;
  }
  if (!p->use_sort &&
      p->smem_bytes > std::min<size_t>(ctx->smem_optin, 227 * 1024)) {
    delete p;
    wbx::set_error("crps: %lld members need more shared memory than one SM has",
                   (long long)d->n_members);
    return WBX_ERR_UNSUPPORTED;
  }
  WBX_CUDA(cudaSetDevice(ctx->device));
  const size_t wbytes = (p->wy.size() + p->wx.size()) * sizeof(double);
  if (wbytes) {
    int rc = p->weights.reserve(wbytes);
    if (rc != WBX_OK) { delete p; return rc; }
    double* base = p->weights.as<double>();
    if (p->has_wy) {
      WBX_CUDA(cudaMemcpyAsync(base, p->wy.data(), p->wy.size() * 8,
                               cudaMemcpyHostToDevice, ctx->stream));
      p->d_wy = base;
    }
    if (p->has_wx) {
      WBX_CUDA(cudaMemcpyAsync(base + p->wy.size(), p->wx.data(),
                               p->wx.size() * 8, cudaMemcpyHostToDevice,
                               ctx->stream));
      p->d_wx = base + p->wy.size();
    }
  }
  if (d->space == WBX_SPACE_DEVICE) {
    int n_cells = 0;
    int rc = wbx::crps_upload_tables(
        ctx, p, 0, p->n_jobs, p->ens.data(), p->target.data(),
        p->has_mask ? p->mask.data() : nullptr, &p->chunk_host[0], &p->tables,
        ctx->stream, &p->params, &p->d_cell_first_job, &n_cells);
    if (rc != WBX_OK) { delete p; return rc; }
    p->params.member_stride = p->member_stride;
    p->params.point_stride = p->point_stride;
    p->grid = wbx::crps_grid(ctx, p, p->params.total_tiles);
  }
  WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = p;
  return WBX_OK;
}

int wbx_crps_plan_destroy(wbx_ctx* ctx, wbx_crps_plan* plan) {
  if (!plan) return WBX_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
  }
  plan->tables.release_idle();
  plan->weights.release_idle();
  delete plan;
  return WBX_OK;
}

static int crps_run_device(wbx_ctx* ctx, wbx_crps_plan* plan, double* d_ws,
                           double* d_w, int accumulate) {
  const size_t rec = (static_cast<size_t>(plan->grid) + plan->n_cells) *
                     wbx::kCrpsWarps * wbx::kCrpsAcc * sizeof(double);
  int rc = ctx->records.reserve(rec);
  if (rc != WBX_OK) return rc;
  wbx::CrpsParams P = plan->params;
  P.records = ctx->records.as<double>();
  for (int k = 0; k < wbx::kCrpsStats; ++k) P.fields[k] = plan->fields[k];
  rc = wbx::crps_launch(ctx, plan, P, plan->grid);
  if (rc != WBX_OK) return rc;
  return wbx::crps_finalize(ctx, plan, P.records, plan->d_cell_first_job,
                            static_cast<int>(plan->n_cells), plan->grid,
                            P.total_tiles, d_ws, d_w, accumulate);
}

static int crps_run_host(wbx_ctx* ctx, wbx_crps_plan* plan, double* d_ws,
                         double* d_w) {
  const size_t slab = static_cast<size_t>(plan->ny * plan->nx);
  const size_t M = static_cast<size_t>(plan->n_members);
  const size_t ebytes = slab * M * 4, tbytes = slab * 4;
  const size_t mbytes = wbx::rup(slab, 16);
  const size_t job_bytes = ebytes + tbytes + (plan->has_mask ? mbytes : 0);
  int64_t per_chunk =
      static_cast<int64_t>((ctx->staging_bytes / 2) / job_bytes);
  per_chunk = std::max<int64_t>(1, std::min<int64_t>(per_chunk, plan->n_jobs));
  for (int b = 0; b < 2; ++b) {
    int rc = ctx->staging[b].reserve(static_cast<size_t>(per_chunk) * job_bytes);
    if (rc != WBX_OK) return rc;
  }
  WBX_CUDA(cudaMemsetAsync(d_ws, 0, plan->n_cells * wbx::kCrpsStats * sizeof(double),
                           ctx->stream));
  WBX_CUDA(cudaMemsetAsync(d_w, 0, plan->n_cells * wbx::kCrpsStats * sizeof(double),
                           ctx->stream));
  const bool member_major = plan->point_stride == 1;
  int buf = 0;
  std::vector<uint64_t> a_ens, a_tgt, a_mask;
  for (int64_t j0 = 0; j0 < plan->n_jobs; j0 += per_chunk, buf ^= 1) {
    const int64_t j1 = std::min<int64_t>(plan->n_jobs, j0 + per_chunk);
    const size_t nj = static_cast<size_t>(j1 - j0);
    unsigned char* sbase = ctx->staging[buf].as<unsigned char>();
    unsigned char* s_ens = sbase;
    unsigned char* s_tgt = s_ens + nj * ebytes;
    unsigned char* s_mask = s_tgt + nj * tbytes;
    WBX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_compute[buf], 0));
    a_ens.resize(nj);
    a_tgt.resize(nj);
    a_mask.resize(nj);
    for (size_t j = 0; j < nj; ++j) {
      const void* src = reinterpret_cast<const void*>(plan->ens[j0 + j]);
      unsigned char* dst = s_ens + j * ebytes;
      if (member_major && static_cast<size_t>(plan->member_stride) != slab) {
        WBX_CUDA(cudaMemcpy2DAsync(dst, slab * 4, src, plan->member_stride * 4,
                                   slab * 4, M, cudaMemcpyHostToDevice,
                                   ctx->copy_stream));
      } else {
        WBX_CUDA(cudaMemcpyAsync(dst, src, ebytes, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
      }
      WBX_CUDA(cudaMemcpyAsync(
          s_tgt + j * tbytes, reinterpret_cast<const void*>(plan->target[j0 + j]),
          tbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      if (plan->has_mask)
        WBX_CUDA(cudaMemcpyAsync(
            s_mask + j * mbytes,
            reinterpret_cast<const void*>(plan->mask[j0 + j]), slab,
            cudaMemcpyHostToDevice, ctx->copy_stream));
      a_ens[j] = reinterpret_cast<uint64_t>(dst);
      a_tgt[j] = reinterpret_cast<uint64_t>(s_tgt + j * tbytes);
      a_mask[j] = reinterpret_cast<uint64_t>(s_mask + j * mbytes);
    }
    wbx::CrpsParams P{};
    const int32_t* d_first = nullptr;
    int n_cells = 0;
    int rc = wbx::crps_upload_tables(
        ctx, plan, j0, j1, a_ens.data(), a_tgt.data(),
        plan->has_mask ? a_mask.data() : nullptr, &plan->chunk_host[buf],
        &ctx->stage_tables[buf], ctx->copy_stream, &P, &d_first, &n_cells);
    if (rc != WBX_OK) return rc;
    P.member_stride = member_major ? static_cast<long long>(slab) : 1;
    P.point_stride = member_major ? 1 : static_cast<long long>(M);
    for (int k = 0; k < wbx::kCrpsStats; ++k)
      P.fields[k] = plan->fields[k]
                        ? plan->fields[k] + static_cast<size_t>(j0) * slab
                        : nullptr;
    WBX_CUDA(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
    const int grid = wbx::crps_grid(ctx, plan, P.total_tiles);
    rc = ctx->records.reserve((static_cast<size_t>(grid) + n_cells) *
                              wbx::kCrpsWarps * wbx::kCrpsAcc * sizeof(double));
    if (rc != WBX_OK) return rc;
    P.records = ctx->records.as<double>();
    WBX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
    rc = wbx::crps_launch(ctx, plan, P, grid);
    if (rc != WBX_OK) return rc;
    rc = wbx::crps_finalize(ctx, plan, P.records, d_first, n_cells, grid,
                            P.total_tiles,
                            d_ws + static_cast<size_t>(P.cell_base) *
                                       wbx::kCrpsStats,
                            d_w + static_cast<size_t>(P.cell_base) *
                                      wbx::kCrpsStats,
                            1);
    if (rc != WBX_OK) return rc;
    WBX_CUDA(cudaEventRecord(ctx->ev_compute[buf], ctx->stream));
  }
  return WBX_OK;
}

int wbx_crps_plan_run(wbx_ctx* ctx, wbx_crps_plan* plan, double* sum_ws,
                      double* sum_w, int32_t out_space, int32_t accumulate) {
  WBX_REQUIRE(ctx && plan && sum_ws && sum_w, "wbx_crps_plan_run: NULL argument");
  WBX_REQUIRE(out_space == WBX_SPACE_DEVICE || out_space == WBX_SPACE_HOST,
              "wbx_crps_plan_run: bad out_space");
  WBX_REQUIRE(!(accumulate && (out_space == WBX_SPACE_HOST ||
                               plan->space == WBX_SPACE_HOST)),
              "wbx_crps_plan_run: accumulate needs device inputs and outputs");
  WBX_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = plan->n_cells * wbx::kCrpsStats * sizeof(double);
  double* d_ws = sum_ws;
  double* d_w = sum_w;
  if (out_space == WBX_SPACE_HOST) {
    int rc = ctx->out_ws.reserve(bytes);
    if (rc != WBX_OK) return rc;
    rc = ctx->out_w.reserve(bytes);
    if (rc != WBX_OK) return rc;
    d_ws = ctx->out_ws.as<double>();
    d_w = ctx->out_w.as<double>();
  }
  int rc = plan->space == WBX_SPACE_DEVICE
               ? crps_run_device(ctx, plan, d_ws, d_w, accumulate)
               : crps_run_host(ctx, plan, d_ws, d_w);
  if (rc != WBX_OK) return rc;
  if (out_space == WBX_SPACE_HOST) {
    WBX_CUDA(cudaMemcpyAsync(sum_ws, d_ws, bytes, cudaMemcpyDeviceToHost,
                             ctx->stream));
    WBX_CUDA(cudaMemcpyAsync(sum_w, d_w, bytes, cudaMemcpyDeviceToHost,
                             ctx->stream));
    WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return WBX_OK;
}

int wbx_crps_plan_run_fields(wbx_ctx* ctx, wbx_crps_plan* plan, double* sum_ws,
                             double* sum_w, int32_t out_space,
                             float* const* fields) {
  WBX_REQUIRE(ctx && plan && fields, "wbx_crps_plan_run_fields: NULL argument");
  for (int k = 0; k < wbx::kCrpsStats; ++k) plan->fields[k] = fields[k];
  const int rc = wbx_crps_plan_run(ctx, plan, sum_ws, sum_w, out_space, 0);
  for (int k = 0; k < wbx::kCrpsStats; ++k) plan->fields[k] = nullptr;
  return rc;
}

int wbx_crps_pointwise(wbx_ctx* ctx, const wbx_crps_point_desc* d,
                       float* skill, float* spread) {
  WBX_REQUIRE(ctx && d, "wbx_crps_pointwise: NULL argument");
  WBX_REQUIRE(d->ndim >= 0 && d->ndim <= WBX_MAX_DIMS, "crps: bad ndim");
  WBX_REQUIRE(d->ens && d->target, "crps: NULL operand");
  WBX_REQUIRE(skill || spread || d->variance || d->unbiased_mse,
              "crps: nothing to compute");
  const bool ens_skipna = (d->flags & WBX_CRPS_SKIPNA_ENSEMBLE) != 0;
  if (!ens_skipna && d->n_members < 2 && spread) {
    wbx::set_error("Cannot estimate CRPS spread with n_ensemble < 2.");
    return WBX_ERR_INVALID;
  }
  WBX_REQUIRE(d->n_members >= 1 && d->n_members <= 1024,
              "crps: n_members out of range");
  wbx::CrpsPointParams P;
  memset(&P, 0, sizeof(P));
  P.n_points = 1;
  for (int i = 0; i < d->ndim; ++i) {
    WBX_REQUIRE(d->size[i] >= 1, "crps: empty dim");
    P.size[i] = d->size[i];
    P.e_stride[i] = d->ens_stride[i];
    P.t_stride[i] = d->target_stride[i];
    P.n_points *= d->size[i];
  }
  P.ens = d->ens;
  P.target = d->target;
  P.member_stride = d->member_stride;
  P.ndim = d->ndim;
  P.n_members = static_cast<int>(d->n_members);
  P.fair = (d->flags & WBX_CRPS_FAIR) ? 1 : 0;
  P.skill = skill;
  P.spread = spread;
  P.variance = d->variance;
  P.umse = d->unbiased_mse;
  const size_t smem = static_cast<size_t>(d->n_members) * wbx::kCrpsPitch * 4;
  WBX_REQUIRE(smem <= std::min<size_t>(ctx->smem_optin, 227 * 1024),
              "crps: too many members for shared memory");
  WBX_CUDA(cudaSetDevice(ctx->device));
  const long long blocks = (P.n_points + wbx::kCrpsThreads - 1) / wbx::kCrpsThreads;
  WBX_REQUIRE(blocks < (1ll << 31), "crps: too many points");
  if (ens_skipna) {
    auto kern = wbx::crps_pointwise_kernel<true>;
    WBX_CUDA(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<static_cast<unsigned>(blocks), wbx::kCrpsThreads, smem,
           ctx->stream>>>(P);
  } else {
    auto kern = wbx::crps_pointwise_kernel<false>;
    WBX_CUDA(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<static_cast<unsigned>(blocks), wbx::kCrpsThreads, smem,
           ctx->stream>>>(P);
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

int wbx_ensemble_mean(wbx_ctx* ctx, const wbx_crps_point_desc* d, float* out) {
  WBX_REQUIRE(ctx && d && out, "wbx_ensemble_mean: NULL argument");
  WBX_REQUIRE(d->ndim >= 0 && d->ndim <= WBX_MAX_DIMS,
              "wbx_ensemble_mean: bad ndim");
  WBX_REQUIRE(d->ens, "wbx_ensemble_mean: NULL operand");
  WBX_REQUIRE(d->n_members >= 1 && d->n_members < (1ll << 31),
              "wbx_ensemble_mean: n_members out of range");
  wbx::CrpsPointParams P;
  memset(&P, 0, sizeof(P));
  P.n_points = 1;
  for (int i = 0; i < d->ndim; ++i) {
    WBX_REQUIRE(d->size[i] >= 1, "wbx_ensemble_mean: empty dim");
    P.size[i] = d->size[i];
    P.e_stride[i] = d->ens_stride[i];
    P.n_points *= d->size[i];
  }
  P.ens = d->ens;
  P.member_stride = d->member_stride;
  P.ndim = d->ndim;
  P.n_members = static_cast<int>(d->n_members);
  // 4 points per thread when the innermost dim is dense and every row /
  // member start stays 16-byte aligned
  bool vec4 = d->ndim >= 1 && d->ens_stride[d->ndim - 1] == 1 &&
              d->size[d->ndim - 1] % 4 == 0 && d->member_stride % 4 == 0 &&
              (reinterpret_cast<uintptr_t>(d->ens) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int i = 0; vec4 && i + 1 < d->ndim; ++i)
    vec4 = d->ens_stride[i] % 4 == 0;
  const bool ens_skipna = (d->flags & WBX_CRPS_SKIPNA_ENSEMBLE) != 0;
  WBX_CUDA(cudaSetDevice(ctx->device));
  const long long threads = vec4 ? P.n_points / 4 : P.n_points;
  const long long blocks = (threads + 255) / 256;
  WBX_REQUIRE(blocks < (1ll << 31), "wbx_ensemble_mean: too many points");
  const unsigned grid = static_cast<unsigned>(blocks);
  if (vec4 && ens_skipna)
    wbx::ensemble_mean_kernel<true, 4><<<grid, 256, 0, ctx->stream>>>(P, out);
  else if (vec4)
    wbx::ensemble_mean_kernel<false, 4><<<grid, 256, 0, ctx->stream>>>(P, out);
  else if (ens_skipna)
    wbx::ensemble_mean_kernel<true, 1><<<grid, 256, 0, ctx->stream>>>(P, out);
  else
    wbx::ensemble_mean_kernel<false, 1><<<grid, 256, 0, ctx->stream>>>(P, out);
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

}  // extern "C"
