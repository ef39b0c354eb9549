// Device code of the fused deterministic-statistics + weighted-reduction path.
//
// One pass over predictions / targets [/ climatology] [/ mask] computes
// Error, AbsoluteError, SquaredError (metrics/deterministic.py:94-123) and,
// with a climatology, SquaredPredictionAnomaly, SquaredTargetAnomaly,
// AnomalyCovariance (deterministic.py:225-259), multiplies by the separable
// weights and reduces them per output cell -- i.e. the whole of
// Aggregator.aggregate_stat_var (aggregation.py:337-366) without ever writing
// a per-gridpoint temporary.
//
// Kernel shape (HBM-bound, 8 or 12 algorithmic bytes per grid point):
//   * persistent grid, one CTA per SM; CTA b owns the contiguous tile range
//     [b*T/G, (b+1)*T/G) of the job-major tile list, so a CTA sees the output
//     cells in non-decreasing order and flushes its accumulators only when the
//     cell changes;
//   * TMA variant: one producer thread streams tiles of every operand into a
//     shared-memory ring with cp.async.bulk (UBLKCP) + mbarrier transaction
//     counts; 16 consumer warps read the ring with 128-bit LDS, evaluate the
//     statistics in f32 exactly as NumPy does, and accumulate weight * value
//     in f64 registers;
//   * LDG variant (unaligned or tiny inputs, and an A/B baseline): same
//     accumulation code fed by ld.global.nc streaming loads;
//   * every (CTA, cell, warp) writes one partial record; a second tiny kernel
//     sums the records of a cell in a fixed order => bit-stable results, no
//     atomics.
#pragma once

#include "common.cuh"

namespace wbx {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kTmaThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kLdgThreads = 256;
constexpr int kLdgWarps = kLdgThreads / 32;
constexpr int kMaxStages = 12;

struct DetParams {
  const uint64_t* pred;
  const uint64_t* target;
  const uint64_t* clim;
  const uint64_t* mask;
  const int32_t* cell;
  const double* w_outer;
  const double* w_y;
  const double* w_x;
  const float* w_xf;   // w_x rounded to f32 (16-byte aligned), or NULL
  long long n_jobs;
  long long total_tiles;
  int cell_base;
  int ny, nx;
  int slab;            // ny * nx
  int tile;            // elements per tile (multiple of 16)
  int tiles_per_slab;
  // (quotient, remainder) of the element strides a consumer thread advances
  // by, divided by nx: between its float4 groups inside a tile and from one
  // tile to the next.  Lets the kernel track (row, column) without dividing.
  int group_dq, group_dr;
  int tile_dq, tile_dr;
  int stat_mask;       // bit s set => statistic slot s is wanted
  double* records;
  // XF launches only (categorical transform of the operands, wbx_det_desc.xform)
  const float* thr_pred;    // [n_jobs] or NULL
  const float* thr_target;  // [n_jobs] or NULL
  int xf_kind;              // WBX_XF_* request
};

// Thresholds of the job a thread is working on (XF launches).
struct XfArgs {
  float thr_p = 0.f, thr_t = 0.f;
  int kind = 0;
};

__device__ __forceinline__ XfArgs xf_for_job(const DetParams& P, long long job) {
  XfArgs xf;
  xf.kind = P.xf_kind;
  xf.thr_p = P.thr_pred ? __ldg(P.thr_pred + job) : 0.f;
  xf.thr_t = P.thr_target ? __ldg(P.thr_target + job) : 0.f;
  return xf;
}

struct StageMeta {
  int cell;
  int len;
  int e0;
  int pad;
  double wo;
};

template <bool CLIM, bool MASK, bool SKIPNA, bool XF = false>
struct AccLayout {
  static_assert(!(CLIM && XF), "XF launches take no climatology");
  static constexpr int kStats = XF ? WBX_NUM_XF_STATS : (CLIM ? 6 : 3);
  static constexpr int kWeights = SKIPNA ? (CLIM ? 4 : 1) : (MASK ? 1 : 0);
  static constexpr int kAcc = kStats + kWeights;
};

// Statistic values of one grid point, f32, rounded after every operation
// exactly like the NumPy ufunc chain of the reference.
template <bool CLIM, bool MASK, bool SKIPNA, bool XF = false>
struct PointStats {
  float s[AccLayout<CLIM, MASK, SKIPNA, XF>::kStats];
  float valid[AccLayout<CLIM, MASK, SKIPNA, XF>::kWeights > 0
                  ? AccLayout<CLIM, MASK, SKIPNA, XF>::kWeights
                  : 1];

  // Categorical statistics of the thresholded operands (wbx_det_desc.xform):
  // contingency table entries as 0/1 values (categorical.py:25-101 applied to
  // wrappers.binarize_thresholds, wrappers.py:88) or the error exceedance
  // indicator (deterministic.py:285-295); NaN inputs give NaN in every slot.
  __device__ __forceinline__ void eval_xf(float p, float t, unsigned char m,
                                          const XfArgs& xf) {
    bool ok;
    if ((xf.kind & 3) == WBX_XF_CONTINGENCY) {
      const bool bp =
          (xf.kind & WBX_XF_PRED_NONZERO) ? (p != 0.f) : (p > xf.thr_p);
      const bool bt =
          (xf.kind & WBX_XF_TARGET_NONZERO) ? (t != 0.f) : (t > xf.thr_t);
      ok = (p == p) && (t == t);
      s[WBX_XF_TRUE_POSITIVES] = (bp && bt) ? 1.f : 0.f;
      s[WBX_XF_FALSE_POSITIVES] = (bp && !bt) ? 1.f : 0.f;
      s[WBX_XF_FALSE_NEGATIVES] = (!bp && bt) ? 1.f : 0.f;
      s[WBX_XF_TRUE_NEGATIVES] = (!bp && !bt) ? 1.f : 0.f;
    } else {
      const float d = fabsf(__fsub_rn(p, t));
      ok = (d == d) && (xf.thr_p == xf.thr_p);
      s[0] = (d > xf.thr_p) ? 1.f : 0.f;
      s[1] = 0.f;
      s[2] = 0.f;
      s[3] = 0.f;
    }
    const bool base = MASK ? (m != 0) : true;
    if constexpr (SKIPNA) {
      const bool good = base && ok;
#pragma unroll
      for (int k = 0; k < WBX_NUM_XF_STATS; ++k) s[k] = good ? s[k] : 0.f;
      valid[0] = good ? 1.f : 0.f;
    } else {
      const float fill = base ? __int_as_float(0x7fc00000) : 0.f;
#pragma unroll
      for (int k = 0; k < WBX_NUM_XF_STATS; ++k)
        s[k] = (base && ok) ? s[k] : fill;
      valid[0] = base ? 1.f : 0.f;
    }
  }

  __device__ __forceinline__ void eval(float p, float t, float c,
                                       unsigned char m,
                                       const XfArgs& xf = XfArgs()) {
    if constexpr (XF) {
      eval_xf(p, t, m, xf);
      return;
    }
    const float d = __fsub_rn(p, t);
    s[0] = d;
    s[1] = fabsf(d);
    s[2] = __fmul_rn(d, d);
    float a = 0.f, b = 0.f;
    if constexpr (CLIM) {
      a = __fsub_rn(p, c);
      b = __fsub_rn(t, c);
      s[3] = __fmul_rn(a, a);
      s[4] = __fmul_rn(b, b);
      s[5] = __fmul_rn(a, b);
    }
    if constexpr (MASK || SKIPNA) {
      const bool base = MASK ? (m != 0) : true;
      if constexpr (SKIPNA) {
        const bool ok0 = base && (d == d);
        s[0] = ok0 ? s[0] : 0.f;
        s[1] = ok0 ? s[1] : 0.f;
        s[2] = ok0 ? s[2] : 0.f;
        valid[0] = ok0 ? 1.f : 0.f;
        if constexpr (CLIM) {
          const bool ok1 = base && (a == a);
          const bool ok2 = base && (b == b);
          const bool ok3 = base && (s[5] == s[5]);
          s[3] = ok1 ? s[3] : 0.f;
          s[4] = ok2 ? s[4] : 0.f;
          s[5] = ok3 ? s[5] : 0.f;
          valid[1] = ok1 ? 1.f : 0.f;
          valid[2] = ok2 ? 1.f : 0.f;
          valid[3] = ok3 ? 1.f : 0.f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < (CLIM ? 6 : 3); ++k) s[k] = base ? s[k] : 0.f;
        valid[0] = base ? 1.f : 0.f;
      }
    }
  }
};

// Four consecutive points of one latitude row (weight w is row-uniform): sum
// the four statistic values in f32, then one f64 FMA per statistic.
template <bool CLIM, bool MASK, bool SKIPNA, bool XF = false>
__device__ __forceinline__ void accum_row4(const float4 p, const float4 t,
                                           const float4 c, const uchar4 m,
                                           const double w, const int stat_mask,
                                           double* acc,
                                           const XfArgs& xf = XfArgs()) {
  using L = AccLayout<CLIM, MASK, SKIPNA, XF>;
  PointStats<CLIM, MASK, SKIPNA, XF> q0, q1, q2, q3;
  q0.eval(p.x, t.x, c.x, m.x, xf);
  q1.eval(p.y, t.y, c.y, m.y, xf);
  q2.eval(p.z, t.z, c.z, m.z, xf);
  q3.eval(p.w, t.w, c.w, m.w, xf);
#pragma unroll
  for (int k = 0; k < L::kStats; ++k) {
    if (stat_mask & (1 << k)) {  // warp-uniform
      const float s4 = __fadd_rn(__fadd_rn(q0.s[k], q1.s[k]),
                                 __fadd_rn(q2.s[k], q3.s[k]));
      acc[k] += static_cast<double>(s4) * w;
    }
  }
#pragma unroll
  for (int k = 0; k < L::kWeights; ++k) {
    const float n4 = (q0.valid[k] + q1.valid[k]) + (q2.valid[k] + q3.valid[k]);
    acc[L::kStats + k] += static_cast<double>(n4) * w;
  }
}

// Four consecutive points of one row whose weight varies along the row
// (w_x: the latitude axis is the fastest one, as in the lon-major WeatherBench
// archives).  The statistic values are f32 roundings; they are multiplied by
// the column weights rounded to f32 (one 16-byte load, one more rounding of
// the same size per term), summed in f32 like the row-uniform case and folded
// with the f64 row weight by one FMA: one f32 -> f64 conversion per group and
// statistic instead of four (the conversion is the slow instruction here, 8
// cycles per warp on the XU pipe).
template <bool CLIM, bool MASK, bool SKIPNA, bool ALIGNED, bool XF = false>
__device__ __forceinline__ void accum_row4_wx(const float4 p, const float4 t,
                                              const float4 c, const uchar4 m,
                                              const double wrow,
                                              const float* __restrict__ wx4,
                                              const double* __restrict__ wx4d,
                                              const int stat_mask,
                                              double* acc,
                                              const XfArgs& xf = XfArgs()) {
  using L = AccLayout<CLIM, MASK, SKIPNA, XF>;
  PointStats<CLIM, MASK, SKIPNA, XF> q0, q1, q2, q3;
  q0.eval(p.x, t.x, c.x, m.x, xf);
  q1.eval(p.y, t.y, c.y, m.y, xf);
  q2.eval(p.z, t.z, c.z, m.z, xf);
  q3.eval(p.w, t.w, c.w, m.w, xf);
  float4 wv;
  if constexpr (ALIGNED) {
    wv = __ldg(reinterpret_cast<const float4*>(wx4));
  } else {  // odd row lengths: the group starts at any column
    wv = make_float4(__ldg(wx4), __ldg(wx4 + 1), __ldg(wx4 + 2), __ldg(wx4 + 3));
  }
  if constexpr (XF) {
    // 0/1 indicator statistics: exact f64 weights keep "the table entries
    // add up to the sum of weights" an identity
    const double w4[4] = {__ldg(wx4d), __ldg(wx4d + 1), __ldg(wx4d + 2),
                          __ldg(wx4d + 3)};
#pragma unroll
    for (int k = 0; k < L::kStats; ++k) {
      if (stat_mask & (1 << k)) {  // warp-uniform
        double s4 = static_cast<double>(q3.s[k]) * w4[3];
        s4 = fma(static_cast<double>(q2.s[k]), w4[2], s4);
        s4 = fma(static_cast<double>(q1.s[k]), w4[1], s4);
        s4 = fma(static_cast<double>(q0.s[k]), w4[0], s4);
        acc[k] = fma(s4, wrow, acc[k]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < L::kStats; ++k) {
      if (stat_mask & (1 << k)) {  // warp-uniform
        const float s4 = __fadd_rn(
            __fadd_rn(__fmul_rn(q0.s[k], wv.x), __fmul_rn(q1.s[k], wv.y)),
            __fadd_rn(__fmul_rn(q2.s[k], wv.z), __fmul_rn(q3.s[k], wv.w)));
        acc[k] = fma(static_cast<double>(s4), wrow, acc[k]);
      }
    }
  }
  // the sums of weights of masked / skipna aggregations stay in f64 (they are
  // the denominators; unmasked launches have none of these)
  if constexpr (L::kWeights > 0) {
    double2 wa, wb;
    if constexpr (ALIGNED) {
      wa = __ldg(reinterpret_cast<const double2*>(wx4d));
      wb = __ldg(reinterpret_cast<const double2*>(wx4d) + 1);
    } else {
      wa = make_double2(__ldg(wx4d), __ldg(wx4d + 1));
      wb = make_double2(__ldg(wx4d + 2), __ldg(wx4d + 3));
    }
#pragma unroll
    for (int k = 0; k < L::kWeights; ++k) {
      double n4 = static_cast<double>(q3.valid[k]) * wb.y;
      n4 = fma(static_cast<double>(q2.valid[k]), wb.x, n4);
      n4 = fma(static_cast<double>(q1.valid[k]), wa.y, n4);
      n4 = fma(static_cast<double>(q0.valid[k]), wa.x, n4);
      acc[L::kStats + k] = fma(n4, wrow, acc[L::kStats + k]);
    }
  }
}

template <bool CLIM, bool MASK, bool SKIPNA, bool XF = false>
__device__ __forceinline__ void accum_point(float p, float t, float c,
                                            unsigned char m, const double w,
                                            const int stat_mask, double* acc,
                                            const XfArgs& xf = XfArgs()) {
  using L = AccLayout<CLIM, MASK, SKIPNA, XF>;
  PointStats<CLIM, MASK, SKIPNA, XF> q;
  q.eval(p, t, c, m, xf);
#pragma unroll
  for (int k = 0; k < L::kStats; ++k)
    if (stat_mask & (1 << k)) acc[k] += static_cast<double>(q.s[k]) * w;
#pragma unroll
  for (int k = 0; k < L::kWeights; ++k)
    acc[L::kStats + k] += static_cast<double>(q.valid[k]) * w;
}

// Weight of element e of a slab: wo * w_y[y] * w_x[x].
struct WeightCursor {
  const double* wy;
  const double* wx;
  const float* wxf;   // w_x in f32 (16-byte aligned table)
  int nx;
  bool wx4_ok;  // w_x present, nx % 4 == 0 (both tables 16-byte aligned)
  __device__ __forceinline__ double row(unsigned y, double wo) const {
    return wy ? wo * __ldg(wy + y) : wo;
  }
  __device__ __forceinline__ double col(unsigned x) const {
    return wx ? __ldg(wx + x) : 1.0;
  }
};

// One float4 group starting at slab element e.  PER_ELEM handles w_x and rows
// whose length is not a multiple of four (the group may straddle rows).
template <bool CLIM, bool MASK, bool SKIPNA, bool PER_ELEM, bool XF = false>
__device__ __forceinline__ void accum_group4(const float4 p, const float4 t,
                                             const float4 c, const uchar4 m,
                                             unsigned y, unsigned x, double wo,
                                             const WeightCursor& wc,
                                             const int stat_mask,
                                             double* acc,
                                             const XfArgs& xf = XfArgs()) {
  if constexpr (!PER_ELEM) {
    accum_row4<CLIM, MASK, SKIPNA, XF>(p, t, c, m, wc.row(y, wo), stat_mask,
                                       acc, xf);
  } else if (wc.wx4_ok) {
    // rows are a multiple of four long: the group stays in its row
    accum_row4_wx<CLIM, MASK, SKIPNA, true, XF>(p, t, c, m, wc.row(y, wo),
                                                wc.wxf + x, wc.wx + x,
                                                stat_mask, acc, xf);
  } else if (wc.wx != nullptr && x + 3u < static_cast<unsigned>(wc.nx)) {
    // odd row length (e.g. 721 latitudes): all but the groups that straddle
    // a row end still share one row weight
    accum_row4_wx<CLIM, MASK, SKIPNA, false, XF>(p, t, c, m, wc.row(y, wo),
                                                 wc.wxf + x, wc.wx + x,
                                                 stat_mask, acc, xf);
  } else {
    const float pp[4] = {p.x, p.y, p.z, p.w};
    const float tt[4] = {t.x, t.y, t.z, t.w};
    const float cc[4] = {c.x, c.y, c.z, c.w};
    const unsigned char mm[4] = {m.x, m.y, m.z, m.w};
    double wrow = wc.row(y, wo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      accum_point<CLIM, MASK, SKIPNA, XF>(pp[i], tt[i], cc[i], mm[i],
                                          wrow * wc.col(x), stat_mask, acc,
                                          xf);
      if (++x == static_cast<unsigned>(wc.nx)) {
        x = 0;
        ++y;
        if (i < 3) wrow = wc.row(y, wo);
      }
    }
  }
}

template <int NACC>
__device__ __forceinline__ void flush_warp(double (&acc)[NACC], double* rec,
                                           int lane) {
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const double v = warp_sum(acc[a]);
    if (lane == 0) rec[a] = v;
    acc[a] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// TMA-ring kernel
// ---------------------------------------------------------------------------
template <bool CLIM, bool MASK, bool SKIPNA, bool PER_ELEM, bool XF = false>
__global__ void __launch_bounds__(kTmaThreads, 1)
    det_reduce_tma_kernel(const DetParams P, const int stages,
                          const int stage_bytes) {
  using L = AccLayout<CLIM, MASK, SKIPNA, XF>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + kMaxStages;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + kMaxStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long t_begin =
      (static_cast<long long>(blockIdx.x) * P.total_tiles) / gridDim.x;
  const long long t_end =
      (static_cast<long long>(blockIdx.x + 1) * P.total_tiles) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  // PDL: let the finalize kernel be scheduled behind us right away, and do not
  // write records before the previous finalize (their last reader) is done.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();

  const int off_t = P.tile * 4;
  const int off_c = P.tile * 8;
  const int off_m = P.tile * 4 * (CLIM ? 3 : 2);

  if (warp == kConsumerWarps) {
    // ---------------- producer: one elected thread drives the TMA engine ---
    if (lane == 0) {
      const uint64_t policy = l2_evict_first_policy();
      long long job = t_begin / P.tiles_per_slab;
      int k = static_cast<int>(t_begin - job * P.tiles_per_slab);
      long long loaded_job = -1;
      const float *pa = nullptr, *ta = nullptr, *ca = nullptr;
      const unsigned char* ma = nullptr;
      int cell = 0;
      double wo = 1.0;
      int s = 0;
      uint32_t ph = 0;
      for (long long g = t_begin; g < t_end; ++g) {
        if (job != loaded_job) {
          pa = reinterpret_cast<const float*>(__ldg(P.pred + job));
          ta = reinterpret_cast<const float*>(__ldg(P.target + job));
          if constexpr (CLIM)
            ca = reinterpret_cast<const float*>(__ldg(P.clim + job));
          if constexpr (MASK)
            ma = reinterpret_cast<const unsigned char*>(__ldg(P.mask + job));
          cell = __ldg(P.cell + job);
          wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
          loaded_job = job;
        }
        const int e0 = k * P.tile;
        const int len = min(P.tile, P.slab - e0);
        mbar_wait(&empty[s], ph ^ 1u);
        StageMeta mt;
        mt.cell = cell;
        mt.len = len;
        mt.e0 = e0;
        mt.pad = 0;
        if constexpr (XF) mt.pad = static_cast<int>(job);  // threshold lookup
        mt.wo = wo;
        meta[s] = mt;
        unsigned char* st = ring + (size_t)s * stage_bytes;
        const uint32_t fbytes = static_cast<uint32_t>(len) * 4u;
        const uint32_t total =
            fbytes * (CLIM ? 3u : 2u) + (MASK ? static_cast<uint32_t>(len) : 0u);
        mbar_expect_tx(&full[s], total);
        bulk_g2s(st, pa + e0, fbytes, &full[s], policy);
        bulk_g2s(st + off_t, ta + e0, fbytes, &full[s], policy);
        if constexpr (CLIM)
          bulk_g2s(st + off_c, ca + e0, fbytes, &full[s], policy);
        if constexpr (MASK)
          bulk_g2s(st + off_m, ma + e0, static_cast<uint32_t>(len), &full[s],
                   policy);
        if (++k == P.tiles_per_slab) {
          k = 0;
          ++job;
        }
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  // ------------------- consumers ------------------------------------------
  double acc[L::kAcc];
#pragma unroll
  for (int a = 0; a < L::kAcc; ++a) acc[a] = 0.0;
  const WeightCursor wc{P.w_y, P.w_x, P.w_xf, P.nx,
                        P.w_xf != nullptr && (P.nx & 3) == 0};
  int cur_cell = -1;
  const int ctid = threadIdx.x;  // 0 .. kConsumerThreads-1
  const unsigned unx = static_cast<unsigned>(P.nx);
  // (row, column) of this thread's first group in a tile that starts a slab.
  const unsigned y_first = static_cast<unsigned>(4 * ctid) / unx;
  const unsigned x_first = static_cast<unsigned>(4 * ctid) - y_first * unx;
  unsigned ty = 0, tx = 0;  // ... in the current tile
  int prev_e0 = -1;
  int s = 0;
  uint32_t ph = 0;
  for (long long g = t_begin; g < t_end; ++g) {
    mbar_wait(&full[s], ph);
    const StageMeta mt = meta[s];
    XfArgs xf;
    if constexpr (XF) xf = xf_for_job(P, mt.pad);
    if (mt.e0 == 0) {
      ty = y_first;
      tx = x_first;
    } else if (prev_e0 >= 0 && mt.e0 == prev_e0 + P.tile) {
      ty += P.tile_dq;
      tx += P.tile_dr;
      if (tx >= unx) {
        tx -= unx;
        ++ty;
      }
    } else {  // first tile of this CTA starts inside a slab
      const unsigned e = static_cast<unsigned>(mt.e0 + 4 * ctid);
      ty = e / unx;
      tx = e - ty * unx;
    }
    prev_e0 = mt.e0;
    if (mt.cell != cur_cell) {
      if (cur_cell >= 0) {
        double* rec = P.records +
                      ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                           kConsumerWarps + warp) * L::kAcc;
        flush_warp<L::kAcc>(acc, rec, lane);
      }
      cur_cell = mt.cell;
    }
    const unsigned char* st = ring + (size_t)s * stage_bytes;
    const float4* sp = reinterpret_cast<const float4*>(st);
    const float4* stt = reinterpret_cast<const float4*>(st + off_t);
    const float4* sc = reinterpret_cast<const float4*>(st + off_c);
    const uchar4* sm = reinterpret_cast<const uchar4*>(st + off_m);
    const int nvec = mt.len >> 2;
    unsigned gy = ty, gx = tx;
#pragma unroll 2
    for (int j = ctid; j < nvec; j += kConsumerThreads) {
      const float4 pv = sp[j];
      const float4 tv = stt[j];
      float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
      uchar4 mv = make_uchar4(1, 1, 1, 1);
      if constexpr (CLIM) cv = sc[j];
      if constexpr (MASK) mv = sm[j];
      accum_group4<CLIM, MASK, SKIPNA, PER_ELEM, XF>(
          pv, tv, cv, mv, gy, gx, mt.wo, wc, P.stat_mask, acc, xf);
      gy += P.group_dq;
      gx += P.group_dr;
      if (gx >= unx) {
        gx -= unx;
        ++gy;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == stages) {
      s = 0;
      ph ^= 1u;
    }
  }
  if (cur_cell >= 0) {
    double* rec = P.records +
                  ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                       kConsumerWarps + warp) * L::kAcc;
    flush_warp<L::kAcc>(acc, rec, lane);
  }
}

// ---------------------------------------------------------------------------
// LDG kernel: same accumulation, direct streaming loads.  VEC = 4 needs
// 16-byte aligned slabs with slab % 4 == 0; VEC = 1 handles anything.
// ---------------------------------------------------------------------------
template <bool CLIM, bool MASK, bool SKIPNA, bool PER_ELEM, int VEC,
          bool XF = false>
__global__ void __launch_bounds__(kLdgThreads)
    det_reduce_ldg_kernel(const DetParams P) {
  using L = AccLayout<CLIM, MASK, SKIPNA, XF>;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long t_begin =
      (static_cast<long long>(blockIdx.x) * P.total_tiles) / gridDim.x;
  const long long t_end =
      (static_cast<long long>(blockIdx.x + 1) * P.total_tiles) / gridDim.x;
  double acc[L::kAcc];
#pragma unroll
  for (int a = 0; a < L::kAcc; ++a) acc[a] = 0.0;
  const WeightCursor wc{P.w_y, P.w_x, P.w_xf, P.nx,
                        P.w_xf != nullptr && (P.nx & 3) == 0};
  int cur_cell = -1;
  long long job = t_begin / P.tiles_per_slab;
  int k = static_cast<int>(t_begin - job * P.tiles_per_slab);
  for (long long g = t_begin; g < t_end; ++g) {
    const float* pa = reinterpret_cast<const float*>(__ldg(P.pred + job));
    const float* ta = reinterpret_cast<const float*>(__ldg(P.target + job));
    const float* ca = nullptr;
    const unsigned char* ma = nullptr;
    if constexpr (CLIM) ca = reinterpret_cast<const float*>(__ldg(P.clim + job));
    if constexpr (MASK)
      ma = reinterpret_cast<const unsigned char*>(__ldg(P.mask + job));
    const int cell = __ldg(P.cell + job);
    const double wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
    XfArgs xf;
    if constexpr (XF) xf = xf_for_job(P, job);
    if (cell != cur_cell) {
      if (cur_cell >= 0) {
        double* rec = P.records +
                      ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                           kLdgWarps + warp) * L::kAcc;
        flush_warp<L::kAcc>(acc, rec, lane);
      }
      cur_cell = cell;
    }
    const int e0 = k * P.tile;
    const int len = min(P.tile, P.slab - e0);
    if constexpr (VEC == 4) {
      const int nvec = len >> 2;
      constexpr int U = 4;
      for (int j0 = threadIdx.x; j0 < nvec; j0 += kLdgThreads * U) {
        float4 pv[U], tv[U], cv[U];
        uchar4 mv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * kLdgThreads;
          if (j < nvec) {
            pv[u] = ldg_stream_f4(pa + e0 + 4 * j);
            tv[u] = ldg_stream_f4(ta + e0 + 4 * j);
            if constexpr (CLIM) cv[u] = ldg_stream_f4(ca + e0 + 4 * j);
            if constexpr (MASK)
              mv[u] = __ldg(reinterpret_cast<const uchar4*>(ma + e0) + j);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * kLdgThreads;
          if (j < nvec) {
            if constexpr (!CLIM) cv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (!MASK) mv[u] = make_uchar4(1, 1, 1, 1);
            const unsigned e = static_cast<unsigned>(e0 + 4 * j);
            const unsigned y = e / static_cast<unsigned>(P.nx);
            accum_group4<CLIM, MASK, SKIPNA, PER_ELEM, XF>(
                pv[u], tv[u], cv[u], mv[u], y,
                e - y * static_cast<unsigned>(P.nx), wo, wc, P.stat_mask, acc,
                xf);
          }
        }
      }
    } else {
      for (int j = threadIdx.x; j < len; j += kLdgThreads) {
        const unsigned e = static_cast<unsigned>(e0 + j);
        const unsigned y = e / static_cast<unsigned>(P.nx);
        const unsigned x = e - y * static_cast<unsigned>(P.nx);
        const float pv = ldg_stream_f1(pa + e);
        const float tv = ldg_stream_f1(ta + e);
        float cv = 0.f;
        unsigned char mv = 1;
        if constexpr (CLIM) cv = ldg_stream_f1(ca + e);
        if constexpr (MASK) mv = __ldg(ma + e);
        accum_point<CLIM, MASK, SKIPNA, XF>(pv, tv, cv, mv,
                                            wc.row(y, wo) * wc.col(x),
                                            P.stat_mask, acc, xf);
      }
    }
    if (++k == P.tiles_per_slab) {
      k = 0;
      ++job;
    }
  }
  if (cur_cell >= 0) {
    double* rec = P.records +
                  ((static_cast<size_t>(blockIdx.x) + (cur_cell - P.cell_base)) *
                       kLdgWarps + warp) * L::kAcc;
    flush_warp<L::kAcc>(acc, rec, lane);
  }
}

// ---------------------------------------------------------------------------
// Deterministic second pass: sum the records of each cell in CTA/warp order.
// ---------------------------------------------------------------------------
struct FinalizeParams {
  const double* records;
  const int32_t* cell_first_job;  // [n_cells + 1], relative to the launch
  const double* cell_w;           // [n_cells] constant sum_weights or NULL
  double* out_ws;                 // [*, 6]  (already offset to cell_base)
  double* out_w;                  // [*, 4]
  long long total_tiles;
  int n_cells;
  int grid_main;
  int tiles_per_slab;
  int warps;
  int n_stats;    // 3 or 6
  int n_weights;  // 0, 1 or 4
  int accumulate;
};

__global__ void __launch_bounds__(128) det_finalize_kernel(
    const FinalizeParams F) {
  // One warp per (cell, slot): lanes stride over the records of the cell in a
  // fixed assignment, then a butterfly sum -- parallel and still bit-stable.
  // PDL: scheduled early, starts once the reduction kernel has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int slots = WBX_NUM_DET_STATS + WBX_NUM_DET_WCLASSES;
  if (warp_global >= F.n_cells * slots) return;
  const int c = warp_global / slots;
  const int slot = warp_global - c * slots;
  const int nacc = F.n_stats + F.n_weights;
  int a = -1;
  double value = 0.0;
  bool constant = false;
  if (slot < WBX_NUM_DET_STATS) {
    if (slot < F.n_stats) a = slot;
  } else {
    const int k = slot - WBX_NUM_DET_STATS;
    if (F.n_weights == 0) {
      constant = true;
      value = F.cell_w[c];
    } else if (F.n_weights == 1) {
      a = F.n_stats;
    } else {
      a = F.n_stats + k;
    }
  }
  if (a >= 0 && !constant) {
    const long long ft =
        static_cast<long long>(F.cell_first_job[c]) * F.tiles_per_slab;
    const long long lt =
        static_cast<long long>(F.cell_first_job[c + 1]) * F.tiles_per_slab - 1;
    const long long G = F.grid_main;
    const int b_lo = static_cast<int>(((ft + 1) * G - 1) / F.total_tiles);
    const int b_hi = static_cast<int>(((lt + 1) * G - 1) / F.total_tiles);
    const int n = (b_hi - b_lo + 1) * F.warps;
    const double* rec =
        F.records + (static_cast<size_t>(b_lo) + c) * F.warps * nacc + a;
    // four independent partial sums keep several loads in flight; the
    // combination order is fixed, so the result stays bit-stable.
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = lane;
    for (; i + 96 < n; i += 128) {
      s0 += __ldcg(rec + static_cast<size_t>(i) * nacc);
      s1 += __ldcg(rec + static_cast<size_t>(i + 32) * nacc);
      s2 += __ldcg(rec + static_cast<size_t>(i + 64) * nacc);
      s3 += __ldcg(rec + static_cast<size_t>(i + 96) * nacc);
    }
    for (; i < n; i += 32) s0 += __ldcg(rec + static_cast<size_t>(i) * nacc);
    value = warp_sum((s0 + s1) + (s2 + s3));
  }
  if (lane == 0) {
    double* dst = slot < WBX_NUM_DET_STATS
                      ? F.out_ws + (size_t)c * WBX_NUM_DET_STATS + slot
                      : F.out_w + (size_t)c * WBX_NUM_DET_WCLASSES +
                            (slot - WBX_NUM_DET_STATS);
    *dst = F.accumulate ? (*dst + value) : value;
  }
}

// ---------------------------------------------------------------------------
// Per-gridpoint statistic values (materialised Statistic.compute output).
// ---------------------------------------------------------------------------
__global__ void det_elementwise_kernel(const int stat_and_flags,
                                       const float* __restrict__ p,
                                       const float* __restrict__ t,
                                       const float* __restrict__ c,
                                       const long long n, float* out) {
  const int stat = stat_and_flags & 0xff;
  const bool add = (stat_and_flags & WBX_EW_ACCUMULATE) != 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       i < n; i += stride) {
    const float pv = p[i], tv = t[i];
    float r;
    switch (stat) {
      case WBX_EW_PASS_PRED: r = pv; break;
      case WBX_EW_PASS_PRED_NAN_TARGET:
        r = (tv == tv) ? pv : __int_as_float(0x7fc00000);
        break;
      case WBX_STAT_ERROR: r = __fsub_rn(pv, tv); break;
      case WBX_STAT_ABS_ERROR: r = fabsf(__fsub_rn(pv, tv)); break;
      case WBX_STAT_SQ_ERROR: {
        const float d = __fsub_rn(pv, tv);
        r = __fmul_rn(d, d);
        break;
      }
      case WBX_STAT_SQ_PRED_ANOM: {
        const float a = __fsub_rn(pv, c[i]);
        r = __fmul_rn(a, a);
        break;
      }
      case WBX_STAT_SQ_TGT_ANOM: {
        const float b = __fsub_rn(tv, c[i]);
        r = __fmul_rn(b, b);
        break;
      }
      default: {
        const float cv = c[i];
        r = __fmul_rn(__fsub_rn(pv, cv), __fsub_rn(tv, cv));
        break;
      }
    }
    out[i] = add ? __fadd_rn(out[i], r) : r;
  }
}

// Per-gridpoint values of one categorical slot (scalar thresholds); the same
// PointStats code as the fused reduction evaluates them.
__global__ void xf_elementwise_kernel(const int xform, const int slot,
                                      const float thr_p, const float thr_t,
                                      const float* __restrict__ p,
                                      const float* __restrict__ t,
                                      const long long n, float* out) {
  XfArgs xf;
  xf.kind = xform;
  xf.thr_p = thr_p;
  xf.thr_t = thr_t;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       i < n; i += stride) {
    const float pv = p[i];
    float r;
    if (slot == WBX_XF_BINARIZED_PRED) {
      r = (pv == pv) ? (pv > thr_p ? 1.f : 0.f) : __int_as_float(0x7fc00000);
    } else {
      PointStats<false, false, false, true> q;
      q.eval(pv, t[i], 0.f, 1, xf);
      r = q.s[slot];
    }
    out[i] = r;
  }
}

}  // namespace wbx
