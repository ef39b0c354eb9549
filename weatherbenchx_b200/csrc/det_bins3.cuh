// Binned aggregation, third generation: a host-compiled reduction schedule.
//
// aggregation.py:320-335 of the reference multiplies the statistic by every
// bin mask inside xr.dot -- work that grows with the number of bins (34 in the
// public benchmark: 17 regions x {all, land}, run_benchmark_evaluation.py:
// 110-132,369).  The host folds all masks over the slab dims into one uint8
// class map (two grid points share a class iff they belong to the same set of
// bins); the kernel sums every statistic per (cell, class) in one pass and the
// host maps class sums to bin sums.
//
// The class map is an operand of the PLAN, not of the data: everything it
// decides can be decided once, on the host, when the plan is made.  The slab is
// cut into S parts of <= 4096 contiguous elements; CTA (s, g) streams part s of
// every job of job group g through the TMA ring.  For every part the host
// lists the part's SLOTS -- (quad of 4 contiguous elements, class, 4-bit
// element selection): one slot per quad whose elements share a class, one per
// class for the few quads a boundary runs through -- sorts them by class and
// deals them out to the 512 consumer threads x 2 passes.  Thirty-two
// consecutive slots (one warp, one pass) then hold a handful of class runs,
// the SEGMENTS, whose lane ranges are static too.
//
// The kernel is therefore the unbinned one plus static per-thread data:
//   * a thread reads the same quads for all jobs; weight (f64 row weight, or
//     four f32 element weights for longitude-major arrays / odd row lengths)
//     and selection live in registers;
//   * per job: f32 statistics, f32 4-sums, one f64 FMA per statistic into the
//     thread's accumulator -- no class logic, no branches that depend on data
//     or on the map (one warp-uniform static branch picks the code with the
//     element selection for warps that hold a boundary quad);
//   * when the output cell changes: a segmented inclusive warp scan (5 shuffle
//     steps, predicated on the static first lane of the segment) leaves every
//     segment's sum in its last lane, which writes it to the segment's record.
//     No shared-memory accumulators, no CTA barriers, every CTA does the same
//     work whatever the map looks like.
//   * the finalize kernel adds, per (cell, class), the records of the class's
//     segments (a host-built list) in a fixed order.
// No atomics, fixed summation orders => bit-stable results.  The class map is
// never read by the GPU at all: 8 / 12 B per point.
#pragma once

#include "det_reduce.cuh"

namespace wbx {

constexpr int kBins3Passes = 2;
constexpr int kBins3Slots = kBins3Passes * kConsumerThreads;  // per part

// slot descriptor bits: [9:0] quad | [13:10] selection | [18:14] first lane of
// the segment | [19] last lane of the segment | [31:20] segment of the part
__host__ __device__ __forceinline__ uint32_t bins3_pack(int quad, int sel,
                                                        int first, int last,
                                                        int seg) {
  return static_cast<uint32_t>(quad) | (static_cast<uint32_t>(sel) << 10) |
         (static_cast<uint32_t>(first) << 14) |
         (static_cast<uint32_t>(last) << 19) | (static_cast<uint32_t>(seg) << 20);
}

struct Bins3Params {
  const uint32_t* slot_desc;  // [S][passes][512]
  const void* slot_w;         // [S][passes][512] double, or float4 (WX)
  const int32_t* seg_base;    // [S + 1] first segment of every part
  int S;                      // slab parts
  int S_cta;                  // parts in flight; CTA (s, g) takes s, s + S_cta, ...
  int J;                      // job groups; grid = S_cta * J
  int part;                   // elements per part (multiple of 16, <= 4096)
  int total_segs;
  int n_cols;                 // record columns: selected statistics (+ weight)
  double* records;            // [n_cells + J][total_segs][n_cols]
};

template <bool CLIM, bool MASK, bool WX>
__global__ void __launch_bounds__(kTmaThreads, 1)
    det_reduce_bins3_kernel(const DetParams P, const Bins3Params B,
                            const int stages, const int stage_bytes) {
  constexpr int NS = CLIM ? 6 : 3;
  constexpr int NA = NS + (MASK ? 1 : 0);
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + kMaxStages;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + kMaxStages);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int s_cta = blockIdx.x % B.S_cta;
  const int grp = blockIdx.x / B.S_cta;
  const long long j_lo = (static_cast<long long>(grp) * P.n_jobs) / B.J;
  const long long j_hi = (static_cast<long long>(grp + 1) * P.n_jobs) / B.J;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int off_t = B.part * 4;
  const int off_c = B.part * 8;
  const int off_m = B.part * 4 * (CLIM ? 3 : 2);

  if (warp == kConsumerWarps) {
    // ---------------- producer ----------------------------------------------
    if (lane == 0) {
      const uint64_t policy = l2_evict_first_policy();
      int s = 0;
      uint32_t ph = 0;
      for (int s_part = s_cta; s_part < B.S; s_part += B.S_cta) {
        const int e0 = s_part * B.part;
        const int len = min(P.slab, e0 + B.part) - e0;
        const uint32_t fbytes = static_cast<uint32_t>(len) * 4u;
        // the tables of the next job are fetched while this one is issued
        long long job = j_lo;
        const float* pa = nullptr;
        const float* ta = nullptr;
        const float* ca = nullptr;
        const unsigned char* ma = nullptr;
        int cell = 0;
        double wo = 1.0;
        auto fetch = [&](long long j) {
          pa = reinterpret_cast<const float*>(__ldg(P.pred + j));
          ta = reinterpret_cast<const float*>(__ldg(P.target + j));
          if constexpr (CLIM)
            ca = reinterpret_cast<const float*>(__ldg(P.clim + j));
          if constexpr (MASK)
            ma = reinterpret_cast<const unsigned char*>(__ldg(P.mask + j));
          cell = __ldg(P.cell + j);
          wo = P.w_outer ? __ldg(P.w_outer + j) : 1.0;
        };
        if (job < j_hi) fetch(job);
        for (; job < j_hi; ++job) {
          const float* pj = pa;
          const float* tj = ta;
          const float* cj = ca;
          const unsigned char* mj = ma;
          StageMeta mt;
          mt.cell = cell;
          mt.len = len;
          mt.e0 = e0;
          mt.pad = 0;
          mt.wo = wo;
          if (job + 1 < j_hi) fetch(job + 1);
          mbar_wait(&empty[s], ph ^ 1u);
          meta[s] = mt;
          unsigned char* st = ring + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], fbytes * (CLIM ? 3u : 2u) +
                                       (MASK ? static_cast<uint32_t>(len) : 0u));
          bulk_g2s(st, pj + e0, fbytes, &full[s], policy);
          bulk_g2s(st + off_t, tj + e0, fbytes, &full[s], policy);
          if constexpr (CLIM)
            bulk_g2s(st + off_c, cj + e0, fbytes, &full[s], policy);
          if constexpr (MASK)
            bulk_g2s(st + off_m, mj + e0, static_cast<uint32_t>(len), &full[s],
                     policy);
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ------------------- consumers ---------------------------------------------
  const int ctid = threadIdx.x;
  const int stat_mask = P.stat_mask;
  double acc[kBins3Passes][NA];
#pragma unroll
  for (int ps = 0; ps < kBins3Passes; ++ps)
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[ps][a] = 0.0;

  int s = 0;
  uint32_t ph = 0;
  for (int s_part = s_cta; s_part < B.S; s_part += B.S_cta) {
    // ---- the static schedule of this part -----------------------------------
    int quad[kBins3Passes], seg_first[kBins3Passes], seg_rec[kBins3Passes];
    unsigned sel[kBins3Passes];
    bool partial[kBins3Passes], live[kBins3Passes];
    double w64[kBins3Passes];
    float4 w32[kBins3Passes];
    const int seg0 = __ldg(B.seg_base + s_part);
#pragma unroll
    for (int ps = 0; ps < kBins3Passes; ++ps) {
      const size_t at =
          (static_cast<size_t>(s_part) * kBins3Passes + ps) * kConsumerThreads +
          ctid;
      const uint32_t d = __ldg(B.slot_desc + at);
      quad[ps] = static_cast<int>(d & 0x3ffu);
      sel[ps] = (d >> 10) & 0xfu;
      seg_first[ps] = static_cast<int>((d >> 14) & 31u);
      // record of the segment this lane closes (-1: it closes none)
      seg_rec[ps] = ((d >> 19) & 1u) ? seg0 + static_cast<int>(d >> 20) : -1;
      if constexpr (WX) {
        w32[ps] = __ldg(reinterpret_cast<const float4*>(B.slot_w) + at);
        w64[ps] = 0.0;
      } else {
        w64[ps] = __ldg(reinterpret_cast<const double*>(B.slot_w) + at);
        w32[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      partial[ps] = __any_sync(kFull, sel[ps] != 0xfu);
      live[ps] = __any_sync(kFull, sel[ps] != 0u);
    }

    // sums of the segments of this warp -> their records, accumulators reset
    auto flush = [&](const int cell) {
      double* rec = B.records +
                    static_cast<size_t>(cell - P.cell_base + grp) *
                        B.total_segs * B.n_cols;
#pragma unroll
      for (int ps = 0; ps < kBins3Passes; ++ps) {
        if (!live[ps]) continue;   // warp-uniform
        int col = 0;
#pragma unroll
        for (int k = 0; k < NA; ++k) {
          if (k >= NS || (stat_mask & (1 << k))) {   // warp-uniform
            double v = acc[ps][k];
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
              const double up = __shfl_up_sync(kFull, v, dlt);
              if (lane - dlt >= seg_first[ps]) v += up;
            }
            if (seg_rec[ps] >= 0)
              rec[static_cast<size_t>(seg_rec[ps]) * B.n_cols + col] = v;
            acc[ps][k] = 0.0;
            ++col;
          }
        }
      }
    };

    int cur_cell = -1;
    for (long long job = j_lo; job < j_hi; ++job) {
      mbar_wait(&full[s], ph);
      const StageMeta mt = meta[s];
      if (mt.cell != cur_cell) {
        if (cur_cell >= 0) flush(cur_cell);
        cur_cell = mt.cell;
      }
      const unsigned char* stg = ring + (size_t)s * stage_bytes;
      const float4* sp = reinterpret_cast<const float4*>(stg);
      const float4* stt = reinterpret_cast<const float4*>(stg + off_t);
      const float4* sc = reinterpret_cast<const float4*>(stg + off_c);
      const uint32_t* sm = reinterpret_cast<const uint32_t*>(stg + off_m);
#pragma unroll
      for (int ps = 0; ps < kBins3Passes; ++ps) {
        if (!live[ps]) continue;   // warp-uniform
        const float4 pv = sp[quad[ps]];
        const float4 tv = stt[quad[ps]];
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t mw = 0x01010101u;
        if constexpr (CLIM) cv = sc[quad[ps]];
        if constexpr (MASK) mw = sm[quad[ps]];
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
        const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
        PointStats<CLIM, MASK, false> q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          q[i].eval(pp[i], tt[i], cc[i],
                    static_cast<unsigned char>(mw >> (8 * i)));
        if (partial[ps]) {
          // a boundary quad in this warp: elements of other classes (they
          // belong to other slots) become exact zeros, NaN included
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool mine = (sel[ps] >> i) & 1u;
#pragma unroll
            for (int k = 0; k < NS; ++k) q[i].s[k] = mine ? q[i].s[k] : 0.f;
            q[i].valid[0] = mine ? q[i].valid[0] : 0.f;
          }
        }
        if constexpr (!WX) {
          const double wj = w64[ps] * mt.wo;
#pragma unroll
          for (int k = 0; k < NS; ++k) {
            if (stat_mask & (1 << k)) {   // warp-uniform
              const float s4 = __fadd_rn(__fadd_rn(q[0].s[k], q[1].s[k]),
                                         __fadd_rn(q[2].s[k], q[3].s[k]));
              acc[ps][k] = fma(static_cast<double>(s4), wj, acc[ps][k]);
            }
          }
          if constexpr (MASK) {
            const float n4 = (q[0].valid[0] + q[1].valid[0]) +
                             (q[2].valid[0] + q[3].valid[0]);
            acc[ps][NS] = fma(static_cast<double>(n4), wj, acc[ps][NS]);
          }
        } else {
          // element weights (f32, rounded once on the host): weighted 4-sum in
          // f32, accumulated in f64
          const float wf[4] = {w32[ps].x, w32[ps].y, w32[ps].z, w32[ps].w};
#pragma unroll
          for (int k = 0; k < NS; ++k) {
            if (stat_mask & (1 << k)) {
              const float s4 = __fadd_rn(
                  __fadd_rn(__fmul_rn(q[0].s[k], wf[0]),
                            __fmul_rn(q[1].s[k], wf[1])),
                  __fadd_rn(__fmul_rn(q[2].s[k], wf[2]),
                            __fmul_rn(q[3].s[k], wf[3])));
              acc[ps][k] = fma(static_cast<double>(s4), mt.wo, acc[ps][k]);
            }
          }
          if constexpr (MASK) {
            const float n4 =
                __fadd_rn(__fadd_rn(__fmul_rn(q[0].valid[0], wf[0]),
                                    __fmul_rn(q[1].valid[0], wf[1])),
                          __fadd_rn(__fmul_rn(q[2].valid[0], wf[2]),
                                    __fmul_rn(q[3].valid[0], wf[3])));
            acc[ps][NS] = fma(static_cast<double>(n4), mt.wo, acc[ps][NS]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (cur_cell >= 0) flush(cur_cell);
  }
}

// records -> out[(cell * n_classes + class)][slot]: one warp per (cell, class,
// record column); lanes stride over the class's segments, groups in order.
struct Bins3FinalizeParams {
  const double* records;
  const int32_t* class_ptr;        // [n_classes + 1] into class_segs
  const int32_t* class_segs;       // segments of every class, ascending
  const int32_t* cell_first_job;   // [n_cells + 1], relative to the launch
  const double* cell_class_w;      // [n_cells * n_classes] or NULL (masked)
  double* out_ws;                  // [n_cells * n_classes * 6]
  double* out_w;                   // [n_cells * n_classes * 4]
  long long n_jobs;
  int n_cells, n_classes, J, total_segs, n_cols, stat_mask, ns, accumulate;
};

__global__ void __launch_bounds__(128) det_bins3_finalize_kernel(
    const Bins3FinalizeParams F) {
  const long long warp_global =
      (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long per_cell = static_cast<long long>(F.n_classes) * (F.n_cols + 1);
  if (warp_global >= F.n_cells * per_cell) return;
  const int c = static_cast<int>(warp_global / per_cell);
  const int rem = static_cast<int>(warp_global - c * per_cell);
  const int cls = rem / (F.n_cols + 1);
  const int col = rem - cls * (F.n_cols + 1);   // col == n_cols: constant weights
  const size_t oc = static_cast<size_t>(c) * F.n_classes + cls;
  if (col == F.n_cols) {
    if (F.cell_class_w && lane == 0) {
      const double v = F.cell_class_w[oc];
      for (int k = 0; k < WBX_NUM_DET_WCLASSES; ++k) {
        double* dst = F.out_w + oc * WBX_NUM_DET_WCLASSES + k;
        *dst = F.accumulate ? (*dst + v) : v;
      }
    }
    return;
  }
  const long long fj = F.cell_first_job[c];
  const long long lj = static_cast<long long>(F.cell_first_job[c + 1]) - 1;
  const int g_lo = static_cast<int>(((fj + 1) * F.J - 1) / F.n_jobs);
  const int g_hi = static_cast<int>(((lj + 1) * F.J - 1) / F.n_jobs);
  const size_t stride_g = static_cast<size_t>(F.total_segs) * F.n_cols;
  const int lo = F.class_ptr[cls], hi = F.class_ptr[cls + 1];
  double sum = 0.0;
  for (int i = lo + lane; i < hi; i += 32) {
    const double* rec = F.records + static_cast<size_t>(c) * stride_g +
                        static_cast<size_t>(F.class_segs[i]) * F.n_cols + col;
    for (int g = g_lo; g <= g_hi; ++g)
      sum += __ldcg(rec + static_cast<size_t>(g) * stride_g);
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    // column -> accumulator: the col-th selected statistic, then the weight
    int a = -1, seen = 0;
    for (int k = 0; k < F.ns; ++k) {
      if (F.stat_mask & (1 << k)) {
        if (seen == col) a = k;
        ++seen;
      }
    }
    if (a >= 0) {
      double* dst = F.out_ws + oc * WBX_NUM_DET_STATS + a;
      *dst = F.accumulate ? (*dst + sum) : sum;
    } else {
      for (int k = 0; k < WBX_NUM_DET_WCLASSES; ++k) {
        double* dst = F.out_w + oc * WBX_NUM_DET_WCLASSES + k;
        *dst = F.accumulate ? (*dst + sum) : sum;
      }
    }
  }
}

}  // namespace wbx
