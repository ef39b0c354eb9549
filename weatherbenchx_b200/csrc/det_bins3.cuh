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
// class for the few quads a boundary runs through -- and sorts them by class.
// Whole blocks of 16 slots of a class go to groups of 8 consecutive consumer
// threads: a thread owns slots u and u + 8 of its block, so
//   * its two slots always belong to the same class (one accumulator set),
//   * the 8 lanes of a group read 8 neighbouring quads (in the usual case of
//     a run of whole quads: 128 contiguous bytes, no bank conflicts),
//   * the 4 groups of a warp hold at most 4 class runs, the SEGMENTS.
// What is left of every class (< 16 slots) follows, two slots per thread, in
// the lanes after the last block: there a segment is any run of lanes, and
// the warps that hold such lanes ("misc" warps, one or two per part) reduce
// with a general segmented scan.  Nothing is padded, so a part of ~3500
// elements fills the 512 threads whatever the number of classes in it.
//
// The kernel is therefore the unbinned one plus static per-thread data:
//   * a thread reads the same two quads for all jobs; weights (one f64 row
//     weight per quad, or element weights for longitude-major arrays / odd
//     row lengths) and selections live in registers;
//   * per job: f32 statistics, f32 4-sums, one f64 FMA per statistic and quad
//     into the thread's accumulators -- no class logic, no branches that depend
//     on data or on the map (one warp-uniform static branch picks the code with
//     the element selection for warps that hold a boundary quad);
//   * when the output cell changes: three butterfly steps over the 8 lanes of
//     a group that reduce ALL accumulators at once (a lane ends up with the
//     group sum of one of them), two more steps predicated on the static first
//     group of the segment leave the segment's sums in its last group, whose
//     lanes write them to the segment's record.  No shared-memory
//     accumulators, no CTA barriers, every CTA does the same work whatever the
//     map looks like.
//   * the finalize kernel adds, per (cell, class), the records of the class's
//     segments (a host-built list) in a fixed order.
// No atomics, fixed summation orders => bit-stable results.  The class map is
// never read by the GPU at all: 8 / 12 B per point.
#pragma once

#include "det_reduce.cuh"

namespace wbx {

constexpr int kBins3Slots = 2 * kConsumerThreads;  // per part

// Slot descriptor of a consumer thread, two words.
//   a: [9:0] first quad | [19:10] second quad | [23:20] selection of the
//      first | [27:24] selection of the second (0: slot unused)
//   b: [4:0] first lane of the thread's segment | [5] closes the segment (in
//      a misc warp: the last lane of the segment; elsewhere: every lane of
//      its last group) | [6] misc warp | [19:7] segment of the part
__host__ __device__ __forceinline__ uint32_t bins3_pack_a(int quad0, int quad1,
                                                          int sel0, int sel1) {
  return static_cast<uint32_t>(quad0) | (static_cast<uint32_t>(quad1) << 10) |
         (static_cast<uint32_t>(sel0) << 20) |
         (static_cast<uint32_t>(sel1) << 24);
}
__host__ __device__ __forceinline__ uint32_t bins3_pack_b(int first_lane,
                                                          int closes, int misc,
                                                          int seg) {
  return static_cast<uint32_t>(first_lane) |
         (static_cast<uint32_t>(closes) << 5) |
         (static_cast<uint32_t>(misc) << 6) | (static_cast<uint32_t>(seg) << 7);
}

struct Bins3Params {
  const uint2* slot_desc;     // [S][512] (bins3_pack_a, bins3_pack_b)
  const int32_t* seg_base;    // [S + 1] first segment of every part
  int S;                      // slab parts
  int S_cta;                  // parts in flight; CTA (s, g) takes s, s + S_cta, ...
  int J;                      // job groups; grid = S_cta * J
  int part;                   // elements per part (multiple of 16, <= 4096)
  int total_segs;
  int n_cols;                 // record columns: selected statistics (+ weight)
  double* records;            // [n_cells + J][total_segs][n_cols]
};

// Sums of the segments of a warp -> their records, accumulators reset.
//
// Warps whose segments are runs of whole 8-lane groups: all accumulators are
// reduced together.  In step m of the butterfly over the 8 lanes of a group a
// lane keeps one half of its columns (the half bit m of its lane number
// names), sends the other half to its partner and adds what it receives, so
// after log2(NC) steps lane l holds column l % NC summed over those steps'
// partners -- NC - 1 shuffles instead of NC * log2(NC); the remaining steps
// (rest of the group, then the groups of the segment) work on that one value.
// Unselected statistics ride along (their sums are never stored).  `rec` is
// the record of the segment the lane's group closes, or NULL.
template <int NS, int NA>
__device__ __forceinline__ void bins3_flush_groups(double (&acc)[NA],
                                                   double* rec,
                                                   const int stat_mask,
                                                   const int lane,
                                                   const int seg_first_lane) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int NC = NA <= 4 ? 4 : 8;
  double v[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) v[k] = k < NA ? acc[k] : 0.0;
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.0;
  // columns 2j, 2j + 1 -> j: lanes with bit m clear keep the even one
  {
    const bool odd = (lane & 1) != 0;
#pragma unroll
    for (int j = 0; j < NC / 2; ++j) {
      const double keep = odd ? v[2 * j + 1] : v[2 * j];
      const double send = odd ? v[2 * j] : v[2 * j + 1];
      v[j] = keep + __shfl_xor_sync(kFull, send, 1);
    }
  }
  {
    const bool odd = (lane & 2) != 0;
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const double keep = odd ? v[2 * j + 1] : v[2 * j];
      const double send = odd ? v[2 * j] : v[2 * j + 1];
      v[j] = keep + __shfl_xor_sync(kFull, send, 2);
    }
  }
  if constexpr (NC == 8) {
    const bool odd = (lane & 4) != 0;
    const double keep = odd ? v[1] : v[0];
    const double send = odd ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(kFull, send, 4);
  } else {
    v[0] += __shfl_xor_sync(kFull, v[0], 4);
  }
  // the groups of a segment: inclusive scan from its first group
  const int group = lane >> 3, first_group = seg_first_lane >> 3;
  const double up8 = __shfl_up_sync(kFull, v[0], 8);
  if (group - 1 >= first_group) v[0] += up8;
  const double up16 = __shfl_up_sync(kFull, v[0], 16);
  if (group - 2 >= first_group) v[0] += up16;
  // lane l of the closing group holds column l % NC (step m picked bit m of
  // the column number)
  const int k = lane & (NC - 1);
  const bool wanted = k < NA && (k >= NS || ((stat_mask >> k) & 1));
  if (rec != nullptr && (lane & 7) < NC && wanted) {
    const int col = __popc(stat_mask & ((1 << (k < NS ? k : NS)) - 1));
    rec[col] = v[0];
  }
}

// Warps that hold the remainders of classes: segments are arbitrary runs of
// lanes.  One segmented inclusive scan per accumulator, all issued step by
// step; `rec` is the record of the segment the lane closes, or NULL.
template <int NS, int NA>
__device__ __forceinline__ void bins3_flush_lanes(double (&acc)[NA], double* rec,
                                                  const int stat_mask,
                                                  const int lane,
                                                  const int seg_first_lane) {
  constexpr unsigned kFull = 0xffffffffu;
  double v[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    v[k] = acc[k];
    acc[k] = 0.0;
  }
#pragma unroll
  for (int dlt = 1; dlt < 32; dlt <<= 1) {
    const bool add = lane - dlt >= seg_first_lane;
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const double up = __shfl_up_sync(kFull, v[k], dlt);
      v[k] += add ? up : 0.0;
    }
  }
  if (rec != nullptr) {
    int col = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k)
      if (k >= NS || ((stat_mask >> k) & 1)) rec[col++] = v[k];
  }
}

// (17 warps: one SM sub-partition holds five of them, which caps the kernel at
// 96 registers per thread)
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
  double acc[NA];
#pragma unroll
  for (int a = 0; a < NA; ++a) acc[a] = 0.0;

  int s = 0;
  uint32_t ph = 0;
  for (int s_part = s_cta; s_part < B.S; s_part += B.S_cta) {
    // ---- the static schedule of this part -----------------------------------
    int quad[2];
    unsigned sel[2];
    // weights: one f64 row weight per quad, or (WX) four f32 element weights
    double wq[2];
    float wf[2][WX ? 4 : 1];
    const int seg0 = __ldg(B.seg_base + s_part);
    const unsigned e_part = static_cast<unsigned>(s_part) * B.part;
    const unsigned unx = static_cast<unsigned>(P.nx);
    const uint2 d = __ldg(B.slot_desc +
                          static_cast<size_t>(s_part) * kConsumerThreads + ctid);
    quad[0] = static_cast<int>(d.x & 0x3ffu);
    quad[1] = static_cast<int>((d.x >> 10) & 0x3ffu);
    sel[0] = (d.x >> 20) & 0xfu;
    sel[1] = (d.x >> 24) & 0xfu;
    const int seg_first_lane = static_cast<int>(d.y & 31u);
    // record of the segment this lane closes (-1: it closes none)
    const int seg_rec = (d.y & 32u) ? seg0 + static_cast<int>(d.y >> 7) : -1;
    const bool misc = (d.y & 64u) != 0u;   // warp-uniform
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const unsigned e = e_part + 4u * static_cast<unsigned>(quad[ps]);
      unsigned y = e / unx, x = e - y * unx;
      if constexpr (WX) {
#pragma unroll
        wq[ps] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double w = (P.w_y ? __ldg(P.w_y + y) : 1.0) *
                           (P.w_x ? __ldg(P.w_x + x) : 1.0);
          wf[ps][i] = static_cast<float>(w);
          if (++x == unx) {
            x = 0;
            if (y + 1 < static_cast<unsigned>(P.ny)) ++y;
          }
        }
      } else {  // rows are a multiple of four long: a quad stays in its row
        wq[ps] = P.w_y ? __ldg(P.w_y + y) : 1.0;
        wf[ps][0] = 0.f;
      }
    }
    // warp-uniform, static: does the warp hold slots at all / boundary quads
    const bool live = __any_sync(kFull, (sel[0] | sel[1]) != 0u);
    const bool partial = __any_sync(
        kFull, (sel[0] != 0u && sel[0] != 0xfu) || (sel[1] != 0u && sel[1] != 0xfu));
    const bool used0 = sel[0] != 0u, used1 = sel[1] != 0u;

    auto flush = [&](const int cell) {
      if (!live) return;   // warp-uniform
      double* rec = nullptr;
      if (seg_rec >= 0)
        rec = B.records + (static_cast<size_t>(cell - P.cell_base + grp) *
                               B.total_segs + seg_rec) * B.n_cols;
      if (misc)   // warp-uniform
        bins3_flush_lanes<NS, NA>(acc, rec, stat_mask, lane, seg_first_lane);
      else
        bins3_flush_groups<NS, NA>(acc, rec, stat_mask, lane, seg_first_lane);
    };

    int cur_cell = -1;
    for (long long job = j_lo; job < j_hi; ++job) {
      mbar_wait(&full[s], ph);
      const StageMeta mt = meta[s];
      // this job's operands are requested before the previous cell is flushed:
      // the shared-memory latency hides behind the shuffles of the flush
      float4 pv[2], tv[2], cv[2];
      uint32_t mw[2];
      {
        const unsigned char* stg = ring + (size_t)s * stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(stg);
        const float4* stt = reinterpret_cast<const float4*>(stg + off_t);
        const float4* sc = reinterpret_cast<const float4*>(stg + off_c);
        const uint32_t* sm = reinterpret_cast<const uint32_t*>(stg + off_m);
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          pv[ps] = sp[quad[ps]];
          tv[ps] = stt[quad[ps]];
          cv[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
          mw[ps] = 0x01010101u;
          if constexpr (CLIM) cv[ps] = sc[quad[ps]];
          if constexpr (MASK) mw[ps] = sm[quad[ps]];
        }
      }
      if (mt.cell != cur_cell) {
        if (cur_cell >= 0) flush(cur_cell);
        cur_cell = mt.cell;
      }
      if (live) {   // warp-uniform
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const float pp[4] = {pv[ps].x, pv[ps].y, pv[ps].z, pv[ps].w};
          const float tt[4] = {tv[ps].x, tv[ps].y, tv[ps].z, tv[ps].w};
          const float cc[4] = {cv[ps].x, cv[ps].y, cv[ps].z, cv[ps].w};
          PointStats<CLIM, MASK, false> q[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            q[i].eval(pp[i], tt[i], cc[i],
                      static_cast<unsigned char>(mw[ps] >> (8 * i)));
          if (partial) {
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
          // an unused slot (the odd one of a class remainder) adds exact zeros
          const bool used = ps == 0 ? used0 : used1;
          const unsigned keep = used ? 0xffffffffu : 0u;
          if constexpr (!WX) {
            const double wj = wq[ps] * mt.wo;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              if (stat_mask & (1 << k)) {   // warp-uniform
                const float s4 = __fadd_rn(__fadd_rn(q[0].s[k], q[1].s[k]),
                                           __fadd_rn(q[2].s[k], q[3].s[k]));
                acc[k] = fma(static_cast<double>(__uint_as_float(
                                 __float_as_uint(s4) & keep)), wj, acc[k]);
              }
            }
            if constexpr (MASK) {
              const float n4 = (q[0].valid[0] + q[1].valid[0]) +
                               (q[2].valid[0] + q[3].valid[0]);
              acc[NS] = fma(static_cast<double>(__uint_as_float(
                                __float_as_uint(n4) & keep)), wj, acc[NS]);
            }
          } else {
            // element weights: the statistic values are f32 roundings, their
            // weighted 4-sum is taken in f32 with the weights rounded to f32
            // (one more rounding of the same size per term), then f64; the
            // sum of weights of a masked aggregation goes the same way
            // (relative error ~1e-8, against the 1e-5 of the parity bar)
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              if (stat_mask & (1 << k)) {
                const float s4 = __fadd_rn(
                    __fadd_rn(__fmul_rn(q[0].s[k], wf[ps][0]),
                              __fmul_rn(q[1].s[k], wf[ps][1])),
                    __fadd_rn(__fmul_rn(q[2].s[k], wf[ps][2]),
                              __fmul_rn(q[3].s[k], wf[ps][3])));
                acc[k] = fma(static_cast<double>(__uint_as_float(
                                 __float_as_uint(s4) & keep)), mt.wo, acc[k]);
              }
            }
            if constexpr (MASK) {
              const float n4 = __fadd_rn(
                  __fadd_rn(__fmul_rn(q[0].valid[0], wf[ps][0]),
                            __fmul_rn(q[1].valid[0], wf[ps][1])),
                  __fadd_rn(__fmul_rn(q[2].valid[0], wf[ps][2]),
                            __fmul_rn(q[3].valid[0], wf[ps][3])));
              acc[NS] = fma(static_cast<double>(__uint_as_float(
                                __float_as_uint(n4) & keep)), mt.wo, acc[NS]);
            }
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
