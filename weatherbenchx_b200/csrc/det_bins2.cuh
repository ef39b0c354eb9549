// Binned aggregation, second generation: every CTA owns a FIXED part of the
// slab for all jobs, so that everything a class map decides is static.
//
// aggregation.py:320-335 of the reference multiplies the statistic by every
// bin mask inside xr.dot -- work that grows with the number of bins (34 in the
// public benchmark: 17 regions x {all, land}, run_benchmark_evaluation.py:
// 110-132,369).  The host folds all masks over the slab dims into one uint8
// class map (two grid points share a class iff they belong to the same set of
// bins); the kernel sums every statistic per (cell, class) in one pass and the
// host maps class sums to bin sums.
//
// The first-generation kernel (det_reduce_bins_kernel) walks the job-major tile
// list like the unbinned kernel, so the class of the four points a lane holds
// changes from step to step and every 128-point warp step ends in one or more
// warp-wide reductions keyed by (row, class) plus serial folds of the lanes
// that straddle a class boundary: 228 warp-instructions per float4 group
// against ~60 unbinned, 0.41-0.48 of the HBM roofline.
//
// Here the slab is cut into S parts of <= 4096 contiguous elements and the
// jobs into J groups.  CTA (s, g) streams part s of every job of group g (one
// ring stage per job), then part s + S_cta of every job, ...  A consumer
// thread owns the SAME 8 contiguous elements for all jobs of a part, so
//   * its classes are loaded once: slot A = the class of its first element,
//     slot B = the other class if a boundary (coastline, region edge) falls
//     inside its 8 elements; elements of a third class (two different
//     boundaries inside one block of 8, a fraction of a percent of the blocks
//     of real masks) are added to the warp's shared sums by their owners, one
//     lane after the other (maps where that is common keep the
//     first-generation kernel);
//   * its weights are loaded once (row weights, or per-element weights for
//     longitude-major arrays / odd row lengths);
//   * the per-point work is the unbinned kernel's (f32 statistics, f32 4-sums,
//     one f64 FMA per statistic and group) into register accumulators of
//     slot A, plus the same for slot B in the few threads that have one;
//   * nothing crosses lanes until the output cell changes: then the threads'
//     accumulators are folded per class (warp shuffles, fixed order) into
//     warp-private shared-memory sums, the 16 warps are combined in a fixed
//     order and ONE record per (part, group, cell) goes to global memory.
// No atomics, fixed summation orders => bit-stable results.  The class map is
// read once per CTA (7 KB), not once per slab: 8 / 12 B per point again.
#pragma once

#include "det_reduce.cuh"

namespace wbx {

struct Bins2Params {
  const unsigned char* class_map;  // [slab] device
  int n_classes;
  int S;             // slab parts (records: [S][n_cells + J][...])
  int S_cta;         // parts handled concurrently; CTA (s, g) takes parts
                     // s, s + S_cta, s + 2 S_cta, ... one after the other
  int J;             // job groups; grid = S_cta * J
  int part;          // elements per part (multiple of 16, <= 4096)
  int n_cells;       // cells of the launch
  double* records;
};

// Static per-thread description of its 8 elements of the current part.
template <bool WX>
struct Bins2Static {
  unsigned sel_a;    // bit i: element i belongs to slot A
  unsigned sel_b;    // bit i: element i belongs to slot B
  unsigned ovf;      // bit i: element i has a third (fourth, ...) class
  uint2 cls8;        // the classes of the 8 elements (one byte each)
  int cls_a, cls_b;
  bool active, has_b;
  double w[WX ? 8 : 2];  // per element, or per float4 group (row weight)
};

template <bool CLIM, bool MASK, bool WX>
__global__ void __launch_bounds__(kTmaThreads, 1)
    det_reduce_bins2_kernel(const DetParams P, const Bins2Params B,
                            const int stages, const int stage_bytes) {
  constexpr int NS = CLIM ? 6 : 3;
  constexpr int NA = NS + (MASK ? 1 : 0);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + kMaxStages;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + kMaxStages);
  double* wacc_all = reinterpret_cast<double*>(meta + kMaxStages);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nacc = B.n_classes * NA;   // doubles per warp / per record
  // block-of-8 permutation of the current part + per-warp counts (see set-up)
  int* order = reinterpret_cast<int*>(wacc_all + kConsumerWarps * nacc);
  int* warp_specials = order + kConsumerThreads;
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
        for (long long job = j_lo; job < j_hi; ++job) {
          const float* pa = reinterpret_cast<const float*>(__ldg(P.pred + job));
          const float* ta =
              reinterpret_cast<const float*>(__ldg(P.target + job));
          mbar_wait(&empty[s], ph ^ 1u);
          StageMeta mt;
          mt.cell = __ldg(P.cell + job);
          mt.len = len;
          mt.e0 = e0;
          mt.pad = 0;
          mt.wo = P.w_outer ? __ldg(P.w_outer + job) : 1.0;
          meta[s] = mt;
          unsigned char* st = ring + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], fbytes * (CLIM ? 3u : 2u) +
                                       (MASK ? static_cast<uint32_t>(len) : 0u));
          bulk_g2s(st, pa + e0, fbytes, &full[s], policy);
          bulk_g2s(st + off_t, ta + e0, fbytes, &full[s], policy);
          if constexpr (CLIM)
            bulk_g2s(st + off_c,
                     reinterpret_cast<const float*>(__ldg(P.clim + job)) + e0,
                     fbytes, &full[s], policy);
          if constexpr (MASK)
            bulk_g2s(st + off_m,
                     reinterpret_cast<const unsigned char*>(
                         __ldg(P.mask + job)) + e0,
                     static_cast<uint32_t>(len), &full[s], policy);
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
  const unsigned unx = static_cast<unsigned>(P.nx);
  const int stat_mask = P.stat_mask;
  double* wacc = wacc_all + static_cast<size_t>(warp) * nacc;
  for (int i = lane; i < nacc; i += 32) wacc[i] = 0.0;
  __syncwarp();
  // register accumulators: [slot A / B][statistic (+ weight)]
  double acc[2][NA];
#pragma unroll
  for (int sl = 0; sl < 2; ++sl)
#pragma unroll
    for (int a = 0; a < NA; ++a) acc[sl][a] = 0.0;
  Bins2Static<WX> S{};

  // fold the threads' accumulators into this warp's per-class sums: one warp
  // reduction per distinct class among the lanes, leaders in lane order
  auto fold_pair = [&](double (&a)[NA], const int cls, const bool valid) {
    unsigned um = __ballot_sync(0xffffffffu, valid);
    while (um) {
      const int leader = __ffs(um) - 1;
      const int lc = __shfl_sync(0xffffffffu, cls, leader);
      const bool mine = valid && cls == lc;
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        if (k >= NS || (stat_mask & (1 << k))) {
          const double tot = warp_sum(mine ? a[k] : 0.0);
          if (lane == 0) wacc[lc * NA + k] += tot;
        }
      }
      um &= ~__ballot_sync(0xffffffffu, mine);
    }
#pragma unroll
    for (int k = 0; k < NA; ++k) a[k] = 0.0;
  };

  auto flush = [&](const int s_part, const int cell) {
    fold_pair(acc[0], S.cls_a, S.active);
    fold_pair(acc[1], S.cls_b, S.active && S.has_b);
    // combine the 16 warps in a fixed order: one record per (part, group, cell)
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    double* rec = B.records +
                  (static_cast<size_t>(s_part) * (B.n_cells + B.J) + grp +
                   (cell - P.cell_base)) * nacc;
    for (int e = ctid; e < nacc; e += kConsumerThreads) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        v += wacc_all[static_cast<size_t>(w) * nacc + e];
        wacc_all[static_cast<size_t>(w) * nacc + e] = 0.0;
      }
      rec[e] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
  };

  int s = 0;
  uint32_t ph = 0;
  for (int s_part = s_cta; s_part < B.S; s_part += B.S_cta) {
    // ---- static set-up for this part ----------------------------------------
    // A thread owns one aligned block of 8 elements.  Blocks that hold more
    // than one class (a coastline, a region edge) need the two-slot code;
    // they are a few percent of the blocks but would drag nearly every warp
    // through it.  So the blocks are permuted once per part: the mixed ones go
    // to the first threads (one or two warps), all other warps are
    // class-uniform and run the plain code.
    const int e_lo = s_part * B.part;
    const int part_len = min(P.slab, e_lo + B.part) - e_lo;
    const int n_blocks = part_len >> 3;
    bool mixed = false;
    if (ctid < n_blocks) {
      const uint2 k8 = __ldg(reinterpret_cast<const uint2*>(
          B.class_map + e_lo + 8 * ctid));
      const unsigned first = (k8.x & 0xffu) * 0x01010101u;
      mixed = k8.x != first || k8.y != first;
    }
    const unsigned mixed_lanes = __ballot_sync(0xffffffffu, mixed);
    if (lane == 0) warp_specials[warp] = __popc(mixed_lanes);
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    int mixed_before = 0, mixed_total = 0;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
      const int c = warp_specials[w];
      mixed_before += w < warp ? c : 0;
      mixed_total += c;
    }
    mixed_before += __popc(mixed_lanes & ((1u << lane) - 1u));
    if (ctid < n_blocks)
      order[mixed ? mixed_before : mixed_total + (ctid - mixed_before)] = ctid;
    else
      order[ctid] = ctid;
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
    const int blk = order[ctid];      // the block this thread owns
    S.active = ctid < n_blocks;
    S.sel_a = S.sel_b = S.ovf = 0u;
    S.cls8 = make_uint2(0u, 0u);
    S.cls_a = S.cls_b = 0;
    S.has_b = false;
#pragma unroll
    for (int i = 0; i < (WX ? 8 : 2); ++i) S.w[i] = 0.0;
    if (S.active) {
      const unsigned e = static_cast<unsigned>(e_lo + 8 * blk);
      const uint2 k8 = __ldg(reinterpret_cast<const uint2*>(B.class_map + e));
      unsigned char cls[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cls[i] = static_cast<unsigned char>(k8.x >> (8 * i));
        cls[4 + i] = static_cast<unsigned char>(k8.y >> (8 * i));
      }
      S.cls8 = k8;
      S.cls_a = cls[0];
      S.cls_b = cls[0];
      S.sel_a = 1u;
#pragma unroll
      for (int i = 1; i < 8; ++i) {
        if (cls[i] == S.cls_a) {
          S.sel_a |= 1u << i;
        } else {
          if (!S.has_b) {
            S.cls_b = cls[i];
            S.has_b = true;
          }
          if (cls[i] == S.cls_b) S.sel_b |= 1u << i;
          else S.ovf |= 1u << i;   // rare: a third class within 8 elements
        }
      }
      if constexpr (WX) {
        unsigned y = e / unx, x = e - y * unx;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          S.w[i] = (P.w_y ? __ldg(P.w_y + y) : 1.0) *
                   (P.w_x ? __ldg(P.w_x + x) : 1.0);
          if (++x == unx) {
            x = 0;
            ++y;
          }
        }
      } else {  // rows are a multiple of four long: a group stays in its row
        S.w[0] = P.w_y ? __ldg(P.w_y + e / unx) : 1.0;
        S.w[1] = P.w_y ? __ldg(P.w_y + (e + 4u) / unx) : 1.0;
      }
    }
    // (static per part, warp-uniform) does this warp hold mixed blocks at all?
    const bool warp_mixed = __any_sync(0xffffffffu, S.has_b);
    const bool warp_ovf = __any_sync(0xffffffffu, S.ovf != 0u);
    int cur_cell = -1;
    for (long long job = j_lo; job < j_hi; ++job) {
      mbar_wait(&full[s], ph);
      const StageMeta mt = meta[s];
      if (mt.cell != cur_cell) {
        if (cur_cell >= 0) flush(s_part, cur_cell);
        cur_cell = mt.cell;
      }
      {
        const unsigned char* stg = ring + (size_t)s * stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(stg);
        const float4* stt = reinterpret_cast<const float4*>(stg + off_t);
        const float4* sc = reinterpret_cast<const float4*>(stg + off_c);
        const uint2* sm = reinterpret_cast<const uint2*>(stg + off_m);
        uint2 m8 = make_uint2(0x01010101u, 0x01010101u);
        if constexpr (MASK) {
          if (S.active) m8 = sm[blk];
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // inactive threads (beyond the part) compute on zeros into nothing
          float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), tv = pv, cv = pv;
          if (S.active) {
            pv = sp[2 * blk + g];
            tv = stt[2 * blk + g];
            if constexpr (CLIM) cv = sc[2 * blk + g];
          }
          const unsigned mw = g == 0 ? m8.x : m8.y;
          const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
          const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
          const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
          PointStats<CLIM, MASK, false> q[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            q[i].eval(pp[i], tt[i], cc[i],
                      static_cast<unsigned char>(mw >> (8 * i)));
          // all-ones where the element belongs to slot A / slot B
          unsigned ma[4], mb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ma[i] = 0u - ((S.sel_a >> (4 * g + i)) & 1u);
            mb[i] = 0u - ((S.sel_b >> (4 * g + i)) & 1u);
          }
          if constexpr (!WX) {
            const double wg = S.w[g] * mt.wo;
            if (!warp_mixed) {   // class-uniform warp: the unbinned code
#pragma unroll
              for (int k = 0; k < NS; ++k) {
                if (stat_mask & (1 << k)) {  // warp-uniform
                  const float s4 = __fadd_rn(__fadd_rn(q[0].s[k], q[1].s[k]),
                                             __fadd_rn(q[2].s[k], q[3].s[k]));
                  acc[0][k] += static_cast<double>(s4) * wg;
                }
              }
              if constexpr (MASK) {
                const float n4 = (q[0].valid[0] + q[1].valid[0]) +
                                 (q[2].valid[0] + q[3].valid[0]);
                acc[0][NS] += static_cast<double>(n4) * wg;
              }
            } else {
#pragma unroll
              for (int k = 0; k < NS; ++k) {
                if (stat_mask & (1 << k)) {
                  float a4[4], b4[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const unsigned bits = __float_as_uint(q[i].s[k]);
                    a4[i] = __uint_as_float(bits & ma[i]);
                    b4[i] = __uint_as_float(bits & mb[i]);
                  }
                  const float sa = __fadd_rn(__fadd_rn(a4[0], a4[1]),
                                             __fadd_rn(a4[2], a4[3]));
                  const float sb = __fadd_rn(__fadd_rn(b4[0], b4[1]),
                                             __fadd_rn(b4[2], b4[3]));
                  acc[0][k] += static_cast<double>(sa) * wg;
                  acc[1][k] += static_cast<double>(sb) * wg;
                }
              }
              if constexpr (MASK) {
                float a4[4], b4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const unsigned bits = __float_as_uint(q[i].valid[0]);
                  a4[i] = __uint_as_float(bits & ma[i]);
                  b4[i] = __uint_as_float(bits & mb[i]);
                }
                acc[0][NS] +=
                    static_cast<double>((a4[0] + a4[1]) + (a4[2] + a4[3])) * wg;
                acc[1][NS] +=
                    static_cast<double>((b4[0] + b4[1]) + (b4[2] + b4[3])) * wg;
              }
            }
          } else {
            // per-element f64 weights (longitude-major arrays, odd row lengths)
            double we[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) we[i] = S.w[4 * g + i] * mt.wo;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              if (stat_mask & (1 << k)) {
                double va = 0.0, vb = 0.0;
#pragma unroll
                if (!warp_mixed) {
#pragma unroll
                  for (int i = 3; i >= 0; --i)
                    va = fma(static_cast<double>(q[i].s[k]), we[i], va);
                } else {
#pragma unroll
                  for (int i = 3; i >= 0; --i) {
                    const unsigned bits = __float_as_uint(q[i].s[k]);
                    va = fma(static_cast<double>(__uint_as_float(bits & ma[i])),
                             we[i], va);
                    vb = fma(static_cast<double>(__uint_as_float(bits & mb[i])),
                             we[i], vb);
                  }
                }
                acc[0][k] += va;
                acc[1][k] += vb;
              }
            }
            if constexpr (MASK) {
              double va = 0.0, vb = 0.0;
#pragma unroll
              for (int i = 3; i >= 0; --i) {
                const unsigned bits = __float_as_uint(q[i].valid[0]);
                va = fma(static_cast<double>(__uint_as_float(bits & ma[i])),
                         we[i], va);
                vb = fma(static_cast<double>(__uint_as_float(bits & mb[i])),
                         we[i], vb);
              }
              acc[0][NS] += va;
              acc[1][NS] += vb;
            }
          }
          // elements of a third class (two different boundaries inside one
          // block of 8: a coastline next to a region edge): their owners add
          // them to the warp's shared sums one lane after the other
          if (warp_ovf) {
            const unsigned og = (S.ovf >> (4 * g)) & 0xfu;
            unsigned om = __ballot_sync(0xffffffffu, og != 0u);
            while (om) {
              const int src = __ffs(om) - 1;
              if (lane == src) {
                const unsigned word = g == 0 ? S.cls8.x : S.cls8.y;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  if (og & (1u << i)) {
                    double wi;
                    if constexpr (WX) wi = S.w[4 * g + i] * mt.wo;
                    else wi = S.w[g] * mt.wo;
                    double* slot = wacc + ((word >> (8 * i)) & 0xffu) * NA;
#pragma unroll
                    for (int k = 0; k < NS; ++k)
                      if (stat_mask & (1 << k))
                        slot[k] += static_cast<double>(q[i].s[k]) * wi;
                    if constexpr (MASK)
                      slot[NS] += static_cast<double>(q[i].valid[0]) * wi;
                  }
                }
              }
              __syncwarp();
              om &= om - 1;
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
    if (cur_cell >= 0) flush(s_part, cur_cell);
  }
}

// records -> out[(cell * n_classes + class)][slot]: one warp per (cell, class,
// accumulator); lanes stride over the slab parts, groups in order.
struct Bins2FinalizeParams {
  const double* records;
  const int32_t* cell_first_job;   // [n_cells + 1], relative to the launch
  const double* cell_class_w;      // [n_cells * n_classes] or NULL (masked)
  double* out_ws;                  // [n_cells * n_classes * 6]
  double* out_w;                   // [n_cells * n_classes * 4]
  long long n_jobs;
  int n_cells, n_classes, S, J, ns, na, accumulate;
};

__global__ void __launch_bounds__(128) det_bins2_finalize_kernel(
    const Bins2FinalizeParams F) {
  const long long warp_global =
      (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long per_cell = static_cast<long long>(F.n_classes) * (F.na + 1);
  if (warp_global >= F.n_cells * per_cell) return;
  const int c = static_cast<int>(warp_global / per_cell);
  const int rem = static_cast<int>(warp_global - c * per_cell);
  const int cls = rem / (F.na + 1);
  const int a = rem - cls * (F.na + 1);   // a == na: constant weights
  const size_t oc = static_cast<size_t>(c) * F.n_classes + cls;
  if (a == F.na) {
    if (F.cell_class_w && lane == 0) {
      const double v = F.cell_class_w[oc];
      for (int k = 0; k < WBX_NUM_DET_WCLASSES; ++k) {
        double* dst = F.out_w + oc * WBX_NUM_DET_WCLASSES + k;
        *dst = F.accumulate ? (*dst + v) : v;
      }
    }
    return;
  }
  const int nacc = F.n_classes * F.na;
  const long long fj = F.cell_first_job[c];
  const long long lj = static_cast<long long>(F.cell_first_job[c + 1]) - 1;
  const int g_lo = static_cast<int>(((fj + 1) * F.J - 1) / F.n_jobs);
  const int g_hi = static_cast<int>(((lj + 1) * F.J - 1) / F.n_jobs);
  const size_t stride_s = static_cast<size_t>(F.n_cells + F.J) * nacc;
  double sum = 0.0;
  for (int sp = lane; sp < F.S; sp += 32) {
    const double* rec = F.records + sp * stride_s +
                        static_cast<size_t>(c) * nacc +
                        static_cast<size_t>(cls) * F.na + a;
    for (int g = g_lo; g <= g_hi; ++g)
      sum += __ldcg(rec + static_cast<size_t>(g) * nacc);
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    if (a < F.ns) {
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
