// Host side of the fused deterministic-statistics reduction: plan building,
// kernel selection, host-space streaming.  C ABI in include/wbx_b200.h.
#include <algorithm>
#include <new>
#include <string.h>

#include <type_traits>

#include "det_reduce.cuh"
#include "det_bins3.cuh"

namespace wbx {

enum Path { kPathTma = 0, kPathLdg4 = 1, kPathLdg1 = 2 };

static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace wbx

struct wbx_det_plan {
  int32_t space = 0, flags = 0;
  int64_t n_jobs = 0, ny = 0, nx = 0, n_cells = 0;
  bool has_clim = false, has_mask = false, skipna = false, per_elem = false;
  bool has_wo = false, has_wy = false, has_wx = false;
  std::vector<uint64_t> pred, target, clim, mask;
  std::vector<int32_t> cell;
  std::vector<double> wo, wy, wx;
  std::vector<float> wxf;
  double sum_wy = 1.0, sum_wx = 1.0;
  int path = 0;
  int tile = 0, tiles_per_slab = 0, stages = 0, stage_bytes = 0;
  size_t smem_bytes = 0;
  int n_stats = 3, n_weights = 0, nacc = 3;
  // device-space plans: tables uploaded once.
  wbx::DevBuf tables;
  wbx::DetParams params{};
  const int32_t* d_cell_first_job = nullptr;
  const double* d_cell_w = nullptr;
  int grid = 0;
  // weights are shared by device- and host-space plans.
  wbx::DevBuf weights;
  const double* d_wy = nullptr;
  const double* d_wx = nullptr;
  const float* d_wxf = nullptr;
  std::vector<unsigned char> chunk_host[2];
  int stat_mask = 0x3f;
  // categorical transform (wbx_det_desc.xform): per-job thresholds
  int xform = 0;
  std::vector<float> thr_pred, thr_target;
  // binned plans (class map over the slab)
  int n_classes = 0;
  std::vector<double> class_w;   // sum of w_y over the points of every class
  // third-generation binned kernel (det_bins3.cuh): the reduction schedule
  // the host compiles from the class map (slots, segments, class -> segments)
  struct Bins3 {
    bool ok = false, wx = false;
    int part = 0, S = 0, total_segs = 0, n_cols = 0;
    wbx::DevBuf tables;
    const uint2* d_desc = nullptr;
    const int32_t* d_seg_base = nullptr;
    const int32_t* d_class_ptr = nullptr;
    const int32_t* d_class_segs = nullptr;
  } bins3;
  int cells_mult() const { return n_classes > 0 ? n_classes : 1; }
};

namespace wbx {

// Launch geometry of the third-generation binned kernel for nj jobs: the S
// slab parts are fixed by the plan; small slabs share the SMs out between
// slab parts and J job groups (S_cta * J <= #SMs).
struct Bins3Geometry {
  int S_cta = 0, J = 0, stage_bytes = 0, stages = 0;
  size_t smem = 0;
  bool ok = false;
};

static Bins3Geometry bins3_geometry(const wbx_ctx* ctx, const wbx_det_plan* plan,
                                    long long nj) {
  Bins3Geometry g;
  const wbx_det_plan::Bins3& b = plan->bins3;
  if (!b.ok) return g;
  const int G = ctx->sm_count;
  g.S_cta = std::min(b.S, G);
  g.J = static_cast<int>(
      std::max<long long>(1, std::min<long long>(nj, G / g.S_cta)));
  g.stage_bytes = static_cast<int>(round_up(
      static_cast<size_t>(b.part) * 4 * (plan->has_clim ? 3 : 2) +
          (plan->has_mask ? b.part : 0), 128));
  const size_t overhead = 2 * kMaxStages * sizeof(uint64_t) +
                          kMaxStages * sizeof(StageMeta) + 128;
  const size_t cap = std::min<size_t>(ctx->smem_optin, 227 * 1024);
  if (overhead + 2 * static_cast<size_t>(g.stage_bytes) > cap) return g;
  g.stages = static_cast<int>(
      std::min<size_t>(kMaxStages, (cap - overhead) / g.stage_bytes));
  g.smem = static_cast<size_t>(g.stages) * g.stage_bytes + overhead;
  g.ok = g.stages >= 2;
  return g;
}

// Host tables of the reduction schedule (see det_bins3.cuh).
struct Bins3Host {
  int part = 0, S = 0, total_segs = 0;
  std::vector<uint32_t> desc;   // [S][512][2]
  std::vector<int32_t> seg_base, class_ptr, class_segs;
};

// Slots, segments and class lists for parts of `part` elements; false when a
// part needs more than the 512 consumer threads.
struct Bins3Slot { int cls, quad, sel; };
struct Bins3Thread { int cls; Bins3Slot a, b; };

static bool bins3_schedule(const unsigned char* cmap, int n_classes,
                           long long slab, int part, Bins3Host* T) {
  using Slot = Bins3Slot;
  using Thread = Bins3Thread;
  const int S = static_cast<int>((slab + part - 1) / part);
  T->part = part;
  T->S = S;
  T->desc.assign(static_cast<size_t>(S) * kConsumerThreads * 2, 0u);
  T->seg_base.assign(S + 1, 0);
  std::vector<std::vector<int32_t>> segs_of(n_classes);
  std::vector<Slot> slots, sorted;
  std::vector<int> count(n_classes + 1);
  int n_seg_total = 0;
  for (int s = 0; s < S; ++s) {
    const long long e_lo = static_cast<long long>(s) * part;
    const int len = static_cast<int>(std::min<long long>(part, slab - e_lo));
    slots.clear();
    for (int q = 0; q < len / 4; ++q) {
      const unsigned char* c4 = cmap + e_lo + 4 * q;
      int done = 0;
      for (int i = 0; i < 4; ++i) {
        if (done & (1 << i)) continue;
        int sel = 0;
        for (int k = i; k < 4; ++k)
          if (c4[k] == c4[i]) sel |= 1 << k;
        done |= sel;
        slots.push_back({c4[i], q, sel});
      }
    }
    // counting sort by class (quads stay ascending inside a class)
    std::fill(count.begin(), count.end(), 0);
    for (const Slot& sl : slots) ++count[sl.cls + 1];
    std::vector<int> start(n_classes + 1, 0);
    for (int c = 0; c < n_classes; ++c) start[c + 1] = start[c] + count[c + 1];
    sorted.resize(slots.size());
    {
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (const Slot& sl : slots) sorted[fill[sl.cls]++] = sl;
    }
    // threads: first the whole blocks of 16 slots of every class (8 threads,
    // slots u and u + 8), then the remainders (two consecutive slots each)
    std::vector<Thread> threads;
    for (int c = 0; c < n_classes; ++c) {
      const int n = start[c + 1] - start[c];
      for (int blk = 0; blk < n / 16; ++blk) {
        for (int l = 0; l < 8; ++l) {
          Thread th;
          th.cls = c;
          th.a = sorted[start[c] + 16 * blk + l];
          th.b = sorted[start[c] + 16 * blk + 8 + l];
          threads.push_back(th);
        }
      }
    }
    const int n_regular = static_cast<int>(threads.size());
    for (int c = 0; c < n_classes; ++c) {
      const int n = start[c + 1] - start[c];
      for (int u = n / 16 * 16; u < n; u += 2) {
        Thread th;
        th.cls = c;
        th.a = sorted[start[c] + u];
        th.b.cls = c;
        th.b.quad = 0;
        th.b.sel = 0;
        if (u + 1 < n) th.b = sorted[start[c] + u + 1];
        threads.push_back(th);
      }
    }
    const int n_threads = static_cast<int>(threads.size());
    if (n_threads > kConsumerThreads) return false;
    int seg_local = 0;
    T->seg_base[s] = n_seg_total;
    for (int t0 = 0; t0 < kConsumerThreads; t0 += 32) {
      // a warp that holds remainder lanes reduces at lane granularity
      const bool misc = t0 + 32 > n_regular && t0 < n_threads;
      int lane = 0;
      while (lane < 32) {
        const size_t at =
            (static_cast<size_t>(s) * kConsumerThreads + t0 + lane) * 2;
        if (t0 + lane >= n_threads) {   // unused lane: selects nothing
          T->desc[at] = bins3_pack_a(0, 0, 0, 0);
          T->desc[at + 1] = bins3_pack_b(lane, 0, misc, 0);
          ++lane;
          continue;
        }
        // the segment: the run of lanes of this class in this warp
        const int cls = threads[t0 + lane].cls;
        int end = lane;
        while (end + 1 < 32 && t0 + end + 1 < n_threads &&
               threads[t0 + end + 1].cls == cls)
          ++end;
        if (seg_local >= 8192) return false;
        for (int l = lane; l <= end; ++l) {
          const Thread& th = threads[t0 + l];
          const size_t al =
              (static_cast<size_t>(s) * kConsumerThreads + t0 + l) * 2;
          T->desc[al] = bins3_pack_a(th.a.quad, th.b.quad, th.a.sel, th.b.sel);
          // regular warps: segments are runs of whole groups and every lane
          // of the last group stores one accumulator
          const bool closes = misc ? l == end : l >= (end & ~7);
          T->desc[al + 1] = bins3_pack_b(lane, closes, misc, seg_local);
        }
        segs_of[cls].push_back(n_seg_total + seg_local);
        ++seg_local;
        lane = end + 1;
      }
    }
    n_seg_total += seg_local;
  }
  T->seg_base[S] = n_seg_total;
  T->total_segs = n_seg_total;
  T->class_ptr.assign(n_classes + 1, 0);
  T->class_segs.clear();
  for (int c = 0; c < n_classes; ++c) {
    T->class_segs.insert(T->class_segs.end(), segs_of[c].begin(),
                         segs_of[c].end());
    T->class_ptr[c + 1] = static_cast<int32_t>(T->class_segs.size());
  }
  return true;
}

// Chooses the part size (all SMs busy in whole rounds when the slab is large,
// at least 512 elements per part when it is small), compiles the schedule and
// uploads it.
static int bins3_build(wbx_ctx* ctx, wbx_det_plan* p,
                       const unsigned char* cmap) {
  const long long slab = p->ny * p->nx;
  const int G = ctx->sm_count;
  const bool wx = p->has_wx || (p->nx % 4) != 0;
  const long long s_min = (slab + 4095) / 4096;
  long long want;
  if (s_min >= G) {
    want = (s_min + G - 1) / G * G;
  } else {
    const long long J = std::max<long long>(
        1, std::min<long long>(p->n_jobs, G / s_min));
    want = std::max<long long>(
        s_min, std::min<long long>(G / J, (slab + 511) / 512));
  }
  Bins3Host T;
  bool ok = false;
  for (int attempt = 0; attempt < 24 && !ok; ++attempt) {
    const int part = static_cast<int>(round_up((slab + want - 1) / want, 16));
    if (part <= 4096)
      ok = bins3_schedule(cmap, p->n_classes, slab, part, &T);
    if (!ok) {
      // more boundary quads than spare slots: smaller parts
      if (part <= 16) break;
      want = want >= G ? want + G : std::max(want + 1, want * 5 / 4);
    }
  }
  if (!ok) return WBX_OK;   // the first-generation kernel serves the plan
  wbx_det_plan::Bins3& b = p->bins3;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += round_up(bytes, 16);
    return o;
  };
  const size_t o_desc = take(T.desc.size() * 4);
  const size_t o_sb = take(T.seg_base.size() * 4);
  const size_t o_cp = take(T.class_ptr.size() * 4);
  const size_t o_cs = take(std::max<size_t>(T.class_segs.size(), 1) * 4);
  std::vector<unsigned char> host(off, 0);
  memcpy(host.data() + o_desc, T.desc.data(), T.desc.size() * 4);
  memcpy(host.data() + o_sb, T.seg_base.data(), T.seg_base.size() * 4);
  memcpy(host.data() + o_cp, T.class_ptr.data(), T.class_ptr.size() * 4);
  if (!T.class_segs.empty())
    memcpy(host.data() + o_cs, T.class_segs.data(), T.class_segs.size() * 4);
  int rc = b.tables.reserve(off);
  if (rc != WBX_OK) return rc;
  unsigned char* base = b.tables.as<unsigned char>();
  WBX_CUDA(cudaMemcpyAsync(base, host.data(), off, cudaMemcpyHostToDevice,
                           ctx->stream));
  WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  b.d_desc = reinterpret_cast<const uint2*>(base + o_desc);
  b.d_seg_base = reinterpret_cast<const int32_t*>(base + o_sb);
  b.d_class_ptr = reinterpret_cast<const int32_t*>(base + o_cp);
  b.d_class_segs = reinterpret_cast<const int32_t*>(base + o_cs);
  b.part = T.part;
  b.S = T.S;
  b.total_segs = T.total_segs;
  b.wx = wx;
  b.n_cols = (p->has_mask ? 1 : 0);
  for (int k = 0; k < (p->has_clim ? 6 : 3); ++k)
    if (p->stat_mask & (1 << k)) ++b.n_cols;
  b.ok = true;
  return WBX_OK;
}

}  // namespace wbx

namespace wbx {

template <bool CLIM, bool MASK, bool SKIPNA, bool PER_ELEM, bool XF = false>
static int launch_variant(wbx_ctx* ctx, const wbx_det_plan* plan,
                          const DetParams& P, int grid) {
  cudaStream_t st = ctx->stream;
  int prc = ctx->prof_begin();
  if (prc != WBX_OK) return prc;
  if (plan->path == kPathTma) {
    auto kern = det_reduce_tma_kernel<CLIM, MASK, SKIPNA, PER_ELEM, XF>;
    WBX_CUDA(cudaFuncSetAttribute(kern,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(plan->smem_bytes)));
    // Programmatic dependent launch: the kernel may be scheduled while the
    // previous kernel of the stream (the finalize of the last step) drains; it
    // waits (griddepcontrol.wait) before it touches the shared record scratch.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->profile ? 0 : 1;
    int stages = plan->stages, stage_bytes = plan->stage_bytes;
    DetParams Pc = P;
    WBX_CUDA(cudaLaunchKernelEx(&cfg, kern, Pc, stages, stage_bytes));
  } else if (plan->path == kPathLdg4) {
    det_reduce_ldg_kernel<CLIM, MASK, SKIPNA, PER_ELEM, 4, XF>
        <<<grid, kLdgThreads, 0, st>>>(P);
  } else {
    det_reduce_ldg_kernel<CLIM, MASK, SKIPNA, true, 1, XF>
        <<<grid, kLdgThreads, 0, st>>>(P);
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return ctx->prof_end();
}

static int launch_bins3(wbx_ctx* ctx, const wbx_det_plan* plan,
                        const DetParams& P, const Bins3Geometry& g,
                        double* records) {
  int prc = ctx->prof_begin();
  if (prc != WBX_OK) return prc;
  const wbx_det_plan::Bins3& b = plan->bins3;
  Bins3Params B;
  B.slot_desc = b.d_desc;
  B.seg_base = b.d_seg_base;
  B.S = b.S;
  B.S_cta = g.S_cta;
  B.J = g.J;
  B.part = b.part;
  B.total_segs = b.total_segs;
  B.n_cols = b.n_cols;
  B.records = records;
#define WBX_BINS3_LAUNCH(A, M, W)                                              \
  do {                                                                         \
    auto kern = det_reduce_bins3_kernel<A, M, W>;                              \
    WBX_CUDA(cudaFuncSetAttribute(kern,                                        \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  static_cast<int>(g.smem)));                  \
    kern<<<g.S_cta * g.J, kTmaThreads, g.smem, ctx->stream>>>(                 \
        P, B, g.stages, g.stage_bytes);                                        \
  } while (0)
  const int key = (plan->has_clim ? 4 : 0) | (plan->has_mask ? 2 : 0) |
                  (b.wx ? 1 : 0);
  switch (key) {
    case 0: WBX_BINS3_LAUNCH(false, false, false); break;
    case 1: WBX_BINS3_LAUNCH(false, false, true); break;
    case 2: WBX_BINS3_LAUNCH(false, true, false); break;
    case 3: WBX_BINS3_LAUNCH(false, true, true); break;
    case 4: WBX_BINS3_LAUNCH(true, false, false); break;
    case 5: WBX_BINS3_LAUNCH(true, false, true); break;
    case 6: WBX_BINS3_LAUNCH(true, true, false); break;
    default: WBX_BINS3_LAUNCH(true, true, true); break;
  }
#undef WBX_BINS3_LAUNCH
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return ctx->prof_end();
}

static size_t bins3_record_bytes(const wbx_det_plan* plan,
                                 const Bins3Geometry& g, int n_cells) {
  return static_cast<size_t>(n_cells + g.J) * plan->bins3.total_segs *
         plan->bins3.n_cols * sizeof(double);
}

static int launch_bins3_finalize(wbx_ctx* ctx, const wbx_det_plan* plan,
                                 const Bins3Geometry& g, const double* records,
                                 const int32_t* d_first, const double* d_cell_w,
                                 int n_cells, long long n_jobs, double* out_ws,
                                 double* out_w, int accumulate) {
  const wbx_det_plan::Bins3& b = plan->bins3;
  Bins3FinalizeParams F;
  F.records = records;
  F.class_ptr = b.d_class_ptr;
  F.class_segs = b.d_class_segs;
  F.cell_first_job = d_first;
  F.cell_class_w = plan->has_mask ? nullptr : d_cell_w;
  F.out_ws = out_ws;
  F.out_w = out_w;
  F.n_jobs = n_jobs;
  F.n_cells = n_cells;
  F.n_classes = plan->n_classes;
  F.J = g.J;
  F.total_segs = b.total_segs;
  F.n_cols = b.n_cols;
  F.stat_mask = plan->stat_mask;
  F.ns = plan->has_clim ? 6 : 3;
  F.accumulate = accumulate;
  if (!accumulate) {
    // statistic slots the launch does not produce stay 0
    WBX_CUDA(cudaMemsetAsync(
        out_ws, 0,
        sizeof(double) * n_cells * plan->n_classes * WBX_NUM_DET_STATS,
        ctx->stream));
    WBX_CUDA(cudaMemsetAsync(
        out_w, 0,
        sizeof(double) * n_cells * plan->n_classes * WBX_NUM_DET_WCLASSES,
        ctx->stream));
  }
  const long long warps =
      static_cast<long long>(n_cells) * plan->n_classes * (b.n_cols + 1);
  const int block = 128;
  const long long blocks = (warps * 32 + block - 1) / block;
  det_bins3_finalize_kernel<<<static_cast<unsigned>(blocks), block, 0,
                              ctx->stream>>>(F);
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

static int launch_main(wbx_ctx* ctx, const wbx_det_plan* plan,
                       const DetParams& P, int grid) {
  if (plan->n_classes > 0) {
    set_error("internal: binned plans have their own launch path");
    return WBX_ERR_INVALID;
  }
  const int key = (plan->xform ? 16 : 0) | (plan->has_clim ? 8 : 0) |
                  (plan->has_mask ? 4 : 0) | (plan->skipna ? 2 : 0) |
                  (plan->per_elem ? 1 : 0);
  switch (key) {
#define WBX_CASE_XF(K, B, C, D) \
  case K:                       \
    return launch_variant<false, B, C, D, true>(ctx, plan, P, grid);
    WBX_CASE_XF(16, false, false, false)
    WBX_CASE_XF(17, false, false, true)
    WBX_CASE_XF(18, false, true, false)
    WBX_CASE_XF(19, false, true, true)
    WBX_CASE_XF(20, true, false, false)
    WBX_CASE_XF(21, true, false, true)
    WBX_CASE_XF(22, true, true, false)
    WBX_CASE_XF(23, true, true, true)
#undef WBX_CASE_XF
#define WBX_CASE(K, A, B, C, D) \
  case K:                       \
    return launch_variant<A, B, C, D>(ctx, plan, P, grid);
    WBX_CASE(0, false, false, false, false)
    WBX_CASE(1, false, false, false, true)
    WBX_CASE(2, false, false, true, false)
    WBX_CASE(3, false, false, true, true)
    WBX_CASE(4, false, true, false, false)
    WBX_CASE(5, false, true, false, true)
    WBX_CASE(6, false, true, true, false)
    WBX_CASE(7, false, true, true, true)
    WBX_CASE(8, true, false, false, false)
    WBX_CASE(9, true, false, false, true)
    WBX_CASE(10, true, false, true, false)
    WBX_CASE(11, true, false, true, true)
    WBX_CASE(12, true, true, false, false)
    WBX_CASE(13, true, true, false, true)
    WBX_CASE(14, true, true, true, false)
    WBX_CASE(15, true, true, true, true)
#undef WBX_CASE
  }
  set_error("internal: bad variant key %d", key);
  return WBX_ERR_INVALID;
}

static void fill_steps(const wbx_det_plan* plan, DetParams* P) {
  const int nx = static_cast<int>(plan->nx);
  const int group = 4 * kConsumerThreads;
  P->group_dq = group / nx;
  P->group_dr = group % nx;
  P->tile_dq = plan->tile / nx;
  P->tile_dr = plan->tile % nx;
  P->stat_mask = plan->stat_mask;
  P->xf_kind = plan->xform;
}

static int grid_for(const wbx_ctx* ctx, const wbx_det_plan* plan,
                    long long total_tiles) {
  long long g = plan->path == kPathTma ? ctx->sm_count : ctx->sm_count * 6ll;
  return static_cast<int>(std::max(1ll, std::min(g, total_tiles)));
}

static int warps_for(const wbx_det_plan* plan) {
  return plan->path == kPathTma ? kConsumerWarps : kLdgWarps;
}

static int launch_finalize(wbx_ctx* ctx, const wbx_det_plan* plan,
                           const double* records, const int32_t* d_first,
                           const double* d_cell_w, int n_cells, int grid_main,
                           long long total_tiles, double* out_ws, double* out_w,
                           int accumulate) {
  FinalizeParams F;
  F.records = records;
  F.cell_first_job = d_first;
  F.cell_w = d_cell_w;
  F.out_ws = out_ws;
  F.out_w = out_w;
  F.total_tiles = total_tiles;
  F.n_cells = n_cells;
  F.grid_main = grid_main;
  F.tiles_per_slab = plan->tiles_per_slab;
  F.warps = warps_for(plan);
  F.n_stats = plan->n_stats;
  F.n_weights = plan->n_weights;
  F.accumulate = accumulate;
  const int slots = WBX_NUM_DET_STATS + WBX_NUM_DET_WCLASSES;
  const long long threads = static_cast<long long>(n_cells) * slots * 32;
  const int block = 128;
  const int grid = static_cast<int>((threads + block - 1) / block);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->profile ? 0 : 1;
    WBX_CUDA(cudaLaunchKernelEx(&cfg, det_finalize_kernel, F));
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

static int ensure_pinned_out(wbx_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_out_cap) return WBX_OK;
  if (ctx->pinned_out) cudaFreeHost(ctx->pinned_out);
  ctx->pinned_out = nullptr;
  ctx->pinned_out_cap = 0;
  WBX_CUDA(cudaHostAlloc(&ctx->pinned_out, bytes, cudaHostAllocDefault));
  ctx->pinned_out_cap = bytes;
  return WBX_OK;
}

// cell prefix (first job of every cell) and constant sum_weights per cell for
// the job range [j0, j1).
static void cell_tables(const wbx_det_plan* plan, int64_t j0, int64_t j1,
                        std::vector<int32_t>* first, std::vector<double>* cw) {
  const int c0 = plan->cell[j0];
  const int c1 = plan->cell[j1 - 1];
  const int n = c1 - c0 + 1;
  first->assign(n + 1, 0);
  cw->assign(n, 0.0);
  for (int64_t j = j0; j < j1; ++j) {
    const int c = plan->cell[j] - c0;
    (*first)[c + 1] = static_cast<int32_t>(j - j0 + 1);
    const double wo = plan->has_wo ? plan->wo[j] : 1.0;
    (*cw)[c] += wo * plan->sum_wy * plan->sum_wx;
  }
  if (plan->n_classes > 0) {
    // constant sum_weights per (cell, class): (sum of w_outer) * class weight
    std::vector<double> wo_sum(n, 0.0);
    for (int64_t j = j0; j < j1; ++j)
      wo_sum[plan->cell[j] - c0] += plan->has_wo ? plan->wo[j] : 1.0;
    cw->assign(static_cast<size_t>(n) * plan->n_classes, 0.0);
    for (int c = 0; c < n; ++c)
      for (int k = 0; k < plan->n_classes; ++k)
        (*cw)[static_cast<size_t>(c) * plan->n_classes + k] =
            wo_sum[c] * plan->class_w[k];
  }
  // cells are dense and non-decreasing, so first[c+1] was set for every c.
}

}  // namespace wbx

using wbx::round_up;

extern "C" {

int wbx_det_plan_create(wbx_ctx* ctx, const wbx_det_desc* d,
                        wbx_det_plan** out) {
  WBX_REQUIRE(ctx && d && out, "wbx_det_plan_create: NULL argument");
  *out = nullptr;
  WBX_REQUIRE(d->space == WBX_SPACE_DEVICE || d->space == WBX_SPACE_HOST,
              "det: bad space %d", d->space);
  WBX_REQUIRE(d->n_jobs >= 1, "det: n_jobs must be >= 1 (got %lld)",
              (long long)d->n_jobs);
  WBX_REQUIRE(d->ny >= 1 && d->nx >= 1, "det: ny, nx must be >= 1");
  WBX_REQUIRE(d->ny * d->nx < (1ll << 30),
              "det: slab of %lld x %lld elements is too large",
              (long long)d->ny, (long long)d->nx);
  WBX_REQUIRE(d->n_jobs < (1ll << 31), "det: too many jobs");
  WBX_REQUIRE(d->pred && d->target && d->cell,
              "det: pred, target and cell tables are required");
  WBX_REQUIRE(d->n_cells >= 1 && d->n_cells <= d->n_jobs,
              "det: n_cells must be in [1, n_jobs]");
  const bool masked = (d->flags & WBX_FLAG_MASKED) != 0;
  WBX_REQUIRE(masked == (d->mask != nullptr),
              "det: WBX_FLAG_MASKED and a mask table must be given together");
  WBX_REQUIRE(d->space == WBX_SPACE_HOST ||
                  !(d->flags & (WBX_FLAG_CLIM_DEVICE | WBX_FLAG_TARGET_DEVICE |
                                WBX_FLAG_MASK_DEVICE)),
              "det: WBX_FLAG_*_DEVICE only apply to WBX_SPACE_HOST plans");
  WBX_REQUIRE(d->cell[0] == 0, "det: cell[0] must be 0");
  for (int64_t j = 1; j < d->n_jobs; ++j) {
    const int step = d->cell[j] - d->cell[j - 1];
    WBX_REQUIRE(step == 0 || step == 1,
                "det: cell[] must be non-decreasing and dense (job %lld)",
                (long long)j);
  }
  WBX_REQUIRE(d->cell[d->n_jobs - 1] == d->n_cells - 1,
              "det: cell[] must end at n_cells - 1");
  for (int64_t j = 0; j < d->n_jobs; ++j) {
    WBX_REQUIRE(d->pred[j] && d->target[j], "det: NULL slab address (job %lld)",
                (long long)j);
    if (d->clim) WBX_REQUIRE(d->clim[j] != 0, "det: NULL clim slab address");
    if (d->mask) WBX_REQUIRE(d->mask[j] != 0, "det: NULL mask slab address");
  }

  if (d->xform != 0) {
    const int kind = d->xform & 3;
    WBX_REQUIRE((d->xform & ~(3 | WBX_XF_PRED_NONZERO | WBX_XF_TARGET_NONZERO)) == 0 &&
                    (kind == WBX_XF_CONTINGENCY || kind == WBX_XF_ERROR_EXCEEDANCE),
                "det: bad xform request %d", d->xform);
    if (kind == WBX_XF_CONTINGENCY) {
      WBX_REQUIRE(((d->xform & WBX_XF_PRED_NONZERO) != 0) == (d->thr_pred == nullptr),
                  "det: xform needs thr_pred unless WBX_XF_PRED_NONZERO is set");
      WBX_REQUIRE(((d->xform & WBX_XF_TARGET_NONZERO) != 0) ==
                      (d->thr_target == nullptr),
                  "det: xform needs thr_target unless WBX_XF_TARGET_NONZERO is "
                  "set");
    } else {
      WBX_REQUIRE(d->xform == WBX_XF_ERROR_EXCEEDANCE && d->thr_pred != nullptr &&
                      d->thr_target == nullptr,
                  "det: error exceedance takes thr_pred only");
    }
    if (d->clim != nullptr || d->n_classes > 0) {
      wbx::set_error("det: xform is not available with clim or class_map");
      return WBX_ERR_UNSUPPORTED;
    }
  } else {
    WBX_REQUIRE(d->thr_pred == nullptr && d->thr_target == nullptr,
                "det: thresholds given without an xform request");
  }

  wbx_det_plan* p = new (std::nothrow) wbx_det_plan();
  if (!p) {
    wbx::set_error("det: out of host memory");
    return WBX_ERR_NOMEM;
  }
  p->space = d->space;
  p->xform = d->xform;
  if (d->thr_pred) p->thr_pred.assign(d->thr_pred, d->thr_pred + d->n_jobs);
  if (d->thr_target)
    p->thr_target.assign(d->thr_target, d->thr_target + d->n_jobs);
  p->flags = d->flags;
  p->n_jobs = d->n_jobs;
  p->ny = d->ny;
  p->nx = d->nx;
  p->n_cells = d->n_cells;
  p->has_clim = d->clim != nullptr;
  p->has_mask = d->mask != nullptr;
  p->skipna = (d->flags & WBX_FLAG_SKIPNA) != 0;
  p->has_wo = d->w_outer != nullptr;
  p->has_wy = d->w_y != nullptr;
  p->has_wx = d->w_x != nullptr;
  p->pred.assign(d->pred, d->pred + d->n_jobs);
  p->target.assign(d->target, d->target + d->n_jobs);
  if (p->has_clim) p->clim.assign(d->clim, d->clim + d->n_jobs);
  if (p->has_mask) p->mask.assign(d->mask, d->mask + d->n_jobs);
  p->cell.assign(d->cell, d->cell + d->n_jobs);
  if (p->has_wo) p->wo.assign(d->w_outer, d->w_outer + d->n_jobs);
  if (p->has_wy) {
    p->wy.assign(d->w_y, d->w_y + d->ny);
    p->sum_wy = 0.0;
    for (double v : p->wy) p->sum_wy += v;
  } else {
    p->sum_wy = static_cast<double>(d->ny);
  }
  if (p->has_wx) {
    p->wx.assign(d->w_x, d->w_x + d->nx);
    p->sum_wx = 0.0;
    for (double v : p->wx) p->sum_wx += v;
  } else {
    p->sum_wx = static_cast<double>(d->nx);
  }
  p->per_elem = p->has_wx || (d->nx % 4) != 0;
  p->stat_mask = d->stat_mask ? (d->stat_mask & 0x3f) : 0x3f;
  if (p->xform) {
    p->stat_mask &= (p->xform & 3) == WBX_XF_CONTINGENCY ? 0xf : 0x1;
  } else if (!p->has_clim) {
    p->stat_mask &= 0x7;
  }
  if (p->stat_mask == 0) {
    delete p;
    wbx::set_error("det: stat_mask selects no statistic this launch can "
                   "evaluate (climatology statistics need clim)");
    return WBX_ERR_INVALID;
  }
  p->n_stats = p->xform ? WBX_NUM_XF_STATS : (p->has_clim ? 6 : 3);
  p->n_weights = p->skipna ? (p->has_clim ? 4 : 1) : (p->has_mask ? 1 : 0);
  p->nacc = p->n_stats + p->n_weights;
  if (d->n_classes > 0) {
    WBX_REQUIRE(d->class_map != nullptr, "det: n_classes > 0 needs a class_map");
    WBX_REQUIRE(d->n_classes <= 256, "det: at most 256 bin classes");
    p->n_classes = d->n_classes;
    const int64_t slab_ = d->ny * d->nx;
    if (p->skipna || (slab_ % 16) != 0) {
      delete p;
      wbx::set_error("det: this binned request needs the generic path "
                     "(skipna, or slab %% 16 != 0)");
      return WBX_ERR_UNSUPPORTED;
    }
    p->class_w.assign(d->n_classes, 0.0);
    for (int64_t e = 0; e < slab_; ++e) {
      const int c = d->class_map[e];
      if (c >= d->n_classes) {
        delete p;
        wbx::set_error("det: class_map value %d >= n_classes", c);
        return WBX_ERR_INVALID;
      }
      p->class_w[c] += (p->has_wy ? d->w_y[e / d->nx] : 1.0) *
                       (p->has_wx ? d->w_x[e % d->nx] : 1.0);
    }
  }

  const int64_t slab = d->ny * d->nx;
  bool aligned = (slab % 4) == 0 && (!p->has_mask || (slab % 16) == 0);
  {
    // operands the kernel reads in place (device space, or the *_DEVICE
    // operands of a host-space plan) must be 16-byte aligned for the TMA path;
    // staged operands land in aligned staging buffers.
    const bool dev = d->space == WBX_SPACE_DEVICE;
    const bool chk_t = dev || (d->flags & WBX_FLAG_TARGET_DEVICE);
    const bool chk_c = p->has_clim && (dev || (d->flags & WBX_FLAG_CLIM_DEVICE));
    const bool chk_m = p->has_mask && (dev || (d->flags & WBX_FLAG_MASK_DEVICE));
    for (int64_t j = 0; j < d->n_jobs && aligned; ++j) {
      aligned = (!dev || (d->pred[j] % 16) == 0) &&
                (!chk_t || (d->target[j] % 16) == 0) &&
                (!chk_c || (d->clim[j] % 16) == 0) &&
                (!chk_m || (d->mask[j] % 16) == 0);
    }
  }
  // tile: 4096 elements (16 KiB per operand) unless the slab is smaller.
  int tile = 4096;
  if (slab < tile) tile = static_cast<int>(round_up(slab, 16));
  p->tile = tile;
  p->tiles_per_slab = static_cast<int>((slab + tile - 1) / tile);
  p->stage_bytes = static_cast<int>(round_up(
      static_cast<size_t>(tile) * 4 * (p->has_clim ? 3 : 2) +
          (p->has_mask ? tile : 0),
      128));
  const size_t overhead =
      2 * wbx::kMaxStages * sizeof(uint64_t) +
      wbx::kMaxStages * sizeof(wbx::StageMeta) + 128;
  const size_t budget = std::min<size_t>(ctx->smem_optin, 227 * 1024) - overhead;
  int stages = static_cast<int>(budget / p->stage_bytes);
  stages = std::min(stages, wbx::kMaxStages);
  // No point in more stages than tiles a CTA will ever see.
  p->stages = stages;
  p->smem_bytes = static_cast<size_t>(stages) * p->stage_bytes + overhead;
  if (d->flags & WBX_FLAG_FORCE_TMA) {
    if (!aligned || stages < 2) {
      delete p;
      wbx::set_error("det: WBX_FLAG_FORCE_TMA but operands are not 16-byte "
                     "aligned / slab %% 4 != 0");
      return WBX_ERR_UNSUPPORTED;
    }
    p->path = wbx::kPathTma;
  } else if (aligned && stages >= 2 && !(d->flags & WBX_FLAG_FORCE_LDG)) {
    p->path = wbx::kPathTma;
  } else if (aligned) {
    p->path = wbx::kPathLdg4;
  } else {
    p->path = wbx::kPathLdg1;
    p->per_elem = true;
  }
  if (p->n_classes > 0 && p->path != wbx::kPathTma) {
    delete p;
    wbx::set_error("det: binned plans need 16-byte aligned operands (TMA path)");
    return WBX_ERR_UNSUPPORTED;
  }

  WBX_CUDA(cudaSetDevice(ctx->device));
  if (p->n_classes > 0) {
    // the class map itself never goes to the GPU: the reduction schedule the
    // host compiles from it does (det_bins3.cuh)
    int rc = wbx::bins3_build(ctx, p, d->class_map);
    if (rc != WBX_OK) { delete p; return rc; }
    if (!p->bins3.ok) {
      delete p;
      wbx::set_error("det: no reduction schedule for this class map");
      return WBX_ERR_UNSUPPORTED;
    }
  }
  // weights (shared by both spaces)
  {
    // layout: w_y | w_x (f64) | w_x rounded to f32, every table 16-byte aligned
    const size_t wx_off = (p->wy.size() + 1) / 2 * 2;
    const size_t wxf_off = (wx_off + p->wx.size() + 1) / 2 * 2;
    const size_t bytes = wxf_off * sizeof(double) +
                         round_up(p->wx.size() * sizeof(float), 16);
    if (p->wy.size() + p->wx.size()) {
      int rc = p->weights.reserve(bytes);
      if (rc != WBX_OK) { delete p; return rc; }
      double* base = p->weights.as<double>();
      if (p->has_wy) {
        WBX_CUDA(cudaMemcpyAsync(base, p->wy.data(),
                                 p->wy.size() * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
        p->d_wy = base;
      }
      if (p->has_wx) {
        WBX_CUDA(cudaMemcpyAsync(base + wx_off, p->wx.data(),
                                 p->wx.size() * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
        p->d_wx = base + wx_off;
        p->wxf.assign(p->wx.begin(), p->wx.end());
        WBX_CUDA(cudaMemcpyAsync(base + wxf_off, p->wxf.data(),
                                 p->wxf.size() * sizeof(float),
                                 cudaMemcpyHostToDevice, ctx->stream));
        p->d_wxf = reinterpret_cast<const float*>(base + wxf_off);
      }
    }
  }

  if (d->space == WBX_SPACE_DEVICE) {
    std::vector<int32_t> first;
    std::vector<double> cw;
    wbx::cell_tables(p, 0, p->n_jobs, &first, &cw);
    const size_t nj = static_cast<size_t>(p->n_jobs);
    size_t off = 0;
    auto take = [&](size_t bytes) {
      size_t o = off;
      off += round_up(bytes, 16);
      return o;
    };
    const size_t o_pred = take(nj * 8), o_tgt = take(nj * 8);
    const size_t o_clim = p->has_clim ? take(nj * 8) : 0;
    const size_t o_mask = p->has_mask ? take(nj * 8) : 0;
    const size_t o_wo = p->has_wo ? take(nj * 8) : 0;
    const size_t o_cell = take(nj * 4);
    const size_t o_first = take(first.size() * 4);
    const size_t o_cw = take(cw.size() * 8);
    const size_t o_thp = p->thr_pred.empty() ? 0 : take(nj * 4);
    const size_t o_tht = p->thr_target.empty() ? 0 : take(nj * 4);
    std::vector<unsigned char>& host = p->chunk_host[0];
    host.assign(off, 0);
    if (!p->thr_pred.empty())
      memcpy(host.data() + o_thp, p->thr_pred.data(), nj * 4);
    if (!p->thr_target.empty())
      memcpy(host.data() + o_tht, p->thr_target.data(), nj * 4);
    memcpy(host.data() + o_pred, p->pred.data(), nj * 8);
    memcpy(host.data() + o_tgt, p->target.data(), nj * 8);
    if (p->has_clim) memcpy(host.data() + o_clim, p->clim.data(), nj * 8);
    if (p->has_mask) memcpy(host.data() + o_mask, p->mask.data(), nj * 8);
    if (p->has_wo) memcpy(host.data() + o_wo, p->wo.data(), nj * 8);
    memcpy(host.data() + o_cell, p->cell.data(), nj * 4);
    memcpy(host.data() + o_first, first.data(), first.size() * 4);
    memcpy(host.data() + o_cw, cw.data(), cw.size() * 8);
    int rc = p->tables.reserve(off);
    if (rc != WBX_OK) { delete p; return rc; }
    unsigned char* base = p->tables.as<unsigned char>();
    WBX_CUDA(cudaMemcpyAsync(base, host.data(), off, cudaMemcpyHostToDevice,
                             ctx->stream));
    WBX_CUDA(cudaStreamSynchronize(ctx->stream));
    wbx::DetParams& P = p->params;
    P.pred = reinterpret_cast<const uint64_t*>(base + o_pred);
    P.target = reinterpret_cast<const uint64_t*>(base + o_tgt);
    P.clim = p->has_clim ? reinterpret_cast<const uint64_t*>(base + o_clim)
                         : nullptr;
    P.mask = p->has_mask ? reinterpret_cast<const uint64_t*>(base + o_mask)
                         : nullptr;
    P.w_outer =
        p->has_wo ? reinterpret_cast<const double*>(base + o_wo) : nullptr;
    P.cell = reinterpret_cast<const int32_t*>(base + o_cell);
    P.w_y = p->d_wy;
    P.w_x = p->d_wx;
    P.w_xf = p->d_wxf;
    P.n_jobs = p->n_jobs;
    P.total_tiles = p->n_jobs * p->tiles_per_slab;
    P.cell_base = 0;
    P.ny = static_cast<int>(p->ny);
    P.nx = static_cast<int>(p->nx);
    P.slab = static_cast<int>(slab);
    P.tile = p->tile;
    P.tiles_per_slab = p->tiles_per_slab;
    P.records = nullptr;
    P.thr_pred = p->thr_pred.empty()
                     ? nullptr
                     : reinterpret_cast<const float*>(base + o_thp);
    P.thr_target = p->thr_target.empty()
                       ? nullptr
                       : reinterpret_cast<const float*>(base + o_tht);
    wbx::fill_steps(p, &P);
    p->d_cell_first_job = reinterpret_cast<const int32_t*>(base + o_first);
    p->d_cell_w = reinterpret_cast<const double*>(base + o_cw);
    p->grid = wbx::grid_for(ctx, p, P.total_tiles);
  } else {
    WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *out = p;
  return WBX_OK;
}

int wbx_det_plan_destroy(wbx_ctx* ctx, wbx_det_plan* plan) {
  if (!plan) return WBX_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
  }
  plan->tables.release_idle();
  plan->weights.release_idle();
  plan->bins3.tables.release_idle();
  delete plan;
  return WBX_OK;
}

static int run_device_space(wbx_ctx* ctx, wbx_det_plan* plan, double* d_ws,
                            double* d_w, int accumulate) {
  if (plan->bins3.ok) {
    const wbx::Bins3Geometry g = wbx::bins3_geometry(ctx, plan, plan->n_jobs);
    if (g.ok) {
      int rc = ctx->records.reserve(wbx::bins3_record_bytes(
          plan, g, static_cast<int>(plan->n_cells)));
      if (rc != WBX_OK) return rc;
      wbx::DetParams P = plan->params;
      P.records = nullptr;
      rc = wbx::launch_bins3(ctx, plan, P, g, ctx->records.as<double>());
      if (rc != WBX_OK) return rc;
      return wbx::launch_bins3_finalize(
          ctx, plan, g, ctx->records.as<double>(), plan->d_cell_first_job,
          plan->d_cell_w, static_cast<int>(plan->n_cells), plan->n_jobs, d_ws,
          d_w, accumulate);
    }
  }
  const int warps = wbx::warps_for(plan);
  const size_t rec_bytes = (static_cast<size_t>(plan->grid) + plan->n_cells) *
                           warps * plan->nacc * sizeof(double);
  int rc = ctx->records.reserve(rec_bytes);
  if (rc != WBX_OK) return rc;
  wbx::DetParams P = plan->params;
  P.records = ctx->records.as<double>();
  rc = wbx::launch_main(ctx, plan, P, plan->grid);
  if (rc != WBX_OK) return rc;
  return wbx::launch_finalize(ctx, plan, P.records, plan->d_cell_first_job,
                              plan->d_cell_w, static_cast<int>(plan->n_cells),
                              plan->grid, P.total_tiles, d_ws, d_w, accumulate);
}

// Host-space: stream chunks of jobs through two staging buffers; copies of
// chunk i+1 overlap the kernel of chunk i.
static int run_host_space(wbx_ctx* ctx, wbx_det_plan* plan, double* d_ws,
                          double* d_w) {
  const size_t slab = static_cast<size_t>(plan->ny * plan->nx);
  const size_t fbytes = slab * 4;
  const size_t mbytes = round_up(slab, 16);
  // WBX_FLAG_CLIM_DEVICE: the climatology rows are already in device memory
  // (it is reused by every chunk of an evaluation); only the fields stream.
  const bool stage_clim =
      plan->has_clim && !(plan->flags & WBX_FLAG_CLIM_DEVICE);
  // WBX_FLAG_TARGET_DEVICE / MASK_DEVICE: targets (and their mask) are kept on
  // the GPU by the caller (rows shared by consecutive chunks); only the
  // predictions cross PCIe.
  const bool stage_tgt = !(plan->flags & WBX_FLAG_TARGET_DEVICE);
  const bool stage_mask =
      plan->has_mask && !(plan->flags & WBX_FLAG_MASK_DEVICE);
  const size_t job_bytes = fbytes * (1 + (stage_tgt ? 1 : 0) +
                                     (stage_clim ? 1 : 0)) +
                           (stage_mask ? mbytes : 0);
  int64_t per_chunk = static_cast<int64_t>((ctx->staging_bytes / 2) / job_bytes);
  per_chunk = std::max<int64_t>(1, std::min<int64_t>(per_chunk, plan->n_jobs));
  for (int b = 0; b < 2; ++b) {
    int rc = ctx->staging[b].reserve(static_cast<size_t>(per_chunk) * job_bytes);
    if (rc != WBX_OK) return rc;
  }
  WBX_CUDA(cudaMemsetAsync(
      d_ws, 0,
      plan->n_cells * plan->cells_mult() * WBX_NUM_DET_STATS * sizeof(double),
      ctx->stream));
  WBX_CUDA(cudaMemsetAsync(
      d_w, 0,
      plan->n_cells * plan->cells_mult() * WBX_NUM_DET_WCLASSES * sizeof(double),
      ctx->stream));
  const int warps = wbx::warps_for(plan);
  int buf = 0;
  for (int64_t j0 = 0; j0 < plan->n_jobs; j0 += per_chunk, buf ^= 1) {
    const int64_t j1 = std::min<int64_t>(plan->n_jobs, j0 + per_chunk);
    const size_t nj = static_cast<size_t>(j1 - j0);
    unsigned char* sbase = ctx->staging[buf].as<unsigned char>();
    unsigned char* s_pred = sbase;
    unsigned char* s_tgt = s_pred + nj * fbytes;
    unsigned char* s_clim = s_tgt + (stage_tgt ? nj * fbytes : 0);
    unsigned char* s_mask = s_clim + (stage_clim ? nj * fbytes : 0);
    // Before overwriting this staging buffer, wait for the kernel that last
    // read it.
    WBX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_compute[buf], 0));
    auto copy_operand = [&](const std::vector<uint64_t>& addr,
                            unsigned char* dst, size_t bytes,
                            size_t stride) -> int {
      size_t j = 0;
      while (j < nj) {
        // merge runs of slabs that are contiguous in host memory.
        size_t run = 1;
        if (stride == bytes) {
          while (j + run < nj &&
                 addr[j0 + j + run] == addr[j0 + j + run - 1] + bytes)
            ++run;
        }
        WBX_CUDA(cudaMemcpyAsync(
            dst + j * stride, reinterpret_cast<const void*>(addr[j0 + j]),
            run * bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        j += run;
      }
      return WBX_OK;
    };
    int rc = copy_operand(plan->pred, s_pred, fbytes, fbytes);
    if (rc != WBX_OK) return rc;
    if (stage_tgt) {
      rc = copy_operand(plan->target, s_tgt, fbytes, fbytes);
      if (rc != WBX_OK) return rc;
    }
    if (stage_clim) {
      rc = copy_operand(plan->clim, s_clim, fbytes, fbytes);
      if (rc != WBX_OK) return rc;
    }
    if (stage_mask) {
      rc = copy_operand(plan->mask, s_mask, slab, mbytes);
      if (rc != WBX_OK) return rc;
    }
    // chunk tables
    std::vector<int32_t> first;
    std::vector<double> cw;
    wbx::cell_tables(plan, j0, j1, &first, &cw);
    size_t off = 0;
    auto take = [&](size_t bytes) {
      size_t o = off;
      off += round_up(bytes, 16);
      return o;
    };
    const size_t o_pred = take(nj * 8), o_tgt = take(nj * 8);
    const size_t o_clim = plan->has_clim ? take(nj * 8) : 0;
    const size_t o_mask = plan->has_mask ? take(nj * 8) : 0;
    const size_t o_wo = plan->has_wo ? take(nj * 8) : 0;
    const size_t o_cell = take(nj * 4);
    const size_t o_first = take(first.size() * 4);
    const size_t o_cw = take(cw.size() * 8);
    const size_t o_thp = plan->thr_pred.empty() ? 0 : take(nj * 4);
    const size_t o_tht = plan->thr_target.empty() ? 0 : take(nj * 4);
    std::vector<unsigned char>& host = plan->chunk_host[buf];
    host.assign(off, 0);
    if (!plan->thr_pred.empty())
      memcpy(host.data() + o_thp, plan->thr_pred.data() + j0, nj * 4);
    if (!plan->thr_target.empty())
      memcpy(host.data() + o_tht, plan->thr_target.data() + j0, nj * 4);
    uint64_t* h_pred = reinterpret_cast<uint64_t*>(host.data() + o_pred);
    uint64_t* h_tgt = reinterpret_cast<uint64_t*>(host.data() + o_tgt);
    uint64_t* h_clim = reinterpret_cast<uint64_t*>(host.data() + o_clim);
    uint64_t* h_mask = reinterpret_cast<uint64_t*>(host.data() + o_mask);
    for (size_t j = 0; j < nj; ++j) {
      h_pred[j] = reinterpret_cast<uint64_t>(s_pred + j * fbytes);
      h_tgt[j] = stage_tgt ? reinterpret_cast<uint64_t>(s_tgt + j * fbytes)
                           : plan->target[j0 + j];
      if (plan->has_clim)
        h_clim[j] = stage_clim
                        ? reinterpret_cast<uint64_t>(s_clim + j * fbytes)
                        : plan->clim[j0 + j];
      if (plan->has_mask)
        h_mask[j] = stage_mask
                        ? reinterpret_cast<uint64_t>(s_mask + j * mbytes)
                        : plan->mask[j0 + j];
    }
    if (plan->has_wo)
      memcpy(host.data() + o_wo, plan->wo.data() + j0, nj * 8);
    memcpy(host.data() + o_cell, plan->cell.data() + j0, nj * 4);
    memcpy(host.data() + o_first, first.data(), first.size() * 4);
    memcpy(host.data() + o_cw, cw.data(), cw.size() * 8);
    rc = ctx->stage_tables[buf].reserve(off);
    if (rc != WBX_OK) return rc;
    unsigned char* tbase = ctx->stage_tables[buf].as<unsigned char>();
    WBX_CUDA(cudaMemcpyAsync(tbase, host.data(), off, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    WBX_CUDA(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));

    wbx::DetParams P{};
    P.pred = reinterpret_cast<const uint64_t*>(tbase + o_pred);
    P.target = reinterpret_cast<const uint64_t*>(tbase + o_tgt);
    P.clim = plan->has_clim ? reinterpret_cast<const uint64_t*>(tbase + o_clim)
                            : nullptr;
    P.mask = plan->has_mask ? reinterpret_cast<const uint64_t*>(tbase + o_mask)
                            : nullptr;
    P.w_outer =
        plan->has_wo ? reinterpret_cast<const double*>(tbase + o_wo) : nullptr;
    P.cell = reinterpret_cast<const int32_t*>(tbase + o_cell);
    P.w_y = plan->d_wy;
    P.w_x = plan->d_wx;
    P.w_xf = plan->d_wxf;
    P.n_jobs = static_cast<long long>(nj);
    P.total_tiles = static_cast<long long>(nj) * plan->tiles_per_slab;
    P.cell_base = plan->cell[j0];
    P.ny = static_cast<int>(plan->ny);
    P.nx = static_cast<int>(plan->nx);
    P.slab = static_cast<int>(slab);
    P.tile = plan->tile;
    P.tiles_per_slab = plan->tiles_per_slab;
    P.thr_pred = plan->thr_pred.empty()
                     ? nullptr
                     : reinterpret_cast<const float*>(tbase + o_thp);
    P.thr_target = plan->thr_target.empty()
                       ? nullptr
                       : reinterpret_cast<const float*>(tbase + o_tht);
    wbx::fill_steps(plan, &P);
    const int grid = wbx::grid_for(ctx, plan, P.total_tiles);
    const int n_cells = static_cast<int>(first.size()) - 1;
    if (plan->bins3.ok) {
      const wbx::Bins3Geometry g =
          wbx::bins3_geometry(ctx, plan, static_cast<long long>(nj));
      if (g.ok) {
        rc = ctx->records.reserve(wbx::bins3_record_bytes(plan, g, n_cells));
        if (rc != WBX_OK) return rc;
        WBX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        rc = wbx::launch_bins3(ctx, plan, P, g, ctx->records.as<double>());
        if (rc != WBX_OK) return rc;
        rc = wbx::launch_bins3_finalize(
            ctx, plan, g, ctx->records.as<double>(),
            reinterpret_cast<const int32_t*>(tbase + o_first),
            reinterpret_cast<const double*>(tbase + o_cw), n_cells,
            static_cast<long long>(nj),
            d_ws + static_cast<size_t>(P.cell_base) * plan->cells_mult() *
                       WBX_NUM_DET_STATS,
            d_w + static_cast<size_t>(P.cell_base) * plan->cells_mult() *
                      WBX_NUM_DET_WCLASSES,
            1);
        if (rc != WBX_OK) return rc;
        WBX_CUDA(cudaEventRecord(ctx->ev_compute[buf], ctx->stream));
        continue;
      }
    }
    const size_t rec_bytes = (static_cast<size_t>(grid) + n_cells) * warps *
                             plan->nacc * sizeof(double);
    // records are reused by consecutive chunks on the same compute stream, so
    // stream order already serialises their use.
    rc = ctx->records.reserve(rec_bytes);
    if (rc != WBX_OK) return rc;
    P.records = ctx->records.as<double>();
    WBX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
    rc = wbx::launch_main(ctx, plan, P, grid);
    if (rc != WBX_OK) return rc;
    rc = wbx::launch_finalize(
        ctx, plan, P.records,
        reinterpret_cast<const int32_t*>(tbase + o_first),
        reinterpret_cast<const double*>(tbase + o_cw), n_cells, grid,
        P.total_tiles,
        d_ws + static_cast<size_t>(P.cell_base) * plan->cells_mult() *
                   WBX_NUM_DET_STATS,
        d_w + static_cast<size_t>(P.cell_base) * plan->cells_mult() *
                  WBX_NUM_DET_WCLASSES,
        1);
    if (rc != WBX_OK) return rc;
    WBX_CUDA(cudaEventRecord(ctx->ev_compute[buf], ctx->stream));
  }
  return WBX_OK;
}

int wbx_det_plan_run(wbx_ctx* ctx, wbx_det_plan* plan, double* sum_ws,
                     double* sum_w, int32_t out_space, int32_t accumulate) {
  WBX_REQUIRE(ctx && plan && sum_ws && sum_w, "wbx_det_plan_run: NULL argument");
  WBX_REQUIRE(out_space == WBX_SPACE_DEVICE || out_space == WBX_SPACE_HOST,
              "wbx_det_plan_run: bad out_space");
  WBX_REQUIRE(!(accumulate && out_space == WBX_SPACE_HOST),
              "wbx_det_plan_run: accumulate needs device outputs");
  WBX_REQUIRE(!(accumulate && plan->space == WBX_SPACE_HOST),
              "wbx_det_plan_run: accumulate is not supported for host-space "
              "plans");
  WBX_CUDA(cudaSetDevice(ctx->device));
  const size_t ws_bytes =
      plan->n_cells * plan->cells_mult() * WBX_NUM_DET_STATS * sizeof(double);
  const size_t w_bytes =
      plan->n_cells * plan->cells_mult() * WBX_NUM_DET_WCLASSES * sizeof(double);
  double* d_ws = sum_ws;
  double* d_w = sum_w;
  if (out_space == WBX_SPACE_HOST) {
    int rc = ctx->out_ws.reserve(ws_bytes);
    if (rc != WBX_OK) return rc;
    rc = ctx->out_w.reserve(w_bytes);
    if (rc != WBX_OK) return rc;
    d_ws = ctx->out_ws.as<double>();
    d_w = ctx->out_w.as<double>();
  }
  int rc = plan->space == WBX_SPACE_DEVICE
               ? run_device_space(ctx, plan, d_ws, d_w, accumulate)
               : run_host_space(ctx, plan, d_ws, d_w);
  if (rc != WBX_OK) return rc;
  if (out_space == WBX_SPACE_HOST) {
    rc = wbx::ensure_pinned_out(ctx, ws_bytes + w_bytes);
    if (rc != WBX_OK) return rc;
    unsigned char* pin = static_cast<unsigned char*>(ctx->pinned_out);
    WBX_CUDA(cudaMemcpyAsync(pin, d_ws, ws_bytes, cudaMemcpyDeviceToHost,
                             ctx->stream));
    WBX_CUDA(cudaMemcpyAsync(pin + ws_bytes, d_w, w_bytes,
                             cudaMemcpyDeviceToHost, ctx->stream));
    WBX_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(sum_ws, pin, ws_bytes);
    memcpy(sum_w, pin + ws_bytes, w_bytes);
  }
  return WBX_OK;
}

int wbx_det_plan_kernel(wbx_ctx* ctx, const wbx_det_plan* plan,
                        int32_t* kernel) {
  WBX_REQUIRE(ctx && plan && kernel, "wbx_det_plan_kernel: NULL argument");
  if (plan->n_classes > 0) {
    (void)ctx;
    *kernel = WBX_KERNEL_BINS_V3;
  } else {
    *kernel = plan->path;
  }
  return WBX_OK;
}

int wbx_bins_schedule_tables(const unsigned char* class_map, int32_t n_classes,
                             int64_t ny, int64_t nx, int32_t part,
                             uint32_t* desc, int32_t* seg_base,
                             int32_t* class_ptr, int32_t* class_segs,
                             int32_t* total_segs) {
  WBX_REQUIRE(class_map && desc && seg_base && class_ptr && class_segs &&
                  total_segs, "wbx_bins_schedule_tables: NULL argument");
  WBX_REQUIRE(n_classes >= 1 && n_classes <= 256 && ny >= 1 && nx >= 1 &&
                  (ny * nx) % 16 == 0 && part >= 16 && part <= 4096 &&
                  part % 16 == 0,
              "wbx_bins_schedule_tables: bad geometry");
  for (int64_t e = 0; e < ny * nx; ++e)
    WBX_REQUIRE(class_map[e] < n_classes,
                "wbx_bins_schedule_tables: class_map value >= n_classes");
  wbx::Bins3Host T;
  if (!wbx::bins3_schedule(class_map, n_classes, ny * nx, part, &T)) {
    wbx::set_error("wbx_bins_schedule_tables: a part of %d elements needs more "
                   "than %d slots", part, wbx::kBins3Slots);
    return WBX_ERR_UNSUPPORTED;
  }
  memcpy(desc, T.desc.data(), T.desc.size() * sizeof(uint32_t));
  memcpy(seg_base, T.seg_base.data(), T.seg_base.size() * sizeof(int32_t));
  memcpy(class_ptr, T.class_ptr.data(), T.class_ptr.size() * sizeof(int32_t));
  if (!T.class_segs.empty())
    memcpy(class_segs, T.class_segs.data(),
           T.class_segs.size() * sizeof(int32_t));
  *total_segs = T.total_segs;
  return WBX_OK;
}

int wbx_det_reduce(wbx_ctx* ctx, const wbx_det_desc* desc, double* sum_ws,
                   double* sum_w) {
  wbx_det_plan* plan = nullptr;
  int rc = wbx_det_plan_create(ctx, desc, &plan);
  if (rc != WBX_OK) return rc;
  rc = wbx_det_plan_run(ctx, plan, sum_ws, sum_w, WBX_SPACE_HOST, 0);
  wbx_det_plan_destroy(ctx, plan);
  return rc;
}

int wbx_det_elementwise(wbx_ctx* ctx, int32_t stat, const float* pred,
                        const float* target, const float* clim, int64_t n,
                        float* out) {
  WBX_REQUIRE(ctx && pred && target && out, "wbx_det_elementwise: NULL argument");
  const int code = stat & 0xff;
  WBX_REQUIRE(stat >= 0 && (stat & ~(0xff | WBX_EW_ACCUMULATE)) == 0 &&
                  (code < WBX_NUM_DET_STATS || code == WBX_EW_PASS_PRED ||
                   code == WBX_EW_PASS_PRED_NAN_TARGET),
              "wbx_det_elementwise: bad statistic %d", stat);
  WBX_REQUIRE(code < WBX_STAT_SQ_PRED_ANOM || code >= WBX_NUM_DET_STATS ||
                  clim != nullptr,
              "wbx_det_elementwise: statistic %d needs a climatology", stat);
  WBX_REQUIRE(n >= 0, "wbx_det_elementwise: negative n");
  if (n == 0) return WBX_OK;
  WBX_CUDA(cudaSetDevice(ctx->device));
  const int block = 256;
  const long long want = (n + block - 1) / block;
  const int grid = static_cast<int>(
      std::min<long long>(want, static_cast<long long>(ctx->sm_count) * 16));
  wbx::det_elementwise_kernel<<<grid, block, 0, ctx->stream>>>(
      stat, pred, target, clim, n, out);
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

int wbx_xf_elementwise(wbx_ctx* ctx, int32_t xform, int32_t slot,
                       float thr_pred, float thr_target, const float* pred,
                       const float* target, int64_t n, float* out) {
  WBX_REQUIRE(ctx && pred && out, "wbx_xf_elementwise: NULL argument");
  const int kind = xform & 3;
  WBX_REQUIRE((xform & ~(3 | WBX_XF_PRED_NONZERO | WBX_XF_TARGET_NONZERO)) == 0 &&
                  (kind == WBX_XF_CONTINGENCY || kind == WBX_XF_ERROR_EXCEEDANCE),
              "wbx_xf_elementwise: bad xform request %d", xform);
  WBX_REQUIRE(slot >= 0 && slot <= WBX_XF_BINARIZED_PRED &&
                  (kind == WBX_XF_CONTINGENCY || slot == 0),
              "wbx_xf_elementwise: bad slot %d", slot);
  WBX_REQUIRE(target != nullptr || slot == WBX_XF_BINARIZED_PRED,
              "wbx_xf_elementwise: slot %d needs the targets", slot);
  WBX_REQUIRE(n >= 0, "wbx_xf_elementwise: negative n");
  if (n == 0) return WBX_OK;
  WBX_CUDA(cudaSetDevice(ctx->device));
  const int block = 256;
  const long long want = (n + block - 1) / block;
  const int grid = static_cast<int>(
      std::min<long long>(want, static_cast<long long>(ctx->sm_count) * 16));
  wbx::xf_elementwise_kernel<<<grid, block, 0, ctx->stream>>>(
      xform, slot, thr_pred, thr_target, pred, target, n, out);
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

}  // extern "C"
