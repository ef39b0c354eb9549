// Generic strided statistic + weighted aggregation (wbx_reduce_generic).
//
// Serves every Aggregator configuration the slab kernel cannot express:
// arbitrary dim order, broadcasting inside reduced dims, N-d weights, bin
// masks (aggregation.py:320-335), already materialised statistics.  One thread
// block per (output cell, chunk of the reduced index range), f64 accumulation,
// partials combined in a fixed order (bit-stable).  When there are very many
// output cells (e.g. only time is reduced, lat/lon kept) one *thread* owns a
// cell instead.
#include <algorithm>
#include <string.h>

#include "common.cuh"

namespace wbx {

struct GenOperand {
  const void* ptr;
  long long ks[WBX_MAX_DIMS];  // strides over kept dims
  long long rs[WBX_MAX_DIMS];  // strides over reduced dims
  int dtype;
};

struct GenParams {
  int nk, nr;
  long long ksize[WBX_MAX_DIMS];
  long long rsize[WBX_MAX_DIMS];
  long long n_cells, n_red;
  int op, skipna, n_factors, chunks;
  GenOperand a, b, c, mask;
  GenOperand f[WBX_MAX_FACTORS];
  double* part_ws;
  double* part_w;
};

__device__ __forceinline__ float gen_stat(int op, float a, float b, float c) {
  switch (op) {
    case WBX_STAT_ERROR: return __fsub_rn(a, b);
    case WBX_STAT_ABS_ERROR: return fabsf(__fsub_rn(a, b));
    case WBX_STAT_SQ_ERROR: {
      const float d = __fsub_rn(a, b);
      return __fmul_rn(d, d);
    }
    case WBX_STAT_SQ_PRED_ANOM: {
      const float x = __fsub_rn(a, c);
      return __fmul_rn(x, x);
    }
    case WBX_STAT_SQ_TGT_ANOM: {
      const float y = __fsub_rn(b, c);
      return __fmul_rn(y, y);
    }
    case WBX_STAT_ANOM_COV:
      return __fmul_rn(__fsub_rn(a, c), __fsub_rn(b, c));
    default: return a;
  }
}

__device__ __forceinline__ double gen_factor(const GenOperand& f,
                                             long long off) {
  switch (f.dtype) {
    case WBX_DTYPE_F64: return static_cast<const double*>(f.ptr)[off];
    case WBX_DTYPE_F32:
      return static_cast<double>(static_cast<const float*>(f.ptr)[off]);
    default:
      return static_cast<const unsigned char*>(f.ptr)[off] ? 1.0 : 0.0;
  }
}

struct GenOffsets {
  long long a, b, c, m;
  long long f[WBX_MAX_FACTORS];
};

__device__ __forceinline__ void gen_offsets_kept(const GenParams& P,
                                                 long long cell,
                                                 GenOffsets* o) {
  o->a = o->b = o->c = o->m = 0;
  for (int k = 0; k < WBX_MAX_FACTORS; ++k) o->f[k] = 0;
  for (int d = P.nk - 1; d >= 0; --d) {
    const long long i = cell % P.ksize[d];
    cell /= P.ksize[d];
    o->a += i * P.a.ks[d];
    o->b += i * P.b.ks[d];
    o->c += i * P.c.ks[d];
    o->m += i * P.mask.ks[d];
    for (int k = 0; k < P.n_factors; ++k) o->f[k] += i * P.f[k].ks[d];
  }
}

__device__ __forceinline__ void gen_accumulate(const GenParams& P,
                                               const GenOffsets& base,
                                               long long r, double* ws,
                                               double* w) {
  long long oa = base.a, ob = base.b, oc = base.c, om = base.m;
  long long of[WBX_MAX_FACTORS];
  for (int k = 0; k < P.n_factors; ++k) of[k] = base.f[k];
  for (int d = P.nr - 1; d >= 0; --d) {
    const long long i = r % P.rsize[d];
    r /= P.rsize[d];
    oa += i * P.a.rs[d];
    ob += i * P.b.rs[d];
    oc += i * P.c.rs[d];
    om += i * P.mask.rs[d];
    for (int k = 0; k < P.n_factors; ++k) of[k] += i * P.f[k].rs[d];
  }
  const float av = static_cast<const float*>(P.a.ptr)[oa];
  float v = av;
  if (P.op >= 0) {
    const float bv = static_cast<const float*>(P.b.ptr)[ob];
    const float cv = P.c.ptr ? static_cast<const float*>(P.c.ptr)[oc] : 0.f;
    v = gen_stat(P.op, av, bv, cv);
  }
  bool valid = true;
  if (P.mask.ptr) valid = static_cast<const unsigned char*>(P.mask.ptr)[om] != 0;
  if (P.skipna) valid = valid && (v == v);
  double f = 1.0;
  for (int k = 0; k < P.n_factors; ++k) f *= gen_factor(P.f[k], of[k]);
  // where(valid, stat, 0) * f  and  valid * f, exactly as the reference's
  // zero-fill + einsum (a NaN weight still poisons the sum, as there).
  *ws += (valid ? static_cast<double>(v) : 0.0) * f;
  *w += (valid ? 1.0 : 0.0) * f;
}

constexpr int kGenThreads = 128;

// one block per (cell, chunk)
__global__ void __launch_bounds__(kGenThreads)
    generic_reduce_block_kernel(const GenParams P) {
  const long long cell = blockIdx.x / P.chunks;
  const int chunk = static_cast<int>(blockIdx.x - cell * P.chunks);
  GenOffsets base;
  gen_offsets_kept(P, cell, &base);
  const long long r0 = (P.n_red * chunk) / P.chunks;
  const long long r1 = (P.n_red * (chunk + 1)) / P.chunks;
  double ws = 0.0, w = 0.0;
  for (long long r = r0 + threadIdx.x; r < r1; r += kGenThreads)
    gen_accumulate(P, base, r, &ws, &w);
  ws = warp_sum(ws);
  w = warp_sum(w);
  __shared__ double s_ws[kGenThreads / 32], s_w[kGenThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_ws[warp] = ws;
    s_w[warp] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tws = 0.0, tw = 0.0;
    for (int i = 0; i < kGenThreads / 32; ++i) {
      tws += s_ws[i];
      tw += s_w[i];
    }
    P.part_ws[blockIdx.x] = tws;
    P.part_w[blockIdx.x] = tw;
  }
}

// one thread per cell (many cells, few reduced elements each)
__global__ void __launch_bounds__(kGenThreads)
    generic_reduce_thread_kernel(const GenParams P) {
  const long long cell =
      blockIdx.x * static_cast<long long>(kGenThreads) + threadIdx.x;
  if (cell >= P.n_cells) return;
  GenOffsets base;
  gen_offsets_kept(P, cell, &base);
  double ws = 0.0, w = 0.0;
  for (long long r = 0; r < P.n_red; ++r) gen_accumulate(P, base, r, &ws, &w);
  P.part_ws[cell] = ws;
  P.part_w[cell] = w;
}

__global__ void generic_combine_kernel(const double* part_ws,
                                       const double* part_w, long long n_cells,
                                       int chunks, double* out_ws,
                                       double* out_w) {
  const long long cell =
      blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (cell >= n_cells) return;
  double ws = 0.0, w = 0.0;
  for (int k = 0; k < chunks; ++k) {
    ws += part_ws[cell * chunks + k];
    w += part_w[cell * chunks + k];
  }
  out_ws[cell] = ws;
  out_w[cell] = w;
}

static void split_strides(const int64_t* stride, const std::vector<int>& kept,
                          const std::vector<int>& red, GenOperand* o) {
  for (int i = 0; i < WBX_MAX_DIMS; ++i) o->ks[i] = o->rs[i] = 0;
  for (size_t i = 0; i < kept.size(); ++i) o->ks[i] = stride[kept[i]];
  for (size_t i = 0; i < red.size(); ++i) o->rs[i] = stride[red[i]];
}

}  // namespace wbx

extern "C" int wbx_reduce_generic(wbx_ctx* ctx, const wbx_generic_desc* d,
                                  double* sum_ws, double* sum_w,
                                  int32_t out_space) {
  using namespace wbx;
  WBX_REQUIRE(ctx && d && sum_ws && sum_w, "wbx_reduce_generic: NULL argument");
  WBX_REQUIRE(d->ndim >= 1 && d->ndim <= WBX_MAX_DIMS,
              "generic: ndim %d out of range", d->ndim);
  WBX_REQUIRE(d->n_factors >= 0 && d->n_factors <= WBX_MAX_FACTORS,
              "generic: n_factors out of range");
  WBX_REQUIRE(d->op >= -1 && d->op < WBX_NUM_DET_STATS, "generic: bad op");
  WBX_REQUIRE(d->a != nullptr, "generic: operand a is NULL");
  WBX_REQUIRE(d->op < 0 || d->b != nullptr, "generic: operand b is NULL");
  WBX_REQUIRE(d->op < WBX_STAT_SQ_PRED_ANOM || d->c != nullptr,
              "generic: statistic %d needs a climatology", d->op);
  WBX_REQUIRE(out_space == WBX_SPACE_DEVICE || out_space == WBX_SPACE_HOST,
              "generic: bad out_space");
  std::vector<int> kept, red;
  GenParams P;
  memset(&P, 0, sizeof(P));
  P.n_cells = 1;
  P.n_red = 1;
  for (int i = 0; i < d->ndim; ++i) {
    WBX_REQUIRE(d->size[i] >= 1, "generic: empty dim %d", i);
    if (d->reduced[i]) {
      P.rsize[red.size()] = d->size[i];
      red.push_back(i);
      P.n_red *= d->size[i];
    } else {
      P.ksize[kept.size()] = d->size[i];
      kept.push_back(i);
      P.n_cells *= d->size[i];
    }
  }
  P.nk = static_cast<int>(kept.size());
  P.nr = static_cast<int>(red.size());
  P.op = d->op;
  P.skipna = (d->flags & WBX_FLAG_SKIPNA) ? 1 : 0;
  P.n_factors = d->n_factors;
  static const int64_t zeros[WBX_MAX_DIMS] = {0};
  P.a.ptr = d->a;
  split_strides(d->a_stride, kept, red, &P.a);
  P.b.ptr = d->b;
  split_strides(d->b ? d->b_stride : zeros, kept, red, &P.b);
  P.c.ptr = d->c;
  split_strides(d->c ? d->c_stride : zeros, kept, red, &P.c);
  P.mask.ptr = d->mask;
  split_strides(d->mask ? d->mask_stride : zeros, kept, red, &P.mask);
  for (int k = 0; k < d->n_factors; ++k) {
    WBX_REQUIRE(d->factor[k] != nullptr, "generic: factor %d is NULL", k);
    WBX_REQUIRE(d->factor_dtype[k] >= 0 && d->factor_dtype[k] <= 2,
                "generic: factor %d has bad dtype", k);
    P.f[k].ptr = d->factor[k];
    P.f[k].dtype = d->factor_dtype[k];
    split_strides(d->factor_stride[k], kept, red, &P.f[k]);
  }
  WBX_CUDA(cudaSetDevice(ctx->device));
  const bool per_thread = P.n_cells >= 16384 || P.n_red <= 64;
  int chunks = 1;
  if (!per_thread) {
    const long long want = 4ll * ctx->sm_count;
    chunks = static_cast<int>(std::max(1ll, (want + P.n_cells - 1) / P.n_cells));
    const long long max_chunks = std::max(1ll, P.n_red / (kGenThreads * 8));
    chunks = static_cast<int>(std::min<long long>(chunks, max_chunks));
  }
  P.chunks = chunks;
  const size_t part = static_cast<size_t>(P.n_cells) * chunks * sizeof(double);
  int rc = ctx->records.reserve(2 * part);
  if (rc != WBX_OK) return rc;
  P.part_ws = ctx->records.as<double>();
  P.part_w = P.part_ws + static_cast<size_t>(P.n_cells) * chunks;
  double* d_ws = sum_ws;
  double* d_w = sum_w;
  const size_t out_bytes = static_cast<size_t>(P.n_cells) * sizeof(double);
  if (out_space == WBX_SPACE_HOST) {
    rc = ctx->out_ws.reserve(out_bytes);
    if (rc != WBX_OK) return rc;
    rc = ctx->out_w.reserve(out_bytes);
    if (rc != WBX_OK) return rc;
    d_ws = ctx->out_ws.as<double>();
    d_w = ctx->out_w.as<double>();
  }
  if (per_thread) {
    const long long blocks = (P.n_cells + kGenThreads - 1) / kGenThreads;
    WBX_REQUIRE(blocks < (1ll << 31), "generic: too many cells");
    generic_reduce_thread_kernel<<<static_cast<unsigned>(blocks), kGenThreads,
                                   0, ctx->stream>>>(P);
  } else {
    const long long blocks = P.n_cells * chunks;
    generic_reduce_block_kernel<<<static_cast<unsigned>(blocks), kGenThreads, 0,
                                  ctx->stream>>>(P);
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  {
    const int block = 128;
    const long long blocks = (P.n_cells + block - 1) / block;
    generic_combine_kernel<<<static_cast<unsigned>(blocks), block, 0,
                             ctx->stream>>>(P.part_ws, P.part_w, P.n_cells,
                                            chunks, d_ws, d_w);
    WBX_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (out_space == WBX_SPACE_HOST) {
    WBX_CUDA(cudaMemcpyAsync(sum_ws, d_ws, out_bytes, cudaMemcpyDeviceToHost,
                             ctx->stream));
    WBX_CUDA(cudaMemcpyAsync(sum_w, d_w, out_bytes, cudaMemcpyDeviceToHost,
                             ctx->stream));
    WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return WBX_OK;
}
