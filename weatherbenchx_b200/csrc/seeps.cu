// SEEPS per-gridpoint field (metrics/categorical.py:104-304).  C ABI in
// include/wbx_b200.h.  The score of a point depends on two per-point
// parameters -- the climatological wet threshold of the valid time and the dry
// fraction p1 of the location -- so it is evaluated as a field by this
// elementwise kernel (12 B read + 4 B written per point; p1 is a slab shared by
// all times and stays in L2) and then aggregated by the fused masked reduction
// like any other statistic.
#include <algorithm>

#include "common.cuh"
#include "seeps_point.h"

namespace wbx {

// One thread per four consecutive points when everything is 16-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(256)
    seeps_elementwise_kernel(const float* __restrict__ pred,
                             const float* __restrict__ target,
                             const float* __restrict__ wet,
                             const float* __restrict__ p1,
                             const long long p1_len, const float dry_threshold,
                             const long long n, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  // position inside the p1 slab, advanced without a division per element:
  // pos < p1_len and step < p1_len, so one conditional subtraction wraps it
  if constexpr (VEC == 4) {
    const long long nvec = n >> 2;
    long long pos = (4 * i) % p1_len;
    const long long step = (4 * stride) % p1_len;
    for (; i < nvec; i += stride) {
      const float4 p = ldg_stream_f4(pred + 4 * i);
      const float4 t = ldg_stream_f4(target + 4 * i);
      const float4 w = ldg_stream_f4(wet + 4 * i);
      // p1_len % 4 == 0 on this path: the group stays inside one p1 slab
      const float4 q = __ldg(reinterpret_cast<const float4*>(p1 + pos));
      float4 r;
      r.x = wbx_seeps_point(p.x, t.x, w.x, q.x, dry_threshold);
      r.y = wbx_seeps_point(p.y, t.y, w.y, q.y, dry_threshold);
      r.z = wbx_seeps_point(p.z, t.z, w.z, q.z, dry_threshold);
      r.w = wbx_seeps_point(p.w, t.w, w.w, q.w, dry_threshold);
      __stcs(reinterpret_cast<float4*>(out) + i, r);
      pos += step;
      if (pos >= p1_len) pos -= p1_len;
    }
  } else {
    long long pos = i % p1_len;
    const long long step = stride % p1_len;
    for (; i < n; i += stride) {
      out[i] = wbx_seeps_point(pred[i], target[i], wet[i], __ldg(p1 + pos),
                               dry_threshold);
      pos += step;
      if (pos >= p1_len) pos -= p1_len;
    }
  }
}

}  // namespace wbx

extern "C" {

int wbx_seeps_elementwise(wbx_ctx* ctx, const float* pred, const float* target,
                          const float* wet_threshold, const float* p1,
                          int64_t p1_len, float dry_threshold, int64_t n,
                          float* out) {
  WBX_REQUIRE(ctx && pred && target && wet_threshold && p1 && out,
              "wbx_seeps_elementwise: NULL argument");
  WBX_REQUIRE(n >= 0, "wbx_seeps_elementwise: negative n");
  WBX_REQUIRE(p1_len >= 1 && (n % p1_len) == 0,
              "wbx_seeps_elementwise: n (%lld) must be a multiple of p1_len "
              "(%lld)", (long long)n, (long long)p1_len);
  if (n == 0) return WBX_OK;
  WBX_CUDA(cudaSetDevice(ctx->device));
  const auto misaligned = [](const void* ptr) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) != 0;
  };
  const bool vec = (p1_len % 4) == 0 && !misaligned(pred) &&
                   !misaligned(target) && !misaligned(wet_threshold) &&
                   !misaligned(p1) && !misaligned(out);
  const int block = 256;
  const long long work = vec ? n / 4 : n;
  const int grid = static_cast<int>(std::max<long long>(
      1, std::min<long long>((work + block - 1) / block,
                             static_cast<long long>(ctx->sm_count) * 16)));
  if (vec) {
    wbx::seeps_elementwise_kernel<4><<<grid, block, 0, ctx->stream>>>(
        pred, target, wet_threshold, p1, p1_len, dry_threshold, n, out);
  } else {
    wbx::seeps_elementwise_kernel<1><<<grid, block, 0, ctx->stream>>>(
        pred, target, wet_threshold, p1, p1_len, dry_threshold, n, out);
  }
  WBX_CUDA(cudaGetLastError());
  ctx->launches++;
  return WBX_OK;
}

}  // extern "C"
