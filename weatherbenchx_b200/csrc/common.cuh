// Shared device/host helpers for libwbx_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/wbx_b200.h"

namespace wbx {

// ---------------------------------------------------------------------------
// Error plumbing (host)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what, const char* file, int line);

#define WBX_CUDA(call)                                                   \
  do {                                                                   \
    cudaError_t err__ = (call);                                          \
    if (err__ != cudaSuccess)                                            \
      return ::wbx::cuda_fail(err__, #call, __FILE__, __LINE__);         \
  } while (0)

#define WBX_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::wbx::set_error(__VA_ARGS__);      \
      return WBX_ERR_INVALID;             \
    }                                     \
  } while (0)

// Growable device buffer owned by a context.
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();       // cudaFree (implicit device synchronisation)
  void release_idle();  // caller guarantees no work uses it: may be recycled
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(ptr); }
};

}  // namespace wbx

struct wbx_ctx {
  int device = 0;
  int sm_count = 0;
  uint64_t hbm_bytes = 0;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t stream = nullptr;  // active compute stream
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_compute[2] = {nullptr, nullptr};
  uint64_t launches = 0;
  uint64_t staging_bytes = 1ull << 30;
  wbx::DevBuf records;     // per-(CTA, cell, warp) partial sums
  wbx::DevBuf out_ws, out_w;  // device result staging for host outputs
  wbx::DevBuf staging[2];  // host-space slab staging (double buffered)
  wbx::DevBuf stage_tables[2];
  void* pinned_out = nullptr;
  size_t pinned_out_cap = 0;
  // job-table ring of the asynchronous one-shot calls (wbx_zonal_spectrum):
  // a slot is reused only after the kernel that read it has finished
  static constexpr int kTableRing = 8;
  wbx::DevBuf ring_tables[kTableRing];
  cudaEvent_t ring_ev[kTableRing] = {};
  int ring_next = 0;
  // optional per-kernel timing (wbx_ctx_profile)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_ms = 0.0;
  uint64_t prof_count = 0;
  int prof_begin();          // records the start event on `stream`
  int prof_end();            // records the stop event on `stream`
  int prof_collect();        // folds finished event pairs into prof_ms
};

// ---------------------------------------------------------------------------
// Device-side primitives: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
namespace wbx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// L2 policy for read-once streams.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
               : "=l"(pol));
  return pol;
}

// global -> shared bulk copy through the TMA engine; completion is signalled
// on `bar` as transaction bytes.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src,
                                         uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile(
      "ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "l"(p));
  return r;
}

__device__ __forceinline__ float2 ldg_stream_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];"
               : "=f"(r.x), "=f"(r.y)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

}  // namespace wbx
#endif  // __CUDACC__
