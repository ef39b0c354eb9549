// Context, error handling and pinned-memory helpers of libwbx_b200.
#include <stddef.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

#include <mutex>
#include <vector>

namespace wbx {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", static_cast<int>(err),
            cudaGetErrorString(err), what, file, line);
  if (err == cudaErrorMemoryAllocation) return WBX_ERR_NOMEM;
  if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver)
    return WBX_ERR_NO_DEVICE;
  return WBX_ERR_CUDA;
}

// Small device blocks (job tables, weight vectors of a plan) are recycled
// through a process-wide free list instead of cudaMalloc / cudaFree: a
// streamed evaluation creates and retires one plan per chunk, and cudaFree
// synchronises the whole device -- including the copies and kernels other
// contexts (pipeline lanes) have in flight.  Only release_idle() feeds the
// list: its caller has synchronised every stream that touched the buffer
// (plan destruction), so a recycled block is never still in use.
namespace {
struct PooledBlock {
  void* ptr;
  size_t cap;
  int device;
};
constexpr size_t kPoolMaxBlockBytes = 4u << 20;
constexpr size_t kPoolMaxBlocks = 256;
std::mutex g_pool_mutex;
std::vector<PooledBlock> g_pool;
}  // namespace

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return WBX_OK;
  release();
  const size_t want = bytes + bytes / 4 + 256;
  if (want <= kPoolMaxBlockBytes) {
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (size_t i = 0; i < g_pool.size(); ++i) {
      if (g_pool[i].device == device && g_pool[i].cap >= want &&
          g_pool[i].cap <= 4 * want) {
        ptr = g_pool[i].ptr;
        cap = g_pool[i].cap;
        g_pool[i] = g_pool.back();
        g_pool.pop_back();
        return WBX_OK;
      }
    }
  }
  cudaError_t err = cudaMalloc(&ptr, want);
  if (err != cudaSuccess) {
    ptr = nullptr;
    return cuda_fail(err, "cudaMalloc", __FILE__, __LINE__);
  }
  cap = want;
  return WBX_OK;
}

void DevBuf::release() {
  if (ptr) cudaFree(ptr);  // synchronises the device: nothing can still use it
  ptr = nullptr;
  cap = 0;
}

void DevBuf::release_idle() {
  if (ptr) {
    bool pooled = false;
    if (cap <= kPoolMaxBlockBytes) {
      int device = 0;
      cudaGetDevice(&device);
      std::lock_guard<std::mutex> lock(g_pool_mutex);
      if (g_pool.size() < kPoolMaxBlocks) {
        g_pool.push_back({ptr, cap, device});
        pooled = true;
      }
    }
    if (!pooled) cudaFree(ptr);
  }
  ptr = nullptr;
  cap = 0;
}

}  // namespace wbx

int wbx_ctx::prof_begin() {
  if (!profile) return WBX_OK;
  if (prof_used == prof_events.size()) {
    if (prof_used >= 4096) {
      int rc = prof_collect();
      if (rc != WBX_OK) return rc;
    }
    if (prof_used == prof_events.size()) {
      cudaEvent_t a, b;
      WBX_CUDA(cudaEventCreate(&a));
      WBX_CUDA(cudaEventCreate(&b));
      prof_events.emplace_back(a, b);
    }
  }
  WBX_CUDA(cudaEventRecord(prof_events[prof_used].first, stream));
  return WBX_OK;
}

int wbx_ctx::prof_end() {
  if (!profile) return WBX_OK;
  WBX_CUDA(cudaEventRecord(prof_events[prof_used].second, stream));
  ++prof_used;
  return WBX_OK;
}

int wbx_ctx::prof_collect() {
  for (size_t i = 0; i < prof_used; ++i) {
    WBX_CUDA(cudaEventSynchronize(prof_events[i].second));
    float ms = 0.f;
    WBX_CUDA(cudaEventElapsedTime(&ms, prof_events[i].first,
                                  prof_events[i].second));
    prof_ms += ms;
    ++prof_count;
  }
  prof_used = 0;
  return WBX_OK;
}

extern "C" {

int wbx_ctx_profile(wbx_ctx* ctx, int32_t enable) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_profile: ctx is NULL");
  if (!enable && ctx->profile) {
    int rc = ctx->prof_collect();
    if (rc != WBX_OK) return rc;
  }
  ctx->profile = enable != 0;
  return WBX_OK;
}

int wbx_ctx_kernel_time(wbx_ctx* ctx, double* total_ms, uint64_t* count,
                        int32_t reset) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_kernel_time: ctx is NULL");
  WBX_CUDA(cudaSetDevice(ctx->device));
  int rc = ctx->prof_collect();
  if (rc != WBX_OK) return rc;
  if (total_ms) *total_ms = ctx->prof_ms;
  if (count) *count = ctx->prof_count;
  if (reset) {
    ctx->prof_ms = 0.0;
    ctx->prof_count = 0;
  }
  return WBX_OK;
}

int wbx_abi_version(void) { return WBX_ABI_VERSION; }

const char* wbx_last_error(void) { return wbx::g_error; }

int wbx_ctx_create(int device, wbx_ctx** out) {
  WBX_REQUIRE(out != nullptr, "wbx_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    wbx::set_error("wbx_ctx_create: no CUDA device available (%s)",
                   cudaGetErrorString(err));
    return WBX_ERR_NO_DEVICE;
  }
  WBX_REQUIRE(device >= 0 && device < count,
              "wbx_ctx_create: device %d out of range [0, %d)", device, count);
  WBX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  WBX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    wbx::set_error(
        "wbx_ctx_create: device %d is sm_%d%d; libwbx_b200 is built for "
        "sm_100a (B200) only",
        device, prop.major, prop.minor);
    return WBX_ERR_UNSUPPORTED;
  }
  wbx_ctx* ctx = new (std::nothrow) wbx_ctx();
  if (!ctx) {
    wbx::set_error("wbx_ctx_create: out of host memory");
    return WBX_ERR_NOMEM;
  }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->hbm_bytes = prop.totalGlobalMem;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  WBX_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  WBX_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    WBX_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
    WBX_CUDA(
        cudaEventCreateWithFlags(&ctx->ev_compute[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < wbx_ctx::kTableRing; ++i)
    WBX_CUDA(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return WBX_OK;
}

int wbx_ctx_destroy(wbx_ctx* ctx) {
  if (!ctx) return WBX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->own_stream);
  cudaStreamSynchronize(ctx->copy_stream);
  ctx->records.release();
  ctx->out_ws.release();
  ctx->out_w.release();
  for (int i = 0; i < 2; ++i) {
    ctx->staging[i].release();
    ctx->stage_tables[i].release();
    if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    if (ctx->ev_compute[i]) cudaEventDestroy(ctx->ev_compute[i]);
  }
  for (int i = 0; i < wbx_ctx::kTableRing; ++i) {
    ctx->ring_tables[i].release();
    if (ctx->ring_ev[i]) cudaEventDestroy(ctx->ring_ev[i]);
  }
  for (auto& pr : ctx->prof_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (ctx->pinned_out) cudaFreeHost(ctx->pinned_out);
  cudaStreamDestroy(ctx->own_stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return WBX_OK;
}

int wbx_ctx_set_stream(wbx_ctx* ctx, void* cuda_stream) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_set_stream: ctx is NULL");
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream)
                            : ctx->own_stream;
  return WBX_OK;
}

int wbx_ctx_synchronize(wbx_ctx* ctx) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_synchronize: ctx is NULL");
  WBX_CUDA(cudaSetDevice(ctx->device));
  WBX_CUDA(cudaStreamSynchronize(ctx->stream));
  return WBX_OK;
}

int wbx_ctx_info(wbx_ctx* ctx, int* sm_count, uint64_t* hbm_bytes,
                 uint64_t* kernel_launches) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_info: ctx is NULL");
  if (sm_count) *sm_count = ctx->sm_count;
  if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
  if (kernel_launches) *kernel_launches = ctx->launches;
  return WBX_OK;
}

int wbx_ctx_set_staging_bytes(wbx_ctx* ctx, uint64_t bytes) {
  WBX_REQUIRE(ctx != nullptr, "wbx_ctx_set_staging_bytes: ctx is NULL");
  WBX_REQUIRE(bytes >= (1ull << 20), "staging budget must be >= 1 MiB");
  ctx->staging_bytes = bytes;
  return WBX_OK;
}

int wbx_host_alloc(size_t bytes, void** out) {
  WBX_REQUIRE(out != nullptr, "wbx_host_alloc: out is NULL");
  *out = nullptr;
  WBX_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return WBX_OK;
}

int wbx_host_free(void* ptr) {
  if (!ptr) return WBX_OK;
  WBX_CUDA(cudaFreeHost(ptr));
  return WBX_OK;
}

int wbx_host_register(void* ptr, size_t bytes) {
  WBX_REQUIRE(ptr != nullptr, "wbx_host_register: ptr is NULL");
  WBX_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return WBX_OK;
}

int wbx_host_unregister(void* ptr) {
  WBX_REQUIRE(ptr != nullptr, "wbx_host_unregister: ptr is NULL");
  WBX_CUDA(cudaHostUnregister(ptr));
  return WBX_OK;
}

// sizeof / offsetof of every descriptor struct, in declaration order.
int wbx_struct_layout(int32_t which, uint64_t* out, int32_t cap) {
  if (!out || cap < 1) {
    wbx::set_error("wbx_struct_layout: NULL / empty output");
    return WBX_ERR_INVALID;
  }
  int n = 0;
#define WBX_PUT(v) do { if (n < cap) out[n] = (uint64_t)(v); ++n; } while (0)
  switch (which) {
    case 0:
      WBX_PUT(sizeof(wbx_det_desc));
      WBX_PUT(offsetof(wbx_det_desc, space));
      WBX_PUT(offsetof(wbx_det_desc, flags));
      WBX_PUT(offsetof(wbx_det_desc, n_jobs));
      WBX_PUT(offsetof(wbx_det_desc, ny));
      WBX_PUT(offsetof(wbx_det_desc, nx));
      WBX_PUT(offsetof(wbx_det_desc, n_cells));
      WBX_PUT(offsetof(wbx_det_desc, pred));
      WBX_PUT(offsetof(wbx_det_desc, target));
      WBX_PUT(offsetof(wbx_det_desc, clim));
      WBX_PUT(offsetof(wbx_det_desc, mask));
      WBX_PUT(offsetof(wbx_det_desc, cell));
      WBX_PUT(offsetof(wbx_det_desc, w_outer));
      WBX_PUT(offsetof(wbx_det_desc, w_y));
      WBX_PUT(offsetof(wbx_det_desc, w_x));
      WBX_PUT(offsetof(wbx_det_desc, stat_mask));
      WBX_PUT(offsetof(wbx_det_desc, n_classes));
      WBX_PUT(offsetof(wbx_det_desc, class_map));
      WBX_PUT(offsetof(wbx_det_desc, xform));
      WBX_PUT(offsetof(wbx_det_desc, reserved));
      WBX_PUT(offsetof(wbx_det_desc, thr_pred));
      WBX_PUT(offsetof(wbx_det_desc, thr_target));
      break;
    case 1:
      WBX_PUT(sizeof(wbx_crps_desc));
      WBX_PUT(offsetof(wbx_crps_desc, space));
      WBX_PUT(offsetof(wbx_crps_desc, flags));
      WBX_PUT(offsetof(wbx_crps_desc, n_jobs));
      WBX_PUT(offsetof(wbx_crps_desc, ny));
      WBX_PUT(offsetof(wbx_crps_desc, nx));
      WBX_PUT(offsetof(wbx_crps_desc, n_members));
      WBX_PUT(offsetof(wbx_crps_desc, member_stride));
      WBX_PUT(offsetof(wbx_crps_desc, point_stride));
      WBX_PUT(offsetof(wbx_crps_desc, n_cells));
      WBX_PUT(offsetof(wbx_crps_desc, ens));
      WBX_PUT(offsetof(wbx_crps_desc, target));
      WBX_PUT(offsetof(wbx_crps_desc, mask));
      WBX_PUT(offsetof(wbx_crps_desc, cell));
      WBX_PUT(offsetof(wbx_crps_desc, w_outer));
      WBX_PUT(offsetof(wbx_crps_desc, w_y));
      WBX_PUT(offsetof(wbx_crps_desc, w_x));
      WBX_PUT(offsetof(wbx_crps_desc, stat_mask));
      WBX_PUT(offsetof(wbx_crps_desc, reserved));
      break;
    case 2:
      WBX_PUT(sizeof(wbx_crps_point_desc));
      WBX_PUT(offsetof(wbx_crps_point_desc, ndim));
      WBX_PUT(offsetof(wbx_crps_point_desc, flags));
      WBX_PUT(offsetof(wbx_crps_point_desc, n_members));
      WBX_PUT(offsetof(wbx_crps_point_desc, member_stride));
      WBX_PUT(offsetof(wbx_crps_point_desc, size));
      WBX_PUT(offsetof(wbx_crps_point_desc, ens_stride));
      WBX_PUT(offsetof(wbx_crps_point_desc, target_stride));
      WBX_PUT(offsetof(wbx_crps_point_desc, ens));
      WBX_PUT(offsetof(wbx_crps_point_desc, target));
      WBX_PUT(offsetof(wbx_crps_point_desc, variance));
      WBX_PUT(offsetof(wbx_crps_point_desc, unbiased_mse));
      break;
    case 3:
      WBX_PUT(sizeof(wbx_spectrum_desc));
      WBX_PUT(offsetof(wbx_spectrum_desc, n_jobs));
      WBX_PUT(offsetof(wbx_spectrum_desc, ny));
      WBX_PUT(offsetof(wbx_spectrum_desc, nx));
      WBX_PUT(offsetof(wbx_spectrum_desc, field));
      WBX_PUT(offsetof(wbx_spectrum_desc, row_scale));
      WBX_PUT(offsetof(wbx_spectrum_desc, spectrum));
      break;
    case 4:
      WBX_PUT(sizeof(wbx_generic_desc));
      WBX_PUT(offsetof(wbx_generic_desc, ndim));
      WBX_PUT(offsetof(wbx_generic_desc, op));
      WBX_PUT(offsetof(wbx_generic_desc, flags));
      WBX_PUT(offsetof(wbx_generic_desc, n_factors));
      WBX_PUT(offsetof(wbx_generic_desc, size));
      WBX_PUT(offsetof(wbx_generic_desc, reduced));
      WBX_PUT(offsetof(wbx_generic_desc, a));
      WBX_PUT(offsetof(wbx_generic_desc, a_stride));
      WBX_PUT(offsetof(wbx_generic_desc, b));
      WBX_PUT(offsetof(wbx_generic_desc, b_stride));
      WBX_PUT(offsetof(wbx_generic_desc, c));
      WBX_PUT(offsetof(wbx_generic_desc, c_stride));
      WBX_PUT(offsetof(wbx_generic_desc, mask));
      WBX_PUT(offsetof(wbx_generic_desc, mask_stride));
      WBX_PUT(offsetof(wbx_generic_desc, factor));
      WBX_PUT(offsetof(wbx_generic_desc, factor_dtype));
      WBX_PUT(offsetof(wbx_generic_desc, factor_stride));
      break;
    default:
      wbx::set_error("wbx_struct_layout: unknown struct %d", which);
      return WBX_ERR_INVALID;
  }
#undef WBX_PUT
  return n;
}

}  // extern "C"
