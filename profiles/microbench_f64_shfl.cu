// Throughput of the instruction forms the binned reduction (csrc/det_bins3.cuh)
// and the CRPS kernels lean on besides plain FP32, in warp-instructions per
// cycle per SM sub-partition (SMSP):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb2 \
//        profiles/microbench_f64_shfl.cu && /tmp/mb2
// Eight independent chains per thread, 1024 threads per CTA, one CTA per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, float a, float b,
                                                 long long* cycles) {
  __shared__ float4 sm[1024];
  float x[CHAINS];
  double d[CHAINS];
  float4 q[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    x[c] = a + c + threadIdx.x;
    d[c] = x[c];
    q[c] = make_float4(x[c], x[c], x[c], x[c]);
  }
  sm[threadIdx.x] = q[0];
  __syncthreads();
  const double A = a, B = b;
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == 0) d[c] = fma(d[c], A, B);                         // DFMA
      if (OP == 1) d[c] = d[c] + A;                                // DADD
      if (OP == 2) {                                               // F2F.F64.F32 + DADD
        d[c] += static_cast<double>(x[c]);
      }
      if (OP == 3) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1);   // SHFL.BFLY
      if (OP == 4) x[c] = __shfl_up_sync(0xffffffffu, x[c], 8);    // SHFL.UP
      if (OP == 5) {                                               // LDS.128
        q[c] = sm[(threadIdx.x + (int)q[c].x) & 1023];
      }
      if (OP == 6) x[c] = (lane & 1) ? x[c] : x[(c + 1) % CHAINS]; // FSEL/SEL
      if (OP == 7) {                                               // F2F.F32.F64
        x[c] = static_cast<float>(d[c]) + x[c];
      }
      if (OP == 8) {                                               // I2F-free: FMUL+F2F+DFMA
        d[c] = fma(static_cast<double>(x[c] * a), A, d[c]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    s += x[c] + static_cast<float>(d[c]) + q[c].x + q[c].y + q[c].z + q[c].w;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_step) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  bench<OP><<<sms, 1024>>>(out, 1.0001f, 0.5f, cyc);
  bench<OP><<<sms, 1024>>>(out, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += h[i];
  mean /= sms;
  const double warp_instr = 8.0 * ITERS * CHAINS * instr_per_step;
  printf("{\"op\": \"%s\", \"cycles\": %.0f, \"warp_instr_per_clk_per_smsp\": "
         "%.3f, \"err\": \"%s\"}\n",
         name, mean, warp_instr / mean, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("DFMA", 1);
  run<1>("DADD", 1);
  run<2>("F2F.F64.F32 + DADD", 2);
  run<3>("SHFL.BFLY", 1);
  run<4>("SHFL.UP", 1);
  run<5>("LDS.128 (dependent index)", 1);
  run<6>("SEL", 1);
  run<7>("F2F.F32.F64 + FADD", 2);
  run<8>("FMUL + F2F.F64.F32 + DFMA", 3);
  return 0;
}
