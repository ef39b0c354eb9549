// Throughput of the FP32 instruction forms the kernels of this repo are bound
// by, in warp-instructions per cycle per SM sub-partition (SMSP):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb \
//        profiles/microbench_fp32_pipes.cu && /tmp/mb
// Each kernel runs 8 independent dependency chains per thread (latency 4 is
// hidden), 1024 threads per SM-resident CTA, one CTA per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

__device__ __forceinline__ uint64_t pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(uint64_t v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, float a, float b,
                                                 long long* cycles) {
  float x[CHAINS];
  uint64_t X[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) {
    x[c] = a + c + threadIdx.x;
    X[c] = pack(x[c], x[c] + 1.f);
  }
  const uint64_t A = pack(a, a), B = pack(b, b);
  const long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == 0) x[c] = fmaf(x[c], a, b);                       // FFMA 3-reg
      if (OP == 1) x[c] = x[c] + a;                               // FADD
      if (OP == 2) x[c] = fmaxf(x[c], x[(c + 1) % CHAINS]);       // FMNMX
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;"    // FFMA2
                                : "+l"(X[c]) : "l"(A), "l"(B));
      if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;"        // FADD2
                                : "+l"(X[c]) : "l"(A));
      if (OP == 5) x[c] = fmaf(x[c], 1.0009765625f, b);           // FFMA imm
      if (OP == 6) x[c] = fabsf(x[c] - a) + b;                    // FADD + FADD|.|
      if (OP == 7) {                                              // FADD + FMNMX mix
        x[c] = x[c] + a;
        x[c] = fmaxf(x[c], b);
      }
      if (OP == 9) asm volatile("max.f32 %0, %0, %1, %2;"        // FMNMX3
                                : "+f"(x[c]) : "f"(x[(c + 1) % CHAINS]), "f"(b));
      if (OP == 10 && (c & 1) == 0) {   // compare-exchange: 2 FMNMX
        const float lo = fminf(x[c], x[c + 1]), hi = fmaxf(x[c], x[c + 1]);
        x[c] = lo + a;   // (+ a keeps the chain from folding: 2 FADD extra)
        x[c + 1] = hi + a;
      }
      if (OP == 11 && (c & 1) == 0) {   // mixed CE: 1 FMNMX + 2 FADD
        const float lo = fminf(x[c], x[c + 1]);
        const float hi = __fsub_rn(__fadd_rn(x[c], x[c + 1]), lo);
        x[c] = lo + a;
        x[c + 1] = hi + a;
      }
      if (OP == 12) {                                             // VIMNMX (s32)
        int v = max(__float_as_int(x[c]), __float_as_int(x[(c + 1) % CHAINS]));
        x[c] = __int_as_float(v);
      }
      if (OP == 8) {                                              // FADD2 + 2 LOP3
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[c]) : "l"(A));
        X[c] &= 0x7fffffff7fffffffull;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c] + lo_of(X[c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double instr_per_step) {
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
  // warps per SMSP = 1024 / 32 / 4 = 8
  const double warp_instr = 8.0 * ITERS * CHAINS * instr_per_step;
  printf("{\"op\": \"%s\", \"cycles\": %.0f, \"warp_instr_per_clk_per_smsp\": "
         "%.3f, \"err\": \"%s\"}\n",
         name, mean, warp_instr / mean, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA 3-reg", 1);
  run<5>("FFMA imm", 1);
  run<1>("FADD", 1);
  run<2>("FMNMX", 1);
  run<3>("FFMA2 (f32x2)", 1);
  run<4>("FADD2 (f32x2)", 1);
  run<6>("FADD + FADD|x| (pair-sum step)", 2);
  run<7>("FADD + FMNMX (two pipes)", 2);
  run<8>("FADD2 + LOP3.64 (packed abs)", 3);
  run<9>("FMNMX3 (3-input max)", 1);
  // per PAIR of chains: (2 FMNMX + 2 FADD) / 2 and (1 FMNMX + 4 FADD) / 2
  run<10>("compare-exchange 2 FMNMX (+2 FADD)", 2);
  run<11>("compare-exchange 1 FMNMX + 2 FADD (+2 FADD)", 2.5);
  run<12>("VIMNMX s32", 1);
  return 0;
}
