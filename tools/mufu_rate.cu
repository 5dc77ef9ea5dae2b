// MUFU throughput probe: tanh.approx.f32 vs tanh.approx.bf16x2 (results per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 977u + i * 131u + 0x3c003c00u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i]));
      else if (MODE == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a[i]));
      else if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      else if (MODE == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
      else if (MODE == 5) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
      else if (MODE == 6) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      else if (MODE == 7) {      // the f16x2 sigmoid pipeline: pack two f32 -> f16x2, tanh, 0.5 t + 0.5 (HFMA2)
        asm volatile("cvt.rn.f16x2.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(0x38003800u));
      }
      else { asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[i])); asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(a[(i + 3) & 7]) : "r"(a[(i + 5) & 7])); }
    }
  }
  unsigned s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  unsigned* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 1) k<1><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 2) k<2><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 3) k<3><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 4) k<4><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 5) k<5><<<148 * 2, 1024>>>(d, iters);
      else if (mode == 6) k<6><<<148 * 2, 1024>>>(d, iters);
      else k<7><<<148 * 2, 1024>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = 148.0 * 2 * 1024 * 8.0 * iters;      // thread-level MUFU instructions
      if (rep) printf("mode %d (%s): %.3f ms, %.2f thread-instr/clk/SM at 1.965 GHz (x2 results for packed)\n", mode,
                      mode == 0 ? "tanh.f32" : mode == 1 ? "tanh.bf16x2" : mode == 2 ? "ex2.bf16x2" : mode == 3 ? "cvt.rn.bf16x2.f32" : mode == 4 ? "tanh.f32 + cvt.bf16x2 (pairs counted once)" : mode == 5 ? "tanh.f16x2" : mode == 6 ? "ex2.f16x2" : "cvt.f16x2 + tanh.f16x2 + fma.f16x2", ms, instr / 148 / (ms * 1e-3 * 1.965e9));
    }
  }
  return 0;
}
