// Micro-benchmark: issue rate of tcgen05.mma (M=128, K=16, bf16) from shared memory for the four
// K-major layout types and several N.  Operand contents are irrelevant (garbage smem); only the
// instruction timing is measured.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) rate_kernel(int N, int layout, int iters, int a_lbo, int a_sbo, long long* out, int nissue, int tap_pitch) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar4[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar4[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if ((threadIdx.x & 31) == 0 && warp < nissue) {
    uint64_t& bar = bar4[warp];
    const uint32_t tmem_w = tmem + (N > 128 ? (warp & 1) * 256 : warp * 128);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // walk K inside the operand like a real mainloop would (4 k-steps of 32 B for swizzled rows)
      // tap_pitch > 0: the A start address walks the 9 taps of a 3x3 neighbourhood (16-byte pixel steps)
      const int tp = i % 9;
      const uint32_t ko = (layout == 0) ? (tap_pitch > 0 ? (uint32_t)((tp / 3) * tap_pitch + tp % 3) * 16u : 0u) : (uint32_t)(i & 3) * 32u;
      uint64_t ad = make_desc(a0 + ko, a_lbo, a_sbo, layout);
      uint64_t bd = make_desc(b0 + ko, layout == 0 ? N * 16 : 0, layout == 0 ? 128 : (layout == 2 ? 1024 : (layout == 4 ? 512 : 256)), layout);
      uint32_t acc = i > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_w), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t1 = clock64();
    if (warp == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  struct Cfg { int layout; int lbo, sbo; const char* name; int tap_pitch; } cfgs[] = {
      {0, 2880, 160, "none  halo(LBO 2880,SBO 160)", 0}, {0, 2880, 160, "none  halo + 3x3 tap starts", 10},
      {0, 2048, 128, "none  dense(LBO 2048,SBO 128)", 0}, {0, 8064, 128, "none  flat(LBO 8064) + tap starts", 7},
      {0, 8192, 128, "none  flat(LBO 8192) + tap starts", 8},
      {2, 0, 1024, "sw128 dense SBO 1024", 0}};
  for (auto& c : cfgs)
    for (int N : {32, 64, 128, 256}) {
      for (int nissue : {1, 2, 4}) { const int grid = 148;
        rate_kernel<<<grid, 128, 64 * 1024>>>(N, c.layout, iters, c.lbo, c.sbo, d, nissue, c.tap_pitch);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-32s N=%3d issuers=%d  cycles per MMA (aggregate) = %7.1f  (err=%d)\n", c.name, N, nissue, (double)mx / iters / nissue, (int)e);
      }
    }
  return 0;
}
