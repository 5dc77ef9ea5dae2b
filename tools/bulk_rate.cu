// cp.async.bulk (linear, non-tensor) global -> shared throughput per SM: copy size, copies in flight, issuing lanes,
// number of CTAs, and whether all CTAs read the SAME source (a weight stream) or their own.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_rate bulk_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every stage = `per_stage` copies of `bytes` each, issued by `lanes` lanes (lane l issues copies l, l + lanes, ...)
__global__ void __launch_bounds__(32) k(const unsigned char* src, size_t span, int same, int bytes, int per_stage, int S, int lanes,
                                        int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[32];
  const int lane = threadIdx.x;
  if (lane == 0) {
    for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const size_t stage_bytes = (size_t)bytes * per_stage;
  const size_t nst = span / stage_bytes;
  size_t pos = same ? 0 : ((size_t)blockIdx.x * 977) % nst;
  long long t0 = clock64();
  for (int it = 0; it < iters + S; ++it) {
    const int s = it % S;
    if (it >= S) {
      uint32_t done = 0, par = ((it / S) - 1) & 1;
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
    }
    if (it < iters) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"((uint32_t)stage_bytes) : "memory");
      const unsigned char* g = src + pos * stage_bytes;
      if (lane < lanes)
        for (int c = lane; c < per_stage; c += lanes)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(smem + (size_t)s * stage_bytes + (size_t)c * bytes)), "l"(g + (size_t)c * bytes), "r"(bytes), "r"(smem_u32(&bar[s])) : "memory");
      pos = pos + 1 == nst ? 0 : pos + 1;
    }
    __syncwarp();
  }
  if (lane == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t span = 32u << 20;                       // 32 MB: L2 resident after the first pass
  unsigned char* d; cudaMalloc(&d, span); cudaMemset(d, 1, span);
  long long* o; cudaMalloc(&o, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 3000;
  struct Cfg { int bytes, per_stage, S, lanes, grid, same; size_t span; };
  const Cfg cfgs[] = {
      {2048, 9, 8, 1, 148, 1, 295 * 1024},  {2048, 9, 8, 9, 148, 1, 295 * 1024}, {2048, 9, 8, 9, 148, 0, span},
      {2048, 9, 8, 9, 16, 1, 295 * 1024},   {2048, 9, 8, 9, 1, 1, 295 * 1024},   {18432, 1, 8, 1, 148, 1, 295 * 1024},
      {18432, 1, 8, 1, 148, 0, span},       {18432, 1, 8, 1, 1, 0, span},        {4096, 9, 4, 9, 148, 1, 590 * 1024},
      {32768, 1, 5, 1, 148, 0, span},       {32768, 1, 5, 1, 16, 0, span},       {1024, 18, 8, 18, 148, 1, 295 * 1024},
  };
  for (const Cfg& c : cfgs) {
    const size_t smem = (size_t)c.bytes * c.per_stage * c.S + 1024;
    for (int rep = 0; rep < 2; ++rep) {
      k<<<c.grid, 32, smem>>>(d, c.span / ((size_t)c.bytes * c.per_stage) * ((size_t)c.bytes * c.per_stage), c.same, c.bytes, c.per_stage, c.S, c.lanes, iters, o);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, o, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < c.grid; ++i) mx = h[i] > mx ? h[i] : mx;
      if (rep) printf("copy %5d B x %2d per stage, %d stages, %2d lanes, %3d CTAs, %s source: %7.1f cycles/stage -> %5.1f B/cycle/SM (err %d)\n",
                      c.bytes, c.per_stage, c.S, c.lanes, c.grid, c.same ? "shared" : "own   ", (double)mx / iters,
                      (double)c.bytes * c.per_stage * iters / mx, (int)e);
    }
  }
  return 0;
}
