// Global store / load issue cost per SM as a function of the per-instruction access pattern, for the epilogue of a
// 128-pixel x 256-channel bf16 NHWC tile (512 B per pixel; a warp owns 32 pixels x 64 channels = 32 lines of 128 B).
//   pattern 0: "row per lane"   lane l writes 32 B at pixel l, sector k (k = 0..3): 32 lines touched per instruction
//   pattern 1: "line per quad"  lanes 4r..4r+3 write the 4 sectors of pixel 8k + r: 8 full lines per instruction
//   pattern 2: contiguous       lane l writes 32 B at 32 l of a dense 1 KB block (what a plain copy does)
// Same byte count in every mode: 16 warps x 4 instructions x 1 KB = 64 KB per tile, tiles strided like the conv kernels
// (CTA b takes tiles b, b + grid, ...).  mode bit 8: loads instead of stores.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stg_rate stg_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(512) k(uint8_t* base, int ntiles, int pattern, int load, unsigned* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, w4 = warp >> 2;              // 32-pixel quarter, 64-channel group
  unsigned acc = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint8_t* tile = base + (size_t)t * 65536;          // 128 px x 512 B
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      size_t off;
      if (pattern == 0) off = (size_t)(q * 32 + lane) * 512 + w4 * 128 + kk * 32;
      else if (pattern == 1) off = (size_t)(q * 32 + kk * 8 + (lane >> 2)) * 512 + w4 * 128 + (lane & 3) * 32;
      else off = (size_t)(warp * 4 + kk) * 1024 + lane * 32;
      uint32_t* p = reinterpret_cast<uint32_t*>(tile + off);
      if (load) {
        uint32_t a, b, c, d, e, f, g, h;
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
        acc += a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
      } else {
        const uint32_t v = (uint32_t)t + lane;
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v), "r"(v), "r"(v), "r"(v), "r"(v),
                     "r"(v), "r"(v), "r"(v) : "memory");
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const int ntiles = 12288;                            // 805 MB = one 256-channel 64x48 map of 512 forwards
  uint8_t* d; cudaMalloc(&d, (size_t)ntiles * 65536); cudaMemset(d, 1, (size_t)ntiles * 65536);
  unsigned* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int load = 0; load < 2; ++load)
    for (int pattern = 0; pattern < 3; ++pattern) {
      for (int grid = 148; grid <= 296; grid += 148) {
        k<<<grid, 512>>>(d, ntiles, pattern, load, sink);
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) k<<<grid, 512>>>(d, ntiles, pattern, load, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%s pattern %d grid %d: %.1f us  %.0f GB/s  %.0f cycles/tile/SM (1.965 GHz)\n", load ? "load " : "store", pattern, grid,
               ms * 1e3, (double)ntiles * 65536 / ms / 1e6, ms * 1e-3 * 1.965e9 / (ntiles / 148.0));
      }
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
