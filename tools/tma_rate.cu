// TMA box-load throughput per SM as a function of the inner (contiguous) box extent.
// Every CTA streams boxes of the same 8 KB payload (128 rows x 32 bf16 channels of an [R, 32] matrix):
//   mode 0: planar (8 ch, 128 rows, 4 planes) -- 16-byte elements, SWIZZLE_NONE (the conv kernels' layout)
//   mode 1: row-major (32 ch, 128 rows)       -- 64-byte rows, SWIZZLE_64B
//   mode 2: row-major (32 ch, 128 rows)       -- 64-byte rows, SWIZZLE_NONE
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k(const __grid_constant__ CUtensorMap map, int mode, int iters, int nrows, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[8];
  const int S = 8;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int row = (blockIdx.x * 977) % (nrows / 128);
    for (int it = 0; it < iters + S; ++it) {
      const int s = it % S;
      if (it >= S) {                                   // wait for the load issued S iterations ago
        uint32_t done = 0, par = ((it / S) - 1) & 1;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(8192) : "memory");
        const uint32_t dst = smem_u32(smem + s * 8192);
        if (mode == 0)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(smem_u32(&bar[s])), "r"(0), "r"(row * 128), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(smem_u32(&bar[s])), "r"(0), "r"(row * 128) : "memory");
        row = (row + 1) % (nrows / 128);
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int nrows = 128 * 4096;                        // 32 MB matrix: L2 resident after the first pass
  void* d; cudaMalloc(&d, (size_t)nrows * 64); cudaMemset(d, 0, (size_t)nrows * 64);
  long long* o; cudaMalloc(&o, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192 + 1024);
  const int iters = 4000;
  for (int mode = 0; mode < 3; ++mode) {
    CUtensorMap m; memset(&m, 0, sizeof(m));
    CUresult r;
    if (mode == 0) {
      cuuint64_t dims[3] = {8, (cuuint64_t)nrows, 4}; cuuint64_t str[2] = {64, 16};
      cuuint32_t box[3] = {8, 128, 4}, es[3] = {1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {32, (cuuint64_t)nrows}; cuuint64_t str[1] = {64};
      cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 32, 8 * 8192 + 1024>>>(m, mode, iters, nrows, o);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      if (rep) printf("mode %d: %.1f cycles per 8 KB box  -> %.1f B/cycle/SM, %.2f smem rows (elements) per cycle (err %d)\n", mode,
                      (double)mx / iters, 8192.0 * iters / mx, (mode == 0 ? 512.0 : 128.0) * iters / mx, (int)e);
    }
  }
  return 0;
}
