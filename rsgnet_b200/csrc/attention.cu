// TRP core (Target-aware Relation Parser, association.py:288-299):
//     y[n,i,:] = sum_j sigmoid(x[n,i,:] . x[n,j,:]) * g[n,j,:]
// as a blockwise fused kernel: the S x S affinity is never written to memory.  There is no softmax,
// so key blocks accumulate independently (no running max / rescale).
//
// One CTA = 64 query positions of one crop, 4 warps x 16 rows; key/value blocks of 64 positions
// stream through a double-buffered cp.async pipeline.  Both contractions run on the warp-level
// tensor path (mma.sync m16n8k16 bf16 -> fp32); P = sigmoid(QK^T) is re-used straight from the
// accumulator registers as the A operand of P.V.  sigmoid(s) = 0.5*tanh(0.5 s)+0.5 costs a single
// MUFU op (tanh.approx), which is what bounds this kernel: S*S sigmoids per crop.
#include "ops.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_sigmoid(float s) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * s));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t;
  t.x = __float2bfloat16_rn(a);
  t.y = __float2bfloat16_rn(b);
  return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int BQ = 64, BKV = 64, ATT_THREADS = 128;

template <int C>
__global__ void __launch_bounds__(ATT_THREADS)
trp_attention_kernel(const bf16* __restrict__ x, int x_cs, int x_co, const bf16* __restrict__ g,
                     int g_cs, int g_co, bf16* __restrict__ y, int y_cs, int y_co, int S) {
  constexpr int PITCH = C * 2 + 16;          // bytes per smem row; the +16 keeps ldmatrix conflict-free
  constexpr int CH = C / 8;                  // 16-byte chunks per row
  constexpr int KS = C / 16;                 // k16 steps of Q.K^T
  constexpr int NTV = C / 8;                 // n8 tiles of P.V
  __shared__ __align__(16) unsigned char sQ[BQ * PITCH];
  __shared__ __align__(16) unsigned char sK[2][BKV * PITCH];
  __shared__ __align__(16) unsigned char sV[2][BKV * PITCH];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y, q0 = blockIdx.x * BQ;
  const bf16* xb = x + (size_t)n * S * x_cs + x_co;
  const bf16* gb = g + (size_t)n * S * g_cs + g_co;

  for (int i = tid; i < BQ * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = q0 + r < S;
    cp_async16(smem_u32(sQ + r * PITCH + c * 16), ok ? xb + (size_t)(q0 + r) * x_cs + c * 8 : xb, ok);
  }
  auto load_kv = [&](int blk, int st) {
    const int k0 = blk * BKV;
    for (int i = tid; i < BKV * CH; i += ATT_THREADS) {
      const int r = i / CH, c = i - r * CH;
      const bool ok = k0 + r < S;
      cp_async16(smem_u32(sK[st] + r * PITCH + c * 16), ok ? xb + (size_t)(k0 + r) * x_cs + c * 8 : xb, ok);
      cp_async16(smem_u32(sV[st] + r * PITCH + c * 16), ok ? gb + (size_t)(k0 + r) * g_cs + c * 8 : gb, ok);
    }
  };
  const int nblk = (S + BKV - 1) / BKV;
  load_kv(0, 0);
  cp_commit();

  uint32_t qf[KS][4];
  float o[NTV][4];
#pragma unroll
  for (int i = 0; i < NTV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

  for (int blk = 0; blk < nblk; ++blk) {
    const int st = blk & 1;
    if (blk + 1 < nblk) load_kv(blk + 1, st ^ 1);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm4(smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + (ks * 2 + (lane >> 4)) * 16),
              qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    // S = Q K^T : 16 x 64 per warp
    float sacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
    const uint32_t kb = smem_u32(sK[st]);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        ldsm4(kb + row * PITCH + (ks * 2 + ((lane >> 3) & 1)) * 16, b0, b1, b2, b3);
        mma16816(sacc[np * 2], qf[ks], b0, b1);
        mma16816(sacc[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // P = sigmoid(S) -> bf16 A fragments; O += P V
    const uint32_t vb = smem_u32(sV[st]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(fast_sigmoid(sacc[2 * kk][0]), fast_sigmoid(sacc[2 * kk][1]));
      pf[1] = pack_bf16(fast_sigmoid(sacc[2 * kk][2]), fast_sigmoid(sacc[2 * kk][3]));
      pf[2] = pack_bf16(fast_sigmoid(sacc[2 * kk + 1][0]), fast_sigmoid(sacc[2 * kk + 1][1]));
      pf[3] = pack_bf16(fast_sigmoid(sacc[2 * kk + 1][2]), fast_sigmoid(sacc[2 * kk + 1][3]));
#pragma unroll
      for (int np = 0; np < NTV / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldsm4t(vb + row * PITCH + (np * 2 + (lane >> 4)) * 16, b0, b1, b2, b3);
        mma16816(o[np * 2], pf, b0, b1);
        mma16816(o[np * 2 + 1], pf, b2, b3);
      }
    }
    __syncthreads();
  }
  cp_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int row = q0 + warp * 16 + gq + hf * 8;
    if (row >= S) continue;
    bf16* dst = y + ((size_t)n * S + row) * y_cs + y_co;
#pragma unroll
    for (int nt = 0; nt < NTV; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8 + tq * 2) = pack_bf16(o[nt][hf * 2], o[nt][hf * 2 + 1]);
  }
}

template <int C>
int launch(cudaStream_t s, const bf16* x, int x_cs, int x_co, const bf16* g, int g_cs, int g_co,
           bf16* y, int y_cs, int y_co, int N, int S) {
  dim3 grid((S + BQ - 1) / BQ, N);
  trp_attention_kernel<C><<<grid, ATT_THREADS, 0, s>>>(x, x_cs, x_co, g, g_cs, g_co, y, y_cs, y_co, S);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

}  // namespace

int attention_launch(cudaStream_t s, const bf16* x, int x_cs, int x_co, const bf16* g, int g_cs,
                     int g_co, bf16* y, int y_cs, int y_co, int N, int S, int C) {
  RSG_REQUIRE(N <= 65535, "attention: at most 65535 crops per launch");
  RSG_REQUIRE(x_cs % 8 == 0 && x_co % 8 == 0 && g_cs % 8 == 0 && g_co % 8 == 0 && y_cs % 2 == 0 && y_co % 2 == 0,
              "attention: channel strides/offsets must be 16-byte aligned");
  if (N == 0 || S == 0) return RSG_OK;
  switch (C) {
    case 16: return launch<16>(s, x, x_cs, x_co, g, g_cs, g_co, y, y_cs, y_co, N, S);
    case 32: return launch<32>(s, x, x_cs, x_co, g, g_cs, g_co, y, y_cs, y_co, N, S);
    case 48: return launch<48>(s, x, x_cs, x_co, g, g_cs, g_co, y, y_cs, y_co, N, S);
    case 64: return launch<64>(s, x, x_cs, x_co, g, g_cs, g_co, y, y_cs, y_co, N, S);
  }
  rsg_set_error("attention: unsupported channel count %d (16/32/48/64)", C);
  return RSG_ERR_ARG;
}
