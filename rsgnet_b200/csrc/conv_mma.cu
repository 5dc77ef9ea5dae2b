// Generic implicit-GEMM convolution on the legacy warp-level tensor path (mma.sync m16n8k16,
// bf16 x bf16 -> fp32).  This is the "any shape" kernel: arbitrary tap lists (3x3, 1x1, the four
// 2x2 phases of ConvTranspose 4x4 s2), input stride 1/2, channel-sliced NHWC inputs/outputs,
// fp32 NCHW heat-map outputs, and the fused epilogue  out = act(conv + bias + sum_r up_r(res_r)).
// The tcgen05 kernel (conv_tc5.cu) takes over the shapes that dominate the FLOPs.
//
// Reference ops subsumed: nn.Conv2d + BatchNorm2d(eval) + ReLU + residual add + nn.Upsample
// (pose_rsgnet.py:38-54, 75-95, 194-249, 261-270), ConvTranspose2d (pose_rsgnet.py:726-737).
//
// GEMM view: D[M = N*Hout*Wout pixels, Cout] = A[M, ntaps*Cin] * W^T, tile 128 x BN x 32,
// 8 warps (4 along M x 2 along N), 3-stage cp.async pipeline with zero-fill for padding.
#include "conv_params.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int THREADS = 256;

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
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                      uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 64-byte rows, 16-byte chunks XOR-swizzled so that ldmatrix (8 rows x 16 B) is conflict-free
__device__ __forceinline__ int swz(int row, int chunk) {
  return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

template <int BN>
__global__ void __launch_bounds__(THREADS)
conv_mma_kernel(const ConvP p) {
  constexpr int WN = BN / 2;          // warp tile along N
  constexpr int NT = WN / 8;          // n8 tiles per warp
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                              // STAGES x 128 x 64 B
  unsigned char* sB = smem + STAGES * BM * 64;           // STAGES x BN x 64 B

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const long long m_blk = (long long)blockIdx.x * BM;
  const int n_blk = blockIdx.y * BN;

  // ---- A-gather bookkeeping: this thread copies chunk `ach` of rows arow and arow+64
  const int ach = tid & 3;
  int a_n[2], a_y[2], a_x[2];
  bool a_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    long long m = m_blk + (tid >> 2) + h * 64;
    a_ok[h] = m < p.M;
    long long mm = a_ok[h] ? m : 0;
    int x = (int)(mm % p.Wout);
    long long r = mm / p.Wout;
    int y = (int)(r % p.Hout);
    a_n[h] = (int)(r / p.Hout);
    a_y[h] = y * p.stride;
    a_x[h] = x * p.stride;
  }
  const int kchunks = p.CinPad / BK;
  const int ksteps = p.ntaps * kchunks;

  auto load_stage = [&](int ks, int st) {
    const int tap = ks / kchunks;
    const int c0 = (ks - tap * kchunks) * BK;
    const int dy = p.dy[tap], dx = p.dx[tap];
    unsigned char* a_dst = sA + st * (BM * 64);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = (tid >> 2) + h * 64;
      const int iy = a_y[h] + dy, ix = a_x[h] + dx;
      const int c = c0 + ach * 8;
      bool ok = a_ok[h] && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win && c < p.Cin;
      const bf16* src = p.in;
      if (ok) src = p.in + ((size_t)((size_t)a_n[h] * p.Hin + iy) * p.Win + ix) * p.in_cs + p.in_co + c;
      cp_async16(smem_u32(a_dst + swz(row, ach)), src, ok);
    }
    unsigned char* b_dst = sB + st * (BN * 64);
    const bf16* wbase = p.w + ((size_t)tap * p.CoutPad + n_blk) * p.CinPad + c0;
#pragma unroll
    for (int i = tid; i < BN * 4; i += THREADS) {
      const int row = i >> 2, ch = i & 3;
      cp_async16(smem_u32(b_dst + swz(row, ch)), wbase + (size_t)row * p.CinPad + ch * 8, true);
    }
  };

  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ksteps) load_stage(s, s);
    cp_commit();
  }

  for (int ks = 0; ks < ksteps; ++ks) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = ks + STAGES - 1;
      if (nxt < ksteps) load_stage(nxt, nxt % STAGES);
      cp_commit();
    }
    const int st = ks % STAGES;
    const uint32_t a_base = smem_u32(sA + st * (BM * 64));
    const uint32_t b_base = smem_u32(sB + st * (BN * 64));
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = wm * 32 + mt * 16 + (lane & 15);
        ldsm4(a_base + swz(row, kk * 2 + (lane >> 4)), af[mt][0], af[mt][1], af[mt][2], af[mt][3]);
      }
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = wn * WN + np * 16 + (lane & 7) + ((lane >> 4) << 3);
        ldsm4(b_base + swz(row, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][np * 2], af[mt], b0, b1);
          mma16816(acc[mt][np * 2 + 1], af[mt], b2, b3);
        }
      }
    }
  }
  cp_wait<0>();

  // ---- epilogue: bias + residual terms + ReLU, bf16 NHWC and/or fp32 NCHW stores
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const long long m = m_blk + wm * 32 + mt * 16 + g + hf * 8;
      if (m >= p.M) continue;
      const int x = (int)(m % p.Wout);
      const long long r = m / p.Wout;
      const int y = (int)(r % p.Hout);
      const int n = (int)(r / p.Hout);
      const int Y = y * p.omul + p.ooy, X = x * p.omul + p.oox;
      const bf16* rp[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < p.nres) {
          const ResP& rr = p.res[q];
          rp[q] = rr.p + ((size_t)((size_t)(rr.bs0 ? 0 : n) * rr.H + (Y >> rr.shift)) * rr.W +
                          (X >> rr.shift)) * rr.cs + rr.co;
        }
      }
      bf16* op = p.out ? p.out + ((size_t)((size_t)n * p.oH + Y) * p.oW + X) * p.out_cs + p.out_co
                       : nullptr;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = n_blk + wn * WN + nt * 8 + tq * 2;
        float v0 = acc[mt][nt][hf * 2] + __ldg(p.bias + col);
        float v1 = acc[mt][nt][hf * 2 + 1] + __ldg(p.bias + col + 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < p.nres && col < p.Cout) {
            __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(rp[q] + col);
            v0 += __bfloat162float(rv.x);
            v1 += __bfloat162float(rv.y);
          }
        }
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (op && col < p.Cout) {
          __nv_bfloat162 o;
          o.x = __float2bfloat16_rn(v0);
          o.y = __float2bfloat16_rn(v1);
          *reinterpret_cast<__nv_bfloat162*>(op + col) = o;
        }
        if (p.out_f32) {
          const size_t plane = (size_t)p.oH * p.oW;
          float* fp = p.out_f32 + ((size_t)n * p.Cout + col) * plane + (size_t)Y * p.oW + X;
          if (col < p.Cout) fp[0] = v0;
          if (col + 1 < p.Cout) fp[plane] = v1;
        }
      }
    }
  }
}

template <int BN>
int launch(const ConvP& p, cudaStream_t s) {
  const size_t smem = (size_t)STAGES * (BM + BN) * 64;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_mma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr_once.done();
  }
  dim3 grid(ceil_div(p.M, BM), p.CoutPad / BN);
  conv_mma_kernel<BN><<<grid, THREADS, smem, s>>>(p);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

}  // namespace

int conv_mma_launch(const ConvP& p, cudaStream_t s) {
  RSG_REQUIRE(p.CinPad % BK == 0 && p.CoutPad % 32 == 0, "conv: CinPad %d / CoutPad %d must be multiples of 32", p.CinPad, p.CoutPad);
  RSG_REQUIRE(p.Cin % 8 == 0 && p.in_cs % 8 == 0 && p.in_co % 8 == 0, "conv: input channels/stride/offset must be multiples of 8");
  if (p.M == 0) return RSG_OK;
  if (p.CoutPad % 128 == 0) return launch<128>(p, s);
  if (p.CoutPad % 96 == 0) return launch<96>(p, s);
  if (p.CoutPad % 64 == 0) return launch<64>(p, s);
  return launch<32>(p, s);
}
