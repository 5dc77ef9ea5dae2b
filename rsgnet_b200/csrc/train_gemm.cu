// Training-step matrix kernels (reference: lib/core/function.py:240-363 drives torch autograd over
// lib/models/pose_rsgnet.py -- cuDNN fprop / dgrad / wgrad and cuBLAS bmm in fp32, TF32 on the tensor cores by default).
//
// Activations are fp32 NHWC, i.e. row-major [pixels, channels] matrices, so every convolution is a GEMM whose A rows are
// GATHERED pixels:
//   forward    Y[m, co]  = sum_tap sum_ci X[src_f(m, tap), ci] * W[tap][ci][co]          (+ bias)
//   dgrad      dX[m, ci] = sum_tap sum_co dY[src_d(m, tap), co] * W[tap][ci][co]         (B read transposed per tap)
//   wgrad      dW[tap][ci][co] = sum_m X[src(m, tap), ci] * dY[m, co]                    (reduction over pixels, split-K)
// ConvTranspose2d is the same three kernels with the gather roles exchanged.  The TRP's S x S products (association.py:
// 288-299) are the batched plain form (no gather) with optional transposes.
//
// Math: mma.sync m16n8k8 TF32 with fp32 accumulation (what the reference's cuDNN path does on an Ampere-or-later GPU:
// torch.backends.cudnn.allow_tf32 defaults to True).  `precise` = 3xTF32 (hi/lo split of both operands: fp32-class
// products), used by the parity tests to separate logic errors from rounding, and always for the TRP's forward products.
// In the production TF32 mode rsg_train_gemm routes every call with a K-major B operand to the tcgen05 kind::tf32 kernels of
// train_tc5.cu; this file keeps the mma.sync implicit GEMM (3xTF32, transposed / odd-shaped products) and the weight-gradient
// kernels (which cannot use tcgen05 for TF32: both operands are MN-major, see wgrad below).
#include "common.cuh"
#include "../../include/rsg_b200.h"

namespace {

constexpr int GM_THREADS = 256;
constexpr int GM_BM = 128, GM_BK = 16;
constexpr int GM_APITCH = GM_BK + 4;          // As[m][k]: bank = (20 g + t) mod 32, distinct for g < 8, t < 4

struct GemmP {
  const float* A; const float* B; float* C; const float* bias;
  int M, Nc, Ca;
  int lda, ldb, ldc;
  long long sA, sB, sC;                       // batch strides (elements)
  long long tapB;                             // tap stride of B (elements)
  int mode;                                   // 0 plain, 1 forward gather, 2 dgrad gather
  int transA, transB, beta;
  int taps, kw;
  int Ha, Wa, Hc, Wc, stride, pad;            // A's pixel grid, C's pixel grid
  int vecA, vecB, precise;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// four consecutive floats starting at p, of which the first `nvalid` exist (others read as 0)
__device__ __forceinline__ float4 load4(const float* p, int nvalid, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid >= 4 && vec) return *reinterpret_cast<const float4*>(p);
  if (nvalid > 0) v.x = p[0];
  if (nvalid > 1) v.y = p[1];
  if (nvalid > 2) v.z = p[2];
  if (nvalid > 3) v.w = p[3];
  return v;
}

// source row of C-grid pixel (n, y, x) for tap (dy, dx); -1 = outside (contributes zero)
__device__ __forceinline__ long long gather_row(int mode, int n, int y, int x, int dy, int dx, int Ha, int Wa, int stride, int pad) {
  int ya, xa;
  if (mode == 1) {
    ya = y * stride - pad + dy;
    xa = x * stride - pad + dx;
    if (ya < 0 || ya >= Ha || xa < 0 || xa >= Wa) return -1;
  } else {
    const int ty = y + pad - dy, tx = x + pad - dx;
    if (ty < 0 || tx < 0) return -1;
    ya = ty / stride; xa = tx / stride;
    if (ya * stride != ty || xa * stride != tx || ya >= Ha || xa >= Wa) return -1;
  }
  return ((long long)n * Ha + ya) * Wa + xa;
}

template <int BN>
__global__ void __launch_bounds__(GM_THREADS) gather_gemm_kernel(const GemmP p) {
  constexpr int WARPS_N = BN / 32, WARPS_M = 8 / WARPS_N, WTM = GM_BM / WARPS_M, MT = WTM / 16;
  constexpr int BPITCH = BN + 8;              // Bs[k][n]: bank = (8 t + g) mod 32
  __shared__ __align__(16) float As[GM_BM * GM_APITCH];
  __shared__ __align__(16) float Bs[GM_BK * BPITCH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * BN;
  const float* Ap = p.A + (long long)blockIdx.z * p.sA;
  const float* Bp = p.B + (long long)blockIdx.z * p.sB;
  float* Cp = p.C + (long long)blockIdx.z * p.sC;

  // rows this thread stages (non-transposed A): r = (tid >> 2) + 64 j, k quad = tid & 3
  int rn[2], ry[2], rx[2];
  bool rok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int m = m0 + (tid >> 2) + 64 * j;
    rok[j] = m < p.M;
    rn[j] = ry[j] = rx[j] = 0;
    if (p.mode != 0 && rok[j]) {
      const int hw = p.Hc * p.Wc;
      rn[j] = m / hw;
      const int rem = m - rn[j] * hw;
      ry[j] = rem / p.Wc;
      rx[j] = rem - ry[j] * p.Wc;
    }
  }
  const int kchunks = (p.Ca + GM_BK - 1) / GM_BK;
  const int niter = p.taps * kchunks;
  float4 ra[2], rb;

  auto fetch = [&](int it) {
    const int tap = it / kchunks, c0 = (it - tap * kchunks) * GM_BK;
    if (!p.transA) {
      const int k = c0 + (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        long long src = -1;
        if (rok[j]) {
          if (p.mode == 0) src = m0 + (tid >> 2) + 64 * j;
          else src = gather_row(p.mode, rn[j], ry[j], rx[j], tap / p.kw, tap % p.kw, p.Ha, p.Wa, p.stride, p.pad);
        }
        const int nv = src < 0 ? 0 : p.Ca - k;
        ra[j] = load4(Ap + (src < 0 ? 0 : src) * p.lda + k, nv, p.vecA != 0);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {       // element (m, k) at A[k * lda + m]: 16 k rows x 32 m quads
        const int k = c0 + (tid >> 5) + 8 * j, m = m0 + (tid & 31) * 4;
        const int nv = k < p.Ca ? p.M - m : 0;
        ra[j] = load4(Ap + (long long)k * p.lda + m, nv, p.vecA != 0);
      }
    }
    const float* Bt = Bp + (long long)tap * p.tapB;
    if (!p.transB) {                      // element (k, n) at B[k * ldb + n]: 16 k rows x BN/4 quads
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < GM_BK * (BN / 4)) {
        const int k = c0 + tid / (BN / 4), n = n0 + (tid % (BN / 4)) * 4;
        const int nv = k < p.Ca ? p.Nc - n : 0;
        rb = load4(Bt + (long long)k * p.ldb + n, nv, p.vecB != 0);
      }
    } else {                              // element (k, n) at B[n * ldb + k]: BN n rows x 4 k quads
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < BN * 4) {
        const int n = n0 + (tid >> 2), k = c0 + (tid & 3) * 4;
        const int nv = n < p.Nc ? p.Ca - k : 0;
        rb = load4(Bt + (long long)n * p.ldb + k, nv, p.vecB != 0);
      }
    }
  };
  auto stage = [&]() {
    if (!p.transA) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        *reinterpret_cast<float4*>(&As[((tid >> 2) + 64 * j) * GM_APITCH + (tid & 3) * 4]) = ra[j];
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = (tid >> 5) + 8 * j, m = (tid & 31) * 4;
        As[(m + 0) * GM_APITCH + k] = ra[j].x;
        As[(m + 1) * GM_APITCH + k] = ra[j].y;
        As[(m + 2) * GM_APITCH + k] = ra[j].z;
        As[(m + 3) * GM_APITCH + k] = ra[j].w;
      }
    }
    if (!p.transB) {
      if (tid < GM_BK * (BN / 4))
        *reinterpret_cast<float4*>(&Bs[(tid / (BN / 4)) * BPITCH + (tid % (BN / 4)) * 4]) = rb;
    } else if (tid < BN * 4) {
      const int n = tid >> 2, k = (tid & 3) * 4;
      Bs[(k + 0) * BPITCH + n] = rb.x;
      Bs[(k + 1) * BPITCH + n] = rb.y;
      Bs[(k + 2) * BPITCH + n] = rb.z;
      Bs[(k + 3) * BPITCH + n] = rb.w;
    }
  };

  float acc[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  fetch(0);
  for (int it = 0; it < niter; ++it) {
    stage();
    __syncthreads();
    if (it + 1 < niter) fetch(it + 1);
#pragma unroll
    for (int k8 = 0; k8 < GM_BK / 8; ++k8) {
      float af[MT][4], bfr[4][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* a = &As[(wm * WTM + mt * 16 + g) * GM_APITCH + k8 * 8 + t];
        af[mt][0] = a[0];
        af[mt][1] = a[8 * GM_APITCH];
        af[mt][2] = a[4];
        af[mt][3] = a[8 * GM_APITCH + 4];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* b = &Bs[(k8 * 8 + t) * BPITCH + wn * 32 + nt * 8 + g];
        bfr[nt][0] = b[0];
        bfr[nt][1] = b[4 * BPITCH];
      }
      if (!p.precise) {
        uint32_t ah[MT][4], bh[4][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) ah[mt][i] = to_tf32(af[mt][i]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { bh[nt][0] = to_tf32(bfr[nt][0]); bh[nt][1] = to_tf32(bfr[nt][1]); }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
      } else {
        uint32_t ah[MT][4], al[MT][4], bh[4][2], bl[4][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ah[mt][i] = to_tf32(af[mt][i]);
            al[mt][i] = to_tf32(af[mt][i] - __uint_as_float(ah[mt][i]));
          }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            bh[nt][i] = to_tf32(bfr[nt][i]);
            bl[nt][i] = to_tf32(bfr[nt][i] - __uint_as_float(bh[nt][i]));
          }
        // The tensor core's accumulator truncates (round-toward-zero) at every accumulation: ~2000 chained accumulations
        // leave a BIAS of ~4e-5 (measured at a reduction length of 5400).  Each k8 step therefore starts from zero and is
        // added to the running sum by the CUDA cores (round-to-nearest): three truncating accumulations per partial sum.
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            mma_tf32(part, al[mt], bh[nt]);
            mma_tf32(part, ah[mt], bl[nt]);
            mma_tf32(part, ah[mt], bh[nt]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] += part[i];
          }
      }
    }
    __syncthreads();
  }

  // epilogue: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wm * WTM + mt * 16 + g + 8 * h;
        if (m >= p.M) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + wn * 32 + nt * 8 + 2 * t + e;
          if (n >= p.Nc) continue;
          float v = acc[mt][nt][2 * h + e];
          if (p.bias) v += p.bias[n];
          float* dst = Cp + (long long)m * p.ldc + n;
          if (p.beta) v += *dst;
          *dst = v;
        }
      }
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad: dW[tap][ci][co] += sum_m X[src(m, tap), ci] * dY[m, co]; 64 x 64 output tile, 64-pixel reduction chunks, split
// over the pixels with fp32 atomics into a zeroed (or accumulating) dW.  The reduction runs over pixels, so in NHWC memory
// both operands are MN-major (channels contiguous), and tcgen05.mma kind::tf32 returns ZEROS as soon as either transpose bit of
// the instruction descriptor is set (measured on B200, profiles/r2_notes.md §9).  kind::f16 accepts MN-major operands:
// train_wgrad5.cu runs the wide 3x3 stride-1 layers on it with bf16 hi / lo splits; the kernels below keep the 1x1 and
// stride-2 convs, odd channel counts, the narrow layers, and the 3xTF32 mode.
struct WgradP {
  const float* X; const float* dY; float* dW;
  int M, Ca, Nc, ldx, ldy;
  int mode, taps, kw, Ha, Wa, Hc, Wc, stride, pad;
  int rows_per_split, tiles_n;
  int vecX, vecY, precise;
};

constexpr int WG_T = 64, WG_PITCH = WG_T + 8, WG_PX = 64;      // 64 x 64 output tile, 64-pixel reduction chunks

__global__ void __launch_bounds__(GM_THREADS, 2) wgrad_kernel(const WgradP p) {
  __shared__ __align__(16) float Xs[WG_PX * WG_PITCH];
  __shared__ __align__(16) float Ys[WG_PX * WG_PITCH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;            // 2 (ci) x 4 (co) warps: 32 x 16 each
  const int ci0 = (blockIdx.x / p.tiles_n) * WG_T, co0 = (blockIdx.x % p.tiles_n) * WG_T;
  const int tap = blockIdx.y, dy = tap / p.kw, dx = tap % p.kw;
  const int mb = blockIdx.z * p.rows_per_split;
  const int me = min(p.M, mb + p.rows_per_split);
  const int hw = p.Hc * p.Wc;
  const int kr = tid >> 4, q4 = (tid & 15) * 4;       // staged elements: pixels kr + 16 j of the chunk, channels q4 .. q4+3
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  float4 rx[4], ry[4];
  auto fetch = [&](int mc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = mc + kr + 16 * j;
      rx[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      ry[j] = rx[j];
      if (m < me) {
        long long src = m;
        if (p.mode != 0) {
          const int n = m / hw, rem = m - n * hw, y = rem / p.Wc, x = rem - y * p.Wc;
          src = gather_row(p.mode, n, y, x, dy, dx, p.Ha, p.Wa, p.stride, p.pad);
        }
        if (src >= 0) rx[j] = load4(p.X + src * p.ldx + ci0 + q4, p.Ca - (ci0 + q4), p.vecX != 0);
        ry[j] = load4(p.dY + (long long)m * p.ldy + co0 + q4, p.Nc - (co0 + q4), p.vecY != 0);
      }
    }
  };
  if (mb < me) fetch(mb);
  for (int mc = mb; mc < me; mc += WG_PX) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<float4*>(&Xs[(kr + 16 * j) * WG_PITCH + q4]) = rx[j];
      *reinterpret_cast<float4*>(&Ys[(kr + 16 * j) * WG_PITCH + q4]) = ry[j];
    }
    __syncthreads();
    if (mc + WG_PX < me) fetch(mc + WG_PX);
#pragma unroll
    for (int k8 = 0; k8 < WG_PX / 8; ++k8) {
      float af[2][4], bfr[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {                // A(row = ci, col = pixel) = Xs[pixel][ci]
        const float* a = &Xs[(k8 * 8 + t) * WG_PITCH + wm * 32 + mt * 16 + g];
        af[mt][0] = a[0];
        af[mt][1] = a[8];
        af[mt][2] = a[4 * WG_PITCH];
        af[mt][3] = a[4 * WG_PITCH + 8];
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float* b = &Ys[(k8 * 8 + t) * WG_PITCH + wn * 16 + nt * 8 + g];
        bfr[nt][0] = b[0];
        bfr[nt][1] = b[4 * WG_PITCH];
      }
      uint32_t ah[2][4], bh[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) ah[mt][i] = to_tf32(af[mt][i]);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { bh[nt][0] = to_tf32(bfr[nt][0]); bh[nt][1] = to_tf32(bfr[nt][1]); }
      if (p.precise) {
        uint32_t al[2][4], bl[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) al[mt][i] = to_tf32(af[mt][i] - __uint_as_float(ah[mt][i]));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) bl[nt][i] = to_tf32(bfr[nt][i] - __uint_as_float(bh[nt][i]));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {          // partial sum from zero, added round-to-nearest (see gather_gemm_kernel)
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            mma_tf32(part, al[mt], bh[nt]);
            mma_tf32(part, ah[mt], bl[nt]);
            mma_tf32(part, ah[mt], bh[nt]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] += part[i];
          }
      } else {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
      }
    }
    __syncthreads();
  }
  float* W = p.dW + (long long)tap * p.Ca * p.Nc;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ci = ci0 + wm * 32 + mt * 16 + g + 8 * h;
        if (ci >= p.Ca) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = co0 + wn * 16 + nt * 8 + 2 * t + e;
          if (co < p.Nc) atomicAdd(W + (long long)ci * p.Nc + co, acc[mt][nt][2 * h + e]);
        }
      }
}


// The same tile with a 3-stage cp.async ring (aligned operands, Ca % 4 == Nc % 4 == 0): the register-staged version above keeps
// ONE chunk of global loads in flight per thread and is bound by that latency, not by barriers or issue slots.
constexpr int WGA_STAGES = 3;

__device__ __forceinline__ void wg_cp16(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16u : 0u) : "memory");
}

__global__ void __launch_bounds__(GM_THREADS, 2) wgrad_async_kernel(const WgradP p) {
  extern __shared__ __align__(16) float wg_smem[];                // [stage][X | Y][WG_PX][WG_PITCH]
  constexpr int TILE = WG_PX * WG_PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int ci0 = (blockIdx.x / p.tiles_n) * WG_T, co0 = (blockIdx.x % p.tiles_n) * WG_T;
  const int tap = blockIdx.y, dy = tap / p.kw, dx = tap % p.kw;
  const int mb = blockIdx.z * p.rows_per_split;
  const int me = min(p.M, mb + p.rows_per_split);
  const int hw = p.Hc * p.Wc;
  const int kr = tid >> 4, q4 = (tid & 15) * 4;
  const bool xok = ci0 + q4 < p.Ca, yok = co0 + q4 < p.Nc;
  const int nchunk = me > mb ? (me - mb + WG_PX - 1) / WG_PX : 0;
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  auto issue = [&](int c) {
    if (c < nchunk) {
      float* Xs = wg_smem + (c % WGA_STAGES) * 2 * TILE;
      float* Ys = Xs + TILE;
      const int mc = mb + c * WG_PX;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = mc + kr + 16 * j;
        long long src = -1;
        if (m < me) {
          src = m;
          if (p.mode != 0) {
            const int n = m / hw, rem = m - n * hw, y = rem / p.Wc, x = rem - y * p.Wc;
            src = gather_row(p.mode, n, y, x, dy, dx, p.Ha, p.Wa, p.stride, p.pad);
          }
        }
        wg_cp16(&Xs[(kr + 16 * j) * WG_PITCH + q4], src >= 0 && xok ? p.X + src * p.ldx + ci0 + q4 : p.X, src >= 0 && xok);
        wg_cp16(&Ys[(kr + 16 * j) * WG_PITCH + q4], m < me && yok ? p.dY + (long long)m * p.ldy + co0 + q4 : p.dY, m < me && yok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  issue(1);
  for (int c = 0; c < nchunk; ++c) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                             // chunk c has landed for everyone; chunk c - 1 is consumed
    issue(c + 2);
    const float* Xs = wg_smem + (c % WGA_STAGES) * 2 * TILE;
    const float* Ys = Xs + TILE;
#pragma unroll
    for (int k8 = 0; k8 < WG_PX / 8; ++k8) {
      uint32_t ah[2][4], bh[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* a = &Xs[(k8 * 8 + t) * WG_PITCH + wm * 32 + mt * 16 + g];
        ah[mt][0] = to_tf32(a[0]);
        ah[mt][1] = to_tf32(a[8]);
        ah[mt][2] = to_tf32(a[4 * WG_PITCH]);
        ah[mt][3] = to_tf32(a[4 * WG_PITCH + 8]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float* b = &Ys[(k8 * 8 + t) * WG_PITCH + wn * 16 + nt * 8 + g];
        bh[nt][0] = to_tf32(b[0]);
        bh[nt][1] = to_tf32(b[4 * WG_PITCH]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* W = p.dW + (long long)tap * p.Ca * p.Nc;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ci = ci0 + wm * 32 + mt * 16 + g + 8 * h;
        if (ci >= p.Ca) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = co0 + wn * 16 + nt * 8 + 2 * t + e;
          if (co < p.Nc) atomicAdd(W + (long long)ci * p.Nc + co, acc[mt][nt][2 * h + e]);
        }
      }
}


// Narrow layers, 3x3 stride-1 convs (the C0 = 32 branch at 64x48 and the heads at 128x96): the per-tap CTAs of wgrad32_kernel
// read X and dY nine times (226 MB of L2 traffic for a 12.6 MB layer: 72 us).  Flat form as in train_tc5.cu: the pixels are one
// array of pitch P = W + 1 (zero column / zero row per image = padding); a CTA takes 128 consecutive positions of dY and the
// 128 + 2 P + 2 halo positions of X ONCE into shared memory (cp.async, two stages), and its nine warps each own one tap:
// warp t multiplies the X rows shifted by dy P + dx with the same dY rows into its own 32 x 32 tile.
constexpr int WF_THREADS = 288, WF_PX = 128, WF_PITCH = 40;

struct WgradFlatP {
  const float* X; const float* dY; float* dW;
  int Nimg, H, W, Ca, Nc, ldx, ldy, P, RPI, NPH;
  long long total, per_split;
  uint32_t magicP, magicR;          // ceil(2^32 / P), ceil(2^32 / RPI): exact multiply-high quotients for every flat position
};

__device__ __forceinline__ int wf_pixel(const WgradFlatP& p, long long gpos) {
  if (gpos < 0 || gpos >= p.total) return -1;
  const int gi = (int)gpos;                                       // total < 2^31 (host check): 32-bit divisions
  const int R = gi / p.P, Xc = gi - R * p.P;
  const int n = R / p.RPI, yy = R - n * p.RPI;
  if (Xc == 0 || yy == 0) return -1;
  return (n * p.H + (yy - 1)) * p.W + (Xc - 1);
}

__global__ void __launch_bounds__(WF_THREADS, 1) wgrad32_flat_kernel(const WgradFlatP p) {
  extern __shared__ __align__(16) float wf_smem[];                // [stage][X halo NPH rows | dY 128 rows][WF_PITCH]
  const int stage_f = (p.NPH + WF_PX) * WF_PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const long long fb = (long long)blockIdx.x * p.per_split;
  long long fe = fb + p.per_split;
  if (fe > p.total) fe = p.total;
  const int nchunk = fe > fb ? (int)((fe - fb + WF_PX - 1) / WF_PX) : 0;
  const int q4 = (tid & 7) * 4, r0 = tid >> 3;                   // staged: rows r0 + 36 j, channels q4 .. q4+3
  const bool xok = q4 < p.Ca, yok = q4 < p.Nc;
  const int dy = warp / 3, dx = warp - 3 * dy, off = dy * p.P + dx;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  auto issue = [&](int c) {
    if (c < nchunk) {
      float* Xs = wf_smem + (c & 1) * stage_f;
      float* Ys = Xs + p.NPH * WF_PITCH;
      const long long f0 = fb + (long long)c * WF_PX;
      // flat position -> pixel by multiply-high division (magic numbers from the host); the position is shifted by one image
      // block so that negative halo positions decode too (image index -1 = outside)
      auto walk = [&](long long g0, int nrows, long long glimit, auto&& emit) {
        long long gp = g0;
        for (int r = r0; r < nrows; r += WF_THREADS / 8, gp += WF_THREADS / 8) {
          const uint32_t gs = (uint32_t)((int)gp + p.RPI * p.P);
          const uint32_t R = __umulhi(gs, p.magicP), Xc = gs - R * (uint32_t)p.P;
          const uint32_t nn = __umulhi(R, p.magicR), yy = R - nn * (uint32_t)p.RPI;
          const int n = (int)nn - 1;
          const bool ok = gp < glimit && n >= 0 && n < p.Nimg && Xc != 0 && yy != 0;
          emit(r, ok, (n * p.H + ((int)yy - 1)) * p.W + ((int)Xc - 1));
        }
      };
      walk(f0 - p.P - 1 + r0, p.NPH, fe + p.P + 1, [&](int r, bool ok, int px) {
        wg_cp16(&Xs[r * WF_PITCH + q4], ok && xok ? p.X + (long long)px * p.ldx + q4 : p.X, ok && xok);
      });
      walk(f0 + r0, WF_PX, fe, [&](int r, bool ok, int px) {
        wg_cp16(&Ys[r * WF_PITCH + q4], ok && yok ? p.dY + (long long)px * p.ldy + q4 : p.dY, ok && yok);
      });
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  for (int c = 0; c < nchunk; ++c) {
    issue(c + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* Xs = wf_smem + (c & 1) * stage_f + off * WF_PITCH;
    const float* Ys = wf_smem + (c & 1) * stage_f + p.NPH * WF_PITCH;
#pragma unroll 4
    for (int k8 = 0; k8 < WF_PX / 8; ++k8) {
      uint32_t ah[2][4], bh[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* a = &Xs[(k8 * 8 + tq) * WF_PITCH + mt * 16 + g];
        ah[mt][0] = to_tf32(a[0]);
        ah[mt][1] = to_tf32(a[8]);
        ah[mt][2] = to_tf32(a[4 * WF_PITCH]);
        ah[mt][3] = to_tf32(a[4 * WF_PITCH + 8]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* b = &Ys[(k8 * 8 + tq) * WF_PITCH + nt * 8 + g];
        bh[nt][0] = to_tf32(b[0]);
        bh[nt][1] = to_tf32(b[4 * WF_PITCH]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* Wt = p.dW + (long long)warp * p.Ca * p.Nc;               // tap = warp
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = mt * 16 + g + 8 * h, co = nt * 8 + 2 * tq + e;
          if (ci < p.Ca && co < p.Nc) atomicAdd(Wt + (long long)ci * p.Nc + co, acc[mt][nt][2 * h + e]);
        }
}

}  // namespace

// train_tc5.cu: the tcgen05 kind::tf32 kernels (K-major B only)
int train_tc5_supported(const float* A, const float* B, int Ca, int lda, int ldb, long long sA, long long sB, long long tapB,
                        int transA, int transB, int precise);
int train_tc5_launch(cudaStream_t s, const float* A, const float* B, float* C, const float* bias, int M, int Nc, int Ca, int lda,
                     int ldb, int ldc, int batch, long long sA, long long sB, long long sC, long long tapB, int mode, int beta,
                     const int* geom);

// train_wgrad5.cu: tcgen05 weight gradients (bf16 hi / lo splits) of the wide 3x3 stride-1 layers; -1 = not covered
int train_wgrad5_launch(cudaStream_t s, const float* X, const float* dY, float* dW, int M, int Ca, int Nc, int ldx, int ldy,
                        const int* geom);

namespace {

// wgrad for narrow layers (Ca <= 32 and Nc <= 32: the C0 = 32 branch at 64x48 and the heads at 128x96, 40 % of all weight-
// gradient time with the 64 x 64 tile above, which is 75 % empty for them): the whole 32 x 32 tile belongs to every warp,
// the 8 warps split the PIXELS of a 128-pixel chunk (16 each), and the 8 partial tiles are summed through shared memory
// before one atomicAdd per element and CTA.
constexpr int W32_PX = 128, W32_PITCH = 40;    // Xs[px][ci]: bank = (8 t + g) mod 32

__global__ void __launch_bounds__(GM_THREADS, 2) wgrad32_kernel(const WgradP p) {
  __shared__ __align__(16) float Xs[W32_PX * W32_PITCH];
  __shared__ __align__(16) float Ys[W32_PX * W32_PITCH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int tap = blockIdx.y, dy = tap / p.kw, dx = tap % p.kw;
  const int mb = blockIdx.z * p.rows_per_split;
  const int me = min(p.M, mb + p.rows_per_split);
  const int hw = p.Hc * p.Wc;
  const int q4 = (tid & 7) * 4, pr0 = tid >> 3;       // staged: pixels pr0 + 32 j, channels q4 .. q4+3
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  float4 rx[4], ry[4];
  auto fetch = [&](int mc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = mc + pr0 + 32 * j;
      rx[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      ry[j] = rx[j];
      if (m < me) {
        long long src = m;
        if (p.mode != 0) {
          const int n = m / hw, rem = m - n * hw, y = rem / p.Wc, x = rem - y * p.Wc;
          src = gather_row(p.mode, n, y, x, dy, dx, p.Ha, p.Wa, p.stride, p.pad);
        }
        if (src >= 0) rx[j] = load4(p.X + src * p.ldx + q4, p.Ca - q4, p.vecX != 0);
        ry[j] = load4(p.dY + (long long)m * p.ldy + q4, p.Nc - q4, p.vecY != 0);
      }
    }
  };
  if (mb < me) fetch(mb);
  for (int mc = mb; mc < me; mc += W32_PX) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<float4*>(&Xs[(pr0 + 32 * j) * W32_PITCH + q4]) = rx[j];
      *reinterpret_cast<float4*>(&Ys[(pr0 + 32 * j) * W32_PITCH + q4]) = ry[j];
    }
    __syncthreads();
    if (mc + W32_PX < me) fetch(mc + W32_PX);
#pragma unroll
    for (int k8 = 0; k8 < 2; ++k8) {
      const int kr = warp * 16 + k8 * 8;
      float af[2][4], bfr[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* a = &Xs[(kr + t) * W32_PITCH + mt * 16 + g];
        af[mt][0] = a[0];
        af[mt][1] = a[8];
        af[mt][2] = a[4 * W32_PITCH];
        af[mt][3] = a[4 * W32_PITCH + 8];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* b = &Ys[(kr + t) * W32_PITCH + nt * 8 + g];
        bfr[nt][0] = b[0];
        bfr[nt][1] = b[4 * W32_PITCH];
      }
      uint32_t ah[2][4], bh[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) ah[mt][i] = to_tf32(af[mt][i]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { bh[nt][0] = to_tf32(bfr[nt][0]); bh[nt][1] = to_tf32(bfr[nt][1]); }
      if (p.precise) {
        uint32_t al[2][4], bl[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) al[mt][i] = to_tf32(af[mt][i] - __uint_as_float(ah[mt][i]));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) bl[nt][i] = to_tf32(bfr[nt][i] - __uint_as_float(bh[nt][i]));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            mma_tf32(part, al[mt], bh[nt]);
            mma_tf32(part, ah[mt], bl[nt]);
            mma_tf32(part, ah[mt], bh[nt]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] += part[i];
          }
      } else {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
      }
    }
    __syncthreads();
  }
  // cross-warp sum through the staging buffers: the 32 x 32 partial tiles of warps 0-4 in Xs, of warps 5-7 in Ys
  static_assert(W32_PX * W32_PITCH >= 5 * 1024, "reduction buffer");
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ci = mt * 16 + g + 8 * h, co = nt * 8 + 2 * t + e;
          float* dst = (warp < 5 ? Xs : Ys) + ((warp < 5 ? warp : warp - 5) * 1024 + ci * 32 + co);
          *dst = acc[mt][nt][2 * h + e];
        }
  __syncthreads();
  float* W = p.dW + (long long)tap * p.Ca * p.Nc;
  for (int i = tid; i < 1024; i += GM_THREADS) {
    const int ci = i >> 5, co = i & 31;
    if (ci >= p.Ca || co >= p.Nc) continue;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += (w < 5 ? Xs : Ys)[(w < 5 ? w : w - 5) * 1024 + i];
    atomicAdd(W + (long long)ci * p.Nc + co, sum);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// C[b] (M x Nc, row pitch ldc) = / += gather(A[b]) (M x taps*Ca) . B[b] (+ bias).  See the header for the argument
// meaning; geom = {Ha, Wa, Hc, Wc, kh, kw, stride, pad} (ignored for mode 0).
extern "C" int rsg_train_gemm(void* stream, const float* A, const float* B, float* C, const float* bias, int M, int Nc,
                              int Ca, int lda, int ldb, int ldc, int batch, long long sA, long long sB, long long sC,
                              int mode, int transA, int transB, int beta, const int* geom, int precise) {
  RSG_REQUIRE(A && B && C, "train gemm: null operand");
  RSG_REQUIRE(M >= 0 && Nc > 0 && Ca > 0 && batch >= 0, "train gemm: bad sizes M=%d Nc=%d Ca=%d", M, Nc, Ca);
  RSG_REQUIRE(mode >= 0 && mode <= 2, "train gemm: mode %d", mode);
  RSG_REQUIRE(mode == 0 || (geom && !transA), "train gemm: gather modes need geom and a row-major A");
  if (M == 0 || batch == 0) return RSG_OK;
  GemmP p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.B = B; p.C = C; p.bias = bias; p.M = M; p.Nc = Nc; p.Ca = Ca; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = sA; p.sB = sB; p.sC = sC; p.mode = mode; p.transA = transA; p.transB = transB; p.beta = beta;
  p.taps = 1; p.kw = 1; p.precise = precise;
  if (mode != 0) {
    p.Ha = geom[0]; p.Wa = geom[1]; p.Hc = geom[2]; p.Wc = geom[3];
    p.taps = geom[4] * geom[5]; p.kw = geom[5]; p.stride = geom[6]; p.pad = geom[7];
    RSG_REQUIRE(p.taps >= 1 && p.stride >= 1 && p.Hc > 0 && p.Wc > 0 && M % (p.Hc * p.Wc) == 0, "train gemm: bad geometry");
  }
  p.tapB = transB ? (long long)Nc * ldb : (long long)Ca * ldb;
  // precise: 0 = TF32 on the tcgen05 kernel where it applies (K-major B), 1 = 3xTF32 on mma.sync, 2 = TF32 on mma.sync
  if (train_tc5_supported(A, B, Ca, lda, ldb, sA, sB, p.tapB, transA, transB, precise))
    return train_tc5_launch((cudaStream_t)stream, A, B, C, bias, M, Nc, Ca, lda, ldb, ldc, batch, sA, sB, sC, p.tapB, mode, beta, geom);
  p.precise = precise == 1;
  p.vecA = al16(A) && lda % 4 == 0 && sA % 4 == 0;
  p.vecB = al16(B) && ldb % 4 == 0 && sB % 4 == 0 && p.tapB % 4 == 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int gx = ceil_div(M, GM_BM);
  if (Nc <= 32) {
    gather_gemm_kernel<32><<<dim3(gx, ceil_div(Nc, 32), batch), GM_THREADS, 0, s>>>(p);
  } else {
    gather_gemm_kernel<64><<<dim3(gx, ceil_div(Nc, 64), batch), GM_THREADS, 0, s>>>(p);
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

// dW[tap][Ca][Nc] += sum_m X[src(m, tap)][0..Ca) (x) dY[m][0..Nc)
extern "C" int rsg_train_wgrad(void* stream, const float* X, const float* dY, float* dW, int M, int Ca, int Nc, int ldx,
                               int ldy, int mode, const int* geom, int precise) {
  RSG_REQUIRE(X && dY && dW, "train wgrad: null operand");
  RSG_REQUIRE(M >= 0 && Ca > 0 && Nc > 0, "train wgrad: bad sizes");
  RSG_REQUIRE(mode >= 0 && mode <= 2 && (mode == 0 || geom), "train wgrad: mode / geometry");
  if (M == 0) return RSG_OK;
  WgradP p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.dY = dY; p.dW = dW; p.M = M; p.Ca = Ca; p.Nc = Nc; p.ldx = ldx; p.ldy = ldy; p.mode = mode;
  p.taps = 1; p.kw = 1; p.Hc = 1; p.Wc = 1; p.precise = precise == 1;
  if (mode != 0) {
    p.Ha = geom[0]; p.Wa = geom[1]; p.Hc = geom[2]; p.Wc = geom[3];
    p.taps = geom[4] * geom[5]; p.kw = geom[5]; p.stride = geom[6]; p.pad = geom[7];
    RSG_REQUIRE(p.taps >= 1 && p.stride >= 1 && p.Hc > 0 && p.Wc > 0 && M % (p.Hc * p.Wc) == 0, "train wgrad: bad geometry");
  }
  p.vecX = al16(X) && ldx % 4 == 0;
  p.vecY = al16(dY) && ldy % 4 == 0;
  if (precise == 0 && mode == 1) {             // wide 3x3 stride-1 layers: tensor pipe (tcgen05, bf16 hi / lo splits)
    const int r = train_wgrad5_launch((cudaStream_t)stream, X, dY, dW, M, Ca, Nc, ldx, ldy, geom);
    if (r >= 0) return r;
  }
  if (Ca <= 32 && Nc <= 32 && !p.precise && mode == 1 && geom[4] == 3 && geom[5] == 3 && geom[6] == 1 && geom[7] == 1 &&
      geom[0] == geom[2] && geom[1] == geom[3] && geom[1] <= 110 && Ca % 4 == 0 && Nc % 4 == 0 && p.vecX && p.vecY &&
      (long long)M * ldx < (1ll << 31) && (long long)M * ldy < (1ll << 31) &&
      (long long)(M / (geom[0] * geom[1])) * (geom[0] + 1) * (geom[1] + 1) < (1ll << 31) &&
      // the multiply-high position decode is exact for positions below 2^32 / P (two image blocks of slack: shift + halo)
      ((long long)(M / (geom[0] * geom[1]) + 2) * (geom[0] + 1) * (geom[1] + 1) + 512) * (geom[1] + 1) < (1ll << 32)) {
    WgradFlatP f;                              // narrow 3x3 stride-1 layers: flat form, X and dY read once
    memset(&f, 0, sizeof(f));
    f.X = X; f.dY = dY; f.dW = dW; f.H = geom[0]; f.W = geom[1]; f.Nimg = M / (f.H * f.W); f.Ca = Ca; f.Nc = Nc; f.ldx = ldx; f.ldy = ldy;
    f.P = f.W + 1; f.RPI = f.H + 1; f.NPH = WF_PX + 2 * f.P + 2;
    f.total = (long long)f.Nimg * f.RPI * f.P;
    f.magicP = (uint32_t)(((1ull << 32) + f.P - 1) / f.P);
    f.magicR = (uint32_t)(((1ull << 32) + f.RPI - 1) / f.RPI);
    long long splits = 2ll * rsg_num_sms(), maxsplit = (f.total + 4 * WF_PX - 1) / (4 * WF_PX);
    if (splits > maxsplit) splits = maxsplit;
    if (splits < 1) splits = 1;
    f.per_split = ((f.total + splits - 1) / splits + WF_PX - 1) / WF_PX * WF_PX;
    splits = (f.total + f.per_split - 1) / f.per_split;
    const size_t smem = 2 * (size_t)(f.NPH + WF_PX) * WF_PITCH * sizeof(float);
    static DeviceOnce once_f;
    if (once_f.first()) {
      RSG_CUDA(cudaFuncSetAttribute(wgrad32_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      once_f.done();
    }
    if (smem <= 200 * 1024) {
      wgrad32_flat_kernel<<<dim3((unsigned)splits), WF_THREADS, smem, (cudaStream_t)stream>>>(f);
      RSG_LAUNCH_CHECK();
      return RSG_OK;
    }
  }
  if (Ca <= 32 && Nc <= 32) {                  // narrow layers: pixel-split warps on one 32 x 32 tile
    long long want = (4ll * rsg_num_sms() + p.taps - 1) / p.taps, maxsplit = (M + 1023) / 1024;
    if (want > maxsplit) want = maxsplit;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    int rows = ceil_div(M, want);
    rows = (rows + W32_PX - 1) / W32_PX * W32_PX;
    p.rows_per_split = rows;
    wgrad32_kernel<<<dim3(1, (unsigned)p.taps, (unsigned)ceil_div(M, rows)), GM_THREADS, 0, (cudaStream_t)stream>>>(p);
    RSG_LAUNCH_CHECK();
    return RSG_OK;
  }
  const int tiles_m = ceil_div(Ca, WG_T);
  p.tiles_n = ceil_div(Nc, WG_T);
  const long long base = (long long)tiles_m * p.tiles_n * p.taps;
  long long want = (4ll * rsg_num_sms() + base - 1) / base;          // ~4 CTAs per SM
  long long maxsplit = (M + 255) / 256;                              // at least 256 pixels per CTA
  if (want > maxsplit) want = maxsplit;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  int rows = ceil_div(M, want);
  rows = (rows + WG_PX - 1) / WG_PX * WG_PX;
  p.rows_per_split = rows;
  const int ksplit = ceil_div(M, rows);
  const dim3 grid((unsigned)(tiles_m * p.tiles_n), (unsigned)p.taps, (unsigned)ksplit);
  if (!p.precise && p.vecX && p.vecY && Ca % 4 == 0 && Nc % 4 == 0) {
    const size_t smem = (size_t)WGA_STAGES * 2 * WG_PX * WG_PITCH * sizeof(float);
    static DeviceOnce once;
    if (once.first()) {
      RSG_CUDA(cudaFuncSetAttribute(wgrad_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      once.done();
    }
    wgrad_async_kernel<<<grid, GM_THREADS, smem, (cudaStream_t)stream>>>(p);
  } else {
    wgrad_kernel<<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(p);
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
