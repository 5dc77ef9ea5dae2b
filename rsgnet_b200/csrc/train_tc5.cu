// Training-step implicit GEMM on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, operands
// gathered into shared memory by cp.async (what cuDNN's TF32 fprop / dgrad kernels do for the reference).
//
//   C[m, n] (+)= sum_tap sum_c A[src(m, tap), c] * B[tap][n][c]  (+ bias[n])
//
// A rows are fp32 NHWC pixels (contiguous channels) GATHERED per tap -- forward: (y*stride - pad + dy, x*stride - pad + dx),
// input gradient / ConvTranspose: ((y + pad - dy) / stride, ...) when exact -- and B is K-major, i.e. [tap][n][c] with the
// reduction index contiguous: the forward pass reads the [tap][Cout][Cin] copy of the weights, the input-gradient pass the
// [tap][Cin][Cout] copy (rsgnet_b200/train/step.py keeps both).  Both operands therefore land in the canonical K-major
// SWIZZLE_NONE UMMA layout with 16-byte cp.async copies: a 16-byte k-group (4 floats) of a row is one row of a core
// matrix.  Shared-memory tile = [8 k-groups][rows + 1][16 B]: the +1 row of pitch makes the 8 lanes that copy one
// 128-byte row segment hit 8 different 16-byte bank groups, while the global side stays coalesced (8 lanes x 16 B = one line).
//
// CTA = one 128 x N output tile (N = Cout rounded up to 16, <= 256), 160 threads:
//   warps 0-3  gather producers (32 floats of K per stage, 4-stage ring; zero fill = conv padding / tile tails), then the
//              epilogue: TMEM -> registers (+ bias, + C when accumulating) -> fp32 rows of C
//   warp 4     TMEM allocation; one elected thread issues 4 MMAs (K = 8) per stage and commits the stage back to the producers
// Two or three CTAs are resident per SM, so one CTA's epilogue overlaps its neighbours' main loops.
#include "umma.cuh"
#include "../../include/rsg_b200.h"

namespace {
using namespace umma;

constexpr int T5_THREADS = 160, T5_PROD = 128, T5_BM = 128, T5_BK = 32, T5_KG = 8, T5_STAGES = 4, T5_LAG = 2;

struct T5P {
  const float* A; const float* B; float* C; const float* bias;
  int M, Nc, Ca, lda, ldb, ldc;
  long long sA, sB, sC, tapB;
  int mode, beta, taps, kw, Ha, Wa, Hc, Wc, stride, pad;
  int n_mma;                       // MMA N: multiple of 16, <= 256
  uint32_t a_plane, b_plane, stage_bytes, tmem_cols;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// 32-byte store (STG.256): one full sector per lane and instruction.  Row-per-lane epilogues write 128-byte rows; with 16-byte
// stores every sector is written by two instructions (partial-sector writes in L2).
__device__ __forceinline__ void stg256(float* p, const float* f) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])),
               "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])),
               "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ long long t5_gather_row(int mode, int n, int y, int x, int dy, int dx, int Ha, int Wa, int stride, int pad) {
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

__global__ void __launch_bounds__(T5_THREADS) gemm_tf32_tc5_kernel(const T5P p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * T5_STAGES + 1];      // full[S] | empty[S] | accumulator full
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * T5_BM, n0 = blockIdx.y * p.n_mma;
  const float* Ap = p.A + (long long)blockIdx.z * p.sA;
  const float* Bp = p.B + (long long)blockIdx.z * p.sB;
  float* Cp = p.C + (long long)blockIdx.z * p.sC;
  const int kchunks = (p.Ca + T5_BK - 1) / T5_BK;
  const int niter = p.taps * kchunks;

  if (tid == 0) {
    for (int i = 0; i < T5_STAGES; ++i) { mbar_init(BAR(i), T5_PROD); mbar_init(BAR(T5_STAGES + i), 1); }
    mbar_init(BAR(2 * T5_STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    // ===================== gather producers =====================
    const int q = tid & 7, rr = tid >> 3;                       // k-group of the stage, first row (rows rr + 16 j)
    int rn[8], ry[8], rx[8];
    uint32_t rok = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + rr + 16 * j;
      rn[j] = m; ry[j] = 0; rx[j] = 0;
      if (m < p.M) {
        rok |= 1u << j;
        if (p.mode != 0) {
          const int hw = p.Hc * p.Wc;
          rn[j] = m / hw;
          const int rem = m - rn[j] * hw;
          ry[j] = rem / p.Wc;
          rx[j] = rem - ry[j] * p.Wc;
        }
      }
    }
    const int nb = p.n_mma >> 4;                                // B rows per thread: rr + 16 j, j < nb
    uint32_t s = 0, ph = 0;
    const float* srcp[8];                                       // source row of this thread's 8 rows for the current tap
    int tap = 0, ch = 0;
    for (int it = 0; it < niter; ++it) {
      if (ch == 0) {                                            // a new tap: the gather is the same for all its channel chunks
        const int dy = tap / p.kw, dx = tap - dy * p.kw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          long long src = -1;
          if ((rok >> j) & 1u) src = p.mode == 0 ? (long long)rn[j] : t5_gather_row(p.mode, rn[j], ry[j], rx[j], dy, dx, p.Ha, p.Wa, p.stride, p.pad);
          srcp[j] = src < 0 ? nullptr : Ap + src * p.lda + q * 4;
        }
      }
      const int c0 = ch * T5_BK, k = c0 + q * 4;
      const bool kok = k < p.Ca;
      mbar_wait(BAR(T5_STAGES + s), ph ^ 1u);
      const uint32_t sa = sbase + s * p.stage_bytes + (uint32_t)q * p.a_plane + (uint32_t)rr * 16u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = kok && srcp[j] != nullptr;
        cp_async16(sa + (uint32_t)j * 256u, ok ? (const void*)(srcp[j] + c0) : (const void*)Ap, ok ? 16u : 0u);
      }
      const uint32_t sb = sbase + s * p.stage_bytes + T5_KG * p.a_plane + (uint32_t)q * p.b_plane + (uint32_t)rr * 16u;
      const float* Bt = Bp + (long long)tap * p.tapB + k;
      for (int j = 0; j < nb; ++j) {
        const int n = n0 + rr + 16 * j;
        const bool ok = kok && n < p.Nc;
        cp_async16(sb + (uint32_t)j * 256u, ok ? (const void*)(Bt + (long long)n * p.ldb) : (const void*)Bp, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= T5_LAG) {                                        // the copies of iteration it - LAG have landed
        asm volatile("cp.async.wait_group %0;" ::"n"(T5_LAG) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(BAR((it - T5_LAG) % T5_STAGES));
      }
      if (++s == T5_STAGES) { s = 0; ph ^= 1u; }
      if (++ch == kchunks) { ch = 0; ++tap; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int it = niter > T5_LAG ? niter - T5_LAG : 0; it < niter; ++it) mbar_arrive(BAR(it % T5_STAGES));

    // ===================== epilogue: row m0 + 32 warp + lane =====================
    mbar_wait(BAR(2 * T5_STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* crow = Cp + (long long)m * p.ldc + n0;
    const bool vec = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(Cp) & 15u) == 0) && (n0 & 3) == 0;
    const bool vec32 = !p.beta && (p.ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(Cp) & 31u) == 0) && (n0 & 7) == 0;
    for (int c = 0; c < p.n_mma; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < p.M && vec32 && n0 + c + 15 < p.Nc) {               // two full 32-byte sectors per lane and instruction
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) + (p.bias ? p.bias[n0 + c + e] : 0.f);
        stg256(crow + c, f);
        stg256(crow + c + 8, f + 8);
      } else if (m < p.M) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c + 4 * g;
          if (n >= p.Nc) break;
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[e] = __uint_as_float(v[4 * g + e]);
            if (p.bias && n + e < p.Nc) f[e] += p.bias[n + e];
          }
          if (vec && n + 3 < p.Nc) {
            float4* dst = reinterpret_cast<float4*>(crow + c + 4 * g);
            if (p.beta) { const float4 o = *dst; f[0] += o.x; f[1] += o.y; f[2] += o.z; f[3] += o.w; }
            *dst = make_float4(f[0], f[1], f[2], f[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.Nc) {
                float* dst = crow + c + 4 * g + e;
                *dst = p.beta ? *dst + f[e] : f[e];
              }
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = desc_hi(128u);
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < niter; ++it) {
      mbar_wait(BAR(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a0 = sbase + s * p.stage_bytes, b0 = a0 + T5_KG * p.a_plane;
#pragma unroll
        for (int k8 = 0; k8 < T5_BK / 8; ++k8) {
          const uint64_t ad = ((uint64_t)hi << 32) | desc_lo(a0 + 2u * k8 * p.a_plane, p.a_plane);
          const uint64_t bd = ((uint64_t)hi << 32) | desc_lo(b0 + 2u * k8 * p.b_plane, p.b_plane);
          umma_tf32(tmem_base, ad, bd, idesc, (it | k8) ? 1u : 0u);
        }
        umma_commit(BAR(T5_STAGES + s));
        if (it == niter - 1) umma_commit(BAR(2 * T5_STAGES));
      }
      __syncwarp();
      if (++s == T5_STAGES) { s = 0; ph ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// "Flat" 3x3 stride-1 pad-1 form (forward and input gradient of every BasicBlock / Bottleneck / head conv): the pixels of
// all images form ONE flat array of pitch P = W + 1 whose column 0 and whose row 0 of every image block of H + 1 rows
// are zero (left / right and top / bottom padding at once, as in conv_ws.cu).  A CTA owns 128 consecutive flat positions;
// the NP = 128 + 2 P + 2 positions around them are gathered ONCE per 32-channel chunk (the per-tap gather above re-reads
// them 9 times), and a tap is only a start offset of the A descriptor: dy P + dx (forward), (2 - dy) P + (2 - dx) (input
// gradient: the same weights index, the mirrored neighbour).
struct T5F {
  const float* A; const float* B; float* C; const float* bias;
  int Nimg, H, W, Ca, Nc, lda, ldb, ldc;
  long long tapB;
  int P, RPI, NP, JH;              // pitch, rows per image block, halo positions, halo positions per producer thread
  long long total;                 // Nimg * RPI * P flat positions
  int flip, beta, n_mma;
  uint32_t a_plane, b_plane, a_bytes, b_stage, tmem_cols;
  uint32_t magicP, magicR;         // ceil(2^32 / P), ceil(2^32 / RPI): exact quotients for every flat position (< 2^32 / P)
};
constexpr int T5F_MAXJ = 22;       // NP <= 16 * 22 = 352: W <= 110

// flat position -> NHWC pixel index, -1 = padding.  32-bit arithmetic on purpose (the host checks total < 2^31): the 64-bit
// divisions of the first version cost ~100 instructions each, 30 per thread and tile -- more than the MMAs of a 32 -> 32 layer.
__device__ __forceinline__ int t5f_pixel(const T5F& p, long long g64) {
  if (g64 < 0 || g64 >= p.total) return -1;
  const uint32_t g = (uint32_t)g64;
  const uint32_t R = __umulhi(g, p.magicP), X = g - R * (uint32_t)p.P;          // exact: g < 2^32 / P (host check)
  const uint32_t n = __umulhi(R, p.magicR), yy = R - n * (uint32_t)p.RPI;
  if (X == 0 || yy == 0) return -1;
  return ((int)n * p.H + ((int)yy - 1)) * p.W + ((int)X - 1);
}

__global__ void __launch_bounds__(T5_THREADS) conv3x3_tf32_tc5_kernel(const T5F p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * T5_STAGES + 3];      // fullB[S] | emptyB[S] | emptyA[2] | accumulator full
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_EA = 2 * T5_STAGES, B_ACC = 2 * T5_STAGES + 2;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t sB = sbase + 2u * p.a_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long f0 = (long long)blockIdx.x * T5_BM;
  const int n0 = blockIdx.y * p.n_mma;
  const int kchunks = (p.Ca + T5_BK - 1) / T5_BK;
  const int niter = 9 * kchunks;

  if (tid == 0) {
    for (int i = 0; i < T5_STAGES; ++i) { mbar_init(BAR(i), T5_PROD); mbar_init(BAR(T5_STAGES + i), 1); }
    mbar_init(BAR(B_EA), 1); mbar_init(BAR(B_EA + 1), 1); mbar_init(BAR(B_ACC), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 4) {
    const int q = tid & 7, rr = tid >> 3;
    int pix[T5F_MAXJ];                                            // source pixel of halo position rr + 16 j (-1 = zero)
#pragma unroll
    for (int j = 0; j < T5F_MAXJ; ++j) {
      const int hp = rr + 16 * j;
      pix[j] = (j < p.JH && hp < p.NP) ? t5f_pixel(p, f0 - p.P - 1 + hp) : -2;      // -2 = beyond the halo: no copy at all
    }
    const int nb = p.n_mma >> 4;
    uint32_t s = 0, ph = 0;
    int it = 0;
    for (int ch = 0; ch < kchunks; ++ch) {
      const int k = ch * T5_BK + q * 4;
      const bool kok = k < p.Ca;
      for (int tap = 0; tap < 9; ++tap, ++it) {
        mbar_wait(BAR(T5_STAGES + s), ph ^ 1u);
        if (tap == 0) {
          mbar_wait(BAR(B_EA + (ch & 1)), ((uint32_t)(ch >> 1) & 1u) ^ 1u);
          const uint32_t sa = sbase + (uint32_t)(ch & 1) * p.a_bytes + (uint32_t)q * p.a_plane + (uint32_t)rr * 16u;
#pragma unroll
          for (int j = 0; j < T5F_MAXJ; ++j) {
            if (pix[j] == -2) continue;
            const bool ok = kok && pix[j] >= 0;
            cp_async16(sa + (uint32_t)j * 256u, ok ? (const void*)(p.A + (long long)pix[j] * p.lda + k) : (const void*)p.A, ok ? 16u : 0u);
          }
        }
        const uint32_t sb = sB + s * p.b_stage + (uint32_t)q * p.b_plane + (uint32_t)rr * 16u;
        const float* Bt = p.B + (long long)tap * p.tapB + k;
        for (int j = 0; j < nb; ++j) {
          const int n = n0 + rr + 16 * j;
          const bool ok = kok && n < p.Nc;
          cp_async16(sb + (uint32_t)j * 256u, ok ? (const void*)(Bt + (long long)n * p.ldb) : (const void*)p.B, ok ? 16u : 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= T5_LAG) {
          asm volatile("cp.async.wait_group %0;" ::"n"(T5_LAG) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(BAR((it - T5_LAG) % T5_STAGES));
        }
        if (++s == T5_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int i = niter > T5_LAG ? niter - T5_LAG : 0; i < niter; ++i) mbar_arrive(BAR(i % T5_STAGES));

    // epilogue: flat position f0 + 32 warp + lane
    mbar_wait(BAR(B_ACC), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = t5f_pixel(p, f0 + warp * 32 + lane);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* crow = p.C + (long long)(m < 0 ? 0 : m) * p.ldc + n0;
    const bool vec = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
    const bool vec32 = !p.beta && (p.ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 31u) == 0) && (n0 & 7) == 0;
    for (int c = 0; c < p.n_mma; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m >= 0 && vec32 && n0 + c + 15 < p.Nc) {                // two full 32-byte sectors per lane and instruction
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) + (p.bias ? p.bias[n0 + c + e] : 0.f);
        stg256(crow + c, f);
        stg256(crow + c + 8, f + 8);
      } else if (m >= 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c + 4 * g;
          if (n >= p.Nc) break;
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[e] = __uint_as_float(v[4 * g + e]);
            if (p.bias && n + e < p.Nc) f[e] += p.bias[n + e];
          }
          if (vec && n + 3 < p.Nc) {
            float4* dst = reinterpret_cast<float4*>(crow + c + 4 * g);
            if (p.beta) { const float4 o = *dst; f[0] += o.x; f[1] += o.y; f[2] += o.z; f[3] += o.w; }
            *dst = make_float4(f[0], f[1], f[2], f[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.Nc) {
                float* dst = crow + c + 4 * g + e;
                *dst = p.beta ? *dst + f[e] : f[e];
              }
          }
        }
      }
    }
  } else {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = desc_hi(128u);
    uint32_t s = 0, ph = 0;
    int it = 0;
    for (int ch = 0; ch < kchunks; ++ch) {
      const uint32_t abase = sbase + (uint32_t)(ch & 1) * p.a_bytes;
      for (int tap = 0; tap < 9; ++tap, ++it) {
        mbar_wait(BAR(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const int dy = tap / 3, dx = tap - 3 * dy;
          const uint32_t off = p.flip ? (uint32_t)((2 - dy) * p.P + (2 - dx)) : (uint32_t)(dy * p.P + dx);
          const uint32_t a0 = abase + off * 16u, b0 = sB + s * p.b_stage;
#pragma unroll
          for (int k8 = 0; k8 < T5_BK / 8; ++k8) {
            const uint64_t ad = ((uint64_t)hi << 32) | desc_lo(a0 + 2u * k8 * p.a_plane, p.a_plane);
            const uint64_t bd = ((uint64_t)hi << 32) | desc_lo(b0 + 2u * k8 * p.b_plane, p.b_plane);
            umma_tf32(tmem_base, ad, bd, idesc, (it | k8) ? 1u : 0u);
          }
          umma_commit(BAR(T5_STAGES + s));
          if (tap == 8) umma_commit(BAR(B_EA + (ch & 1)));
          if (it == niter - 1) umma_commit(BAR(B_ACC));
        }
        __syncwarp();
        if (++s == T5_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}



// The same flat 3x3 form for layers with at most 32 input channels (the C0 = 32 branch at 64x48 and the heads at 128x96: the
// convs on the critical path of a step): ONE channel chunk, so the halo and the weights of all nine taps (36 KB for 32 -> 32)
// are fetched in a single cp.async batch -- one barrier round trip instead of nine -- and the 36 MMAs are issued back to back.
__global__ void __launch_bounds__(T5_THREADS) conv3x3_tf32_small_kernel(const T5F p, const int mtiles) {
  // persistent: a CTA keeps the weights of all nine taps and walks over the 128-position tiles blockIdx.x, blockIdx.x +
  // gridDim.x, ...; per tile only the halo is fetched.  Two halo buffers and two TMEM accumulators: while the MMAs of tile i
  // run, the producers fetch the halo of tile i + 1 and write out tile i - 1 (one buffer / one accumulator made the CTA a
  // serial chain fetch -> MMA -> epilogue: tensor pipe 22 % busy)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];                       // operands landed [2] | accumulator full [2]
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t sB = sbase + 2u * p.a_bytes;                     // weights: [tap][8 k-groups][n_mma + 1][16 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * p.n_mma;
  if (tid == 0) {
    mbar_init(bar0, T5_PROD);
    mbar_init(bar0 + 8u, T5_PROD);
    mbar_init(bar0 + 16u, 1);
    mbar_init(bar0 + 24u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t acc_cols = p.tmem_cols >> 1;                     // second accumulator

  if (warp < 4) {
    const int q = tid & 7, rr = tid >> 3, k = q * 4;
    const bool kok = k < p.Ca;
    const uint32_t sa = sbase + (uint32_t)q * p.a_plane + (uint32_t)rr * 16u;
    const int nb = p.n_mma >> 4;
    for (int tap = 0; tap < 9; ++tap) {                            // the weights: once per CTA, in the first tile's batch
      const uint32_t sb = sB + (uint32_t)tap * p.b_stage + (uint32_t)q * p.b_plane + (uint32_t)rr * 16u;
      const float* Bt = p.B + (long long)tap * p.tapB + k;
      for (int j = 0; j < nb; ++j) {
        const int n = n0 + rr + 16 * j;
        const bool ok = kok && n < p.Nc;
        cp_async16(sb + (uint32_t)j * 256u, ok ? (const void*)(Bt + (long long)n * p.ldb) : (const void*)p.B, ok ? 16u : 0u);
      }
    }
    const bool vec = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
    const bool vec32 = (p.ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 31u) == 0) && (n0 & 7) == 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    // halo position rr + 16 j <-> flat position f0 - P - 1 + rr + 16 j, decoded by multiply-high division with the two
    // constants (magic numbers from the host; the position is shifted by one image block so that the first tile's negative
    // positions decode too: image index -1 = outside).  ncu / clock64 history: two integer divisions per position WERE the
    // kernel (38 M instructions, IPC 1.15); a loop-carried incremental update was 2450 of a tile's 8500 cycles.
    auto fetch = [&](int tile, uint32_t buf) {
      const uint32_t gs0 = (uint32_t)(tile * T5_BM - p.P - 1 + rr + p.RPI * p.P);
      const uint32_t dst = sa + buf * p.a_bytes;
      for (int j = 0; j < p.JH; ++j) {
        const int hp = rr + 16 * j;
        if (hp >= p.NP) break;
        const uint32_t gs = gs0 + 16u * (uint32_t)j;
        const uint32_t R = __umulhi(gs, p.magicP), X = gs - R * (uint32_t)p.P;
        const uint32_t nn = __umulhi(R, p.magicR), yy = R - nn * (uint32_t)p.RPI;
        const int n = (int)nn - 1;
        const bool ok = kok && n >= 0 && n < p.Nimg && X != 0 && yy != 0;
        const int px = (n * p.H + ((int)yy - 1)) * p.W + ((int)X - 1);
        cp_async16(dst + (uint32_t)j * 256u, ok ? (const void*)(p.A + (long long)px * p.lda + k) : (const void*)p.A, ok ? 16u : 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto epilogue = [&](int tile, uint32_t buf) {
      const int m = t5f_pixel(p, (long long)tile * T5_BM + warp * 32 + lane);
      float* crow = p.C + (long long)(m < 0 ? 0 : m) * p.ldc + n0;
      const uint32_t ta = taddr + buf * acc_cols;
      for (int c = 0; c < p.n_mma; c += 16) {
        uint32_t v[16];
        tmem_ld16(ta + (uint32_t)c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m >= 0) {
          if (vec32 && n0 + c + 15 < p.Nc) {                       // two full sectors per lane
            float f[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]) + (p.bias ? p.bias[n0 + c + e] : 0.f);
            stg256(crow + c, f);
            stg256(crow + c + 8, f + 8);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = n0 + c + 4 * g;
              if (n >= p.Nc) break;
              float f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[e] = __uint_as_float(v[4 * g + e]);
                if (p.bias && n + e < p.Nc) f[e] += p.bias[n + e];
              }
              if (vec && n + 3 < p.Nc) {
                *reinterpret_cast<float4*>(crow + c + 4 * g) = make_float4(f[0], f[1], f[2], f[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n + e < p.Nc) crow[c + 4 * g + e] = f[e];
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // a later tile's MMAs overwrite this accumulator
    };
    fetch(blockIdx.x, 0u);
    uint32_t it = 0;
    int prev = -1;
    for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1u;
      asm volatile("cp.async.wait_group 0;" ::: "memory");          // halo of `tile` (and, first time, the weights) landed
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar0 + 8u * b);                                   // -> the MMAs of `tile` may start
      if (prev >= 0) {                                              // MMAs of the previous tile done: its halo buffer is free
        mbar_wait(bar0 + 16u + 8u * (b ^ 1u), ((it - 1u) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int next = tile + (int)gridDim.x;
      if (next < mtiles) fetch(next, b ^ 1u);
      if (prev >= 0) epilogue(prev, b ^ 1u);
      prev = tile;
    }
    if (prev >= 0) {
      const uint32_t lb = (it - 1u) & 1u;
      mbar_wait(bar0 + 16u + 8u * lb, ((it - 1u) >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue(prev, lb);
    }
  } else {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = desc_hi(128u);
    const int kk = (p.Ca + 7) >> 3;                                // K = 8 steps that hold channels
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1u;
      mbar_wait(bar0 + 8u * b, (it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t abase = sbase + b * p.a_bytes, d = tmem_base + b * acc_cols;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - 3 * dy;
          const uint32_t off = p.flip ? (uint32_t)((2 - dy) * p.P + (2 - dx)) : (uint32_t)(dy * p.P + dx);
          const uint32_t a0 = abase + off * 16u, b0 = sB + (uint32_t)tap * p.b_stage;
          for (int k8 = 0; k8 < kk; ++k8) {
            const uint64_t ad = ((uint64_t)hi << 32) | desc_lo(a0 + 2u * k8 * p.a_plane, p.a_plane);
            const uint64_t bd = ((uint64_t)hi << 32) | desc_lo(b0 + 2u * k8 * p.b_plane, p.b_plane);
            umma_tf32(d, ad, bd, idesc, (tap | k8) ? 1u : 0u);
          }
        }
        umma_commit(bar0 + 16u + 8u * b);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace

// 1 when the tcgen05 kernel covers this call of rsg_train_gemm (else the mma.sync kernel runs)
int train_tc5_supported(const float* A, const float* B, int Ca, int lda, int ldb, long long sA, long long sB, long long tapB,
                        int transA, int transB, int precise) {
  if (precise || transA || !transB) return 0;
  if ((Ca & 3) || (lda & 3) || (ldb & 3) || (sA & 3) || (sB & 3) || (tapB & 3)) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(B) & 15u)) return 0;
  return 1;
}

int train_tc5_launch(cudaStream_t s, const float* A, const float* B, float* C, const float* bias, int M, int Nc, int Ca, int lda,
                     int ldb, int ldc, int batch, long long sA, long long sB, long long sC, long long tapB, int mode, int beta,
                     const int* geom) {
  T5P p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.B = B; p.C = C; p.bias = bias; p.M = M; p.Nc = Nc; p.Ca = Ca; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = sA; p.sB = sB; p.sC = sC; p.tapB = tapB; p.mode = mode; p.beta = beta; p.taps = 1; p.kw = 1;
  if (mode != 0) {
    p.Ha = geom[0]; p.Wa = geom[1]; p.Hc = geom[2]; p.Wc = geom[3];
    p.taps = geom[4] * geom[5]; p.kw = geom[5]; p.stride = geom[6]; p.pad = geom[7];
  }
  // every 3x3 stride-1 pad-1 convolution (forward: mode 1, input gradient: mode 2) takes the flat form
  if (mode != 0 && batch == 1 && !beta && geom[4] == 3 && geom[5] == 3 && geom[6] == 1 && geom[7] == 1 && geom[0] == geom[2] &&
      geom[1] == geom[3] && geom[1] <= 110 && (long long)M * lda < (1ll << 31) &&
      (long long)(M / (geom[0] * geom[1])) * (geom[0] + 1) * (geom[1] + 1) < (1ll << 31) &&
      // the multiply-high position decode is exact for positions below 2^32 / P (two image blocks of slack: shift + halo)
      ((long long)(M / (geom[0] * geom[1]) + 2) * (geom[0] + 1) * (geom[1] + 1) + 512) * (geom[1] + 1) < (1ll << 32)) {
    T5F f;
    memset(&f, 0, sizeof(f));
    f.A = A; f.B = B; f.C = C; f.bias = bias; f.H = geom[0]; f.W = geom[1]; f.Nimg = M / (f.H * f.W); f.Ca = Ca; f.Nc = Nc;
    f.lda = lda; f.ldb = ldb; f.ldc = ldc; f.tapB = tapB; f.P = f.W + 1; f.RPI = f.H + 1;
    f.NP = T5_BM + 2 * f.P + 2; f.JH = (f.NP + 15) / 16;
    f.total = (long long)f.Nimg * f.RPI * f.P;
    f.flip = mode == 2; f.beta = beta;
    f.magicP = (uint32_t)(((1ull << 32) + f.P - 1) / f.P);
    f.magicR = (uint32_t)(((1ull << 32) + f.RPI - 1) / f.RPI);
    const int mtiles = (int)((f.total + T5_BM - 1) / T5_BM);
    int nm = (Nc + 15) / 16 * 16;
    if (nm > 256) nm = 256;
    int nt = (Nc + nm - 1) / nm;
    nm = ((Nc + nt - 1) / nt + 15) / 16 * 16;
    while ((long long)mtiles * nt < 2ll * rsg_num_sms() && nm >= 64 && nm % 32 == 0) { nm /= 2; nt = (Nc + nm - 1) / nm; }   // small maps: more CTAs
    f.n_mma = nm;
    f.a_plane = (uint32_t)(f.NP + 1) * 16u;
    f.b_plane = (uint32_t)(nm + 1) * 16u;
    f.a_bytes = (T5_KG * f.a_plane + 127u) & ~127u;
    f.b_stage = (T5_KG * f.b_plane + 127u) & ~127u;
    uint32_t cols = 32;
    while (cols < (uint32_t)nm) cols <<= 1;
    f.tmem_cols = cols;
    static DeviceOnce once_f;
    if (once_f.first()) {
      RSG_CUDA(cudaFuncSetAttribute(conv3x3_tf32_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      RSG_CUDA(cudaFuncSetAttribute(conv3x3_tf32_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      once_f.done();
    }
    if (Ca <= 32 && nm <= 64) {                // one channel chunk: everything in one cp.async batch
      const size_t smem_s = 128 + 2 * (size_t)f.a_bytes + 9 * (size_t)f.b_stage;      // two halo buffers + the nine taps' weights
      if (smem_s <= 160 * 1024) {
        T5F fs = f;
        fs.tmem_cols = 2 * cols;                 // two accumulators
        int occ = 0;
        RSG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv3x3_tf32_small_kernel, T5_THREADS, smem_s));
        if (occ < 1) occ = 1;
        int gx = occ * rsg_num_sms() / nt;     // persistent: exactly the CTAs that are resident at once
        if (gx > mtiles) gx = mtiles;
        if (gx < 1) gx = 1;
        conv3x3_tf32_small_kernel<<<dim3((unsigned)gx, (unsigned)nt, 1), T5_THREADS, smem_s, s>>>(fs, mtiles);
        RSG_LAUNCH_CHECK();
        return RSG_OK;
      }
    }
    const size_t smem = 128 + 2 * (size_t)f.a_bytes + (size_t)T5_STAGES * f.b_stage;
    if (smem <= 220 * 1024) {
      conv3x3_tf32_tc5_kernel<<<dim3((unsigned)mtiles, (unsigned)nt, 1), T5_THREADS, smem, s>>>(f);
      RSG_LAUNCH_CHECK();
      return RSG_OK;
    }
  }
  int n_mma = (Nc + 15) / 16 * 16;
  if (n_mma > 256) n_mma = 256;
  // balance the column tiles when Cout > 256 (e.g. 600 -> 3 x 208)
  const int ntiles = (Nc + n_mma - 1) / n_mma;
  n_mma = ((Nc + ntiles - 1) / ntiles + 15) / 16 * 16;
  p.n_mma = n_mma;
  p.a_plane = (T5_BM + 1) * 16u;
  p.b_plane = (uint32_t)(n_mma + 1) * 16u;
  p.stage_bytes = (T5_KG * (p.a_plane + p.b_plane) + 127u) & ~127u;
  uint32_t cols = 32;
  while (cols < (uint32_t)n_mma) cols <<= 1;
  p.tmem_cols = cols;
  // the ring is walked taps x ceil(Ca / 32) times: a short reduction (the 64 -> 256 1x1 convs of layer1: two iterations of
  // 49 KB) only touches its first stages, and asking for all four cost the second resident CTA per SM
  const long long niter = (long long)p.taps * ((Ca + T5_BK - 1) / T5_BK);
  const size_t smem = 128 + (size_t)(niter < T5_STAGES ? niter : T5_STAGES) * p.stage_bytes;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(gemm_tf32_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.done();
  }
  gemm_tf32_tc5_kernel<<<dim3((unsigned)ceil_div(M, T5_BM), (unsigned)ntiles, (unsigned)batch), T5_THREADS, smem, s>>>(p);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

