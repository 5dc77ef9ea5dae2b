// TRP core (Target-aware Relation Parser, association.py:288-299) on the Blackwell tensor pipe:
//     y[n,i,:] = sum_j sigmoid(x[n,i,:] . x[n,j,:]) * g[n,j,:]
// The warp-level mma.sync version (attention.cu) ran exactly at the legacy-MMA issue ceiling
// (~296 TFLOP/s on B200); both contractions are tcgen05 MMAs here, which leaves the S*S sigmoids
// (one MUFU.TANH each, 16 per clock per SM) as the only bound.  There is no softmax, so key blocks
// accumulate independently and nothing is rescaled.
//
// One CTA = 128 query positions of one crop (the M dimension), walking the crop's keys in blocks of 128:
//   MMA1  S_blk[128 q x 128 keys] = Q . K^T         A = Q, B = K: both K-major planes [C/8][128][8 ch]
//   P = sigmoid(S_blk)  TMEM -> registers (8 warps) -> bf16 -> shared memory as [keys/8][128 q][8 keys]
//   MMA2  O[128 q x C] += P . G_blk                 A = P (K-major over keys), B = G: MN-major -- the
//         same planar [C/8][128 keys][8 ch] box TMA delivers is the canonical MN-major no-swizzle
//         layout (8 channels contiguous, consecutive keys 16 B apart: LBO = 128 B, SBO = one plane)
// S_blk is double-buffered in TMEM and P in shared memory, so MMA1 of block j+1 and the sigmoids of
// block j overlap MMA2 of block j-1.  Keys past the end of the crop arrive as zero rows of G (TMA
// out-of-bounds fill), so they contribute nothing and need no mask.
//
// Warps: 0 = TMA producer, 1 = MMA1 issuer, 2 = MMA2 issuer (one elected lane each), 3..18 = two groups of
// 8 sigmoid / epilogue warps (group = key-block parity, TMEM lane quarter = warp % 4, key-column half).
#include "ops.cuh"
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int AT_BQ = 128, AT_BK = 64;
constexpr int AT_THREADS = 352;                 // 3 + 8 warps; two CTAs per SM cover each other's barrier round trips
constexpr int AT_W0 = 3;                        // first sigmoid warp
constexpr int AT_SW = 8;                        // sigmoid warps
constexpr int AT_KV_STAGES = 3;

struct AttP {
  bf16* y;
  float* y32;                         // non-null: y as dense fp32 [N,S,C] instead (feeds the fp32 W + GroupNorm tail)
  int y_cs, y_co;
  int S, C, nblk;
  uint32_t q_bytes, kv_tile_bytes;    // one [C/8][128][8] tile
  int skip;                           // debug: bit0 no K/G loads after the first ring fill, bit1 no MUFU, bit2 no MMA2
};

struct AttMaps { CUtensorMap q, x, g; };     // q: 128-row boxes of x; x, g: AT_BK-row boxes
constexpr uint32_t AT_TMEM_COLS = 256;          // S buffers at columns 0 / AT_BK, O at 2 * AT_BK

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float sigmoid_mufu(float s) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * s));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(AT_THREADS, 2)
trp_attention_tc5_kernel(const __grid_constant__ AttMaps maps, const AttP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: 0 Q | KV full x3 | KV empty x3 | S full x2 | S empty x2 | P full x2 | P empty x2 | O full
  __shared__ __align__(8) uint64_t bars[1 + 2 * AT_KV_STAGES + 9];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_Q = 0, B_KVF = 1, B_KVE = 1 + AT_KV_STAGES, B_SF = 1 + 2 * AT_KV_STAGES, B_SE = B_SF + 2,
                B_PF = B_SE + 2, B_PE = B_PF + 2, B_OF = B_PE + 2;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  // layout: Q | KV stage s: K tile, G tile | P buffers (2 x 128 x 128 bf16)
  const uint32_t kv0 = p.q_bytes, p0 = kv0 + AT_KV_STAGES * 2u * p.kv_tile_bytes;
  constexpr uint32_t P_BYTES = AT_BQ * AT_BK * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y, q0 = blockIdx.x * AT_BQ;

  if (threadIdx.x == 0) {
    mbar_init(BAR(B_Q), 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) { mbar_init(BAR(B_KVF + i), 1); mbar_init(BAR(B_KVE + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_SF + i), 1); mbar_init(BAR(B_SE + i), AT_SW);
      mbar_init(BAR(B_PF + i), AT_SW); mbar_init(BAR(B_PE + i), 1);
    }
    mbar_init(BAR(B_OF), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_o = tmem_base + 2u * AT_BK;      // S buffers at columns 0 / AT_BK, O behind them
  const int nblk = p.nblk;
  const int c8 = p.C >> 3;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.g) : "memory");
      mbar_arrive_expect_tx(BAR(B_Q), p.q_bytes);
      tma_load_4d(sbase, &maps.q, BAR(B_Q), 0, q0, 0, n);
      uint32_t s = 0, ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(BAR(B_KVE + s), ph ^ 1u);
        const uint32_t dst = sbase + kv0 + s * 2u * p.kv_tile_bytes;
        if ((p.skip & 1) && j >= AT_KV_STAGES) {
          mbar_arrive(BAR(B_KVF + s));
        } else {
          mbar_arrive_expect_tx(BAR(B_KVF + s), 2u * p.kv_tile_bytes);
          tma_load_4d(dst, &maps.x, BAR(B_KVF + s), 0, j * AT_BK, 0, n);
          tma_load_4d(dst + p.kv_tile_bytes, &maps.g, BAR(B_KVF + s), 0, j * AT_BK, 0, n);
        }
        if (++s == AT_KV_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < AT_W0) {
    // ===================== MMA issuers (warp 1: MMA1, warp 2: MMA2) =====================
    // instruction descriptors: f32 accumulate, bf16 x bf16, M = 128; MMA1 N = 128 (both K-major),
    // MMA2 N = C with B MN-major (bit 16)
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BK >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(p.C >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_k = desc_hi(128u);                          // K-major operands: SBO = 128 B (8 rows x 16 B)
    const uint32_t lo_plane = ((2048u >> 4) & 0x3FFFu) << 16;     // Q, P: LBO = one plane of 128 rows = 2048 B
    const uint32_t lo_kplane = (((uint32_t)AT_BK * 16u >> 4) & 0x3FFFu) << 16;   // K: plane of AT_BK keys
    const uint32_t hi_g = desc_hi((uint32_t)AT_BK * 16u);         // G (MN-major): SBO = one channel plane of AT_BK keys
    const uint32_t lo_g = ((128u >> 4) & 0x3FFFu) << 16;          // ... LBO = 8 keys x 16 B
    const int ks1 = p.C >> 4;                                     // k16 steps of MMA1
    const uint32_t q16 = sbase >> 4;
    const uint32_t kv16 = (sbase + kv0) >> 4, kvs16 = (2u * p.kv_tile_bytes) >> 4;
    if (warp == 1) {
      // ---- MMA1: S[j&1] = Q . K_j^T.  Its own warp: the issuing thread's serial instruction stream (waits,
      // descriptor math, ~10 cycles per instruction) is what bounds the block rate, so the two contractions
      // are issued by two threads.
      mbar_wait(BAR(B_Q), 0);
      uint32_t s = 0, phs = 0;
      for (int j = 0; j < nblk; ++j) {
        const uint32_t b = (uint32_t)j & 1u;
        mbar_wait(BAR(B_KVF + s), phs);
        mbar_wait(BAR(B_SE + b), (((uint32_t)j >> 1) & 1u) ^ 1u);   // sigmoid warps drained this S buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t k16a = kv16 + s * kvs16;
        if (elect_one()) {
          for (int k = 0; k < ks1; ++k)
            umma_f16(tmem_base + b * (uint32_t)AT_BK, ((uint64_t)hi_k << 32) | ((q16 + (uint32_t)k * 256u) | lo_plane),
                     ((uint64_t)hi_k << 32) | ((k16a + (uint32_t)k * (2u * AT_BK)) | lo_kplane), idesc1, k > 0);
          umma_commit(BAR(B_SF + b));
        }
        __syncwarp();
        if (++s == AT_KV_STAGES) { s = 0; phs ^= 1u; }
      }
    } else {
      // ---- MMA2: O += P_j . G_j; releases the P buffer and the K/G stage (MMA1 of block j finished long before:
      // P_j was computed from its result)
      const uint32_t g16_0 = kv16 + (p.kv_tile_bytes >> 4);
      const uint32_t p16_0 = (sbase + p0) >> 4;
      const int nk2 = (p.skip & 4) ? 1 : AT_BK / 16;
      uint32_t s = 0;
      for (int j = 0; j < nblk; ++j) {
        const uint32_t b = (uint32_t)j & 1u;
        mbar_wait(BAR(B_PF + b), ((uint32_t)j >> 1) & 1u);           // P_j is in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t p16 = (p16_0 + b * (P_BYTES >> 4)) | lo_plane;
        const uint32_t g16 = (g16_0 + s * kvs16) | lo_g;
        if (elect_one()) {
          if (nk2 == AT_BK / 16) {
#pragma unroll
            for (int k = 0; k < AT_BK / 16; ++k)                     // 16 keys per step: two P planes, two 8-key groups of G
              umma_f16(tmem_o, ((uint64_t)hi_k << 32) | (p16 + (uint32_t)k * 256u), ((uint64_t)hi_g << 32) | (g16 + (uint32_t)k * 16u),
                       idesc2, (j | k) != 0);
          } else {
            umma_f16(tmem_o, ((uint64_t)hi_k << 32) | p16, ((uint64_t)hi_g << 32) | g16, idesc2, j != 0);
          }
          umma_commit(BAR(B_PE + b));                                // P buffer free
          umma_commit(BAR(B_KVE + s));                               // K/G stage free
          if (j == nblk - 1) umma_commit(BAR(B_OF));
        }
        __syncwarp();
        if (++s == AT_KV_STAGES) s = 0;
      }
    }
  } else {
    // ===================== sigmoid warps (3 .. 10) =====================
    const int grp = 0;
    const int q = warp & 3;                            // TMEM lane quarter
    const int half = (warp - AT_W0) >> 2;              // key columns [32 half, 32 half + 32)
    const int row = q * 32 + lane;                     // query row inside the block
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int j = 0; j < nblk; ++j) {
      const uint32_t b = (uint32_t)j & 1u, ph = ((uint32_t)j >> 1) & 1u;
      mbar_wait(BAR(B_SF + b), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[2][16];
      tmem_ld16(tq + b * (uint32_t)AT_BK + (uint32_t)(half * 32), v[0]);
      tmem_ld16(tq + b * (uint32_t)AT_BK + (uint32_t)(half * 32 + 16), v[1]);
      mbar_wait(BAR(B_PE + b), ph ^ 1u);               // MMA2 of block j-2 has consumed this P buffer
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // this warp's reads of S are done: MMA1 of block j+2 may overwrite the buffer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_SE + b));
      unsigned char* const pb = sgen + p0 + b * P_BYTES + (size_t)row * 16;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 32 + cc * 16;
        const uint32_t* vv = v[cc];
        uint4 o0, o1;
        if (p.skip & 2) {
          o0.x = pack_bf16x2(__uint_as_float(vv[0]), __uint_as_float(vv[1])); o0.y = pack_bf16x2(__uint_as_float(vv[2]), __uint_as_float(vv[3]));
          o0.z = pack_bf16x2(__uint_as_float(vv[4]), __uint_as_float(vv[5])); o0.w = pack_bf16x2(__uint_as_float(vv[6]), __uint_as_float(vv[7]));
          o1.x = pack_bf16x2(__uint_as_float(vv[8]), __uint_as_float(vv[9])); o1.y = pack_bf16x2(__uint_as_float(vv[10]), __uint_as_float(vv[11]));
          o1.z = pack_bf16x2(__uint_as_float(vv[12]), __uint_as_float(vv[13])); o1.w = pack_bf16x2(__uint_as_float(vv[14]), __uint_as_float(vv[15]));
        } else {
          o0.x = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[0])), sigmoid_mufu(__uint_as_float(vv[1])));
          o0.y = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[2])), sigmoid_mufu(__uint_as_float(vv[3])));
          o0.z = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[4])), sigmoid_mufu(__uint_as_float(vv[5])));
          o0.w = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[6])), sigmoid_mufu(__uint_as_float(vv[7])));
          o1.x = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[8])), sigmoid_mufu(__uint_as_float(vv[9])));
          o1.y = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[10])), sigmoid_mufu(__uint_as_float(vv[11])));
          o1.z = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[12])), sigmoid_mufu(__uint_as_float(vv[13])));
          o1.w = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[14])), sigmoid_mufu(__uint_as_float(vv[15])));
        }
        // plane (c0/8 + i) holds keys c0+8i .. +7 of all 128 rows: consecutive lanes -> consecutive 16 bytes
        *reinterpret_cast<uint4*>(pb + (size_t)(c0 >> 3) * 2048) = o0;
        *reinterpret_cast<uint4*>(pb + (size_t)((c0 >> 3) + 1) * 2048) = o1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_PF + b));
    }
    // ---- output: O (128 x C fp32 in TMEM) -> bf16 rows of y; the four warps of half 0 do it
    if (half == 0 && grp == 0) {
      mbar_wait(BAR(B_OF), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int grow = q0 + row;
      bf16* dst = p.y + ((size_t)n * p.S + grow) * p.y_cs + p.y_co;
      float* dst32 = p.y32 ? p.y32 + ((size_t)n * p.S + grow) * p.C : nullptr;
      for (int c0 = 0; c0 < p.C; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tq + 2u * AT_BK + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (dst32) {
          // y is a sum of S sigmoid-weighted terms: |mean| >> spread over the positions, and GroupNorm removes the mean
          // right after the 1x1 W conv -- a bf16 y would leave 2^-9 |mean| of rounding noise on a signal of that spread
          if (grow < p.S) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              reinterpret_cast<uint4*>(dst32 + c0)[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
        } else if (grow < p.S) {
          uint4 o0, o1;
          o0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
          o0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
          o0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
          o0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
          o1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
          o1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
          o1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
          o1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
          reinterpret_cast<uint4*>(dst + c0)[0] = o0;
          reinterpret_cast<uint4*>(dst + c0)[1] = o1;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(AT_TMEM_COLS) : "memory");
  }
  (void)c8;
}

// ------------------------------------------------------------------------------------------------------------------
// Two query tiles per CTA (256 query positions, one CTA per SM, 16 sigmoid warps).  The 128-query kernel above is bound
// by the TMA element rate, not by the MUFU: every 64-key block costs one K and one G tile of 16-byte elements (512
// elements = 512 cycles of the SM's TMA engine, tools/tma_rate.cu) for 8192 sigmoids (512 MUFU cycles), and with two
// CTAs per SM the blocks took 820 cycles each (skip modes: 1.40 ms with the MUFU work removed, 1.85 ms with it).  Here a
// K / G tile serves two query tiles: 512 TMA cycles per 16384 sigmoids (1024 MUFU cycles).
// TMEM (512 columns): S[t][b] at (2t + b) * 64, O[t] at 256 + 64t (C columns each).
// ------------------------------------------------------------------------------------------------------------------
constexpr int AT2_THREADS = 608;                // 3 + 16 warps
constexpr int AT2_KV_STAGES = 4;
constexpr int AT2_NB = 3;                       // S / P buffers per query tile: MMA1 runs up to two key blocks ahead of the sigmoids

__global__ void __launch_bounds__(AT2_THREADS, 1)
trp_attention_tc5x2_kernel(const __grid_constant__ AttMaps maps, const AttP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: Q | KVF[4] | KVE[4] | SF[t][b] (2 NB) | SE | PF | PE | OF
  __shared__ __align__(8) uint64_t bars[1 + 2 * AT2_KV_STAGES + 8 * AT2_NB + 1];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_Q = 0, B_KVF = 1, B_KVE = 1 + AT2_KV_STAGES, B_SF = 1 + 2 * AT2_KV_STAGES, B_SE = B_SF + 2 * AT2_NB,
                B_PF = B_SE + 2 * AT2_NB, B_PE = B_PF + 2 * AT2_NB, B_OF = B_PE + 2 * AT2_NB;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  // layout: Q (two 128-row tiles) | KV stage s: K tile, G tile | P buffers [t][b] (2 NB x 128 x 64 bf16)
  const uint32_t kv0 = 2u * p.q_bytes, p0 = kv0 + AT2_KV_STAGES * 2u * p.kv_tile_bytes;
  constexpr uint32_t P_BYTES = AT_BQ * AT_BK * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y, q0 = blockIdx.x * 2 * AT_BQ;

  if (threadIdx.x == 0) {
    mbar_init(BAR(B_Q), 1);
    for (int i = 0; i < AT2_KV_STAGES; ++i) { mbar_init(BAR(B_KVF + i), 1); mbar_init(BAR(B_KVE + i), 1); }
    for (int i = 0; i < 2 * AT2_NB; ++i) {
      mbar_init(BAR(B_SF + i), 1); mbar_init(BAR(B_SE + i), AT_SW);
      mbar_init(BAR(B_PF + i), AT_SW); mbar_init(BAR(B_PE + i), 1);
    }
    mbar_init(BAR(B_OF), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_o = tmem_base + (uint32_t)(2 * AT2_NB) * AT_BK;        // O[t] at tmem_o + 64 t
  const int nblk = p.nblk;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.g) : "memory");
      mbar_arrive_expect_tx(BAR(B_Q), 2u * p.q_bytes);
      tma_load_4d(sbase, &maps.q, BAR(B_Q), 0, q0, 0, n);
      tma_load_4d(sbase + p.q_bytes, &maps.q, BAR(B_Q), 0, q0 + AT_BQ, 0, n);
      uint32_t s = 0, ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(BAR(B_KVE + s), ph ^ 1u);
        const uint32_t dst = sbase + kv0 + s * 2u * p.kv_tile_bytes;
        mbar_arrive_expect_tx(BAR(B_KVF + s), 2u * p.kv_tile_bytes);
        tma_load_4d(dst, &maps.x, BAR(B_KVF + s), 0, j * AT_BK, 0, n);
        tma_load_4d(dst + p.kv_tile_bytes, &maps.g, BAR(B_KVF + s), 0, j * AT_BK, 0, n);
        if (++s == AT2_KV_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < AT_W0) {
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BK >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(p.C >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi_k = desc_hi(128u);
    const uint32_t lo_plane = ((2048u >> 4) & 0x3FFFu) << 16;
    const uint32_t lo_kplane = (((uint32_t)AT_BK * 16u >> 4) & 0x3FFFu) << 16;
    const uint32_t hi_g = desc_hi((uint32_t)AT_BK * 16u);
    const uint32_t lo_g = ((128u >> 4) & 0x3FFFu) << 16;
    const int ks1 = p.C >> 4;
    const uint32_t q16 = sbase >> 4, qt16 = p.q_bytes >> 4;
    const uint32_t kv16 = (sbase + kv0) >> 4, kvs16 = (2u * p.kv_tile_bytes) >> 4;
    if (warp == 1) {
      // ---- MMA1: S[t][j&1] = Q_t . K_j^T for both query tiles
      mbar_wait(BAR(B_Q), 0);
      uint32_t s = 0, phs = 0, b = 0, pj = 0;        // b = j % NB, pj = (j / NB) & 1, carried incrementally
      for (int j = 0; j < nblk; ++j, b = (b + 1 == AT2_NB ? 0u : b + 1), pj ^= (b == 0)) {
        mbar_wait(BAR(B_KVF + s), phs);
        const uint32_t k16a = kv16 + s * kvs16;
#pragma unroll
        for (uint32_t t = 0; t < 2; ++t) {
          mbar_wait(BAR(B_SE + AT2_NB * t + b), pj ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            for (int k = 0; k < ks1; ++k)
              umma_f16(tmem_base + ((uint32_t)AT2_NB * t + b) * (uint32_t)AT_BK, ((uint64_t)hi_k << 32) | ((q16 + t * qt16 + (uint32_t)k * 256u) | lo_plane),
                       ((uint64_t)hi_k << 32) | ((k16a + (uint32_t)k * (2u * AT_BK)) | lo_kplane), idesc1, k > 0);
            umma_commit(BAR(B_SF + AT2_NB * t + b));
          }
          __syncwarp();
        }
        if (++s == AT2_KV_STAGES) { s = 0; phs ^= 1u; }
      }
    } else {
      // ---- MMA2: O[t] += P[t][b] . G_j; releases the P buffers and, after both tiles, the K/G stage
      const uint32_t g16_0 = kv16 + (p.kv_tile_bytes >> 4);
      const uint32_t p16_0 = (sbase + p0) >> 4;
      uint32_t s = 0, b = 0, pj = 0;
      for (int j = 0; j < nblk; ++j, b = (b + 1 == AT2_NB ? 0u : b + 1), pj ^= (b == 0)) {
        const uint32_t g16 = (g16_0 + s * kvs16) | lo_g;
#pragma unroll
        for (uint32_t t = 0; t < 2; ++t) {
          mbar_wait(BAR(B_PF + AT2_NB * t + b), pj);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t p16 = (p16_0 + ((uint32_t)AT2_NB * t + b) * (P_BYTES >> 4)) | lo_plane;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < AT_BK / 16; ++k)
              umma_f16(tmem_o + 64u * t, ((uint64_t)hi_k << 32) | (p16 + (uint32_t)k * 256u), ((uint64_t)hi_g << 32) | (g16 + (uint32_t)k * 16u),
                       idesc2, (j | k) != 0);
            umma_commit(BAR(B_PE + AT2_NB * t + b));
            if (t == 1) {
              umma_commit(BAR(B_KVE + s));
              if (j == nblk - 1) umma_commit(BAR(B_OF));
            }
          }
          __syncwarp();
        }
        if (++s == AT2_KV_STAGES) s = 0;
      }
    }
  } else {
    // ===================== sigmoid warps (3 .. 18): tile t = (warp - 3) / 8 =====================
    const int t = (warp - AT_W0) >> 3;
    const int q = warp & 3;                            // TMEM lane quarter
    const int half = ((warp - AT_W0) >> 2) & 1;        // key columns [32 half, 32 half + 32)
    const int row = q * 32 + lane;                     // query row inside the tile
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    // (software-pipelining the TMEM load of block j+1 under the sigmoids of block j was measured slower: 1662 vs 1594 us)
    uint32_t b = 0, ph = 0;
    for (int j = 0; j < nblk; ++j, b = (b + 1 == AT2_NB ? 0u : b + 1), ph ^= (b == 0)) {
      const uint32_t sb = (uint32_t)AT2_NB * (uint32_t)t + b;
      mbar_wait(BAR(B_SF + sb), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[2][16];
      tmem_ld16(tq + sb * (uint32_t)AT_BK + (uint32_t)(half * 32), v[0]);
      tmem_ld16(tq + sb * (uint32_t)AT_BK + (uint32_t)(half * 32 + 16), v[1]);
      mbar_wait(BAR(B_PE + sb), ph ^ 1u);              // MMA2 of block j-NB has consumed this P buffer
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_SE + sb));
      unsigned char* const pb = sgen + p0 + sb * P_BYTES + (size_t)row * 16;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 32 + cc * 16;
        const uint32_t* vv = v[cc];
        uint4 o0, o1;
        o0.x = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[0])), sigmoid_mufu(__uint_as_float(vv[1])));
        o0.y = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[2])), sigmoid_mufu(__uint_as_float(vv[3])));
        o0.z = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[4])), sigmoid_mufu(__uint_as_float(vv[5])));
        o0.w = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[6])), sigmoid_mufu(__uint_as_float(vv[7])));
        o1.x = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[8])), sigmoid_mufu(__uint_as_float(vv[9])));
        o1.y = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[10])), sigmoid_mufu(__uint_as_float(vv[11])));
        o1.z = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[12])), sigmoid_mufu(__uint_as_float(vv[13])));
        o1.w = pack_bf16x2(sigmoid_mufu(__uint_as_float(vv[14])), sigmoid_mufu(__uint_as_float(vv[15])));
        *reinterpret_cast<uint4*>(pb + (size_t)(c0 >> 3) * 2048) = o0;
        *reinterpret_cast<uint4*>(pb + (size_t)((c0 >> 3) + 1) * 2048) = o1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_PF + sb));
    }
    // ---- output: O[t] (128 x C fp32 in TMEM) -> y rows; the four warps of half 0 of each tile
    if (half == 0) {
      mbar_wait(BAR(B_OF), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int grow = q0 + t * AT_BQ + row;
      bf16* dst = p.y + ((size_t)n * p.S + grow) * p.y_cs + p.y_co;
      float* dst32 = p.y32 ? p.y32 + ((size_t)n * p.S + grow) * p.C : nullptr;
      for (int c0 = 0; c0 < p.C; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tq + (uint32_t)(2 * AT2_NB) * AT_BK + 64u * (uint32_t)t + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (grow >= p.S) continue;
        if (dst32) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            reinterpret_cast<uint4*>(dst32 + c0)[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
          uint4 o0, o1;
          o0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
          o0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
          o0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
          o0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
          o1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
          o1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
          o1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
          o1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
          reinterpret_cast<uint4*>(dst + c0)[0] = o0;
          reinterpret_cast<uint4*>(dst + c0)[1] = o1;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// [N, S, C] bf16 view (channel stride cs, offset co) as (8 ch, S, C/8, N); box = (8, rows, C/8, 1)
int make_att_map(const bf16* base, int cs, int co, int N, int S, int C, int rows, CUtensorMap* m) {
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {8, (cuuint64_t)S, (cuuint64_t)(C / 8), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cs * 2, 16, (cuuint64_t)S * cs * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)rows, (cuuint32_t)(C / 8), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(base + co), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with %d", (int)r);
  return RSG_OK;
}

}  // namespace

int attention_tc5_launch(cudaStream_t s, const bf16* x, int x_cs, int x_co, const bf16* g, int g_cs, int g_co,
                         bf16* y, int y_cs, int y_co, float* y32, int N, int S, int C, int* handled) {
  *handled = 0;
  static const bool legacy = rsg_dbg_env("RSG_ATT_LEGACY") != nullptr;      // A/B switch: the mma.sync kernel
  if (legacy) return RSG_OK;
  if (C % 16 != 0 || C < 16 || C > 64 || N > 65535) return RSG_OK;
  if (x_cs % 8 != 0 || x_co % 8 != 0 || g_cs % 8 != 0 || g_co % 8 != 0 || y_cs % 8 != 0 || y_co % 8 != 0) return RSG_OK;
  if (((uintptr_t)x | (uintptr_t)g | (uintptr_t)y | (uintptr_t)y32) % 16 != 0) return RSG_OK;
  if (!y && !y32) return RSG_OK;
  *handled = 1;
  if (N == 0 || S == 0) return RSG_OK;
  AttP p;
  memset(&p, 0, sizeof(p));
  p.y = y; p.y32 = y32; p.y_cs = y_cs; p.y_co = y_co; p.S = S; p.C = C;
  p.nblk = (S + AT_BK - 1) / AT_BK;
  p.q_bytes = (uint32_t)AT_BQ * C * 2;
  p.kv_tile_bytes = (uint32_t)AT_BK * C * 2;
  { const char* e = rsg_dbg_env("RSG_ATT_SKIP"); p.skip = e ? atoi(e) : 0; }
  AttMaps maps;
  memset(&maps, 0, sizeof(maps));
  { int rc = make_att_map(x, x_cs, x_co, N, S, C, AT_BQ, &maps.q); if (rc) return rc; }
  { int rc = make_att_map(x, x_cs, x_co, N, S, C, AT_BK, &maps.x); if (rc) return rc; }
  { int rc = make_att_map(g, g_cs, g_co, N, S, C, AT_BK, &maps.g); if (rc) return rc; }
  const size_t smem = 128 + p.q_bytes + AT_KV_STAGES * 2u * p.kv_tile_bytes + 2u * AT_BQ * AT_BK * 2u;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(trp_attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.done();
  }
  static const bool one_tile = rsg_dbg_env("RSG_ATT_1TILE") != nullptr;
  if (S > AT_BQ && !one_tile) {
    // two query tiles per CTA: a K / G tile is loaded once per 256 query positions
    const size_t smem2 = 128 + 2u * p.q_bytes + AT2_KV_STAGES * 2u * p.kv_tile_bytes + (size_t)(2 * AT2_NB) * AT_BQ * AT_BK * 2u;
    static DeviceOnce attr2_once;
    if (attr2_once.first()) {
      RSG_CUDA(cudaFuncSetAttribute(trp_attention_tc5x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr2_once.done();
    }
    dim3 grid2((S + 2 * AT_BQ - 1) / (2 * AT_BQ), N);
    trp_attention_tc5x2_kernel<<<grid2, AT2_THREADS, smem2, s>>>(maps, p);
    RSG_LAUNCH_CHECK();
    return RSG_OK;
  }
  dim3 grid((S + AT_BQ - 1) / AT_BQ, N);
  trp_attention_tc5_kernel<<<grid, AT_THREADS, smem, s>>>(maps, p);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
