// Fused HRNet Bottleneck (pose_rsgnet.py:57-95; layer1 = 4 of them, :619):
//   out = relu(bn3(conv3_1x1(relu(bn2(conv2_3x3(relu(bn1(conv1_1x1(x)))))))) + residual)
// with 64 planes and 256 output channels, BN folded, in ONE kernel.  Unfused, a layer1 block moves 3.4 GB per 512
// forwards (the 256-channel map is read by conv1, read again as the residual and written by conv3; the two 64-channel
// intermediates are written and read) and its two 1x1 convs are bound by the epilogue's global traffic and by the TMA
// element rate (profiles/r1_notes.md §7-§8).  Fused, the block reads x once through TMA (+ the residual, an L2 hit when
// it is x itself) and writes the output once; both intermediates live in shared memory.
//
// A CTA owns 16 x 8 output pixels.  Everything is a flat pixel array of pitch 10 (the 18 x 10 halo patch, 180 pixels):
//   input       K-chunks of 64 channels, ONE TMA box each, S-deep ring.  Default ("wide"): a 4-D box of 128-byte
//               pixel rows, [180 px][64 ch] SWIZZLE_128B = the K-major swizzled UMMA layout (conv1 is a 1x1 conv: no
//               tap shifts, so the swizzle atoms stay aligned); RSG_BNECK_PLANAR=1: the conv kernels' planar 5-D box
//               [8 planes][180 px][8 ch] (16-byte TMA elements: 453 instead of 425 us).  Out-of-image = zero fill.
//   conv1 (1x1) on ALL 180 patch pixels (conv2 needs the halo): 2 M tiles of 128 consecutive flat pixels,
//               N = 64, K = Cin accumulated over the chunks                               -> TMEM acc1 [2][64 cols]
//   epilogue 1  +bias1, ReLU, ZERO outside the image (conv2's padding) -> bf16 -> mid1 [8][180 px][8 ch]
//   conv2 (3x3) 128 output pixels = 16 rows of 8 (SBO = one pitch), tap (dy,dx) = start offset (1+dy)*10 + (1+dx)
//                                                                                           -> TMEM acc2 [64 cols]
//   epilogue 2  +bias2, ReLU -> bf16 -> mid2 [8][128 px][8 ch]
//   conv3 (1x1) N = 256, K = 64                                                            -> TMEM acc3 [256 cols]
//   epilogue 3  +bias3 + residual (global, prefetched one iteration early) -> ReLU -> bf16 NHWC stores, one full
//               128-byte line per lane quad (16x256b TMEM loads + channel-permuted W3)
// The residual is a plain NHWC tensor: x itself for the identity blocks (ncu: an L2 hit, DRAM read 903 MB for the
// 805 MB map), the output of the separate 1x1 downsample conv for the first block of layer1 (Cin = 64).
//
// Every buffer exists once (weights 136 KB + ring + mid1 + mid2 fill the 227 KB; TMEM 448 of 512 columns), and the
// three stages of DIFFERENT tiles overlap: while tile i is in conv1, tile i-1 is in conv2 and tile i-2 in conv3 /
// epilogue 3.  Measured (512 forwards, 64x48): 425 us per identity block against 811 us for the three convs; the
// clock64 timeline (RSG_BNECK_TIMELINE=1) shows a ~9000-cycle tile period made of the TMA round trip of the two
// stages (~4500) and the epilogue warps' global stores / residual loads (~3000) that sit between two conv1 hand-offs;
// DRAM runs at 3.9 TB/s (47 %), the tensor pipe at 42 %.
//
// Warps (640 threads, one persistent CTA per SM): 0 = conv1 issuer (+ TMEM, weights), 1 = conv2 issuer, 2 = TMA
// producer, 3 = conv3 issuer, 4..19 = epilogue warps (all three epilogues, software-pipelined by one / two tiles).
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int BN_THREADS = 640;
constexpr int BN_PITCH = 10, BN_PX = 180;            // 18 x 10 patch
constexpr uint32_t BN_STAGE = 8u * BN_PX * 16u;      // 23040 B: 64 channels of the patch (also mid1)
constexpr uint32_t BN_MID2 = 8u * 128u * 16u;        // 16384 B
constexpr uint32_t BN_W2 = 9u * 64u * 64u * 2u;      // 73728 B
constexpr uint32_t BN_W3 = 64u * 256u * 2u;          // 32768 B
constexpr int BN_MAX_S = 4;
constexpr uint32_t TM_ACC2 = 128u, TM_ACC3 = 256u;   // TMEM columns: acc1 = [0,128), acc2 = [128,192), acc3 = [256,512)

struct BnP {
  const bf16 *w1, *w2, *w3;            // [Cin/8][64][8], [9][8][64][8], [8][256][8]  (tcgen05 packing, NS = Cout of each conv;
                                       // w3's rows in accumulator-column order, see epilogue 3)
  const float *b1, *b2, *b3;
  int Cin, H, W, N, nchunks, S;
  int tiles_x, tiles_y;
  uint32_t tiles_per_img, magic_tpi, magic_tx;
  int ntiles;
  const bf16* res; int res_cs, res_co;
  bf16* out; int out_cs, out_co;
  uint32_t w1_bytes, stage_bytes;
  int wide;                            // 1: input chunks as [184 px][64 ch] rows of 128 B (SWIZZLE_128B), 0: planar [8][180 px][8 ch]
  int v32;                             // bit0: output rows 32-byte aligned, bit1: residual rows too
  long long* dbg;                      // debug timeline (CTA 0): [tile][16] clock64 stamps, or nullptr
  int skip;                            // debug: bit0 no stores, bit1 no residual loads, bit2 conv2 one tap, bit4 no TMA, bit5 L2 prefetch of the next tile
};

#define BN_STAMP(tile, slot) do { if (p.dbg && blockIdx.x == 0 && (tile) < 32u) p.dbg[(tile) * 16u + (slot)] = clock64(); } while (0)

// bias + ReLU + bf16 pack of 8 accumulator values
__device__ __forceinline__ uint4 bias_relu_pack8(const uint32_t* v, const float* b) {
  const float4 b0 = reinterpret_cast<const float4*>(b)[0], b1 = reinterpret_cast<const float4*>(b)[1];
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
  h[0] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[0]) + b0.x, 0.f), fmaxf(__uint_as_float(v[1]) + b0.y, 0.f));
  h[1] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[3]) + b0.w, 0.f));
  h[2] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[5]) + b1.y, 0.f));
  h[3] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[7]) + b1.w, 0.f));
  return o;
}

__global__ void __launch_bounds__(BN_THREADS, 1)
conv_bneck_kernel(const __grid_constant__ CUtensorMap in_map, const BnP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: 0 weights | full[S] | empty[S] | acc1 full, acc1 empty | mid1 full, mid1 empty | acc2 full, acc2 empty |
  //           mid2 full, mid2 empty | acc3 full, acc3 empty
  __shared__ __align__(8) uint64_t bars[1 + 2 * BN_MAX_S + 10];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sB1[64], sB2[64], sB3[256];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 1, B_EMPTY = 1 + p.S, B_A1F = 1 + 2 * p.S, B_A1E = B_A1F + 1, B_M1F = B_A1F + 2, B_M1E = B_A1F + 3,
            B_A2F = B_A1F + 4, B_A2E = B_A1F + 5, B_M2F = B_A1F + 6, B_M2E = B_A1F + 7, B_A3F = B_A1F + 8, B_A3E = B_A1F + 9;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B stages: 1024-byte aligned
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  // layout: W1 | W2 | W3 | ring | mid1 | mid2.  conv1's second M tile reads 76 pixels past the end of each plane (rows
  // that are never used); for the last plane of the last stage that lands in mid1, inside the allocation.
  const uint32_t off_w2 = p.w1_bytes, off_w3 = off_w2 + BN_W2, off_st = off_w3 + BN_W3,
                 off_m1 = off_st + (uint32_t)p.S * p.stage_bytes, off_m2 = off_m1 + BN_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1);
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
    mbar_init(BAR(B_A1F), 1); mbar_init(BAR(B_A1E), 16);
    mbar_init(BAR(B_M1F), 16); mbar_init(BAR(B_M1E), 1);
    mbar_init(BAR(B_A2F), 1); mbar_init(BAR(B_A2E), 16);
    mbar_init(BAR(B_M2F), 16); mbar_init(BAR(B_M2E), 1);
    mbar_init(BAR(B_A3F), 1); mbar_init(BAR(B_A3E), 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { sB1[i] = p.b1[i]; sB2[i] = p.b2[i]; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sB3[i] = p.b3[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int first = blockIdx.x, step = gridDim.x;
  const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

  if (warp == 2) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
      uint32_t s = 0, ph = 0;
      // Only S = 2 chunk stages fit beside the weights.  Prefetching the next tile's boxes into L2
      // (cp.async.bulk.prefetch.tensor) was measured: the TMA wait per tile drops from ~5000 to ~3500 cycles but the
      // kernel does not get faster (438 vs 428 us) -- the epilogue warps' global stores are the longer chain
      // (profiles/r1_notes.md §12); RSG_BNECK_PREFETCH=1 (skip bit 5) re-enables it for experiments.
      auto coords = [&](int t, int& n, int& ty, int& tx) {
        n = (int)fastdiv((uint32_t)t, p.magic_tpi);
        const int rem = t - n * (int)p.tiles_per_img;
        ty = (int)fastdiv((uint32_t)rem, p.magic_tx);
        tx = rem - ty * p.tiles_x;
      };
      const bool pf = (p.skip & 32) && !(p.skip & 16);
      for (int t = first; t < p.ntiles; t += step) {
        int n, ty, tx, n2 = 0, ty2 = 0, tx2 = 0;
        coords(t, n, ty, tx);
        const bool pf2 = pf && t + step < p.ntiles;
        if (pf2) coords(t + step, n2, ty2, tx2);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(BAR(B_EMPTY + s), ph ^ 1u);
          if (p.skip & 16) {
            mbar_arrive(BAR(B_FULL + s));
          } else {
            mbar_arrive_expect_tx(BAR(B_FULL + s), BN_STAGE);
            if (p.wide) tma_load_4d(sbase + off_st + s * p.stage_bytes, &in_map, BAR(B_FULL + s), c * 64, tx * 8 - 1, ty * 16 - 1, n);
            else tma_load_5d(sbase + off_st + s * p.stage_bytes, &in_map, BAR(B_FULL + s), 0, tx * 8 - 1, ty * 16 - 1, c * 8, n);
          }
          if (pf2) {
            if (p.wide) tma_prefetch_4d(&in_map, c * 64, tx2 * 8 - 1, ty2 * 16 - 1, n2);
            else tma_prefetch_5d(&in_map, 0, tx2 * 8 - 1, ty2 * 16 - 1, c * 8, n2);
          }
          if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < 4) {
    // ===================== MMA issuers =====================
    // warp-uniform control flow, only the tcgen05 instructions are predicated on one elected lane (UTCHMMA takes its
    // descriptors from uniform registers; profiles/r1_notes.md §2)
    if (warp == 0 && elect_one()) {
      const uint32_t total = p.w1_bytes + BN_W2 + BN_W3;
      mbar_arrive_expect_tx(BAR(0), total);
      const unsigned char* src[3] = {reinterpret_cast<const unsigned char*>(p.w1), reinterpret_cast<const unsigned char*>(p.w2),
                                     reinterpret_cast<const unsigned char*>(p.w3)};
      const uint32_t dst[3] = {0u, off_w2, off_w3}, len[3] = {p.w1_bytes, BN_W2, BN_W3};
#pragma unroll
      for (int k = 0; k < 3; ++k)
        for (uint32_t off = 0; off < len[k]; off += 32768u) {
          const uint32_t nb = len[k] - off < 32768u ? len[k] - off : 32768u;
          bulk_load(sbase + dst[k] + off, src[k] + off, nb, BAR(0));
        }
    }
    __syncwarp();
    mbar_wait(BAR(0), 0);
    if (warp == 0) {
      // ---- conv1: 1x1, Cin -> 64, on the 180 patch pixels (2 M tiles)
      // A operand: planar = K-major SWIZZLE_NONE (8-pixel core matrices 128 B apart, planes LBO apart; M tile = +128 px,
      // k16 = +2 planes); wide = K-major SWIZZLE_128B (pixel rows of 128 B; M tile = +128 rows, k16 = +32 B)
      const uint32_t hiA = p.wide ? desc_hi_sw128(1024u) : desc_hi(128u), hiB = desc_hi(128u);
      const uint32_t a_lbo = p.wide ? (1u << 16) : ((uint32_t)BN_PX << 16);
      const uint32_t a_mt = p.wide ? 1024u : 128u, a_kk = p.wide ? 2u : 2u * BN_PX;          // 16-byte units
      const uint32_t w16 = (sbase >> 4) | (64u << 16);                     // LBO = 64 rows x 16 B
      uint32_t s = 0, ph = 0, i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        mbar_wait(BAR(B_A1E), (i & 1u) ^ 1u);
        if (lane == 0) BN_STAMP(i, 0);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(BAR(B_FULL + s), ph);
          if (lane == 0 && c == p.nchunks - 1) BN_STAMP(i, 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a16 = ((sbase + off_st + s * p.stage_bytes) >> 4) | a_lbo;
          const uint32_t b16 = w16 + (uint32_t)c * 512u;                   // 8 planes x 1024 B per chunk
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem_base + (uint32_t)mt * 64u, ((uint64_t)hiA << 32) | (a16 + (uint32_t)mt * a_mt + (uint32_t)kk * a_kk),
                         ((uint64_t)hiB << 32) | (b16 + (uint32_t)kk * 128u), idesc64, (c | kk) ? 1u : 0u);
            }
            umma_commit(BAR(B_EMPTY + s));
            if (c == p.nchunks - 1) umma_commit(BAR(B_A1F));
          }
          __syncwarp();
          if (lane == 0 && c == p.nchunks - 1) BN_STAMP(i, 2);
          if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ---- conv2: 3x3, 64 -> 64, on mid1
      const uint32_t hiA = desc_hi((uint32_t)BN_PITCH * 16u), hiB = desc_hi(128u);   // 8-pixel groups one pitch apart
      const uint32_t w16 = ((sbase + off_w2) >> 4) | (64u << 16);
      const uint32_t a16 = ((sbase + off_m1) >> 4) | ((uint32_t)BN_PX << 16);
      uint32_t toff[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) toff[tp] = (uint32_t)((tp / 3) * BN_PITCH + tp % 3);
      uint32_t i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        mbar_wait(BAR(B_M1F), i & 1u);
        mbar_wait(BAR(B_A2E), (i & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) BN_STAMP(i, 5);
        if (elect_one()) {
          if (p.skip & 4) {
            umma_f16(tmem_base + TM_ACC2, ((uint64_t)hiA << 32) | (a16 + toff[4]), ((uint64_t)hiB << 32) | (w16 + 4u * 512u), idesc64, 0u);
          } else {
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(tmem_base + TM_ACC2, ((uint64_t)hiA << 32) | (a16 + toff[tp] + (uint32_t)kk * (2u * BN_PX)),
                         ((uint64_t)hiB << 32) | (w16 + (uint32_t)tp * 512u + (uint32_t)kk * 128u), idesc64, (tp | kk) ? 1u : 0u);
            }
          }
          umma_commit(BAR(B_M1E));
          umma_commit(BAR(B_A2F));
        }
        __syncwarp();
        if (lane == 0) BN_STAMP(i, 6);
      }
    } else {
      // ---- conv3: 1x1, 64 -> 256, on mid2
      const uint32_t hiA = desc_hi(128u), hiB = desc_hi(128u);
      const uint32_t w16 = ((sbase + off_w3) >> 4) | (256u << 16);         // LBO = 256 rows x 16 B
      const uint32_t a16 = ((sbase + off_m2) >> 4) | (128u << 16);         // LBO = 128 px x 16 B
      uint32_t i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        mbar_wait(BAR(B_M2F), i & 1u);
        mbar_wait(BAR(B_A3E), (i & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) BN_STAMP(i, 9);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base + TM_ACC3, ((uint64_t)hiA << 32) | (a16 + (uint32_t)kk * 256u),
                     ((uint64_t)hiB << 32) | (w16 + (uint32_t)kk * 512u), idesc256, kk ? 1u : 0u);
          umma_commit(BAR(B_M2E));
          umma_commit(BAR(B_A3F));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps 4 .. 19 =====================
    const int ew = warp - 4, q = warp & 3, w4 = ew >> 2;       // TMEM lane quarter (hardware rule: warp % 4), piece index
    const int m = q * 32 + lane, hy = m >> 3, wx = m & 7;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    pdl_wait();

    auto tile_coords = [&](int t, int& n, int& ty, int& tx) {
      n = (int)fastdiv((uint32_t)t, p.magic_tpi);
      const int rem = t - n * (int)p.tiles_per_img;
      ty = (int)fastdiv((uint32_t)rem, p.magic_tx);
      tx = rem - ty * p.tiles_x;
    };

    // ---- epilogue 1: piece = (M tile, 32-column half) of conv1's accumulators -> mid1
    auto epilogue1 = [&](uint32_t i, int t) {
      const int mt = w4 >> 1, cb = (w4 & 1) * 32;
      const int o = mt * 128 + m;                              // flat position in the pitch-10 patch
      const int r = o / BN_PITCH, c = o - r * BN_PITCH;
      int n, ty, tx;
      tile_coords(t, n, ty, tx);
      const int gy = ty * 16 - 1 + r, gx = tx * 8 - 1 + c;
      const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      mbar_wait(BAR(B_A1F), i & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (threadIdx.x == 128) BN_STAMP(i, 3);
      uint32_t v[32];
      tmem_ld32(tq + (uint32_t)(mt * 64 + cb), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_A1E));
      mbar_wait(BAR(B_M1E), (i & 1u) ^ 1u);                    // conv2 of the previous tile has consumed mid1
      if (o < BN_PX) {
        unsigned char* dst = sgen + off_m1 + (uint32_t)o * 16u + (uint32_t)(cb >> 3) * (BN_PX * 16u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 ov = make_uint4(0, 0, 0, 0);
          if (inside) ov = bias_relu_pack8(v + 8 * k, sB1 + cb + 8 * k);
          *reinterpret_cast<uint4*>(dst + (uint32_t)k * (BN_PX * 16u)) = ov;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_M1F));
      if (threadIdx.x == 128) BN_STAMP(i, 4);
    };

    // ---- epilogue 2: 16 columns of conv2's accumulator -> mid2
    auto epilogue2 = [&](uint32_t j) {
      const int cb = w4 * 16;
      mbar_wait(BAR(B_A2F), j & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (threadIdx.x == 128) BN_STAMP(j, 7);
      uint32_t v[16];
      tmem_ld16(tq + TM_ACC2 + (uint32_t)cb, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_A2E));
      mbar_wait(BAR(B_M2E), (j & 1u) ^ 1u);                    // conv3 of the previous tile has consumed mid2
      unsigned char* dst = sgen + off_m2 + (uint32_t)m * 16u + (uint32_t)(cb >> 3) * 2048u;
      *reinterpret_cast<uint4*>(dst) = bias_relu_pack8(v, sB2 + cb);
      *reinterpret_cast<uint4*>(dst + 2048u) = bias_relu_pack8(v + 8, sB2 + cb + 8);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_M2F));
      if (threadIdx.x == 128) BN_STAMP(j, 8);
    };

    // ---- epilogue 3: 64 columns of conv3's accumulator + residual -> global.
    // A row-per-lane epilogue (32x32b TMEM loads) touches 32 different 128-byte lines with every 32-byte-per-lane store
    // instruction and is bound by the LSU at ~2 cycles per line (tools/stg_rate.cu: 4.7 TB/s, 4090 cycles per tile).
    // Here the accumulator is read as 16x256b blocks: lanes 4r..4r+3 hold pixel r (and r + 8) of the block, and -- with
    // W3's output channels permuted at pack time (column 8g + 2j + e = channel 16j + 2g + e inside each group of 64) --
    // lane 4r+j owns channels [16j, 16j+16): one store instruction writes 8 full lines (6.3 TB/s in the same probe).
    // The residual loads use the same mapping and are issued one whole iteration early (before epilogues 1 and 2 of
    // the younger tiles), so their latency is never waited for.
    const int cb3 = w4 * 64 + 16 * (lane & 3);               // this lane's 16 channels
    const int wx3 = lane >> 2;
    auto prefetch3 = [&](int t, uint4 (&pre)[8]) {
      int n, ty, tx;
      tile_coords(t, n, ty, tx);
      const int y0 = ty * 16 + q * 4, x = tx * 8 + wx3;      // the lane's four pixels: rows y0 .. y0+3, column x
      if (x < p.W && !(p.skip & 2)) {
        const bf16* rp = p.res + p.res_co + cb3 + (((size_t)n * p.H + (size_t)y0) * p.W + (size_t)x) * (size_t)p.res_cs;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (y0 + k < p.H) ldg32(rp + (size_t)k * p.W * p.res_cs, (p.v32 & 2) != 0, pre[2 * k], pre[2 * k + 1]);
      }
    };
    auto epilogue3 = [&](uint32_t j, int t, const uint4 (&pre)[8]) {
      int n, ty, tx;
      tile_coords(t, n, ty, tx);
      const int y0 = ty * 16 + q * 4, x = tx * 8 + wx3;
      mbar_wait(BAR(B_A3F), j & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (threadIdx.x == 128) BN_STAMP(j, 10);
      bf16* const outp = p.out + p.out_co + cb3 + (((size_t)n * p.H + (size_t)y0) * p.W + (size_t)x) * (size_t)p.out_cs;
      const float4* bp = reinterpret_cast<const float4*>(sB3 + cb3);
#pragma unroll
      for (int b16 = 0; b16 < 2; ++b16) {
        uint32_t v[32];
        tmem_ld_16x256b_x8(tmem_base + ((uint32_t)(q * 32 + b16 * 16) << 16) + TM_ACC3 + (uint32_t)(w4 * 64), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (b16 == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_A3E));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = b16 * 2 + h;                           // pixel row y0 + k (MMA row q*32 + 8k + wx3)
          float f[16];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {                     // channels 4 g4 .. 4 g4 + 3 = columns (g, e) = (2 g4, 0..1), (2 g4 + 1, 0..1)
            const float4 bb = bp[g4];
            f[4 * g4 + 0] = __uint_as_float(v[4 * (2 * g4) + 2 * h + 0]) + bb.x;
            f[4 * g4 + 1] = __uint_as_float(v[4 * (2 * g4) + 2 * h + 1]) + bb.y;
            f[4 * g4 + 2] = __uint_as_float(v[4 * (2 * g4 + 1) + 2 * h + 0]) + bb.z;
            f[4 * g4 + 3] = __uint_as_float(v[4 * (2 * g4 + 1) + 2 * h + 1]) + bb.w;
          }
          add_res8(f, pre[2 * k]);
          add_res8(f + 8, pre[2 * k + 1]);
          uint4 o0, o1;
          __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
          __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h0[e] = __floats2bfloat162_rn(fmaxf(f[2 * e], 0.f), fmaxf(f[2 * e + 1], 0.f));
            h1[e] = __floats2bfloat162_rn(fmaxf(f[8 + 2 * e], 0.f), fmaxf(f[8 + 2 * e + 1], 0.f));
          }
          if (y0 + k < p.H && x < p.W && !(p.skip & 1))
            stg32(outp + (size_t)k * p.W * p.out_cs, (p.v32 & 1) != 0, o0, o1);
        }
      }
      if (threadIdx.x == 128) BN_STAMP(j, 11);
    };

    // tile i is in epilogue 1 while tile i-1 is in epilogue 2 and tile i-2 in epilogue 3: conv2 / conv3 of those tiles
    // had a whole iteration to complete, so only conv1 (TMA-fed) is waited for
    const int nloc = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;
    for (int i = 0; i < nloc + 2; ++i) {
      uint4 pre[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pre[k] = make_uint4(0, 0, 0, 0);
      if (i >= 2) prefetch3(first + (i - 2) * step, pre);
      if (i < nloc) epilogue1((uint32_t)i, first + i * step);
      if (i >= 1 && i - 1 < nloc) epilogue2((uint32_t)(i - 1));
      if (i >= 2) epilogue3((uint32_t)(i - 2), first + (i - 2) * step, pre);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// 1 when the fused kernel covers a Bottleneck (Cin -> planes -> planes -> Cout) on H x W maps.
extern "C" int rsg_bottleneck_supported(int Cin, int planes, int Cout, int H, int W) {
  if (rsg_dbg_env("RSG_DISABLE_BNECK")) return 0;
  return planes == 64 && Cout == 256 && Cin >= 64 && Cin <= 256 && Cin % 64 == 0 && H >= 16 && W >= 8 ? 1 : 0;
}

int conv_bneck_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, int N, int H, int W, int Cin, const bf16* w1,
                      const float* b1, const bf16* w2, const float* b2, const bf16* w3, const float* b3, const bf16* res,
                      int res_cs, int res_co, bf16* out, int out_cs, int out_co) {
  RSG_REQUIRE(Cin >= 64 && Cin <= 256 && Cin % 64 == 0 && H >= 16 && W >= 8, "bottleneck: Cin=%d on %dx%d is not covered", Cin, H, W);
  RSG_REQUIRE(in_cs % 8 == 0 && in_co % 8 == 0 && out_cs % 8 == 0 && out_co % 8 == 0 && res_cs % 8 == 0 && res_co % 8 == 0,
              "bottleneck: channel strides/offsets must be multiples of 8");
  RSG_REQUIRE(in && res && out && w1 && w2 && w3 && b1 && b2 && b3, "bottleneck: null pointer");
  if (N == 0) return RSG_OK;
  BnP k;
  memset(&k, 0, sizeof(k));
  k.w1 = w1; k.w2 = w2; k.w3 = w3; k.b1 = b1; k.b2 = b2; k.b3 = b3;
  k.Cin = Cin; k.H = H; k.W = W; k.N = N; k.nchunks = Cin / 64;
  k.w1_bytes = (uint32_t)Cin * 64u * 2u;
  // wide rows (one 128-byte TMA row per pixel instead of eight 16-byte elements): the planar boxes of a DRAM-resident
  // 256-channel map arrive at ~11 B/cycle/SM with the two stages that fit (timeline: 8500 cycles per tile waiting for TMA)
  k.wide = rsg_dbg_env("RSG_BNECK_PLANAR") ? 0 : 1;
  k.stage_bytes = k.wide ? 184u * 128u : BN_STAGE;
  const size_t fixed = 1024 + (size_t)k.w1_bytes + BN_W2 + BN_W3 + BN_STAGE + BN_MID2;
  int S = (int)((225 * 1024 - (long long)fixed) / (long long)k.stage_bytes);
  if (S > BN_MAX_S) S = BN_MAX_S;
  { const char* e = rsg_dbg_env("RSG_BNECK_S"); if (e && atoi(e) >= 1 && atoi(e) <= S) S = atoi(e); }
  RSG_REQUIRE(S >= 2, "bottleneck: shared memory budget");
  k.S = S;
  k.tiles_x = (W + 7) / 8; k.tiles_y = (H + 15) / 16;
  k.tiles_per_img = (uint32_t)(k.tiles_x * k.tiles_y);
  const long long nt = (long long)k.tiles_per_img * N;
  RSG_REQUIRE(nt * k.tiles_per_img < (1ll << 32), "bottleneck: too many tiles");
  k.ntiles = (int)nt;
  k.magic_tpi = k.tiles_per_img > 1 ? (uint32_t)(((1ull << 32) + k.tiles_per_img - 1) / k.tiles_per_img) : 0u;
  k.magic_tx = k.tiles_x > 1 ? (uint32_t)(((1ull << 32) + k.tiles_x - 1) / k.tiles_x) : 0u;
  k.res = res; k.res_cs = res_cs; k.res_co = res_co;
  k.out = out; k.out_cs = out_cs; k.out_co = out_co;
  k.v32 = ((out_cs % 16 == 0 && out_co % 16 == 0 && ((uintptr_t)out & 31) == 0) ? 1 : 0) |
          ((res_cs % 16 == 0 && res_co % 16 == 0 && ((uintptr_t)res & 31) == 0) ? 2 : 0);
  { const char* e = rsg_dbg_env("RSG_BNECK_SKIP"); k.skip = e ? atoi(e) : 0; }
  if (rsg_dbg_env("RSG_BNECK_PREFETCH")) k.skip |= 32;
  static long long* dbg_buf = nullptr;
  if (rsg_dbg_env("RSG_BNECK_TIMELINE")) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 32 * 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 32 * 16 * sizeof(long long), s);
    k.dbg = dbg_buf;
  }
  const size_t smem = fixed + (size_t)S * k.stage_bytes;
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (k.wide) {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, (cuuint64_t)H * W * in_cs * 2};
    cuuint32_t box[4] = {64, BN_PITCH, 18, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(in + in_co), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (bottleneck, wide) failed with %d", (int)r);
  } else {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cin / 8), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, 16, (cuuint64_t)H * W * in_cs * 2};
    cuuint32_t box[5] = {8, BN_PITCH, 18, 8, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(in + in_co), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (bottleneck) failed with %d", (int)r);
  }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_bneck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    attr_once.done();
  }
  static const bool dbg = rsg_dbg_env("RSG_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[bneck] Cin=%d %dx%d S=%d smem=%zu tiles=%d\n", Cin, H, W, S, smem, k.ntiles);
  int gx = rsg_num_sms();
  if (gx > k.ntiles) gx = k.ntiles;
  RSG_CUDA(launch_pdl(conv_bneck_kernel, dim3((unsigned)gx), dim3(BN_THREADS), smem, s, map, k));
  if (k.dbg) {
    static int dumped = 0;
    if (dumped++ == 3) {
      long long h[32 * 16];
      cudaStreamSynchronize(s);
      cudaMemcpy(h, k.dbg, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[bneck timeline, CTA 0, cycles since conv1 start of tile 4] tile: c1_start c1_lastfull c1_issued e1_accok e1_done | "
                      "c2_start c2_issued e2_accok e2_done | c3_start e3_accok e3_done\n");
      const long long t0 = h[4 * 16];
      for (int i = 4; i < 16; ++i) {
        fprintf(stderr, "%2d:", i);
        for (int j = 0; j < 12; ++j) fprintf(stderr, " %7lld", h[i * 16 + j] ? h[i * 16 + j] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return RSG_OK;
}
