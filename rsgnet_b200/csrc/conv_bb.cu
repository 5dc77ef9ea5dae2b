// Fused HRNet BasicBlock (pose_rsgnet.py:25-54): out = relu(bn2(conv2(relu(bn1(conv1(x))))) + x), two 3x3
// stride-1 convs with C = 32 / 48 channels, BN folded, in ONE kernel.  The unfused pair is HBM-bound (500 MB per
// block at C = 32, 512 forwards: x read twice -- as input and as residual --, the intermediate written and read
// with its halo, the output written); fused, a block reads x once and writes the output once (200 MB): the
// intermediate lives in shared memory and the residual is the centre of the input patch that is already there.
//
// A CTA owns 16 x 8 output pixels.  Everything is a "flat" pixel array of pitch 12:
//   input patch   20 x 12 pixels (2-pixel halo), ONE 5-D TMA box, [C/8][240][8 ch]; out-of-image = zero fill
//   conv1         produces the 18 x 10 intermediate at flat positions o = r*12 + c (216 rows = two M tiles of 128,
//                 70 % useful); its A operand is the input patch with M rows = consecutive flat pixels (SBO = 128 B)
//                 and tap (dy,dx) = start offset (1+dy)*12 + (1+dx)
//   epilogue 1    TMEM -> +bias1, ReLU, ZERO outside the image (conv2's padding) -> bf16 -> shared memory
//                 [C/8][224][8 ch] at the same flat positions (consecutive lanes = consecutive 16 bytes)
//   conv2         128 output pixels = 16 rows of 8 (SBO = one pitch = 192 B), taps = the same start offsets
//   epilogue 2    TMEM -> +bias2 + x (read from the input patch in shared memory) -> ReLU -> bf16 NHWC stores
// Double-buffered conv1 / conv2 accumulators in TMEM and a double-buffered intermediate let conv1 of tile i+1 run
// while epilogue 1 of tile i converts and conv2 of tile i-1 drains.
//
// Warps (640 threads, one persistent CTA per SM): 0 and 3 = conv1 MMA issuers (+ TMEM, weights), 1 = conv2 MMA issuer,
// 2 = TMA producer, 4..19 = epilogue warps: each converts one 32-row x C/2-column piece of conv1's accumulator per
// tile (epilogue 1) and, for the tiles of its group's parity, one piece of conv2's accumulator two tiles later.
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int BB_THREADS = 640;
constexpr int BB_PITCH = 12, BB_IN_ROWS = 20, BB_IN_PX = BB_PITCH * BB_IN_ROWS;     // 240
constexpr int BB_MID_PX = 224;                                                       // >= 18 * 12, 16-byte planes
constexpr int BB_MAX_S = 8;

struct BbP {
  const bf16* w1; const bf16* w2;      // [tap][C/8][C][8] each (the tcgen05 packing with NS = C)
  const float* b1; const float* b2;
  int C, H, W, N;
  int tiles_x, tiles_y;
  uint32_t tiles_per_img, magic_tpi, magic_tx;
  int ntiles, S, NB;
  bf16* out;
  int out_cs, out_co;
  uint32_t w_bytes, stage_bytes, mid_bytes, tmem_cols;
  int skip;                            // debug: bit0 no stores, bit1 conv1 one tap, bit2 conv2 one tap, bit3 no epilogue-1 math/stores, bit4 no TMA
};

// 9 taps x K2 k16 steps, fully unrolled from register-resident offsets: the issuing thread is bound by the latency
// of its own instruction stream (profiles/r1_notes.md §6)
template <int K2>
__device__ __forceinline__ void issue_taps(uint32_t d, uint32_t a16, uint32_t w16, const uint32_t (&toff)[9], uint32_t a_k16,
                                           uint32_t b_tap16, uint32_t b_k16, uint32_t hiA, uint32_t hiB, uint32_t idesc,
                                           bool one_tap = false) {
  if (one_tap) {        // debug skip mode
    umma_f16(d, ((uint64_t)hiA << 32) | (a16 + toff[4]), ((uint64_t)hiB << 32) | (w16 + 4u * b_tap16), idesc, 0u);
    return;
  }
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
#pragma unroll
    for (int kk = 0; kk < K2; ++kk)
      umma_f16(d, ((uint64_t)hiA << 32) | (a16 + toff[tp] + (uint32_t)kk * a_k16),
               ((uint64_t)hiB << 32) | (w16 + (uint32_t)tp * b_tap16 + (uint32_t)kk * b_k16), idesc, (tp | kk) ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(BB_THREADS, 1)
conv_bb_kernel(const __grid_constant__ CUtensorMap in_map, const BbP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: 0 weights | full[S] | empty[S] | acc1 full x2 | acc1 empty x2 | mid full x2 | mid empty x2 | acc2 full x2 | acc2 empty x2
  __shared__ __align__(8) uint64_t bars[1 + 2 * BB_MAX_S + 24];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sB1[64], sB2[64];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // NB = 2 or 4 buffers for each of: conv1 accumulators, intermediate patch, conv2 accumulators.  The hand-offs of one
  // tile (conv1 -> epilogue 1 -> conv2 -> epilogue 2) are ~1800 cycles of barrier latency; four tiles in flight hide it.
  const int NB = p.NB;
  const uint32_t nbm = (uint32_t)NB - 1u, nbs = NB == 4 ? 2u : 1u;
  const int B_FULL = 1, B_EMPTY = 1 + p.S, B_A1F = 1 + 2 * p.S, B_A1E = B_A1F + NB, B_MF = B_A1E + NB, B_ME = B_MF + NB,
            B_A2F = B_ME + NB, B_A2E = B_A2F + NB;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  // layout: W1 | W2 | stages | mid x2
  const uint32_t off_w2 = p.w_bytes, off_st = 2u * p.w_bytes, off_mid = off_st + (uint32_t)p.S * p.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, c8 = C >> 3, k2n = C >> 4;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1);
    // a patch is free when both conv1 issuers have consumed it (two commits) AND the 8 epilogue-2 warps hold their residual
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 10); }
    for (int i = 0; i < NB; ++i) {
      mbar_init(BAR(B_A1F + i), 2); mbar_init(BAR(B_A1E + i), 16);
      mbar_init(BAR(B_MF + i), 16); mbar_init(BAR(B_ME + i), 1);
      mbar_init(BAR(B_A2F + i), 1); mbar_init(BAR(B_A2E + i), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sB1[i] = p.b1[i]; sB2[i] = p.b2[i]; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tm_acc2 = tmem_base + 2u * (uint32_t)NB * (uint32_t)C;   // acc1: [buf][tile] x C columns, acc2: [buf] x C behind
  if (threadIdx.x == 0) pdl_launch_dependents();
  const int first = blockIdx.x, step = gridDim.x;

  if (warp == 2) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
      uint32_t s = 0, ph = 0;
      for (int t = first; t < p.ntiles; t += step) {
        const int n = (int)fastdiv((uint32_t)t, p.magic_tpi);
        const int rem = t - n * (int)p.tiles_per_img;
        const int ty = (int)fastdiv((uint32_t)rem, p.magic_tx), tx = rem - ty * p.tiles_x;
        mbar_wait(BAR(B_EMPTY + s), ph ^ 1u);
        if (p.skip & 16) { mbar_arrive(BAR(B_FULL + s)); if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; } continue; }
        mbar_arrive_expect_tx(BAR(B_FULL + s), p.stage_bytes);
        tma_load_5d(sbase + off_st + s * p.stage_bytes, &in_map, BAR(B_FULL + s), 0, tx * 8 - 2, ty * 16 - 2, 0, n);
        if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // ===================== MMA issuers: warps 0 and 3 = conv1 (M tile 0 / 1), warp 1 = conv2 =====================
    // One thread cannot issue conv1's 36 thin MMAs (40 cycles of pipe each) fast enough: ~60-70 cycles of its own
    // instruction latency per MMA; three issuers of 18 MMAs each keep the pipe full.
    if (warp == 0 && elect_one()) {
      mbar_arrive_expect_tx(BAR(0), 2u * p.w_bytes);
      for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
        const uint32_t nb = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
        bulk_load(sbase + off, reinterpret_cast<const unsigned char*>(p.w1) + off, nb, BAR(0));
        bulk_load(sbase + off_w2 + off, reinterpret_cast<const unsigned char*>(p.w2) + off, nb, BAR(0));
      }
    }
    __syncwarp();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
    mbar_wait(BAR(0), 0);
    const uint32_t hiB = desc_hi(128u);
    const uint32_t lbo_b16 = ((uint32_t)C * 16u) >> 4;                   // weights: LBO = C rows x 16 B
    const uint32_t b_tap16 = (uint32_t)c8 * lbo_b16, b_k16 = 2u * lbo_b16;
    uint32_t toff[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) toff[tp] = (uint32_t)((tp / 3) * BB_PITCH + tp % 3);   // (1+dy)*12 + (1+dx)
    if (warp != 1) {
      const uint32_t mt = warp == 0 ? 0u : 1u;
      const uint32_t hiA = desc_hi(128u);                                // M rows = consecutive flat pixels
      const uint32_t lboA = (uint32_t)BB_IN_PX, a_k16 = 2u * lboA;       // 16-byte units
      const uint32_t w16 = (sbase >> 4) | (lbo_b16 << 16);
      uint32_t s = 0, ph = 0, i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
        mbar_wait(BAR(B_FULL + s), ph);
        mbar_wait(BAR(B_A1E + b), u ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = ((sbase + off_st + s * p.stage_bytes) >> 4) | (lboA << 16);
        if (elect_one()) {
          const uint32_t d = tmem_base + (b * 2u + mt) * (uint32_t)C;
          if (k2n == 2) issue_taps<2>(d, a16 + 128u * mt, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 2) != 0);
          else issue_taps<3>(d, a16 + 128u * mt, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 2) != 0);
          umma_commit(BAR(B_A1F + b));
          umma_commit(BAR(B_EMPTY + s));
        }
        __syncwarp();
        if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
      }
    } else {
      const uint32_t hiA = desc_hi((uint32_t)BB_PITCH * 16u);            // 8-pixel groups one pitch apart
      const uint32_t lboA = (uint32_t)BB_MID_PX, a_k16 = 2u * lboA;
      const uint32_t w16 = ((sbase + off_w2) >> 4) | (lbo_b16 << 16);
      uint32_t i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
        mbar_wait(BAR(B_MF + b), u);
        mbar_wait(BAR(B_A2E + b), u ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = ((sbase + off_mid + b * p.mid_bytes) >> 4) | (lboA << 16);
        if (elect_one()) {
          const uint32_t d = tm_acc2 + b * (uint32_t)C;
          if (k2n == 2) issue_taps<2>(d, a16, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 4) != 0);
          else issue_taps<3>(d, a16, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 4) != 0);
          umma_commit(BAR(B_ME + b));
          umma_commit(BAR(B_A2F + b));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4 .. 19 =====================
    // Every warp converts one 32-row x C/2-column piece of conv1's accumulator per tile (epilogue 1: 2 M tiles x 4
    // lane quarters x 2 column halves = 16 pieces), and the 8 warps of group (tile parity) finish tile i-2
    // (epilogue 2) right after: conv2 of that tile completed long ago, so nobody waits on the tensor pipe, and the
    // per-warp instruction streams of the two epilogues are balanced (1 + 1/2 piece per tile).
    const int ew = warp - 4, q = warp & 3, grp = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int ncol = C >> 1, cb = half * ncol;           // this warp's columns [cb, cb + ncol), 16 or 24
    // --- epilogue-1 constants
    const int mt = grp;
    const int o = mt * 128 + q * 32 + lane;              // flat position in the pitch-12 array
    const int r = o / BB_PITCH, c = o - r * BB_PITCH;
    const bool valid = r < 18 && c < 10;
    // --- epilogue-2 constants
    const int m = q * 32 + lane, hy = m >> 3, wx = m & 7;
    const uint32_t res_px = (uint32_t)((hy + 2) * BB_PITCH + wx + 2) * 16u;
    bf16* const outp = p.out + p.out_co + cb;
    pdl_wait();

    auto ld_cols = [&](uint32_t taddr, int c0, uint32_t* v) {      // 16 columns, or the last 8 of a 24-column piece
      if (ncol - c0 >= 16) tmem_ld16(taddr + (uint32_t)c0, v);
      else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr + (uint32_t)c0));
#pragma unroll
        for (int j = 8; j < 16; ++j) v[j] = 0;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    };
    uint32_t s2 = 0, ph2 = 0;                            // stage ring position of epilogue 2

    auto epilogue2 = [&](uint32_t j, int t) {            // CTA-local tile j (global tile t), handled by group j & 1
      const uint32_t b = j & nbm, u = (j >> nbs) & 1u;
      const int n = (int)fastdiv((uint32_t)t, p.magic_tpi);
      const int rem = t - n * (int)p.tiles_per_img;
      const int ty = (int)fastdiv((uint32_t)rem, p.magic_tx), tx = rem - ty * p.tiles_x;
      const int y = ty * 16 + hy, x = tx * 8 + wx;
      const bool ok = y < p.H && x < p.W;
      mbar_wait(BAR(B_FULL + s2), ph2);                  // (complete long ago) residual = centre of the input patch
      const unsigned char* resp = sgen + off_st + s2 * p.stage_bytes + res_px;
      uint4 rv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k * 8 < ncol) rv[k] = *reinterpret_cast<const uint4*>(resp + (size_t)((cb >> 3) + k) * (BB_IN_PX * 16));
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_EMPTY + s2));     // the patch is free once all 8 warps hold their residual
      mbar_wait(BAR(B_A2F + b), u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tm_acc2 + ((uint32_t)(q * 32) << 16) + b * (uint32_t)C + (uint32_t)cb;
      const uint32_t ooff = (((uint32_t)n * (uint32_t)p.H + (uint32_t)y) * (uint32_t)p.W + (uint32_t)x) * (uint32_t)p.out_cs;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = ci * 16;
        if (c0 >= ncol) break;
        uint32_t v[16];
        ld_cols(taddr, c0, v);
        if (c0 + 16 >= ncol) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_A2E + b));
        }
        const int nv = ncol - c0 >= 16 ? 16 : 8;
        float f[16];
        const float4* bp = reinterpret_cast<const float4*>(sB2 + cb + c0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = (j4 * 4 < nv) ? bp[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
          f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + bb.x;
          f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + bb.y;
          f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + bb.z;
          f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + bb.w;
        }
        add_res8(f, rv[ci * 2]);
        if (nv == 16) add_res8(f + 8, rv[ci * 2 + 1 < 3 ? ci * 2 + 1 : 2]);
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          h0[k] = __floats2bfloat162_rn(fmaxf(f[2 * k], 0.f), fmaxf(f[2 * k + 1], 0.f));
          h1[k] = __floats2bfloat162_rn(fmaxf(f[8 + 2 * k], 0.f), fmaxf(f[8 + 2 * k + 1], 0.f));
        }
        if (ok && !(p.skip & 1)) {
          uint4* op = reinterpret_cast<uint4*>(outp + ooff + c0);
          op[0] = o0;
          if (nv == 16) op[1] = o1;
        }
      }
    };
    // epilogue 2 walks the stage ring over the tiles of ITS parity only: advance by two slots per handled tile
    auto advance2 = [&]() {
      for (int k = 0; k < 2; ++k)
        if (++s2 == (uint32_t)p.S) { s2 = 0; ph2 ^= 1u; }
    };
    if (grp == 1) { if (++s2 == (uint32_t)p.S) { s2 = 0; ph2 ^= 1u; } }      // group 1 starts at local tile 1

    uint32_t i = 0;
    int t = first;
    for (; t < p.ntiles; t += step, ++i) {
      // ---------------- epilogue 1 of tile i ----------------
      const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
      const int n = (int)fastdiv((uint32_t)t, p.magic_tpi);
      const int rem = t - n * (int)p.tiles_per_img;
      const int ty = (int)fastdiv((uint32_t)rem, p.magic_tx), tx = rem - ty * p.tiles_x;
      const int gy = ty * 16 - 1 + r, gx = tx * 8 - 1 + c;
      const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      mbar_wait(BAR(B_A1F + b), u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (b * 2u + (uint32_t)mt) * (uint32_t)C + (uint32_t)cb;
      mbar_wait(BAR(B_ME + b), u ^ 1u);                  // conv2 of tile i-2 has consumed this buffer
      unsigned char* dst = sgen + off_mid + b * p.mid_bytes + (uint32_t)o * 16u + (size_t)(cb >> 3) * (BB_MID_PX * 16);
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = ci * 16;
        if (c0 >= ncol) break;
        uint32_t v[16];
        ld_cols(taddr, c0, v);
        if (c0 + 16 >= ncol) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_A1E + b));
        }
        const int nv = ncol - c0 >= 16 ? 16 : 8;
        if (valid && !(p.skip & 8)) {
          uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
          if (inside) {
            float f[16];
            const float4* bp = reinterpret_cast<const float4*>(sB1 + cb + c0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bb = (j4 * 4 < nv) ? bp[j4] : make_float4(0.f, 0.f, 0.f, 0.f);
              f[4 * j4 + 0] = fmaxf(__uint_as_float(v[4 * j4 + 0]) + bb.x, 0.f);
              f[4 * j4 + 1] = fmaxf(__uint_as_float(v[4 * j4 + 1]) + bb.y, 0.f);
              f[4 * j4 + 2] = fmaxf(__uint_as_float(v[4 * j4 + 2]) + bb.z, 0.f);
              f[4 * j4 + 3] = fmaxf(__uint_as_float(v[4 * j4 + 3]) + bb.w, 0.f);
            }
            __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
            __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              h0[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
              h1[k] = __floats2bfloat162_rn(f[8 + 2 * k], f[8 + 2 * k + 1]);
            }
          }
          *reinterpret_cast<uint4*>(dst + (size_t)(c0 >> 3) * (BB_MID_PX * 16)) = o0;
          if (nv == 16) *reinterpret_cast<uint4*>(dst + (size_t)((c0 >> 3) + 1) * (BB_MID_PX * 16)) = o1;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_MF + b));
      // ---------------- epilogue 2 of tile i-2 (this group's parity) ----------------
      if (i >= 2 && ((i - 2) & 1u) == (uint32_t)grp) {
        epilogue2(i - 2, t - 2 * step);
        advance2();
      }
    }
    // drain: the last two tiles
    for (uint32_t j = i >= 2 ? i - 2 : 0; j < i; ++j) {
      if ((j & 1u) == (uint32_t)grp) {
        epilogue2(j, first + (int)j * step);
        advance2();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace

// 1 when the fused kernel covers a C -> C BasicBlock on H x W maps.
extern "C" int rsg_basic_block_supported(int C, int H, int W) {
  if (getenv("RSG_DISABLE_BB")) return 0;
  return (C == 32 || C == 48) && H >= 16 && W >= 8 ? 1 : 0;
}

int conv_bb_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, int N, int H, int W, int C, const bf16* w1,
                   const float* b1, const bf16* w2, const float* b2, bf16* out, int out_cs, int out_co) {
  RSG_REQUIRE(rsg_basic_block_supported(C, H, W) || getenv("RSG_DISABLE_BB"), "basic block: C=%d on %dx%d is not covered", C, H, W);
  RSG_REQUIRE(in_cs % 8 == 0 && in_co % 8 == 0 && out_cs % 8 == 0 && out_co % 8 == 0, "basic block: channel strides/offsets must be multiples of 8");
  if (N == 0) return RSG_OK;
  BbP k;
  memset(&k, 0, sizeof(k));
  k.w1 = w1; k.w2 = w2; k.b1 = b1; k.b2 = b2; k.C = C; k.H = H; k.W = W; k.N = N;
  k.tiles_x = (W + 7) / 8; k.tiles_y = (H + 15) / 16;
  k.tiles_per_img = (uint32_t)(k.tiles_x * k.tiles_y);
  const long long nt = (long long)k.tiles_per_img * N;
  RSG_REQUIRE(nt * k.tiles_per_img < (1ll << 32), "basic block: too many tiles");
  k.ntiles = (int)nt;
  k.magic_tpi = k.tiles_per_img > 1 ? (uint32_t)(((1ull << 32) + k.tiles_per_img - 1) / k.tiles_per_img) : 0u;
  k.magic_tx = k.tiles_x > 1 ? (uint32_t)(((1ull << 32) + k.tiles_x - 1) / k.tiles_x) : 0u;
  k.out = out; k.out_cs = out_cs; k.out_co = out_co;
  k.w_bytes = 9u * C * C * 2u;
  k.stage_bytes = (uint32_t)(C / 8) * BB_IN_PX * 16u;
  k.mid_bytes = (uint32_t)(C / 8) * BB_MID_PX * 16u;
  int NB = 12 * C <= 512 ? 4 : 2;                    // TMEM: 3 * NB * C columns
  { const char* e = getenv("RSG_BB_NB"); if (e && (atoi(e) == 2 || atoi(e) == 4) && 3 * atoi(e) * C <= 512) NB = atoi(e); }
  k.NB = NB;
  int S = (int)((220 * 1024 - 2 * (int)k.w_bytes - NB * (int)k.mid_bytes) / (int)k.stage_bytes);
  if (S > BB_MAX_S) S = BB_MAX_S;
  RSG_REQUIRE(S >= 2, "basic block: shared memory budget");
  k.S = S;
  uint32_t cols = 32;
  while (cols < 3u * NB * C) cols <<= 1;
  k.tmem_cols = cols;
  { const char* e = getenv("RSG_BB_SKIP"); k.skip = e ? atoi(e) : 0; }
  const size_t smem = 128 + 2 * (size_t)k.w_bytes + (size_t)S * k.stage_bytes + NB * (size_t)k.mid_bytes + 1024;
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, 16, (cuuint64_t)H * W * in_cs * 2};
    cuuint32_t box[5] = {8, BB_PITCH, BB_IN_ROWS, (cuuint32_t)(C / 8), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(in + in_co), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (basic block) failed with %d", (int)r);
  }
  static bool attr_done = false;
  if (!attr_done) {
    RSG_CUDA(cudaFuncSetAttribute(conv_bb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    attr_done = true;
  }
  int gx = rsg_num_sms();
  if (gx > k.ntiles) gx = k.ntiles;
  RSG_CUDA(launch_pdl(conv_bb_kernel, dim3((unsigned)gx), dim3(BB_THREADS), smem, s, map, k));
  return RSG_OK;
}
