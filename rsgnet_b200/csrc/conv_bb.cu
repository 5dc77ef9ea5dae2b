// Fused HRNet BasicBlock (pose_rsgnet.py:25-54): out = relu(bn2(conv2(relu(bn1(conv1(x))))) + x), two 3x3
// stride-1 convs with C = 32 / 48 channels, BN folded, in ONE kernel.  The unfused pair is HBM-bound (500 MB per
// block at C = 32, 512 forwards: x read twice -- as input and as residual --, the intermediate written and read
// with its halo, the output written); fused, a block reads x once and writes the output once (200 MB): the
// intermediate lives in shared memory and the residual is the centre of the input patch that is already there.
//
// A CTA owns 16 x TWT output pixels (TWT = 16, or 8 where shared memory / the map width do not allow 16).
// Everything is a "flat" pixel array of pitch P = TWT + 4:
//   input patch   20 x P pixels (2-pixel halo), ONE 5-D TMA box, [C/8][20 P][8 ch]; out-of-image = zero fill
//   conv1         produces the 18 x (TWT+2) intermediate at flat positions o = r*P + c (MT1 = 3 / 2 M tiles of 128
//                 rows, 84 % / 70 % useful); its A operand is the input patch with M rows = consecutive flat pixels
//                 (SBO = 128 B) and tap (dy,dx) = start offset (1+dy)*P + (1+dx)
//   epilogue 1    TMEM -> +bias1, ReLU, ZERO outside the image (conv2's padding) -> bf16 -> shared memory
//                 [C/8][MT1*128][8 ch] at the same flat positions (consecutive lanes = consecutive 16 bytes)
//   conv2         per 8-pixel column strip: 128 output pixels = 16 rows of 8 (SBO = one pitch), the same tap offsets
//   epilogue 2    TMEM -> +bias2 + x (read from the input patch in shared memory) -> ReLU -> bf16 NHWC stores
// NB tiles are in flight in TMEM / the intermediate buffers, so conv1 of tile i+1 runs while epilogue 1 of tile i
// converts and conv2 of tile i-1 drains; ncu: the tensor pipe (operand fetch of N = C MMAs) is >90 % busy.
//
// Warps (640 threads, one persistent CTA per SM): 0 and 3 = conv1 MMA issuers (+ TMEM, weights), 1 = conv2 MMA issuer,
// 2 = TMA producer, 4..19 = epilogue warps: pieces of 32 rows x C/2 columns of conv1's accumulators (epilogue 1) and,
// NB/2 tiles later, of conv2's accumulators (epilogue 2), spread so that the per-warp instruction streams balance.
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int BB_THREADS = 640;
constexpr int BB_IN_ROWS = 20;
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

template <int TWT>
__global__ void __launch_bounds__(BB_THREADS, 1)
conv_bb_kernel(const __grid_constant__ CUtensorMap in_map, const BbP p) {
  constexpr int PITCH = TWT + 4, IN_PX = BB_IN_ROWS * PITCH, MID_W = TWT + 2;
  constexpr int MT1 = (18 * PITCH + 127) / 128, MID_PX = MT1 * 128, MT2 = TWT / 8;
  constexpr int P1 = 2 * MT1;                          // epilogue-1 pieces per TMEM lane quarter: (M tile, column half)
  constexpr int E2W = MT2 == 2 ? 16 : 8;               // epilogue-2 warps per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: 0 weights | full[S] | empty[S] | acc1 full | acc1 empty | mid full | mid empty | acc2 full | acc2 empty (NB each)
  __shared__ __align__(8) uint64_t bars[1 + 2 * BB_MAX_S + 24];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sB1[64], sB2[64];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // NB = 2 or 4 buffers for each of: conv1 accumulators, intermediate patch, conv2 accumulators.  The hand-offs of one
  // tile (conv1 -> epilogue 1 -> conv2 -> epilogue 2) are ~1800 cycles of barrier latency; several tiles in flight hide it.
  const int NB = p.NB;
  const uint32_t nbm = (uint32_t)NB - 1u, nbs = NB == 4 ? 2u : 1u, lag = (uint32_t)NB >> 1;
  const int B_FULL = 1, B_EMPTY = 1 + p.S, B_A1F = 1 + 2 * p.S, B_A1E = B_A1F + NB, B_MF = B_A1E + NB, B_ME = B_MF + NB,
            B_A2F = B_ME + NB, B_A2E = B_A2F + NB;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  // layout: W1 | W2 | stages | mid x NB
  const uint32_t off_w2 = p.w_bytes, off_st = 2u * p.w_bytes, off_mid = off_st + (uint32_t)p.S * p.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, c8 = C >> 3, k2n = C >> 4;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1);
    // a patch is free when both conv1 issuers have consumed it (two commits) AND the epilogue-2 warps hold their residual
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), E2W + 2); }
    for (int i = 0; i < NB; ++i) {
      mbar_init(BAR(B_A1F + i), 2); mbar_init(BAR(B_A1E + i), 16);
      mbar_init(BAR(B_MF + i), 16); mbar_init(BAR(B_ME + i), 1);
      mbar_init(BAR(B_A2F + i), 1); mbar_init(BAR(B_A2E + i), E2W);
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
  const uint32_t tm_acc2 = tmem_base + (uint32_t)(NB * MT1) * (uint32_t)C;   // acc1: [buf][M tile] x C, acc2: [buf][strip] x C
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
        tma_load_5d(sbase + off_st + s * p.stage_bytes, &in_map, BAR(B_FULL + s), 0, tx * TWT - 2, ty * 16 - 2, 0, n);
        if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // ===================== MMA issuers: warps 0 and 3 = conv1 (split over the M tiles), warp 1 = conv2 =====================
    // One thread cannot issue all of conv1's thin MMAs (40 cycles of pipe each) fast enough: ~60-70 cycles of its own
    // instruction latency per MMA; three issuers keep the pipe full.
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
    for (int tp = 0; tp < 9; ++tp) toff[tp] = (uint32_t)((tp / 3) * PITCH + tp % 3);   // (1+dy)*P + (1+dx)
    if (warp != 1) {
      // M tiles [mt0, mt1) of conv1: warp 0 takes the first ceil(MT1/2), warp 3 the rest
      constexpr int SPLIT = (MT1 + 1) / 2;
      const int mt0 = warp == 0 ? 0 : SPLIT, mt1 = warp == 0 ? SPLIT : MT1;
      const uint32_t hiA = desc_hi(128u);                                // M rows = consecutive flat pixels
      const uint32_t lboA = (uint32_t)IN_PX, a_k16 = 2u * lboA;          // 16-byte units
      const uint32_t w16 = (sbase >> 4) | (lbo_b16 << 16);
      uint32_t s = 0, ph = 0, i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
        mbar_wait(BAR(B_FULL + s), ph);
        mbar_wait(BAR(B_A1E + b), u ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = ((sbase + off_st + s * p.stage_bytes) >> 4) | (lboA << 16);
        if (elect_one()) {
          for (int mt = mt0; mt < mt1; ++mt) {
            const uint32_t d = tmem_base + (b * (uint32_t)MT1 + (uint32_t)mt) * (uint32_t)C;
            if (k2n == 2) issue_taps<2>(d, a16 + 128u * (uint32_t)mt, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 2) != 0);
            else issue_taps<3>(d, a16 + 128u * (uint32_t)mt, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 2) != 0);
          }
          umma_commit(BAR(B_A1F + b));
          umma_commit(BAR(B_EMPTY + s));
        }
        __syncwarp();
        if (++s == (uint32_t)p.S) { s = 0; ph ^= 1u; }
      }
    } else {
      const uint32_t hiA = desc_hi((uint32_t)PITCH * 16u);               // 8-pixel groups one pitch apart
      const uint32_t lboA = (uint32_t)MID_PX, a_k16 = 2u * lboA;
      const uint32_t w16 = ((sbase + off_w2) >> 4) | (lbo_b16 << 16);
      uint32_t i = 0;
      for (int t = first; t < p.ntiles; t += step, ++i) {
        const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
        mbar_wait(BAR(B_MF + b), u);
        mbar_wait(BAR(B_A2E + b), u ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = ((sbase + off_mid + b * p.mid_bytes) >> 4) | (lboA << 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < MT2; ++k) {                                // 8-pixel column strips
            const uint32_t d = tm_acc2 + (b * (uint32_t)MT2 + (uint32_t)k) * (uint32_t)C;
            if (k2n == 2) issue_taps<2>(d, a16 + 8u * (uint32_t)k, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 4) != 0);
            else issue_taps<3>(d, a16 + 8u * (uint32_t)k, w16, toff, a_k16, b_tap16, b_k16, hiA, hiB, idesc, (p.skip & 4) != 0);
          }
          umma_commit(BAR(B_ME + b));
          umma_commit(BAR(B_A2F + b));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4 .. 19 =====================
    // Pieces are 32 rows (the warp's TMEM lane quarter) x C/2 columns.  Epilogue 1 of a tile has P1 = 2*MT1 pieces per
    // quarter, spread over the quarter's 4 warps (rotated by tile parity when P1 is not a multiple of 4); epilogue 2
    // of tile i - NB/2 follows (conv2 of that tile completed long ago, nobody waits on the tensor pipe): 2*MT2 pieces
    // per quarter -- one per warp for 16-wide tiles, alternating warp pairs by tile parity for 8-wide tiles.
    const int ew = warp - 4, q = warp & 3, w4 = ew >> 2;
    const int ncol = C >> 1;                             // columns per piece: 16 or 24
    const int m = q * 32 + lane, hy = m >> 3, wx = m & 7;
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
    auto tile_coords = [&](int t, int& n, int& ty, int& tx) {
      n = (int)fastdiv((uint32_t)t, p.magic_tpi);
      const int rem = t - n * (int)p.tiles_per_img;
      ty = (int)fastdiv((uint32_t)rem, p.magic_tx);
      tx = rem - ty * p.tiles_x;
    };

    // ---- epilogue 2 of CTA-local tile j (global tile t); `mine` = this warp has a piece of it
    uint32_t s2 = 0, ph2 = 0;                            // stage ring position of tile j, carried incrementally
    auto advance2 = [&]() { if (++s2 == (uint32_t)p.S) { s2 = 0; ph2 ^= 1u; } };
    auto epilogue2 = [&](uint32_t j, int t) {            // called for every j in order
      int piece;                                         // (strip, half) = (piece >> 1, piece & 1)
      if (MT2 == 2) piece = w4;
      else { if ((uint32_t)(w4 >> 1) != (j & 1u)) { advance2(); return; } piece = w4 & 1; }
      const int strip = piece >> 1, cb = (piece & 1) * ncol;
      const uint32_t b = j & nbm, u = (j >> nbs) & 1u;
      int n, ty, tx;
      tile_coords(t, n, ty, tx);
      const int y = ty * 16 + hy, x = tx * TWT + 8 * strip + wx;
      const bool ok = y < p.H && x < p.W;
      mbar_wait(BAR(B_FULL + s2), ph2);                  // (complete long ago) residual = centre of the input patch
      const unsigned char* resp = sgen + off_st + s2 * p.stage_bytes + (uint32_t)((hy + 2) * PITCH + 8 * strip + wx + 2) * 16u;
      uint4 rv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k * 8 < ncol) rv[k] = *reinterpret_cast<const uint4*>(resp + (size_t)((cb >> 3) + k) * (IN_PX * 16));
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_EMPTY + s2));     // the patch is free once all epilogue-2 warps hold their residual
      mbar_wait(BAR(B_A2F + b), u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tm_acc2 + ((uint32_t)(q * 32) << 16) + (b * (uint32_t)MT2 + (uint32_t)strip) * (uint32_t)C + (uint32_t)cb;
      bf16* const outp = p.out + p.out_co + cb +
                         (((size_t)n * p.H + (size_t)y) * p.W + (size_t)x) * (size_t)p.out_cs;
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
          uint4* op = reinterpret_cast<uint4*>(outp + c0);
          op[0] = o0;
          if (nv == 16) op[1] = o1;
        }
      }
      advance2();
    };

    uint32_t i = 0;
    int t = first;
    for (; t < p.ntiles; t += step, ++i) {
      // ---------------- epilogue 1 of tile i ----------------
      const uint32_t b = i & nbm, u = (i >> nbs) & 1u;
      int n, ty, tx;
      tile_coords(t, n, ty, tx);
      mbar_wait(BAR(B_A1F + b), u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mbar_wait(BAR(B_ME + b), u ^ 1u);                  // conv2 of tile i-NB has consumed this buffer
      const int slot = (P1 % 4 == 0) ? w4 : ((w4 + 2 * (int)(i & 1u)) & 3);
      const int npieces = (P1 - slot + 3) / 4;           // pieces slot, slot+4, ...
      for (int pi = 0; pi < npieces; ++pi) {
        const int piece = slot + 4 * pi, mt = piece >> 1, cb = (piece & 1) * ncol;
        const int o = mt * 128 + q * 32 + lane;          // flat position in the pitch-P array
        const int r = o / PITCH, c = o - r * PITCH;
        const bool valid = r < 18 && c < MID_W;
        const int gy = ty * 16 - 1 + r, gx = tx * TWT - 1 + c;
        const bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (b * (uint32_t)MT1 + (uint32_t)mt) * (uint32_t)C + (uint32_t)cb;
        unsigned char* dst = sgen + off_mid + b * p.mid_bytes + (uint32_t)o * 16u + (size_t)(cb >> 3) * (MID_PX * 16);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c0 = ci * 16;
          if (c0 >= ncol) break;
          uint32_t v[16];
          ld_cols(taddr, c0, v);
          if (pi == npieces - 1 && c0 + 16 >= ncol) {    // all of this warp's reads of conv1's accumulators are done
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
            *reinterpret_cast<uint4*>(dst + (size_t)(c0 >> 3) * (MID_PX * 16)) = o0;
            if (nv == 16) *reinterpret_cast<uint4*>(dst + (size_t)((c0 >> 3) + 1) * (MID_PX * 16)) = o1;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_MF + b));
      // ---------------- epilogue 2 of tile i - lag ----------------
      if (i >= lag) epilogue2(i - lag, t - (int)lag * step);
    }
    // drain: the last `lag` tiles
    for (uint32_t j = i >= lag ? i - lag : 0; j < i; ++j) epilogue2(j, first + (int)j * step);
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
  if (rsg_dbg_env("RSG_DISABLE_BB")) return 0;
  return (C == 32 || C == 48) && H >= 16 && W >= 8 ? 1 : 0;
}

int conv_bb_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, int N, int H, int W, int C, const bf16* w1,
                   const float* b1, const bf16* w2, const float* b2, bf16* out, int out_cs, int out_co) {
  RSG_REQUIRE(rsg_basic_block_supported(C, H, W) || rsg_dbg_env("RSG_DISABLE_BB"), "basic block: C=%d on %dx%d is not covered", C, H, W);
  RSG_REQUIRE(in_cs % 8 == 0 && in_co % 8 == 0 && out_cs % 8 == 0 && out_co % 8 == 0, "basic block: channel strides/offsets must be multiples of 8");
  if (N == 0) return RSG_OK;
  BbP k;
  memset(&k, 0, sizeof(k));
  k.w1 = w1; k.w2 = w2; k.b1 = b1; k.b2 = b2; k.C = C; k.H = H; k.W = W; k.N = N;
  k.w_bytes = 9u * C * C * 2u;
  // 16-wide tiles (conv1 on 324 useful rows of 384 instead of 180 of 256) where the map and shared memory allow
  int twt = 16;
  { const char* e = rsg_dbg_env("RSG_BB_TW"); if (e && (atoi(e) == 8 || atoi(e) == 16)) twt = atoi(e); }
  int NB = 2, S = 0;
  for (;; twt = 8) {
    const int pitch = twt + 4, in_px = BB_IN_ROWS * pitch, mt1 = (18 * pitch + 127) / 128, mt2 = twt / 8;
    k.stage_bytes = (uint32_t)(C / 8) * in_px * 16u;
    k.mid_bytes = (uint32_t)(C / 8) * mt1 * 128u * 16u;
    NB = (2 * (mt1 + mt2) * 2 * C <= 512 && twt == 8) ? 4 : 2;            // TMEM: NB * (MT1 + MT2) * C columns
    { const char* e = rsg_dbg_env("RSG_BB_NB"); if (e && (atoi(e) == 2 || atoi(e) == 4) && atoi(e) * (mt1 + mt2) * C <= 512) NB = atoi(e); }
    S = (int)((220 * 1024 - 2 * (int)k.w_bytes - NB * (int)k.mid_bytes) / (int)k.stage_bytes);
    if (S > BB_MAX_S) S = BB_MAX_S;
    const bool fits = S >= 3 && NB * (mt1 + mt2) * C <= 512 && (twt == 8 || W >= 16);
    if (fits || twt == 8) {
      uint32_t cols = 32;
      while (cols < (uint32_t)(NB * (mt1 + mt2) * C)) cols <<= 1;
      k.tmem_cols = cols;
      break;
    }
  }
  RSG_REQUIRE(S >= 2, "basic block: shared memory budget");
  k.S = S; k.NB = NB;
  k.tiles_x = (W + twt - 1) / twt; k.tiles_y = (H + 15) / 16;
  k.tiles_per_img = (uint32_t)(k.tiles_x * k.tiles_y);
  const long long nt = (long long)k.tiles_per_img * N;
  RSG_REQUIRE(nt * k.tiles_per_img < (1ll << 32), "basic block: too many tiles");
  k.ntiles = (int)nt;
  k.magic_tpi = k.tiles_per_img > 1 ? (uint32_t)(((1ull << 32) + k.tiles_per_img - 1) / k.tiles_per_img) : 0u;
  k.magic_tx = k.tiles_x > 1 ? (uint32_t)(((1ull << 32) + k.tiles_x - 1) / k.tiles_x) : 0u;
  k.out = out; k.out_cs = out_cs; k.out_co = out_co;
  { const char* e = rsg_dbg_env("RSG_BB_SKIP"); k.skip = e ? atoi(e) : 0; }
  const size_t smem = 128 + 2 * (size_t)k.w_bytes + (size_t)S * k.stage_bytes + NB * (size_t)k.mid_bytes + 1024;
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_cs * 2, (cuuint64_t)W * in_cs * 2, 16, (cuuint64_t)H * W * in_cs * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)(twt + 4), BB_IN_ROWS, (cuuint32_t)(C / 8), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(in + in_co), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (basic block) failed with %d", (int)r);
  }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_bb_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    RSG_CUDA(cudaFuncSetAttribute(conv_bb_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    attr_once.done();
  }
  static const bool dbg = rsg_dbg_env("RSG_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[bb] C=%d %dx%d tile 16x%d NB=%d S=%d smem=%zu tiles=%d tmem=%u\n", C, H, W, twt, NB, S, smem, k.ntiles, k.tmem_cols);
  int gx = rsg_num_sms();
  if (gx > k.ntiles) gx = k.ntiles;
  if (twt == 16) RSG_CUDA(launch_pdl(conv_bb_kernel<16>, dim3((unsigned)gx), dim3(BB_THREADS), smem, s, map, k));
  else RSG_CUDA(launch_pdl(conv_bb_kernel<8>, dim3((unsigned)gx), dim3(BB_THREADS), smem, s, map, k));
  return RSG_OK;
}
