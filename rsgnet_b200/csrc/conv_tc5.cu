// Blackwell-native implicit-GEMM convolution: tcgen05.mma (single-thread issue, fp32 accumulators
// in TMEM) for stride-1 convolutions whose taps lie in the 3x3 neighbourhood (3x3 p1 and 1x1).
//
// "Halo-resident" formulation.  A CTA owns a 16 x 8 pixel patch of one crop (= the 128 rows of the
// MMA M dimension).  The (16+2) x (8+2) halo patch of KC input channels is brought into shared
// memory ONCE, laid out as [KC/8][18][10][8 ch] -- exactly the canonical K-major / no-swizzle UMMA
// operand layout (8 consecutive pixels x 16 bytes = one core matrix, SBO = one halo row = 160 B,
// LBO = one 8-channel plane = 2880 B).  Each of the 9 taps is then just a different START ADDRESS
// of the A descriptor into that same patch ((1+dy)*10 + (1+dx) pixels further), so the activation
// tile is read from L2 once (1.4x with halo) instead of 9x, and the zero padding of the
// convolution is the zero-fill of the out-of-bounds copies.  Weights (BN folded, packed
// [slice][tap][Cin/8][NS][8]) are TMA-bulk-copied once per CTA and stay resident while the CTA
// walks over its tiles; blockIdx.y selects a slice of NS output channels.
//
// The halo patch is ONE 5-D TMA tensor load per stage (box 8 ch x 10 px x 18 rows x KC/8 planes x
// 1 crop over the NHWC activation viewed as (8, W, H, C/8, N)); a cp.async (LDGSTS) gather with the
// same layout was measured at 1.2-2 TB/s per chip against >5 TB/s for the TMA box once the other
// bottlenecks were gone (profiles/r1_notes.md).
//
// Warp roles (608 threads, one persistent CTA per SM): warps 0-1 = MMA issuers alternating tiles
// (warp 0 also allocates TMEM and bulk-copies the weights), warp 2 = TMA producer (one elected
// thread), warps 3..18 = two epilogue groups of 8 warps alternating tiles (TMEM -> registers -> +bias
// +residual terms -> ReLU -> bf16 NHWC stores).  Pipelines: S-deep ring of halo stages (full/empty
// mbarriers, one stage = one K-chunk of one tile, half of the ring per MMA warp) and up to 8 TMEM
// accumulators (full/empty mbarriers).
//
// Reference ops subsumed: Conv2d(3x3|1x1, s1) + BatchNorm2d(eval) [+ residual adds] [+ ReLU]
// (pose_rsgnet.py:38-54 BasicBlock, :75-95 Bottleneck, :261-270 fuse sum, heads :965-1003).
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int TH = 16, TW = 8;                 // pixel patch = 128 MMA rows
constexpr int NTHREADS = 640;                  // one fat persistent CTA per SM (20 warps = 5 per scheduler -> 96 registers/thread)
constexpr int MAX_MMA_WARPS = 3;               // warps 0..2 may issue MMAs (p.NMMA = 2 or 3 of them do), round-robin over tiles
constexpr int EPI_WARP0 = 4;                   // warp 3 = TMA producer; epilogue warps 4..19: two groups of 8
constexpr int NEPI = 2;
struct Tc5P {
  const bf16* w;        // [nslices][ntaps][Cin/8][NS][8]
  const float* bias;    // [CoutPad]
  int Cin, NS, Cout;    // NS = output channels per CTA (blockIdx.y selects the slice)
  int KC, nchunks;      // channels per halo stage, Cin / KC
  int S;                // halo ring depth
  int NACC;             // TMEM accumulator ring depth
  int NMMA;             // MMA-issuing warps: 2 or 3 (S is a multiple of it: every issuer owns S/NMMA ring slots)
  int EW;               // epilogue warps: 4, or 8 (two column halves per TMEM lane quarter)
  int ntaps;
  // per-tap A descriptor pieces: low word = (start offset inside the stage | LBO << 16) in 16-byte units,
  // high word = SBO | version, k16 step in 16-byte units.  Stride 1: every tap addresses the one halo
  // patch; stride 2: the patch of its (row parity, column parity) phase.
  uint32_t tap_lo[9], tap_hi[9], tap_kstep[9];
  int halo;             // 1 for 3x3, 0 for 1x1 (stride 1)
  int stride;           // 1, or 2: four phase patches per stage, loaded with TMA element strides of 2
  uint32_t ph_off[4];   // stride 2: byte offset of each phase patch inside a stage
  int H, W;             // iteration space (= output of the conv before the omul/oo mapping)
  int oH, oW, omul, ooy, oox;   // out pixel = (y*omul+ooy, x*omul+oox) in an oH x oW map
  int psC;              // > 0: pixel shuffle -- output column c belongs to phase c / psC (oy = 2y + (ph>>1),
                        // ox = 2x + (ph&1)) and channel c % psC: ConvTranspose(4,2,1) as ONE 3x3 conv
  const bf16* in;       // NHWC, channel stride in_cs, first channel in_co
  int in_cs, in_co;
  int tiles_x, tiles_y; // per image
  uint32_t tiles_per_img, magic_tpi, magic_tx;   // __umulhi(t, magic) == t / divisor (host-checked range)
  long long ntiles;
  bf16* out;
  int out_cs, out_co;
  int nres;
  ResP res[4];
  int relu;
  int v32;              // bit0: output rows 32-byte aligned, bit1: residual term 0 too (LDG/STG.256)
  uint32_t w_bytes, stage_bytes, tx_bytes, tmem_cols;
  int wstream;          // 1: the weights of a K chunk travel with its halo stage ([tap][KC/8][NS][8] at a_bytes) instead of
                        // staying resident (layers whose slice does not fit beside >= 4 stages: 256->64 stride 2)
  uint32_t a_bytes, wtap_bytes;   // wstream: offset of the weights inside a stage, bytes per tap (KC/8 * NS * 16)
  long long* dbg;       // debug timeline (CTA 0): [tile][8] clock64 stamps, or nullptr
  int skip;             // debug: bit0 no halo loads, bit1 no MMAs, bit2 no residual loads, bit3 no stores
};

constexpr int MAX_STAGES = 12;
constexpr int MAX_ACC = 8;
constexpr int MAX_CHUNKS = 32;

struct Tc5Maps { CUtensorMap m[4]; };

// 9 taps x K2 k16 steps issued back to back from per-tap offsets held in registers.  The issuing thread is
// bound by the latency of its own instruction stream (tools/umma_rate.cu: 137 cycles per MMA for a plain
// loop, 203 with a tap-table lookup), so the hot shapes run fully unrolled.
template <int K2>
__device__ __forceinline__ void issue9(uint32_t d_tmem, uint32_t a_stage, uint32_t w_chunk, const uint32_t (&alo9)[9],
                                       const uint32_t (&blo9)[9], uint32_t hiA, uint32_t hiB, uint32_t a_kstep,
                                       uint32_t b_kstep, uint32_t idesc, uint32_t acc) {
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
#pragma unroll
    for (int kc = 0; kc < K2; ++kc) {
      umma_f16(d_tmem, ((uint64_t)hiA << 32) | (a_stage + alo9[tp] + (uint32_t)kc * a_kstep),
               ((uint64_t)hiB << 32) | (w_chunk + blo9[tp] + (uint32_t)kc * b_kstep), idesc, (tp | kc) ? 1u : acc);
    }
  }
}

// WSTREAM is a template parameter so that the resident-weights kernel (every layer but the stride-2 transitions) is
// compiled exactly as before
template <bool WSTREAM>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc5_kernel(const __grid_constant__ Tc5Maps maps, const Tc5P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * MAX_STAGES + 2 * MAX_ACC];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sBias[128];       // bias of this CTA's NS output channels
  // bars: 0 weights | 1..S halo full | 1+S..2S halo empty | then acc full x NACC, acc empty x NACC
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 1, B_EMPTY = 1 + p.S, B_ACCF = 1 + 2 * p.S, B_ACCE = 1 + 2 * p.S + p.NACC;

  unsigned char* sW = smem;
  unsigned char* sH = smem + (WSTREAM ? 0u : p.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lbo_b = (uint32_t)p.NS * 16u, sbo_b = 128u;
  const int slice = blockIdx.y;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1);
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
    for (int i = 0; i < p.NACC; ++i) { mbar_init(BAR(B_ACCF + i), 1); mbar_init(BAR(B_ACCE + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.NS; i += blockDim.x) sBias[i] = p.bias[slice * p.NS + i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) pdl_launch_dependents();       // the next kernel may begin its prologue as SMs free up

  const long long first = blockIdx.x, step = gridDim.x;

  const int NMMA = p.NMMA;
  if (warp == MAX_MMA_WARPS) {
    // ===================== halo producer (TMA) =====================
    const int ptid = lane;
    // Lane w < NMMA feeds the ring of MMA warp w (tiles w, w+NMMA, ...): one 5-D TMA box per stage,
    // (8 ch, 10 px, 18 rows, KC/8 planes, 1 crop), lands in shared memory as [KC/8][18][10][8ch];
    // out-of-bounds coordinates are zero-filled, which is the convolution's zero padding.  A single
    // thread's serial instruction stream costs ~10 cycles per instruction here, so ONE producer thread
    // bounded the kernel at ~720 cycles per tile (profiles/r1_notes.md); the lanes run the same code on
    // different tiles in lockstep.
    if (ptid < NMMA) {
      pdl_wait();                                      // activations of the previous kernel
      if (ptid == 0) {
        const int nph = p.stride == 2 ? 4 : 1;
        for (int i = 0; i < nph; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
      }
      const uint32_t S2 = (uint32_t)p.S / (uint32_t)NMMA, nmma = (uint32_t)NMMA, w = (uint32_t)ptid;
      uint32_t j = 0, phase = 0;                      // ring position / phase, carried incrementally
      uint32_t it = w;                                // debug index = CTA-local tile number (nchunks == 1)
      for (int t = (int)first + (int)w * (int)step; t < (int)p.ntiles; t += (int)nmma * (int)step) {
        const int n = (int)fastdiv((uint32_t)t, p.magic_tpi);
        const int rem = t - n * (int)p.tiles_per_img;
        const int ty = (int)fastdiv((uint32_t)rem, p.magic_tx), tx = rem - ty * p.tiles_x;
        for (int c = 0; c < p.nchunks; ++c) {
          const int s = (int)(j * nmma + w);
          mbar_wait(BAR(B_EMPTY + s), phase ^ 1u);
          if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && it < 64 && c == 0) p.dbg[it * 8 + 0] = clock64();
          if (!(p.skip & 1)) {
            mbar_arrive_expect_tx(BAR(B_FULL + s), p.tx_bytes + (WSTREAM ? (uint32_t)p.ntaps * p.wtap_bytes : 0u));
            const uint32_t dst = smem_u32(sH + (size_t)s * p.stage_bytes);
            if (WSTREAM) {
              // this chunk's planes of every tap: [slice][tap][Cin/8][NS][8] -> [tap][KC/8][NS][8] inside the stage
              const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.w) +
                                          ((size_t)blockIdx.y * p.ntaps * (p.Cin >> 3) + (size_t)c * (p.KC >> 3)) * p.NS * 16u;
              for (int tp = 0; tp < p.ntaps; ++tp)
                bulk_load(dst + p.a_bytes + (uint32_t)tp * p.wtap_bytes, wsrc + (size_t)tp * (p.Cin >> 3) * p.NS * 16u,
                          p.wtap_bytes, BAR(B_FULL + s));
            }
            if (p.stride == 1) {
              tma_load_5d(dst, &maps.m[0], BAR(B_FULL + s), 0, tx * TW - p.halo, ty * TH - p.halo, c * (p.KC >> 3), n);
            } else {
              // phase (row parity, column parity): even rows/cols start at the tile origin, odd ones one
              // element earlier (the -1 taps); the box walks the input with element strides of 2
#pragma unroll
              for (int ph4 = 0; ph4 < 4; ++ph4)
                tma_load_5d(dst + p.ph_off[ph4], &maps.m[ph4], BAR(B_FULL + s), 0, 2 * tx * TW - (ph4 & 1),
                            2 * ty * TH - (ph4 >> 1), c * (p.KC >> 3), n);
            }
          } else {
            mbar_arrive(BAR(B_FULL + s));
          }
          if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && it < 64 && c == 0) p.dbg[it * 8 + 1] = clock64();
          if (++j == S2) { j = 0; phase ^= 1u; }
        }
        it += nmma;
      }
    }
  } else if (warp < MAX_MMA_WARPS) {
   if (warp < NMMA) {
    // ===================== weight bulk copy + MMA issuers =====================
    // Two issuing warps, alternating tiles: one warp sustains one UTCHMMA per ~86 cycles whatever N
    // is, two together reach the operand-fetch floor (40 / 48 / 64 cycles for N = 32 / 64 / 128;
    // tools/umma_rate.cu, profiles/r1_notes.md).
    // The whole warp runs this loop with warp-uniform values and only the tcgen05 instructions
    // are predicated on one elected lane: UTCHMMA takes its descriptors from UNIFORM registers, and
    // issuing from inside `if (lane == 0)` made the compiler wrap every MMA in a ~20-instruction
    // R2UR "waterfall" loop (~150 cycles per MMA, profiles/r1_notes.md).
    if (warp == 0 && !WSTREAM && elect_one()) {
      mbar_arrive_expect_tx(BAR(0), p.w_bytes);
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.w) + (size_t)slice * p.w_bytes;
      for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
        uint32_t nb = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
        bulk_load(smem_u32(sW + off), wsrc + off, nb, BAR(0));
      }
    }
    __syncwarp();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NS >> 3) << 17) | ((128u >> 4) << 24);
    if (!WSTREAM) mbar_wait(BAR(0), 0);
    const int kc2n = p.KC >> 4;
    const uint32_t hiB = desc_hi(sbo_b);
    const uint32_t a0 = smem_u32(sH), w0 = smem_u32(sW);
    const uint32_t b_kstep = (2u * lbo_b) >> 4;     // per k16 step, in 16-byte units
    const uint32_t b_tapstep = ((uint32_t)((WSTREAM ? p.KC : p.Cin) >> 3) * lbo_b) >> 4;
    const uint32_t b_chunkstep = WSTREAM ? 0u : ((uint32_t)(p.KC >> 3) * lbo_b) >> 4;
    const uint32_t lo_lbo_b = ((lbo_b >> 4) & 0x3FFFu) << 16;
    // fast path: 3x3 stride 1 (one halo patch: the descriptor high word and k step are the same for every tap)
    const bool fast9 = p.ntaps == 9 && p.stride == 1 && !(p.skip & 2) && (kc2n == 1 || kc2n == 2 || kc2n == 4);
    uint32_t alo9[9], blo9[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      alo9[tp] = p.tap_lo[tp < p.ntaps ? tp : 0];
      blo9[tp] = (uint32_t)tp * b_tapstep;
    }
    const uint32_t hiA0 = p.tap_hi[0], a_kstep0 = p.tap_kstep[0];
    // ring / accumulator positions and phases are carried incrementally (no divisions on this latency-bound
    // instruction stream).  Each issuing warp consumes its OWN ring of S/2 stages (slots warp, warp+2, ...):
    // with one shared ring a warp could wait for phase k+1 of a slot before phase k had completed, and an
    // mbarrier parity wait cannot tell "not yet" from "one phase ago".
    const uint32_t S2 = (uint32_t)p.S / (uint32_t)NMMA, nmma = (uint32_t)NMMA;
    uint32_t js = 0, sph = 0;                       // position / phase in this warp's stage ring
    uint32_t b = (uint32_t)warp, aph = 0;           // accumulator index (NACC >= NMMA) / phase
    uint32_t tl = (uint32_t)warp;
    for (uint32_t t = (uint32_t)first + (uint32_t)warp * (uint32_t)step; t < (uint32_t)p.ntiles;
         t += nmma * (uint32_t)step, tl += nmma) {
      mbar_wait(BAR(B_ACCE + b), aph ^ 1u);          // accumulator drained by the epilogue
      const uint32_t d_tmem = tmem_base + b * (uint32_t)p.NS;
      if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && tl < 64) p.dbg[tl * 8 + 2] = clock64();
      for (int c = 0; c < p.nchunks; ++c) {
        const int s = (int)(js * nmma + (uint32_t)warp);
        mbar_wait(BAR(B_FULL + s), sph);               // halo chunk landed
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && tl < 64) p.dbg[tl * 8 + 3] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // low descriptor words: (address >> 4) | LBO << 16; adding 16-byte offsets never carries out of
        // the 14-bit address field (shared memory is < 256 KB)
        const uint32_t a_stage = (a0 + (uint32_t)s * p.stage_bytes) >> 4;
        const uint32_t w_chunk = WSTREAM ? (((a0 + (uint32_t)s * p.stage_bytes + p.a_bytes) >> 4) | lo_lbo_b)
                                           : (((w0 >> 4) + (uint32_t)c * b_chunkstep) | lo_lbo_b);
        if (elect_one()) {
          uint32_t acc = c > 0;
          const int ntp = fast9 ? 0 : ((p.skip & 2) ? 1 : p.ntaps);
          if (fast9) {
            if (kc2n == 1) issue9<1>(d_tmem, a_stage, w_chunk, alo9, blo9, hiA0, hiB, a_kstep0, b_kstep, idesc, acc);
            else if (kc2n == 2) issue9<2>(d_tmem, a_stage, w_chunk, alo9, blo9, hiA0, hiB, a_kstep0, b_kstep, idesc, acc);
            else issue9<4>(d_tmem, a_stage, w_chunk, alo9, blo9, hiA0, hiB, a_kstep0, b_kstep, idesc, acc);
          }
          for (int tp = 0; tp < ntp; ++tp) {
            uint32_t alo = a_stage + p.tap_lo[tp];
            uint32_t blo = w_chunk + (uint32_t)tp * b_tapstep;
            const uint32_t hiA = p.tap_hi[tp], a_kstep = p.tap_kstep[tp];
            for (int kc = 0; kc < kc2n; ++kc) {
              const uint64_t ad = ((uint64_t)hiA << 32) | alo;
              const uint64_t bd = ((uint64_t)hiB << 32) | blo;
              umma_f16(d_tmem, ad, bd, idesc, acc);
              acc = 1;
              alo += a_kstep;
              blo += b_kstep;
            }
          }
          umma_commit(BAR(B_EMPTY + s));        // stage free once these MMAs retire
          if (c == p.nchunks - 1) umma_commit(BAR(B_ACCF + b));   // accumulator ready
        }
        __syncwarp();
        if (++js == S2) { js = 0; sph ^= 1u; }
      }
      if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && tl < 64) p.dbg[tl * 8 + 4] = clock64();
      b += nmma;
      if (b >= (uint32_t)p.NACC) { b -= (uint32_t)p.NACC; aph ^= 1u; }
    }
   }
  } else {
    // ===================== epilogue (warps 5 .. 20) =====================
    // Two groups of 8 warps that alternate tiles (like the MMA warps).  Inside a group: TMEM lane
    // quarter q = warp & 3 (hardware rule) and a column half.
    // These warps are what bounds the kernel once loads and MMA issue are out of the way: every
    // warp runs a serial ~6.6 cycles/instruction stream per tile (profiles/r1_notes.md), so the
    // code below is written for instruction count: magic-number tile decomposition, 32-bit element
    // offsets, 16-column blocks, packed bf16x2 ReLU, first residual term prefetched before the wait.
    const int eidx = (warp - EPI_WARP0) >> 2;   // 0..3 for each lane quarter
    const int egroup = eidx & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;              // MMA row = pixel inside the patch
    const int hy = row >> 3, wx = row & 7;
    const int ncol = p.NS >> 1;                 // columns handled by this warp
    const int cbeg = (eidx >> 1) ? ncol : 0;
    const int gch0 = slice * p.NS + cbeg;       // first global output channel of this warp
    const int nres = (p.skip & 4) ? 0 : p.nres;
    const float* biasp = sBias + cbeg;
    bf16* const outp = p.out + p.out_co;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
    const uint32_t bar_accf = opaque(BAR(B_ACCF)), bar_acce = opaque(BAR(B_ACCE));
    pdl_wait();                                        // residual terms are read, outputs written: previous kernel done
    uint32_t b = (uint32_t)egroup, aph = 0, tl = (uint32_t)egroup;      // accumulator index / phase, carried incrementally
    for (uint32_t t = (uint32_t)first + (uint32_t)egroup * (uint32_t)step; t < (uint32_t)p.ntiles;
         t += (uint32_t)NEPI * (uint32_t)step, tl += NEPI) {
      const uint32_t n = fastdiv(t, p.magic_tpi);                  // t / tiles_per_img
      const uint32_t rem = t - n * p.tiles_per_img;
      const uint32_t ty = fastdiv(rem, p.magic_tx);                // rem / tiles_x
      const uint32_t tx = rem - ty * (uint32_t)p.tiles_x;
      const int y = (int)ty * TH + hy, x = (int)tx * TW + wx;
      const bool ok = y < p.H && x < p.W;
      const uint32_t ooff = ((n * (uint32_t)p.oH + (uint32_t)(y * p.omul + p.ooy)) * (uint32_t)p.oW +
                             (uint32_t)(x * p.omul + p.oox)) * (uint32_t)p.out_cs;
      const bf16* r0p = nullptr;
      // the thread's whole residual row segment (<= 64 columns) is fetched before waiting for the
      // accumulator, so its L2 / HBM latency overlaps the MMAs instead of every 16-column step
      uint4 pre[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] = make_uint4(0, 0, 0, 0);
      if (nres > 0 && ok && gch0 < p.Cout) {
        const ResP& rr = p.res[0];
        r0p = rr.p + rr.co + gch0 +
              (((rr.bs0 ? 0u : n) * (uint32_t)rr.H + (uint32_t)(y >> rr.shift)) * (uint32_t)rr.W + (uint32_t)(x >> rr.shift)) * (uint32_t)rr.cs;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j * 8 < ncol && !(j & 1)) ldg32(r0p + j * 8, (p.v32 & 2) != 0, pre[j], pre[j | 1]);
      }
      mbar_wait(bar_accf + 8u * b, aph);
      if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x == 32 * (EPI_WARP0 + 1) || threadIdx.x == 32 * (EPI_WARP0 + 5)) && tl < 64) p.dbg[tl * 8 + 5] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tq + b * (uint32_t)p.NS;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c0 = ci * 16;
        if (c0 >= ncol) break;
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 16 >= ncol) {
          // all of this warp's TMEM reads for the tile are done: release the accumulator
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce + 8u * b);
          if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x == 32 * (EPI_WARP0 + 1) || threadIdx.x == 32 * (EPI_WARP0 + 5)) && tl < 64) p.dbg[tl * 8 + 6] = clock64();
        }
        if (!ok || gch0 + c0 >= p.Cout) continue;        // padded output channels are never stored
        float f[16];
        const float4* bp = reinterpret_cast<const float4*>(biasp + c0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = bp[j4];
          f[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + bb.x;
          f[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + bb.y;
          f[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + bb.z;
          f[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + bb.w;
        }
        if (nres > 0) {
          if (c0 < 64) { add_res8(f, pre[(c0 >> 3) & 7]); add_res8(f + 8, pre[((c0 >> 3) + 1) & 7]); }
          else {
            add_res8(f, __ldg(reinterpret_cast<const uint4*>(r0p + c0)));
            add_res8(f + 8, __ldg(reinterpret_cast<const uint4*>(r0p + c0) + 1));
          }
          for (int qi = 1; qi < nres; ++qi) {
            const ResP& rr = p.res[qi];
            const bf16* rp = rr.p + rr.co + gch0 + c0 +
                (((rr.bs0 ? 0u : n) * (uint32_t)rr.H + (uint32_t)(y >> rr.shift)) * (uint32_t)rr.W + (uint32_t)(x >> rr.shift)) * (uint32_t)rr.cs;
            add_res8(f, __ldg(reinterpret_cast<const uint4*>(rp)));
            add_res8(f + 8, __ldg(reinterpret_cast<const uint4*>(rp) + 1));
          }
        }
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          h0[k2] = __floats2bfloat162_rn(f[2 * k2], f[2 * k2 + 1]);
          h1[k2] = __floats2bfloat162_rn(f[8 + 2 * k2], f[8 + 2 * k2 + 1]);
        }
        if (p.relu) {
          const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) { h0[k2] = __hmax2(h0[k2], z); h1[k2] = __hmax2(h1[k2], z); }
        }
        if (!(p.skip & 8)) {
          const uint32_t gc = (uint32_t)(gch0 + c0);
          uint32_t off = ooff + gc;
          if (p.psC) {                                       // pixel shuffle: column block -> output phase
            const uint32_t ph = gc / (uint32_t)p.psC;
            off = ooff + ((ph >> 1) * (uint32_t)p.oW + (ph & 1)) * (uint32_t)p.out_cs + (gc - ph * (uint32_t)p.psC);
          }
          stg32(outp + off, (p.v32 & 1) != 0, o0, o1);
        }
      }
      if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x == 32 * (EPI_WARP0 + 1) || threadIdx.x == 32 * (EPI_WARP0 + 5)) && tl < 64) p.dbg[tl * 8 + 7] = clock64();
      b += (uint32_t)NEPI;
      if (b >= (uint32_t)p.NACC) { b -= (uint32_t)p.NACC; aph ^= 1u; }
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// box = (8 ch, bw px, bh rows, KC/8 planes, 1 crop) over the NHWC activation viewed as (8, W, H, C/8, N),
// walked with element stride `es` along W and H (es = 2: every other pixel -- one phase of a stride-2 conv)
int make_patch_map(const ConvP& p, int bw, int bh, int KC, int es, CUtensorMap* m) {
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t esz = 2;
  cuuint64_t dims[5] = {8, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)(p.Cin / 8), (cuuint64_t)p.N};
  cuuint64_t strides[4] = {(cuuint64_t)p.in_cs * esz, (cuuint64_t)p.Win * p.in_cs * esz, 16,
                           (cuuint64_t)p.Hin * p.Win * p.in_cs * esz};
  // with an element stride the box extent counts SOURCE elements: bw loaded pixels span bw * es of them
  cuuint32_t box[5] = {8, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)(KC / 8), 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(p.in + p.in_co), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (box %d x %d, stride %d)", (int)r, bw, bh, es);
  return RSG_OK;
}

// stride-2 phase patches: phase = 2 * (odd rows) + (odd cols); odd phases carry one extra leading row / column
// (the -1 taps)
inline uint32_t desc_hi_host(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
inline int ph_rows(int ph) { return TH + (ph >> 1); }
inline int ph_cols(int ph) { return TW + (ph & 1); }
inline uint32_t ph_bytes(int ph, int KC) { return ((uint32_t)(KC / 8) * ph_rows(ph) * ph_cols(ph) * 16u + 127u) & ~127u; }

inline int stage_bytes_of(int mode, int kc) {
  if (mode == 2) return (int)(ph_bytes(0, kc) + ph_bytes(1, kc) + ph_bytes(2, kc) + ph_bytes(3, kc));
  return (TH + 2 * mode) * (TW + 2 * mode) * kc * 2;
}

}  // namespace

// Shape -> (NS, KC, S): shared with the host-side packer through rsg_conv_tc5_config().
// mode: 0 = 1x1 stride 1, 1 = taps in the 3x3 neighbourhood, stride 1 (halo patch), 2 = 3x3 neighbourhood,
// stride 2 (four phase patches).
// *stream = 1: the slice's weights do not fit beside >= 4 stride-2 stages, so every stage carries the weights of its own
// K chunk (same w_tc5 packing; the producer bulk-copies ntaps pieces per stage).
static int tc5_pick(int Cin, int CoutPad, int ntaps, int mode, int* NS, int* KC, int* S, int* stream) {
  *stream = 0;
  if (Cin % 16 != 0 || CoutPad % 32 != 0 || ntaps < 1 || ntaps > 9 || mode < 0 || mode > 2) return 0;
  const int budget = 196 * 1024;
  const int cands[4] = {128, 96, 64, 32};
  // prefer the widest slice that still leaves >= 4 stages; among the K chunkings the largest with >= 4 stages
  for (int min_s = 4; min_s >= 2; min_s -= 2) {
    if (min_s < 4 && mode == 2 && !rsg_dbg_env("RSG_TC5_NO_WSTREAM")) {
      // stride 2 with fewer than four resident-weight stages measured 2x slower than the generic kernel: stream the
      // weights with the halo stages instead (widest slice, largest chunk with >= 4 stages)
      for (int i = 0; i < 4; ++i) {
        const int ns = cands[i];
        if (ns > CoutPad || CoutPad % ns != 0) continue;
        for (int kc = 32; kc >= 16; kc -= 16) {
          if (Cin % kc != 0 || Cin / kc > MAX_CHUNKS) continue;
          const long long stage = stage_bytes_of(mode, kc) + (long long)ntaps * (kc / 8) * ns * 16;
          if (4 * stage > budget) continue;
          int s = (int)(budget / stage) & ~1;
          if (s > MAX_STAGES) s = MAX_STAGES;
          *NS = ns; *KC = kc; *S = s; *stream = 1;
          return 1;
        }
      }
    }
    for (int i = 0; i < 4; ++i) {
      const int ns = cands[i];
      if (ns > CoutPad || CoutPad % ns != 0) continue;
      const long long wb = (long long)ntaps * Cin * ns * 2;
      for (int kc = 64; kc >= 16; kc -= 16) {
        if (Cin % kc != 0 || Cin / kc > MAX_CHUNKS) continue;
        { const char* e = rsg_dbg_env("RSG_TC5_KC"); if (e && atoi(e) != kc && Cin % atoi(e) == 0) continue; }
        const int stage = stage_bytes_of(mode, kc);
        if (wb + (long long)min_s * stage > budget) continue;
        int s = MAX_STAGES;                     // even: the MMA warps own equal parts of the ring
        { const char* e = rsg_dbg_env("RSG_TC5_S"); if (e && atoi(e) >= 2 && atoi(e) <= MAX_STAGES) s = atoi(e) & ~1; }
        while (s > 2 && wb + (long long)s * stage > budget) s -= 2;
        *NS = ns; *KC = kc; *S = s;
        return 1;
      }
    }
  }
  return 0;
}

extern "C" int rsg_conv_tc5_config(int Cin, int CoutPad, int ntaps, int mode, int* NS, int* KC, int* S) {
  int stream = 0;
  return tc5_pick(Cin, CoutPad, ntaps, mode, NS, KC, S, &stream);
}

int conv_tc5_launch(const ConvP& p, cudaStream_t s, int* handled) {
  *handled = 0;
  static const bool disabled = rsg_dbg_env("RSG_DISABLE_TC5") != nullptr;   // A/B switch for debugging
  if (disabled) return RSG_OK;
  if (!p.w_tc5 || !p.out || p.out_f32) return RSG_OK;
  if (p.stride != 1 && p.stride != 2) return RSG_OK;
  if (p.stride == 2 && rsg_dbg_env("RSG_TC5_NO_S2")) return RSG_OK;
  if (p.omul < 1 || p.oH < p.Hout * p.omul || p.oW < p.Wout * p.omul) return RSG_OK;
  if (p.Hout != (p.Hin - 1) / p.stride + 1 || p.Wout != (p.Win - 1) / p.stride + 1) return RSG_OK;
  if (p.omul != 1 && p.nres != 0) return RSG_OK;
  if (p.psC && (p.psC % 16 != 0 || p.Cout != 4 * p.psC || p.omul != 2 || p.ooy != 0 || p.oox != 0)) return RSG_OK;
  if (p.Cout % 16 != 0 || p.in_cs % 8 != 0 || p.in_co % 8 != 0 || p.out_cs % 8 != 0 || p.out_co % 8 != 0) return RSG_OK;
  int halo = 0;
  for (int t = 0; t < p.ntaps && t < 16; ++t) {
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return RSG_OK;
    if (p.dy[t] != 0 || p.dx[t] != 0) halo = 1;
  }
  const int mode = p.stride == 2 ? 2 : halo;
  for (int q = 0; q < p.nres; ++q)
    if (p.res[q].cs % 8 != 0 || p.res[q].co % 8 != 0) return RSG_OK;
  int NS, KC, S, wstream = 0;
  if (!tc5_pick(p.Cin, p.CoutPad, p.ntaps, mode, &NS, &KC, &S, &wstream)) return RSG_OK;
  // a stride-2 layer whose weights leave room for only two phase-patch stages (256->64: 147 KB per 32-channel
  // slice) measured 2x slower than the generic kernel (tc5_pick streams the weights instead where it can)
  if (p.stride == 2 && S < 4 && !p.force) return RSG_OK;
  if (p.M == 0) { *handled = 1; return RSG_OK; }

  Tc5P k;
  memset(&k, 0, sizeof(k));
  k.w = p.w_tc5; k.bias = p.bias; k.Cin = p.Cin; k.NS = NS; k.Cout = p.Cout;
  k.KC = KC; k.nchunks = p.Cin / KC; k.S = S;
  k.EW = 8;
  k.ntaps = p.ntaps;
  k.stride = p.stride;
  if (p.stride == 1) {
    const uint32_t hw = TW + 2 * halo, hh = TH + 2 * halo, lbo16 = hw * hh;       // 16-byte units
    for (int t = 0; t < p.ntaps; ++t) {
      k.tap_lo[t] = (uint32_t)((halo + p.dy[t]) * (int)hw + (halo + p.dx[t])) | (lbo16 << 16);
      k.tap_hi[t] = desc_hi_host(hw * 16u);
      k.tap_kstep[t] = 2u * lbo16;
    }
    k.stage_bytes = hw * hh * (uint32_t)KC * 2u;
    k.tx_bytes = k.stage_bytes;
  } else {
    uint32_t off = 0;
    k.tx_bytes = 0;
    for (int ph = 0; ph < 4; ++ph) {
      k.ph_off[ph] = off;
      off += ph_bytes(ph, KC);
      k.tx_bytes += (uint32_t)(KC / 8) * ph_rows(ph) * ph_cols(ph) * 16u;
    }
    k.stage_bytes = off;
    for (int t = 0; t < p.ntaps; ++t) {
      // input pixel (2y+dy, 2x+dx): even offsets live in the even phase at patch index y, -1 in the odd
      // phase at patch index y (it starts one element earlier), +1 in the odd phase at index y+1
      const int ph = (p.dy[t] != 0 ? 2 : 0) + (p.dx[t] != 0 ? 1 : 0);
      const uint32_t cols = ph_cols(ph), lbo16 = (uint32_t)ph_rows(ph) * cols;
      k.tap_lo[t] = ((k.ph_off[ph] >> 4) + (uint32_t)((p.dy[t] == 1) * (int)cols + (p.dx[t] == 1))) | (lbo16 << 16);
      k.tap_hi[t] = desc_hi_host(cols * 16u);
      k.tap_kstep[t] = 2u * lbo16;
    }
  }
  k.halo = halo; k.H = p.Hout; k.W = p.Wout;
  k.oH = p.oH; k.oW = p.oW; k.omul = p.omul; k.ooy = p.ooy; k.oox = p.oox; k.psC = p.psC;
  k.in = p.in; k.in_cs = p.in_cs; k.in_co = p.in_co;
  k.tiles_x = (p.Wout + TW - 1) / TW; k.tiles_y = (p.Hout + TH - 1) / TH;
  k.ntiles = (long long)k.tiles_x * k.tiles_y * p.N;
  if (k.ntiles >= (1ll << 31)) return RSG_OK;
  k.tiles_per_img = (uint32_t)(k.tiles_x * k.tiles_y);
  // magic = ceil(2^32 / d): __umulhi(t, magic) == t / d exactly while t * d < 2^32 (d == 1 needs no division)
  if ((unsigned long long)k.ntiles * k.tiles_per_img >= (1ull << 32)) return RSG_OK;
  k.magic_tpi = k.tiles_per_img > 1 ? (uint32_t)(((1ull << 32) + k.tiles_per_img - 1) / k.tiles_per_img) : 0u;
  k.magic_tx = k.tiles_x > 1 ? (uint32_t)(((1ull << 32) + k.tiles_x - 1) / k.tiles_x) : 0u;
  // maps much smaller than the 16x8 patch waste most MMA rows (8x6: 37%): generic kernel instead
  if (!p.psC && !p.force && (double)p.Hout * p.Wout < 0.6 * 128.0 * k.tiles_x * k.tiles_y) return RSG_OK;
  k.out = p.out; k.out_cs = p.out_cs; k.out_co = p.out_co;
  k.nres = p.nres;
  for (int q = 0; q < p.nres; ++q) k.res[q] = p.res[q];
  k.relu = p.relu;
  {
    auto al32 = [](const void* ptr, int cs, int co) { return ((uintptr_t)ptr % 32 == 0) && cs % 16 == 0 && co % 16 == 0; };
    k.v32 = (al32(p.out, p.out_cs, p.out_co) && (!p.psC || p.psC % 16 == 0) ? 1 : 0) |
            (p.nres > 0 && al32(p.res[0].p, p.res[0].cs, p.res[0].co) ? 2 : 0);
    if (rsg_dbg_env("RSG_NO_V32")) k.v32 = 0;
  }
  { const char* e = rsg_dbg_env("RSG_TC5_SKIP"); k.skip = e ? atoi(e) : 0; }
  static long long* dbg_buf = nullptr;
  if (rsg_dbg_env("RSG_TC5_TIMELINE")) {
    if (!dbg_buf) { cudaMalloc(&dbg_buf, 64 * 8 * sizeof(long long)); }
    cudaMemsetAsync(dbg_buf, 0, 64 * 8 * sizeof(long long), s);
    k.dbg = dbg_buf;
  }
  k.w_bytes = (uint32_t)p.ntaps * p.Cin * NS * 2;
  k.wstream = wstream;
  if (wstream) {
    k.a_bytes = k.stage_bytes;                             // multiple of 128 (phase patches are padded to 128)
    k.wtap_bytes = (uint32_t)(KC / 8) * NS * 16u;
    k.stage_bytes += (uint32_t)p.ntaps * k.wtap_bytes;
  }
  const size_t smem = (wstream ? 0 : (size_t)k.w_bytes) + (size_t)S * k.stage_bytes;
  // accumulator ring: MMA completion + barrier wake-up latency is ~1-2k cycles per tile hand-off, so
  // two accumulators leave the tensor pipe idle (profiles/r1_notes.md); use up to 8
  const uint32_t col_budget = 512u;                       // one CTA per SM owns all of TMEM
  int nacc = (int)(col_budget / (uint32_t)NS);
  if (nacc > MAX_ACC) nacc = MAX_ACC;
  // thin layers are bound by the per-warp MMA issue rate (~110 cycles per MMA): four issuing warps when
  // the rings can be split four ways
  // three issuers when the ring splits three ways (keeps the tensor pipe fed while the others sit in their
  // per-tile barrier round trips), else two
  int nmma = S >= 6 ? 3 : 2;
  { const char* e = rsg_dbg_env("RSG_TC5_NMMA"); if (e && (atoi(e) == 2 || atoi(e) == 3)) nmma = atoi(e); }
  if (S < nmma) nmma = 2;
  // Tile tl is filled by MMA warp tl % NMMA into accumulator tl % NACC and drained by epilogue group tl % NEPI.  NACC
  // must be a multiple of BOTH, so that an accumulator always belongs to one issuer and one group, in order: with
  // NACC = 3 and two groups (NS = 128: four accumulators fit, three issuers) group 1 could reach its parity-1 wait for
  // tile 5 on accumulator 2 before tile 2 had even been committed -- an mbarrier parity wait cannot tell "not yet" from
  // "one phase ago" -- read garbage, release the accumulator early and shift the phase accounting until the CTA hung
  // (seen once in 20 steps under the overlapped H2D pipeline).  Three issuers therefore need six accumulators.
  if (nmma == 3 && nacc < 6) nmma = 2;
  S = S / nmma * nmma;
  k.S = S;
  k.NMMA = nmma;
  nacc = nacc / (nmma * NEPI) * (nmma * NEPI);           // nmma = 2: multiples of 2 (NEPI = 2), nmma = 3: 6
  if (nmma == 2) nacc = (int)(col_budget / (uint32_t)NS) >= 8 ? 8 : ((int)(col_budget / (uint32_t)NS) >= 4 ? 4 : 2);
  { const char* e = rsg_dbg_env("RSG_TC5_NACC"); if (e && atoi(e) >= 2 && atoi(e) % (nmma == 3 ? 6 : 2) == 0 && atoi(e) * NS <= 512 && atoi(e) <= MAX_ACC) nacc = atoi(e); }
  k.NACC = nacc;
  uint32_t cols = 32;
  while (cols < (uint32_t)(nacc * NS)) cols <<= 1;
  k.tmem_cols = cols;
  const int nslices = p.CoutPad / NS;
  const int threads = NTHREADS;

  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RSG_CUDA(cudaFuncSetAttribute(conv_tc5_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    RSG_CUDA(cudaFuncSetAttribute(conv_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    RSG_CUDA(cudaFuncSetAttribute(conv_tc5_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_once.done();
  }
  const int occ = 1;
  static const bool dbg = rsg_dbg_env("RSG_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[tc5] Cin=%d Cout=%d taps=%d NS=%d KC=%d S=%d NACC=%d smem=%zu tiles=%lld\n", p.Cin, p.CoutPad, p.ntaps, NS, KC, S, k.NACC, smem, k.ntiles);
  long long gx = ((long long)rsg_num_sms() * occ + nslices - 1) / nslices;
  if (gx > k.ntiles) gx = k.ntiles;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)nslices);
  Tc5Maps maps;
  memset(&maps, 0, sizeof(maps));
  if (p.stride == 1) {
    int rc = make_patch_map(p, TW + 2 * halo, TH + 2 * halo, KC, 1, &maps.m[0]);
    if (rc) return rc;
  } else {
    for (int ph = 0; ph < 4; ++ph) {
      int rc = make_patch_map(p, ph_cols(ph), ph_rows(ph), KC, 2, &maps.m[ph]);
      if (rc) return rc;
    }
  }
  if (wstream) RSG_CUDA(launch_pdl(conv_tc5_kernel<true>, grid, dim3(threads), smem, s, maps, k));
  else RSG_CUDA(launch_pdl(conv_tc5_kernel<false>, grid, dim3(threads), smem, s, maps, k));
  if (k.dbg) {
    static int dumped = 0;
    if (dumped++ == 3) {
      long long h[64 * 8];
      cudaStreamSynchronize(s);
      cudaMemcpy(h, k.dbg, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = h[0];
      fprintf(stderr, "[tc5 timeline] tile: prod_empty_ok prod_issued | mma_acc_ok mma_full_ok mma_committed | epi_accf_ok epi_ld_done epi_done (cycles since first stamp)\n");
      for (int i = 0; i < 24; ++i) {
        fprintf(stderr, "%2d:", i);
        for (int j = 0; j < 8; ++j) fprintf(stderr, " %8lld", h[i * 8 + j] ? h[i * 8 + j] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  *handled = 1;
  return RSG_OK;
}
