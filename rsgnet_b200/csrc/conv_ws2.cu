// conv_ws.cu's flat implicit GEMM on CTA PAIRS (tcgen05.mma.cta_group::2): the GEMM-shaped layers of stage 3/4
// (128->128 @16x12, 256->256 @8x6) with the weight stream SPLIT over the two SMs of a cluster.
//
// Why (measured on the 1-CTA kernel, profiles/r2_notes.md): with N = 128 an M = 128 MMA needs 64 cycles of math and
// 64 cycles of shared-memory operand fetch (4 KB of A + 4 KB of B), and the TMA writes of the streamed weights
// (44 KB per 16-channel chunk) take another 15 % of the same shared-memory bandwidth: 84 cycles per MMA; all 512 TMEM
// columns hold ONE supertile, so the epilogue of supertile i and the MMAs of supertile i+1 cannot overlap either.
// A CTA pair executes one M = 256 MMA on two SMs: each CTA supplies its own 128 pixel rows of A and HALF of the
// weights (64 of the 128 output channels; the hardware exchanges the halves), so per SM the operand fetch drops to
// 4 KB + 2 KB = 48 cycles (math-bound), the weight stream per SM halves, and a supertile needs only 2 of the 4
// accumulators: the other two take the next supertile while the 16 epilogue warps drain this one.
//
// Geometry: as in conv_ws.cu, nimg whole images land as a pitch-(W+1) flat pixel array per 8-channel plane (first row
// and column = the zero padding, TMA out-of-bounds fill) and an MMA row block is 128 consecutive flat pixels; a CTA's
// supertile is T = 2 such blocks (256 flat pixels: one 16x12 image, four 8x6 images).  CTA r of pair-unit u owns the
// images [(2u + r) nimg, (2u + r + 1) nimg).
//
// Roles per CTA (608 threads): warps 0-15 = epilogue of its own 128-lane accumulators, warp 16 = TMA producer (its own
// activation planes + its half of the weights, into its own shared memory).  Warps 17-18 (the HIGHEST warp ids: the
// warp scheduler prefers higher ids among eligible warps, and the issuing threads are the latency-critical ones): in
// the LEADER CTA (cluster rank 0) they issue the MMAs for both SMs (tile = warp - 17, two instruction streams); in the
// peer CTA warp 17 forwards
// "my stage is full" to the leader (a remote mbarrier arrive), since non-tensor bulk copies can only signal a barrier
// of the CTA they write to.  tcgen05.commit multicasts "stage free" / "accumulator full" to both CTAs; the epilogue
// warps of both CTAs release an accumulator with arrives on the leader's barrier (remote for the peer).
//
// Reference ops subsumed: the same as conv_ws.cu (Conv2d 3x3|1x1 s1 + BatchNorm2d(eval) [+ residual] [+ ReLU] of the
// low-resolution branches, pose_rsgnet.py:25-54 inside HighResolutionModule :108-272).
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int W2_THREADS = 608;               // 19 warps
constexpr int W2_EPI_WARP0 = 0;                // warps 0-15: epilogue
constexpr int W2_PROD_WARP = 16;               // TMA producer + TMEM allocation
constexpr int W2_MMA_WARP0 = 17;               // warps 17, 18: MMA issuers (leader) / stage forwarder (peer)
constexpr int W2_MAX_S = 8;
constexpr int W2_NS = 128;                    // output channels per pair (64 per CTA's half of B)
constexpr int W2_SMEM_BUDGET = 225 * 1024;

struct Ws2P {
  const bf16* w;        // [slice][half][chunk][tap][KC/8][64][8]: the weights of a chunk are ONE contiguous block
  const float* bias;
  int Cin, Cout;
  int KC, nchunks, S, nimg;
  int ntaps;
  int tapoff[9];        // (1+dy)*P + (1+dx), pixels (= 16-byte units)
  int H, W, P, pitch;   // P = W + 1, pitch = (H + 1) * P pixels per image
  uint32_t magic_pitch, magic_P;
  int N, nunits;        // pair-units: 2 * nimg images each
  bf16* out;
  int out_cs, out_co;
  int nres;
  ResP res[4];
  int relu;
  int v32;
  uint32_t plane_bytes, a_bytes, b_off, b_tap_bytes, stage_bytes;
  int ctawait;          // debug builds: 1 = plain (cta-scope) waits on the barriers the partner CTA signals
  int skip;             // debug builds: bit0 no loads, bit1 one tap, bit2 no residual loads / stores
  long long* dbg;       // debug builds: globaltimer stamps of pair 0 [2 CTAs][64]
};

#define W2_STAMP(i) do { if (p.dbg && pair == 0 && slice == 0) p.dbg[rank * 128 + (i)] = gtime2(); } while (0)
__device__ __forceinline__ long long gtime2() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of THIS CTA -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from the partner CTA (release.cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("[conv_ws2] cluster mbarrier wait timed out: smem 0x%x parity %u block (%d,%d) thread %d\n", bar, parity,
         (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
  __trap();
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// arrive on the barrier at this shared::cta offset in BOTH CTAs of the pair once all MMAs issued so far have retired
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W2_THREADS, 1)
conv_ws2_kernel(const __grid_constant__ CUtensorMap in_map, const Ws2P p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // barriers: FULL[S] (own loads) | EMPTY[S] | PFULL[S] (leader: the peer's stage is full) | ACCF[4] | ACCE[4] (leader)
  __shared__ __align__(8) uint64_t bars[3 * W2_MAX_S + 8];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sBias[W2_NS];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = p.S, B_PFULL = 2 * p.S, B_ACCF = 3 * p.S, B_ACCE = 3 * p.S + 4;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int slice = blockIdx.y;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 2); mbar_init(BAR(B_PFULL + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(B_ACCF + i), 1); mbar_init(BAR(B_ACCE + i), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < W2_NS; i += blockDim.x) sBias[i] = p.bias[slice * W2_NS + i];
  {   // the gap between the activation planes and the weights is read by the bottom-right taps: keep it zero
    const uint32_t gap16 = (p.b_off - p.a_bytes) >> 4;
    for (uint32_t i = threadIdx.x; i < gap16 * (uint32_t)p.S; i += blockDim.x) {
      const uint32_t s = i / gap16, j = i - s * gap16;
      *reinterpret_cast<uint4*>(sgen + (size_t)s * p.stage_bytes + p.a_bytes + 16u * j) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W2_PROD_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) { pdl_launch_dependents(); W2_STAMP(0); }

  if (warp == W2_PROD_WARP) {
    // ===================== producer: own activation planes (one TMA box) + own half of the weights =====================
    // Two copies per stage, issued by two lanes: the activation box (tensor TMA) and ONE bulk copy with the weights of
    // all taps of the chunk -- the weights are packed chunk-major for exactly this reason.  A TMA operation costs
    // ~100 cycles of engine time whatever its size (tools/bulk_rate.cu: nine 2 KB copies 957 cycles, one 18 KB copy
    // 542), and one thread needs ~130 cycles to issue one; ten copies per chunk ran the whole kernel at the
    // producer's rate.  A copy that completes before lane 0 has armed the barrier only drives the transaction count
    // negative for a moment (the pending arrival keeps the phase open).
    pdl_wait();
    if (lane == 0) {
      W2_STAMP(1);
      asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
    }
    const uint32_t w_chunk_bytes = (uint32_t)p.ntaps * p.b_tap_bytes;
    const unsigned char* wsl = reinterpret_cast<const unsigned char*>(p.w) +
                               ((size_t)slice * 2u + rank) * (size_t)p.nchunks * w_chunk_bytes;
    const uint32_t tx_bytes = p.a_bytes + w_chunk_bytes;
    uint32_t s = 0, sph = 0;
    int ui = 0;
    for (int u = pair; u < p.nunits; u += npairs, ++ui) {
      const int img0 = (2 * u + (int)rank) * p.nimg;
      for (int c = 0; c < p.nchunks; ++c, s = (s + 1 == (uint32_t)p.S ? 0u : s + 1), sph ^= (s == 0)) {
        mbar_wait(BAR(B_EMPTY + s), sph ^ 1u);
        if (ui == 2 && c < 8 && lane == 0) W2_STAMP(64 + 2 * c);
        const uint32_t dst = sbase + s * p.stage_bytes;
        if (p.skip & 1) { if (lane == 0) mbar_arrive(BAR(B_FULL + s)); continue; }
        if (lane == 0) {
          mbar_arrive_expect_tx(BAR(B_FULL + s), tx_bytes);
          tma_load_5d(dst, &in_map, BAR(B_FULL + s), 0, -1, -1, img0, c * (p.KC >> 3));
        } else if (lane == 1) {
          bulk_load(dst + p.b_off, wsl + (size_t)c * w_chunk_bytes, w_chunk_bytes, BAR(B_FULL + s));
        }
        if (ui == 2 && c < 8 && lane == 0) W2_STAMP(65 + 2 * c);
      }
    }
  } else if (warp >= W2_MMA_WARP0) {
    if (rank != 0) {
      // ===================== peer CTA: forward "stage full" to the leader =====================
      if (warp == W2_MMA_WARP0 && lane == 0) {
        const uint32_t pfull0 = mapa(BAR(B_PFULL), 0);
        uint32_t s = 0, sph = 0;
        for (int u = pair; u < p.nunits; u += npairs)
          for (int c = 0; c < p.nchunks; ++c, s = (s + 1 == (uint32_t)p.S ? 0u : s + 1), sph ^= (s == 0)) {
            mbar_wait(BAR(B_FULL + s), sph);
            // TMA writes (async proxy) completed on this barrier; the remote arrive (release.cluster) orders them
            // before the leader's acquire-wait
            mbar_arrive_remote(pfull0 + 8u * s);
          }
      }
    } else {
      // ===================== leader CTA: MMA issuers (tile = warp) =====================
      // M = 256 (128 rows per CTA), N = 128, bf16 x bf16 -> f32, both operands K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(W2_NS >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t hiA = desc_hi(128u), hiB = desc_hi(128u);
      const uint32_t lo_lbo_a = ((p.plane_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t lo_lbo_b = ((64u * 16u >> 4) & 0x3FFFu) << 16;             // B half: 64 rows x 16 B per 8-channel plane
      const uint32_t a_kstep = (2u * p.plane_bytes) >> 4, b_kstep = (2u * 64u * 16u) >> 4;
      const uint32_t b_tapstep = p.b_tap_bytes >> 4;
      const int kc2n = p.KC >> 4;
      const bool fast9 = p.ntaps == 9 && kc2n == 1 && !(p.skip & 2);
      uint32_t aoff[9], boff[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        aoff[tp] = (uint32_t)p.tapoff[tp < p.ntaps ? tp : 0];
        boff[tp] = (uint32_t)tp * b_tapstep;
      }
      const uint32_t t = (uint32_t)(warp - W2_MMA_WARP0);
      uint32_t st = 0, s = 0, sph = 0;
      for (int u = pair; u < p.nunits; u += npairs, ++st) {
        const uint32_t acc_i = 2u * (st & 1u) + t;                               // accumulator buffer (st & 1), tile t
        const uint32_t d_tmem = tmem_base + acc_i * (uint32_t)W2_NS;
        for (int c = 0; c < p.nchunks; ++c, s = (s + 1 == (uint32_t)p.S ? 0u : s + 1), sph ^= (s == 0)) {
          mbar_wait(BAR(B_FULL + s), sph);
          if (st == 2 && c < 8 && warp == W2_MMA_WARP0 && lane == 0) W2_STAMP(80 + 3 * c);
          if (c == 0 && warp == W2_MMA_WARP0 && lane == 0 && st < 6) W2_STAMP(8 + 4 * st);
          if (p.ctawait) mbar_wait(BAR(B_PFULL + s), sph); else mbar_wait_cluster(BAR(B_PFULL + s), sph);
          if (st == 2 && c < 8 && warp == W2_MMA_WARP0 && lane == 0) W2_STAMP(81 + 3 * c);
          if (c == 0 && warp == W2_MMA_WARP0 && lane == 0 && st < 6) W2_STAMP(9 + 4 * st);
          if (c == 0) { if (p.ctawait) mbar_wait(BAR(B_ACCE + acc_i), ((st >> 1) & 1u) ^ 1u); else mbar_wait_cluster(BAR(B_ACCE + acc_i), ((st >> 1) & 1u) ^ 1u); }      // both CTAs drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (c == 0 && warp == W2_MMA_WARP0 && lane == 0 && st < 6) W2_STAMP(10 + 4 * st);
          const uint32_t stage16 = (sbase + s * p.stage_bytes) >> 4;
          const uint32_t b_stage = (stage16 + (p.b_off >> 4)) | lo_lbo_b;
          const uint32_t a_tile = (stage16 + 128u * t) | lo_lbo_a;
          if (elect_one()) {
            uint32_t acc = c > 0;
            if (fast9) {
#pragma unroll
              for (int tp = 0; tp < 9; ++tp)
                umma2_f16(d_tmem, ((uint64_t)hiA << 32) | (a_tile + aoff[tp]), ((uint64_t)hiB << 32) | (b_stage + boff[tp]),
                          idesc, tp == 0 ? acc : 1u);
            } else {
              const int ntp = (p.skip & 2) ? 1 : p.ntaps;
              for (int tp = 0; tp < ntp; ++tp) {
                uint32_t alo = a_tile + (uint32_t)p.tapoff[tp];
                uint32_t blo = b_stage + (uint32_t)tp * b_tapstep;
                for (int kc = 0; kc < kc2n; ++kc) {
                  umma2_f16(d_tmem, ((uint64_t)hiA << 32) | alo, ((uint64_t)hiB << 32) | blo, idesc, acc);
                  acc = 1;
                  alo += a_kstep;
                  blo += b_kstep;
                }
              }
            }
            if (c == p.nchunks - 1) umma2_commit(BAR(B_ACCF + acc_i));           // accumulator complete, in both CTAs
            umma2_commit(BAR(B_EMPTY + s));                                       // stage free (two issuers -> count 2), both CTAs
          }
          __syncwarp();
          if (st == 2 && c < 8 && warp == W2_MMA_WARP0 && lane == 0) W2_STAMP(82 + 3 * c);
          if (c == p.nchunks - 1 && warp == W2_MMA_WARP0 && lane == 0 && st < 6) W2_STAMP(11 + 4 * st);
        }
      }
    }
  } else {
    // ===================== epilogue (warps 0 .. 15): this CTA's 128 lanes of both tiles =====================
    // The unified L1 / shared-memory data pipe is the contended resource of this kernel (MMA operand fetch 96 B/clk +
    // TMA writes): a row-per-lane epilogue (every lane its own 128-byte line: 32 wavefronts per instruction) took
    // ~45 % of the pipe's cycles and slowed the MMAs down (measured: loads and stores cost ADDITIVE time).  The
    // accumulator is therefore read as 16x256b blocks: lanes 4r..4r+3 hold pixel rows r and r+8 of the block and --
    // with the output channels permuted at pack time (column 8g + 2j + e = channel 16j + 2g + e inside each group of
    // 64) -- lane 4r+j owns channels [16j, 16j+16): one load / store instruction touches 8 full lines.
    const int eidx = warp >> 2;                     // 0..3
    const int t = eidx & 1;                         // tile of the supertile
    const int q = warp & 3;                         // TMEM lane quarter
    const int cbeg = (eidx >> 1) * 64;              // 64-column half of the 128 output channels
    const int j4 = lane & 3, r8 = lane >> 2;
    const int gch = slice * W2_NS + cbeg + 16 * j4; // this lane's 16 output channels
    const bool ch_ok = gch < p.Cout;
    const float4* bp = reinterpret_cast<const float4*>(sBias + cbeg + 16 * j4);
    bf16* const outp = p.out + p.out_co + gch;
    const uint32_t bar_accf = opaque(BAR(B_ACCF));
    const uint32_t bar_acce_leader = mapa(BAR(B_ACCE), 0);      // shared::cluster address (own CTA for the leader)
    const int nres = (p.skip & 4) ? 0 : p.nres;
    pdl_wait();
    uint32_t st = 0;
    for (int u = pair; u < p.nunits; u += npairs, ++st) {
      const uint32_t acc_i = 2u * (st & 1u) + (uint32_t)t;
      // the lane's four pixels: MMA rows q*32 + 8k + r8 (k = 0..3) of tile t
      uint32_t ooff[4], nn[4], yy[4], xx[4];
      bool ok[4];
      uint4 pre[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t o = (uint32_t)t * 128u + (uint32_t)(q * 32 + 8 * k + r8);
        const uint32_t i = fastdiv(o, p.magic_pitch);
        const uint32_t r = o - i * (uint32_t)p.pitch;
        yy[k] = fastdiv(r, p.magic_P);
        xx[k] = r - yy[k] * (uint32_t)p.P;
        nn[k] = (uint32_t)(2 * u + (int)rank) * (uint32_t)p.nimg + i;
        ok[k] = ch_ok && i < (uint32_t)p.nimg && nn[k] < (uint32_t)p.N && yy[k] < (uint32_t)p.H && xx[k] < (uint32_t)p.W;
        ooff[k] = ((nn[k] * (uint32_t)p.H + yy[k]) * (uint32_t)p.W + xx[k]) * (uint32_t)p.out_cs;
        pre[2 * k] = make_uint4(0, 0, 0, 0);
        pre[2 * k + 1] = make_uint4(0, 0, 0, 0);
        if (nres > 0 && ok[k]) {                     // residual fetched before the accumulator is ready
          const ResP& rr = p.res[0];
          const bf16* rp = rr.p + rr.co + gch +
              (((rr.bs0 ? 0u : nn[k]) * (uint32_t)rr.H + (yy[k] >> rr.shift)) * (uint32_t)rr.W + (xx[k] >> rr.shift)) * (uint32_t)rr.cs;
          ldg32(rp, (p.v32 & 2) != 0, pre[2 * k], pre[2 * k + 1]);
        }
      }
      mbar_wait(bar_accf + 8u * acc_i, (st >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == W2_EPI_WARP0 && lane == 0 && st < 6) W2_STAMP(32 + 2 * st);
#pragma unroll
      for (int b16 = 0; b16 < 2; ++b16) {
        uint32_t v[32];
        tmem_ld_16x256b_x8(tmem_base + ((uint32_t)(q * 32 + b16 * 16) << 16) + acc_i * (uint32_t)W2_NS + (uint32_t)cbeg, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (b16 == 1) {                              // this warp's reads of the accumulator are done
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(bar_acce_leader + 8u * acc_i);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = b16 * 2 + h;
          if (!ok[k]) continue;
          float f[16];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {           // channels 4 g4 .. 4 g4 + 3 = columns (g, e) = (2 g4, 0..1), (2 g4 + 1, 0..1)
            const float4 bb = bp[g4];
            f[4 * g4 + 0] = __uint_as_float(v[4 * (2 * g4) + 2 * h + 0]) + bb.x;
            f[4 * g4 + 1] = __uint_as_float(v[4 * (2 * g4) + 2 * h + 1]) + bb.y;
            f[4 * g4 + 2] = __uint_as_float(v[4 * (2 * g4 + 1) + 2 * h + 0]) + bb.z;
            f[4 * g4 + 3] = __uint_as_float(v[4 * (2 * g4 + 1) + 2 * h + 1]) + bb.w;
          }
          if (nres > 0) {
            add_res8(f, pre[2 * k]);
            add_res8(f + 8, pre[2 * k + 1]);
            for (int qi = 1; qi < nres; ++qi) {
              const ResP& rr = p.res[qi];
              const bf16* rp = rr.p + rr.co + gch +
                  (((rr.bs0 ? 0u : nn[k]) * (uint32_t)rr.H + (yy[k] >> rr.shift)) * (uint32_t)rr.W + (xx[k] >> rr.shift)) * (uint32_t)rr.cs;
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
          if (!(p.skip & 4)) stg32(outp + ooff[k], (p.v32 & 1) != 0, o0, o1);
        }
      }
      if (warp == W2_EPI_WARP0 && lane == 0 && st < 6) W2_STAMP(33 + 2 * st);
    }
  }
  // ---- teardown: nobody may leave (or free TMEM) while the partner can still signal or write into this CTA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) W2_STAMP(2);
  if (warp == W2_PROD_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

long long* g_ws2_dbg = nullptr;
void ws2_dump_timeline() {
  static long long h[32 * 256];
  if (cudaDeviceSynchronize() != cudaSuccess) return;
  if (cudaMemcpy(h, g_ws2_dbg, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  for (int l = 12; l < 14; ++l) {
    const long long* r = h + l * 256;
    const long long t0 = r[1] ? r[1] : r[0];
    for (int cta = 0; cta < 2; ++cta) {
      const long long* q = r + cta * 128;
      fprintf(stderr, "[ws2 launch %d cta %d, ns since producer start] start %lld prod %lld end %lld |", l, cta, q[0] - t0, q[1] - t0, q[2] - t0);
      for (int st = 0; st < 6; ++st) fprintf(stderr, " u%d: full %lld pfull %lld acce %lld issued %lld |", st, q[8 + 4 * st] ? q[8 + 4 * st] - t0 : -1, q[9 + 4 * st] ? q[9 + 4 * st] - t0 : -1, q[10 + 4 * st] ? q[10 + 4 * st] - t0 : -1, q[11 + 4 * st] ? q[11 + 4 * st] - t0 : -1);
      fprintf(stderr, " epi:");
      for (int st = 0; st < 6; ++st) fprintf(stderr, " %lld-%lld", q[32 + 2 * st] ? q[32 + 2 * st] - t0 : -1, q[33 + 2 * st] ? q[33 + 2 * st] - t0 : -1);
      fprintf(stderr, "\n   unit 2 per chunk: producer empty_ok->issued:");
      for (int c = 0; c < 8; ++c) fprintf(stderr, " %lld->%lld", q[64 + 2 * c] - t0, q[65 + 2 * c] - t0);
      fprintf(stderr, "\n   unit 2 per chunk: mma full_ok/pfull_ok/issued:");
      for (int c = 0; c < 8; ++c) fprintf(stderr, " %lld/%lld/%lld", q[80 + 3 * c] - t0, q[81 + 3 * c] - t0, q[82 + 3 * c] - t0);
      fprintf(stderr, "\n");
    }
  }
}

struct Ws2Cfg {
  int KC, nimg, S;
  uint32_t plane_bytes, a_bytes, b_off, b_tap_bytes, stage_bytes;
};

bool ws2_config(int Cin, int CoutPad, int ntaps, int H, int W, Ws2Cfg* c) {
  if (Cin % 16 != 0 || CoutPad % W2_NS != 0 || ntaps < 1 || ntaps > 9 || H < 1 || W < 1) return false;
  const int P = W + 1, pitch = (H + 1) * P;
  if (P + 1 > 64 || P > 256 || H + 1 > 256) return false;          // TMA box limits
  int nimg = 256 / pitch;                                          // T = 2 row blocks of 128 flat pixels per CTA
  if (nimg < 1) return false;
  if (nimg > 64) nimg = 64;
  if (nimg * pitch * 10 < 256 * 7) return false;                   // < 70 % of the MMA rows used: the 1-CTA kernel packs better
  c->nimg = nimg;
  c->KC = 16;                                                      // fixed: the host packs the weights in 16-channel chunks
  c->plane_bytes = (uint32_t)nimg * pitch * 16u;
  c->a_bytes = (uint32_t)(c->KC / 8) * c->plane_bytes;
  c->b_off = (c->a_bytes + (uint32_t)(P + 1) * 16u + 127u) & ~127u;
  c->b_tap_bytes = (uint32_t)(c->KC / 8) * 64u * 16u;
  c->stage_bytes = (c->b_off + (uint32_t)ntaps * c->b_tap_bytes + 127u) & ~127u;
  // junk rows of the last row block read up to (256 - nimg*pitch + 2P + 2) pixels past the last plane: inside the stage
  if ((uint32_t)(256 - nimg * pitch + 2 * P + 2) * 16u > c->stage_bytes - c->a_bytes) return false;
  int S = (W2_SMEM_BUDGET - 128) / (int)c->stage_bytes;
  if (S > W2_MAX_S) S = W2_MAX_S;
  if (S < 3) return false;
  c->S = S;
  return true;
}

int make_flat_map2(const ConvP& p, const Ws2Cfg& c, CUtensorMap* m) {
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t es = 2;
  cuuint64_t dims[5] = {8, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.N, (cuuint64_t)(p.Cin / 8)};
  cuuint64_t strides[4] = {(cuuint64_t)p.in_cs * es, (cuuint64_t)p.Win * p.in_cs * es,
                           (cuuint64_t)p.Hin * p.Win * p.in_cs * es, 16};
  cuuint32_t box[5] = {8, (cuuint32_t)(p.Win + 1), (cuuint32_t)(p.Hin + 1), (cuuint32_t)c.nimg, (cuuint32_t)(c.KC / 8)};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(p.in + p.in_co), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (flat map, pair kernel) failed with %d", (int)r);
  return RSG_OK;
}

}  // namespace

// 1 when the CTA-pair kernel covers a stride-1 conv of this shape; the host then packs w_tc5 as
// [CoutPad/128][2 halves][Cin/16 chunks][ntaps][2][64][8] with the 64 output channels of a half in accumulator-column
// order (column 8g + 2j + e = channel 16j + 2g + e) and sets engine = 4.
extern "C" int rsg_conv_ws2_config(int Cin, int CoutPad, int ntaps, int H, int W) {
  Ws2Cfg c;
  return ws2_config(Cin, CoutPad, ntaps, H, W, &c) ? 1 : 0;
}

int conv_ws2_launch(const ConvP& p, cudaStream_t s, int* handled) {
  *handled = 0;
  if (!p.w_tc5 || !p.out || p.out_f32) return RSG_OK;
  if (p.stride != 1 || p.omul != 1 || p.ooy != 0 || p.oox != 0) return RSG_OK;
  if (p.Hout != p.Hin || p.Wout != p.Win || p.oH != p.Hin || p.oW != p.Win) return RSG_OK;
  if (p.Cout % 16 != 0 || p.in_cs % 8 != 0 || p.in_co % 8 != 0 || p.out_cs % 8 != 0 || p.out_co % 8 != 0) return RSG_OK;
  for (int t = 0; t < p.ntaps && t < 16; ++t)
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return RSG_OK;
  for (int q = 0; q < p.nres; ++q)
    if (p.res[q].cs % 8 != 0 || p.res[q].co % 8 != 0) return RSG_OK;
  Ws2Cfg c;
  if (!ws2_config(p.Cin, p.CoutPad, p.ntaps, p.Hin, p.Win, &c)) return RSG_OK;
  if (p.M == 0) { *handled = 1; return RSG_OK; }
  if ((long long)p.N * (p.Hin + 1) * (p.Win + 1) >= (1ll << 31)) return RSG_OK;
  if (rsg_num_sms() < 2) return RSG_OK;

  Ws2P k;
  memset(&k, 0, sizeof(k));
  k.w = p.w_tc5; k.bias = p.bias; k.Cin = p.Cin; k.Cout = p.Cout;
  k.KC = c.KC; k.nchunks = p.Cin / c.KC; k.S = c.S; k.nimg = c.nimg;
  k.ntaps = p.ntaps;
  k.H = p.Hin; k.W = p.Win; k.P = p.Win + 1; k.pitch = (p.Hin + 1) * k.P;
  for (int t = 0; t < p.ntaps; ++t) k.tapoff[t] = (1 + p.dy[t]) * k.P + (1 + p.dx[t]);
  k.magic_pitch = k.pitch > 1 ? (uint32_t)(((1ull << 32) + k.pitch - 1) / k.pitch) : 0u;
  k.magic_P = k.P > 1 ? (uint32_t)(((1ull << 32) + k.P - 1) / k.P) : 0u;
  k.N = p.N; k.nunits = (p.N + 2 * c.nimg - 1) / (2 * c.nimg);
  k.out = p.out; k.out_cs = p.out_cs; k.out_co = p.out_co;
  k.nres = p.nres;
  for (int q = 0; q < p.nres; ++q) k.res[q] = p.res[q];
  k.relu = p.relu;
  {
    auto al32 = [](const void* ptr, int cs, int co) { return ((uintptr_t)ptr % 32 == 0) && cs % 16 == 0 && co % 16 == 0; };
    k.v32 = (al32(p.out, p.out_cs, p.out_co) ? 1 : 0) | (p.nres > 0 && al32(p.res[0].p, p.res[0].cs, p.res[0].co) ? 2 : 0);
  }
  k.plane_bytes = c.plane_bytes; k.a_bytes = c.a_bytes; k.b_off = c.b_off; k.b_tap_bytes = c.b_tap_bytes;
  k.stage_bytes = c.stage_bytes;
  k.skip = rsg_dbg_int("RSG_WS2_SKIP", 0);
  k.ctawait = rsg_dbg_int("RSG_WS2_CTAWAIT", 0);
  static long long* dbg_buf = nullptr;
  static int dbg_launch = 0;
  if (rsg_dbg_env("RSG_WS2_TIMELINE")) {
    if (!dbg_buf) {
      cudaMalloc(&dbg_buf, 32 * 256 * sizeof(long long));
      cudaMemset(dbg_buf, 0, 32 * 256 * sizeof(long long));
      atexit(ws2_dump_timeline);
      g_ws2_dbg = dbg_buf;
    }
    k.dbg = dbg_buf + (size_t)(dbg_launch++ % 32) * 256;
  }
  const size_t smem = (size_t)c.S * c.stage_bytes + 128;
  const int nslices = p.CoutPad / W2_NS;

  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_ws2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_SMEM_BUDGET));
    RSG_CUDA(cudaFuncSetAttribute(conv_ws2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_once.done();
  }
  static const bool dbg = rsg_dbg_env("RSG_DEBUG") != nullptr;
  int npairs = rsg_num_sms() / 2 / nslices;
  if (npairs > k.nunits) npairs = k.nunits;
  if (npairs < 1) npairs = 1;
  if (dbg) fprintf(stderr, "[ws2] Cin=%d Cout=%d taps=%d %dx%d nimg=%d S=%d stage=%u smem=%zu units=%d pairs=%d x %d slices\n", p.Cin,
                   p.CoutPad, p.ntaps, p.Hin, p.Win, c.nimg, c.S, c.stage_bytes, smem, k.nunits, npairs, nslices);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  { int rc = make_flat_map2(p, c, &map); if (rc) return rc; }
  // cluster dimensions are compiled into the kernel (__cluster_dims__(2,1,1)); grid.x = 2 CTAs per pair
  RSG_CUDA(launch_pdl(conv_ws2_kernel, dim3((unsigned)(2 * npairs), (unsigned)nslices), dim3(W2_THREADS), smem, s, map, k));
  *handled = 1;
  return RSG_OK;
}
