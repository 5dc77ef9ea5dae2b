// Weight-streaming tcgen05 convolution for the GEMM-shaped end of HRNet: stride-1 3x3 / 1x1 convs with
// many channels on small maps (128->128 @16x12, 256->256 @8x6 for W32; 192/384 for W48), where the
// folded weights (0.3 - 1.2 MB) cannot stay resident in shared memory and conv_tc5.cu's 16x8 pixel
// patch wastes most MMA rows.
//
// "Flat" implicit GEMM.  nimg whole images are brought into shared memory by ONE 5-D TMA box per
// K-chunk, starting at (x,y) = (-1,-1), so that every image arrives as (H+1) rows of P = W+1 pixels
// whose first row and first column are the convolution's zero padding (TMA out-of-bounds fill):
//     plane[k8][img][r = y+1][c = x+1][8 ch]      (one plane per 8 input channels)
// Seen as a flat pixel array with pitch P, the input of tap (dy,dx) for output pixel o = img*pitch +
// y*P + x is simply flat[o + (1+dy)*P + (1+dx)]: the zero column doubles as the right padding of the
// previous row and the zero row as the bottom padding of the previous image.  The 128 rows of an MMA
// are therefore 128 CONSECUTIVE flat pixels (K-major SWIZZLE_NONE: core matrix = 8 pixels x 16 B,
// SBO = 128 B, LBO = one plane), each tap is a start-address offset, and images of any size pack into
// M tiles without patch quantisation (8x6: 75 % useful rows instead of 37 %).  Rows that fall on the
// padding positions compute junk and are not stored.
//
// A supertile = nimg images = T <= 4 M tiles, each with its own TMEM accumulator of NS columns
// (T * NS <= 512).  The K loop runs over chunks of KC input channels; one stage of the ring holds the
// activation planes of the chunk plus the weights of all taps for that chunk ([tap][KC/8][NS][8],
// bulk copies), and all T tiles consume the stage, so the weight stream is amortised over 512 pixels.
//
// Warp roles (608 threads, one persistent CTA per SM): warps 0-1 issue MMAs (tiles t = warp mod 2, so
// two instruction streams feed the tensor pipe, profiles/r1_notes.md §1), warp 2 = TMA producer and
// owner of the TMEM allocation, warps 3-18 = two epilogue groups of 8 warps (tile parity), each warp one
// TMEM lane quarter and one half of the NS columns: +bias, +residual terms, ReLU, bf16 NHWC stores.
//
// Reference ops subsumed: Conv2d(3x3|1x1, s1, bias=False) + BatchNorm2d(eval) [+ residual] [+ ReLU] of
// the stage-3/4 low-resolution branches (pose_rsgnet.py:25-54 BasicBlock inside HighResolutionModule
// :108-272).
#include "umma.cuh"

namespace {
using namespace umma;

constexpr int WS_THREADS = 608;                // 19 warps: 104 registers per thread
constexpr int WS_EPI_WARP0 = 3;
constexpr int WS_MAX_S = 8;
constexpr int WS_MAX_T = 4;
constexpr int WS_SMEM_BUDGET = 225 * 1024;

struct WsP {
  const bf16* w;        // [slice][chunk][tap][KC/8][NS][8]: the weights of a 16-channel chunk are one contiguous block
  const float* bias;
  int Cin, NS, Cout;
  int KC, nchunks, S, T, nimg;
  int nph;              // 1 (stride 1) or 4 phase-plane sets (stride 2): one TMA box each
  int ntaps;
  int tapoff[9];        // [phase block +] (1+dy)*P + (1+dx), pixels (= 16-byte units)
  int H, W, P, pitch;   // OUTPUT map size; P = W + 1, pitch = (H + 1) * P pixels per image
  uint32_t magic_pitch, magic_P;
  int N, nsuper;
  bf16* out;
  int out_cs, out_co;
  int nres;
  ResP res[4];
  int relu;
  int v32;              // bit0: output rows 32-byte aligned, bit1: residual term 0 too (LDG/STG.256)
  uint32_t plane_bytes, phase_bytes, a_bytes, b_off, b_tap_bytes, stage_bytes, tmem_cols;   // phase_bytes: one phase's planes, padded to 128 B
  long long* dbg;       // debug timeline of CTA (0,0): globaltimer stamps [16] or nullptr
  int skip;             // debug: bit0 no loads, bit1 one tap only, bit2 no residual loads / stores
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define WS_STAMP(i) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[i] = gtime(); } while (0)

__global__ void __launch_bounds__(WS_THREADS, 1)
conv_ws_kernel(const __grid_constant__ CUtensorMap in_map, const WsP p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WS_MAX_S + 2 * WS_MAX_T];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float sBias[256];
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = p.S, B_ACCF = 2 * p.S, B_ACCE = 2 * p.S + p.T;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;        // TMA destinations: 128-byte aligned
  unsigned char* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.y;

  if (threadIdx.x == 0) {
    WS_STAMP(0);
    for (int i = 0; i < p.S; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 2); }
    for (int i = 0; i < p.T; ++i) { mbar_init(BAR(B_ACCF + i), 1); mbar_init(BAR(B_ACCE + i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.NS; i += blockDim.x) sBias[i] = p.bias[slice * p.NS + i];
  // the gap between the activation planes and the weights of every stage is what the bottom-right
  // taps of the last image read: keep it zero (TMA never writes there)
  {
    const uint32_t gap16 = (p.b_off - p.a_bytes) >> 4;
    for (uint32_t i = threadIdx.x; i < gap16 * (uint32_t)p.S; i += blockDim.x) {
      const uint32_t s = i / gap16, j = i - s * gap16;
      *reinterpret_cast<uint4*>(sgen + (size_t)s * p.stage_bytes + p.a_bytes + 16u * j) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const int first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0) WS_STAMP(1);
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 2) {
    // ===================== producer: one TMA box + ntaps bulk copies per stage =====================
    // The activation box (tensor TMA, lane 0) and the weights of ALL taps of the chunk as one contiguous block (bulk
    // copies of at most 32 KB, lanes 1..): the weights are packed chunk-major for this.  A TMA operation costs ~100
    // cycles of engine time whatever its size and ~130 cycles of a single thread's time to issue
    // (tools/bulk_rate.cu); ten copies per chunk ran the kernel at the producer's rate.  A copy completing before the
    // barrier is armed only drives the transaction count negative for a moment.
    pdl_wait();                                      // activations of the previous kernel
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
    const uint32_t w_chunk_bytes = (uint32_t)p.ntaps * p.b_tap_bytes;
    const unsigned char* wsl = reinterpret_cast<const unsigned char*>(p.w) + (size_t)slice * p.nchunks * (size_t)w_chunk_bytes;
    const uint32_t tx_bytes = (uint32_t)p.nph * (uint32_t)(p.KC >> 3) * p.plane_bytes + w_chunk_bytes;   // bytes the copies deliver (phase blocks are padded)
    const uint32_t piece = 32768u, my_off = (uint32_t)(lane - 1) * piece;
    const uint32_t my_bytes = (lane >= 1 && my_off < w_chunk_bytes) ? (w_chunk_bytes - my_off < piece ? w_chunk_bytes - my_off : piece) : 0u;
    uint32_t it = 0, s = 0, sph = 0;              // ring slot / phase, carried incrementally (no divisions)
    for (int u = first; u < p.nsuper; u += step) {
      for (int c = 0; c < p.nchunks; ++c, ++it, s = (s + 1 == (uint32_t)p.S ? 0u : s + 1), sph ^= (s == 0)) {
        mbar_wait(BAR(B_EMPTY + s), sph ^ 1u);
        const uint32_t dst = sbase + s * p.stage_bytes;
        if (p.skip & 1) { if (lane == 0) mbar_arrive(BAR(B_FULL + s)); continue; }
        if (lane == 0) {
          mbar_arrive_expect_tx(BAR(B_FULL + s), tx_bytes);
          if (p.nph == 1) tma_load_5d(dst, &in_map, BAR(B_FULL + s), 0, -1, -1, u * p.nimg, c * (p.KC >> 3));
          if (it == 0) WS_STAMP(2);
          WS_STAMP(3);
        }
        if (p.nph == 4 && lane >= 28) {
          // stride 2: phase (py, px) = the input pixels (2y' + py, 2x' + px), loaded with an element stride of 2 from
          // (px - 2, py - 2) so that plane row / column 0 is y' / x' = -1 (out of bounds: the zero padding)
          const int ph = lane - 28;
          tma_load_5d(dst + (uint32_t)ph * p.phase_bytes, &in_map, BAR(B_FULL + s), 0, (ph & 1) - 2,
                      (ph >> 1) - 2, u * p.nimg, c * (p.KC >> 3));
        } else if (my_bytes) {
          bulk_load(dst + p.b_off + my_off, wsl + (size_t)c * w_chunk_bytes + my_off, my_bytes, BAR(B_FULL + s));
        }
      }
    }
  } else if (warp < 2) {
    // ===================== MMA issuers =====================
    // Warp-uniform control flow, tcgen05 instructions predicated on one elected lane (UTCHMMA takes its
    // descriptors from uniform registers; see conv_tc5.cu).
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NS >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hiA = desc_hi(128u), hiB = desc_hi(128u);
    const uint32_t lo_lbo_a = ((p.plane_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t lo_lbo_b = (((uint32_t)p.NS * 16u >> 4) & 0x3FFFu) << 16;
    const uint32_t a_kstep = (2u * p.plane_bytes) >> 4, b_kstep = (2u * (uint32_t)p.NS * 16u) >> 4;
    const uint32_t b_tapstep = p.b_tap_bytes >> 4;
    const int kc2n = p.KC >> 4;
    // The issuing thread is bound by the LATENCY of its own instruction stream (tools/umma_rate.cu: one
    // thread sustains one MMA per 137 cycles with a 10-instruction loop body, 203 with a tap lookup), so
    // the hot shape (9 taps, one k16 step per chunk) runs fully unrolled from per-tap offsets kept in
    // registers: two adds + the descriptor moves per MMA.
    const bool fast9 = p.ntaps == 9 && kc2n == 1 && !(p.skip & 2);
    uint32_t aoff[9], boff[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      aoff[tp] = (uint32_t)p.tapoff[tp < p.ntaps ? tp : 0];
      boff[tp] = (uint32_t)tp * b_tapstep;
    }
    uint32_t it = 0, st = 0, s = 0, sph = 0;
    for (int u = first; u < p.nsuper; u += step, ++st) {
      for (int c = 0; c < p.nchunks; ++c, ++it, s = (s + 1 == (uint32_t)p.S ? 0u : s + 1), sph ^= (s == 0)) {
        mbar_wait(BAR(B_FULL + s), sph);
        if (warp == 0 && lane == 0) { if (it == 0) WS_STAMP(4); if (it == (uint32_t)p.nchunks) WS_STAMP(6); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage16 = (sbase + s * p.stage_bytes) >> 4;
        const uint32_t b_stage = (stage16 + (p.b_off >> 4)) | lo_lbo_b;
        if (elect_one()) {
          for (int t = warp; t < p.T; t += 2) {
            if (c == 0) {                                   // accumulator drained by the epilogue
              mbar_wait(BAR(B_ACCE + t), (st & 1) ^ 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.NS);
            const uint32_t a_tile = (stage16 + 128u * (uint32_t)t) | lo_lbo_a;
            uint32_t acc = c > 0;
            if (fast9) {
#pragma unroll
              for (int tp = 0; tp < 9; ++tp) {
                umma_f16(d_tmem, ((uint64_t)hiA << 32) | (a_tile + aoff[tp]), ((uint64_t)hiB << 32) | (b_stage + boff[tp]),
                         idesc, tp == 0 ? acc : 1u);
              }
              if (c == p.nchunks - 1) umma_commit(BAR(B_ACCF + t));
              continue;
            }
            const int ntp = (p.skip & 2) ? 1 : p.ntaps;
            for (int tp = 0; tp < ntp; ++tp) {
              uint32_t alo = a_tile + (uint32_t)p.tapoff[tp];
              uint32_t blo = b_stage + (uint32_t)tp * b_tapstep;
              for (int kc = 0; kc < kc2n; ++kc) {
                umma_f16(d_tmem, ((uint64_t)hiA << 32) | alo, ((uint64_t)hiB << 32) | blo, idesc, acc);
                acc = 1;
                alo += a_kstep;
                blo += b_kstep;
              }
            }
            if (c == p.nchunks - 1) umma_commit(BAR(B_ACCF + t));     // accumulator of tile t complete
          }
          umma_commit(BAR(B_EMPTY + s));                    // stage free once this warp's MMAs retire
        }
        __syncwarp();
        if (warp == 0 && lane == 0) { if (it == (uint32_t)p.nchunks - 1) WS_STAMP(5); WS_STAMP(7); }
      }
    }
  } else if (warp >= WS_EPI_WARP0) {
    // ===================== epilogue (warps 3 .. 18) =====================
    const int eidx = (warp - WS_EPI_WARP0) >> 2;    // 0..3
    const int egroup = eidx & 1;
    const int q = warp & 3;                         // TMEM lane quarter (hardware rule: warp % 4)
    const int ncol = p.NS >> 1;
    const int cbeg = (eidx >> 1) ? ncol : 0;
    const int gch0 = slice * p.NS + cbeg;
    const int nres = (p.skip & 4) ? 0 : p.nres;
    const float* biasp = sBias + cbeg;
    bf16* const outp = p.out + p.out_co + gch0;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
    const uint32_t bar_accf = opaque(BAR(B_ACCF)), bar_acce = opaque(BAR(B_ACCE));
    pdl_wait();                                        // residual terms are read, outputs written: previous kernel done
    uint32_t st = 0;
    for (int u = first; u < p.nsuper; u += step, ++st) {
      for (int t = egroup; t < p.T; t += 2) {
        const uint32_t o = (uint32_t)t * 128u + (uint32_t)(q * 32 + lane);
        const uint32_t i = fastdiv(o, p.magic_pitch);
        const uint32_t r = o - i * (uint32_t)p.pitch;
        const uint32_t y = fastdiv(r, p.magic_P);
        const uint32_t x = r - y * (uint32_t)p.P;
        const uint32_t n = (uint32_t)u * (uint32_t)p.nimg + i;
        const bool ok = i < (uint32_t)p.nimg && n < (uint32_t)p.N && y < (uint32_t)p.H && x < (uint32_t)p.W;
        const uint32_t ooff = ((n * (uint32_t)p.H + y) * (uint32_t)p.W + x) * (uint32_t)p.out_cs;
        const bf16* r0p = nullptr;
        // the whole residual row segment of this thread (first 64 columns) is fetched BEFORE waiting for the
        // accumulator: the epilogue warps are idle during the K loop, so the L2 latency is free
        uint4 pre[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pre[j] = make_uint4(0, 0, 0, 0);
        if (nres > 0 && ok && gch0 < p.Cout) {
          const ResP& rr = p.res[0];
          r0p = rr.p + rr.co + gch0 +
                (((rr.bs0 ? 0u : n) * (uint32_t)rr.H + (y >> rr.shift)) * (uint32_t)rr.W + (x >> rr.shift)) * (uint32_t)rr.cs;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j * 8 < ncol && !(j & 1)) ldg32(r0p + j * 8, (p.v32 & 2) != 0, pre[j], pre[j | 1]);
        }
        mbar_wait(bar_accf + 8u * (uint32_t)t, st & 1);
        if (warp == WS_EPI_WARP0 && lane == 0 && st == 0 && t == egroup) WS_STAMP(8);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tq + (uint32_t)(t * p.NS);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c0 = ci * 16;
          if (c0 >= ncol) break;
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c0 + 16 >= ncol) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8u * (uint32_t)t);
          }
          if (!ok || gch0 + c0 >= p.Cout) continue;
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
                  (((rr.bs0 ? 0u : n) * (uint32_t)rr.H + (y >> rr.shift)) * (uint32_t)rr.W + (x >> rr.shift)) * (uint32_t)rr.cs;
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
          if (!(p.skip & 4)) {
            stg32(outp + ooff + c0, (p.v32 & 1) != 0, o0, o1);
          }
        }
        if (warp == WS_EPI_WARP0 && lane == 0) { if (st == 0 && t == egroup) WS_STAMP(9); WS_STAMP(10); }
      }
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) WS_STAMP(11);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

long long* g_dbg_buf = nullptr;
void ws_dump_timeline() {
  long long h[64 * 16];
  if (cudaDeviceSynchronize() != cudaSuccess) return;
  if (cudaMemcpy(h, g_dbg_buf, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  long long t0 = 0;
  for (int i = 0; i < 64; ++i) if (h[i * 16] && (!t0 || h[i * 16] < t0)) t0 = h[i * 16];
  fprintf(stderr, "[ws timeline of CTA(0,0), ns] start prologue_done first_issue last_issue | mma_first_full mma_unit0_last_issued "
                  "mma_unit1_first_full mma_last_issued | epi_first_accf epi_first_tile_done epi_last_done | teardown\n");
  for (int i = 0; i < 64; ++i) {
    if (!h[i * 16]) continue;
    fprintf(stderr, "%2d:", i);
    for (int j = 0; j < 12; ++j) fprintf(stderr, " %7lld", h[i * 16 + j] ? h[i * 16 + j] - t0 : -1);
    fprintf(stderr, "\n");
  }
}

struct WsCfg {
  int NS, KC, nimg, T, S, nph;
  uint32_t phase_bytes;
  uint32_t plane_bytes, a_bytes, b_off, b_tap_bytes, stage_bytes;
};

// H, W: the OUTPUT map ((Hin - 1) / stride + 1 ...); stride 2 keeps four phase-plane sets of the input per stage
bool ws_config(int Cin, int CoutPad, int ntaps, int H, int W, int stride, WsCfg* c) {
  if (Cin % 16 != 0 || ntaps < 1 || ntaps > 9 || H < 1 || W < 1 || (stride != 1 && stride != 2)) return false;
  c->nph = stride == 2 ? 4 : 1;
  int NS = 0;
  if (CoutPad % 128 == 0) NS = 128;
  else if (CoutPad <= 256 && CoutPad % 16 == 0) NS = CoutPad;
  if (NS < 64) return false;
  const int P = W + 1, pitch = (H + 1) * P;
  // one whole image must fit into the T <= 4 accumulators of a supertile (T * NS <= 512 TMEM columns): a narrower
  // slice when it does not (W48's 192 -> 192 @24x18: pitch 475 needs T = 4, i.e. NS = 96 x 2 slices)
  auto fits = [&](int ns) { int t = 512 / ns; if (t > WS_MAX_T) t = WS_MAX_T; return t * 128 >= pitch; };
  if (!fits(NS)) {
    const int cands[3] = {128, 96, 64};
    for (int i = 0; i < 3; ++i)
      if (cands[i] < NS && CoutPad % cands[i] == 0 && fits(cands[i])) { NS = cands[i]; break; }
  }
  int tmax = 512 / NS;
  if (tmax > WS_MAX_T) tmax = WS_MAX_T;
  if (P + 1 > 64 || P * stride > 256 || (H + 1) * stride > 256) return false;
  int nimg = tmax * 128 / pitch;
  if (nimg < 1) return false;
  if (nimg > 64) nimg = 64;
  c->NS = NS; c->nimg = nimg;
  c->KC = 16;                                    // fixed: the host packs the weights in 16-channel chunks
  c->T = (nimg * pitch + 127) / 128;
  c->plane_bytes = (uint32_t)nimg * pitch * 16u;
  c->phase_bytes = ((uint32_t)(c->KC / 8) * c->plane_bytes + 127u) & ~127u;       // TMA destinations are 128-byte aligned
  c->a_bytes = c->nph == 1 ? (uint32_t)(c->KC / 8) * c->plane_bytes : 4u * c->phase_bytes;
  c->b_off = (c->a_bytes + (uint32_t)(P + 1) * 16u + 127u) & ~127u;
  c->b_tap_bytes = (uint32_t)(c->KC / 8) * NS * 16u;
  c->stage_bytes = (c->b_off + (uint32_t)ntaps * c->b_tap_bytes + 127u) & ~127u;
  // junk rows of the last tile read up to (T*128 - nimg*pitch + 2P + 2) pixels past the last plane:
  // that must stay inside the stage (gap + weights)
  if ((uint32_t)(c->T * 128 - nimg * pitch + 2 * P + 2) * 16u > c->stage_bytes - c->a_bytes) return false;
  int S = (WS_SMEM_BUDGET - 128) / (int)c->stage_bytes;
  if (S > WS_MAX_S) S = WS_MAX_S;
  if (S < 2) return false;
  c->S = S;
  return true;
}

int make_flat_map(const ConvP& p, const WsCfg& c, int stride, CUtensorMap* m) {
  EncodeTiledFn enc = tensor_map_encoder();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t es = 2;
  // NHWC activation viewed as (8 ch, W, H, N, C/8): the image index sits INSIDE the channel-plane index so
  // that a box lands as [KC/8][nimg][H+1][W+1][8ch]
  cuuint64_t dims[5] = {8, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)p.N, (cuuint64_t)(p.Cin / 8)};
  cuuint64_t strides[4] = {(cuuint64_t)p.in_cs * es, (cuuint64_t)p.Win * p.in_cs * es,
                           (cuuint64_t)p.Hin * p.Win * p.in_cs * es, 16};
  // with an element stride the box extent counts SOURCE elements: Wout + 1 loaded pixels span (Wout + 1) * stride of them
  cuuint32_t box[5] = {8, (cuuint32_t)((p.Wout + 1) * stride), (cuuint32_t)((p.Hout + 1) * stride), (cuuint32_t)c.nimg,
                       (cuuint32_t)(c.KC / 8)};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(p.in + p.in_co), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (flat map, stride %d) failed with %d", stride, (int)r);
  return RSG_OK;
}

}  // namespace

// Shape -> output channels per CTA: the host packer lays w_tc5 out as [CoutPad/NS][Cin/16 chunks][ntaps][2][NS][8].
// Returns 0 when the weight-streaming kernel does not cover the shape.
extern "C" int rsg_conv_ws_config2(int Cin, int CoutPad, int ntaps, int Hin, int Win, int stride, int* NS) {
  if (stride != 1 && stride != 2) return 0;
  WsCfg c;
  if (!ws_config(Cin, CoutPad, ntaps, (Hin - 1) / stride + 1, (Win - 1) / stride + 1, stride, &c)) return 0;
  if (NS) *NS = c.NS;
  return 1;
}
extern "C" int rsg_conv_ws_config(int Cin, int CoutPad, int ntaps, int H, int W, int* NS) {
  return rsg_conv_ws_config2(Cin, CoutPad, ntaps, H, W, 1, NS);
}

int conv_ws_launch(const ConvP& p, cudaStream_t s, int* handled) {
  *handled = 0;
  static const bool disabled = rsg_dbg_env("RSG_DISABLE_WS") != nullptr;
  if (disabled) return RSG_OK;
  if (!p.w_tc5 || !p.out || p.out_f32) return RSG_OK;
  if ((p.stride != 1 && p.stride != 2) || p.omul != 1 || p.ooy != 0 || p.oox != 0) return RSG_OK;
  if (p.Hout != (p.Hin - 1) / p.stride + 1 || p.Wout != (p.Win - 1) / p.stride + 1 || p.oH != p.Hout || p.oW != p.Wout) return RSG_OK;
  if (p.Cout % 16 != 0 || p.in_cs % 8 != 0 || p.in_co % 8 != 0 || p.out_cs % 8 != 0 || p.out_co % 8 != 0) return RSG_OK;
  for (int t = 0; t < p.ntaps && t < 16; ++t)
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return RSG_OK;
  for (int q = 0; q < p.nres; ++q)
    if (p.res[q].cs % 8 != 0 || p.res[q].co % 8 != 0) return RSG_OK;
  WsCfg c;
  if (!ws_config(p.Cin, p.CoutPad, p.ntaps, p.Hout, p.Wout, p.stride, &c)) return RSG_OK;
  if (p.M == 0) { *handled = 1; return RSG_OK; }
  if ((long long)p.N * (p.Hin + 1) * (p.Win + 1) >= (1ll << 31)) return RSG_OK;

  WsP k;
  memset(&k, 0, sizeof(k));
  k.w = p.w_tc5; k.bias = p.bias; k.Cin = p.Cin; k.NS = c.NS; k.Cout = p.Cout;
  k.KC = c.KC; k.nchunks = p.Cin / c.KC; k.S = c.S; k.T = c.T; k.nimg = c.nimg; k.nph = c.nph;
  k.ntaps = p.ntaps;
  k.H = p.Hout; k.W = p.Wout; k.P = p.Wout + 1; k.pitch = (p.Hout + 1) * k.P;
  for (int t = 0; t < p.ntaps; ++t) {
    if (p.stride == 1) k.tapoff[t] = (1 + p.dy[t]) * k.P + (1 + p.dx[t]);
    else {
      // input pixel (2y + dy, 2x + dx): dy = 0 -> even-row phase at y' = y; dy = -1 -> odd-row phase at y' = y - 1; dy = +1 -> odd-row
      // phase at y' = y (the same for x); plane row / column 0 is y' / x' = -1
      const int ph = (p.dy[t] != 0 ? 2 : 0) + (p.dx[t] != 0 ? 1 : 0);
      const int phase_px = ph * (int)(c.phase_bytes / 16u);
      k.tapoff[t] = phase_px + (1 + (p.dy[t] == -1 ? -1 : 0)) * k.P + (1 + (p.dx[t] == -1 ? -1 : 0));
    }
  }
  // magic = ceil(2^32 / d): exact for n * d < 2^32 (n < 512 here)
  k.magic_pitch = k.pitch > 1 ? (uint32_t)(((1ull << 32) + k.pitch - 1) / k.pitch) : 0u;
  k.magic_P = k.P > 1 ? (uint32_t)(((1ull << 32) + k.P - 1) / k.P) : 0u;
  k.N = p.N; k.nsuper = (p.N + c.nimg - 1) / c.nimg;
  k.out = p.out; k.out_cs = p.out_cs; k.out_co = p.out_co;
  k.nres = p.nres;
  for (int q = 0; q < p.nres; ++q) k.res[q] = p.res[q];
  k.relu = p.relu;
  {
    auto al32 = [](const void* ptr, int cs, int co) { return ((uintptr_t)ptr % 32 == 0) && cs % 16 == 0 && co % 16 == 0; };
    k.v32 = (al32(p.out, p.out_cs, p.out_co) ? 1 : 0) | (p.nres > 0 && al32(p.res[0].p, p.res[0].cs, p.res[0].co) ? 2 : 0);
    if (rsg_dbg_env("RSG_NO_V32")) k.v32 = 0;
  }
  { const char* e = rsg_dbg_env("RSG_WS_SKIP"); k.skip = e ? atoi(e) : 0; }
  static int dbg_launch = 0;
  if (rsg_dbg_env("RSG_WS_TIMELINE")) {
    if (!g_dbg_buf) {
      cudaMalloc(&g_dbg_buf, 64 * 16 * sizeof(long long));
      cudaMemset(g_dbg_buf, 0, 64 * 16 * sizeof(long long));
      atexit(ws_dump_timeline);
    }
    k.dbg = g_dbg_buf + (size_t)(dbg_launch++ % 64) * 16;
  }
  k.plane_bytes = c.plane_bytes; k.phase_bytes = c.phase_bytes; k.a_bytes = c.a_bytes; k.b_off = c.b_off; k.b_tap_bytes = c.b_tap_bytes;
  k.stage_bytes = c.stage_bytes;
  uint32_t cols = 32;
  while (cols < (uint32_t)(c.T * c.NS)) cols <<= 1;
  k.tmem_cols = cols;
  const size_t smem = (size_t)c.S * c.stage_bytes + 128;
  const int nslices = p.CoutPad / c.NS;

  static DeviceOnce attr_once;
  if (attr_once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(conv_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BUDGET));
    RSG_CUDA(cudaFuncSetAttribute(conv_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_once.done();
  }
  static const bool dbg = rsg_dbg_env("RSG_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "[ws] Cin=%d Cout=%d taps=%d %dx%d NS=%d nimg=%d T=%d S=%d stage=%u smem=%zu nsuper=%d\n", p.Cin,
                   p.CoutPad, p.ntaps, p.Hin, p.Win, c.NS, c.nimg, c.T, c.S, c.stage_bytes, smem, k.nsuper);
  int gx = rsg_num_sms() / nslices;
  if (gx > k.nsuper) gx = k.nsuper;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)nslices);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  { int rc = make_flat_map(p, c, p.stride, &map); if (rc) return rc; }
  RSG_CUDA(launch_pdl(conv_ws_kernel, grid, dim3(WS_THREADS), smem, s, map, k));
  *handled = 1;
  return RSG_OK;
}
