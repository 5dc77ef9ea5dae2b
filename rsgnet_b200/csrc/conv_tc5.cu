// Blackwell-native implicit-GEMM convolution: tcgen05.mma (single-thread issue, fp32 accumulators
// in TMEM) fed by TMA, for stride-1 convolutions whose taps lie in the 3x3 neighbourhood
// (3x3 p1 and 1x1) with weights that fit in shared memory.
//
// "Halo-resident" formulation.  A CTA owns a 16 x 8 pixel patch of one crop (= the 128 rows of the
// MMA M dimension).  ONE 5-D TMA box load brings the (16+2) x (8+2) halo patch of all Cin channels
// into shared memory laid out as [Cin/8][18][10][8 ch] -- which is exactly the canonical
// K-major / no-swizzle UMMA operand layout (8 consecutive pixels x 16 bytes = one core matrix,
// SBO = one halo row = 160 B, LBO = one 8-channel plane = 2880 B).  Each of the 9 taps is then just
// a different START ADDRESS of the A descriptor into that same patch ((1+dy)*10 + (1+dx) pixels
// further), so the activation tile is read from L2 once (1.4x with halo) instead of 9x, and the
// zero padding of the convolution is the TMA out-of-bounds fill.  Weights ([tap][Cin/8][Cout][8],
// BN folded) are bulk-copied once per CTA and stay resident while the CTA walks over its tiles.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> +bias +residual terms -> ReLU -> bf16 NHWC stores).
// Pipelines: 2 halo buffers (full/empty mbarriers) and 2 TMEM accumulators (full/empty mbarriers).
//
// Reference ops subsumed: Conv2d(3x3|1x1, s1) + BatchNorm2d(eval) [+ residual adds] [+ ReLU]
// (pose_rsgnet.py:38-54 BasicBlock, :75-95 Bottleneck, :261-270 fuse sum, heads :965-1003).
#include <cuda.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "conv_params.cuh"

namespace {

constexpr int TH = 16, TW = 8;                 // pixel patch = 128 MMA rows
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 26;      // bounded waits: a protocol bug traps, never hangs

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base_offset = 0, layout_type = SWIZZLE_NONE (0)
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

struct Tc5P {
  const bf16* w;        // [ntaps][Cin/8][N][8]
  const float* bias;    // [N]
  int Cin, N, Cout;
  int ntaps;
  int8_t dy[9], dx[9];
  int halo;             // 1 for 3x3, 0 for 1x1
  int H, W, Nimg;
  int tiles_x, tiles_y; // per image
  long long ntiles;
  bf16* out;
  int out_cs, out_co;
  int nres;
  ResP res[4];
  int relu;
  uint32_t w_bytes, halo_bytes, tmem_cols;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc5_kernel(const __grid_constant__ CUtensorMap in_map, const Tc5P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[9];
  __shared__ uint32_t tmem_base_slot;
  // bars: 0 weights, 1-2 halo full, 3-4 halo empty, 5-6 acc full, 7-8 acc empty
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  unsigned char* sW = smem;
  unsigned char* sH[2] = {smem + p.w_bytes, smem + p.w_bytes + p.halo_bytes};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW_ = TW + 2 * p.halo, HH_ = TH + 2 * p.halo;     // halo patch extent
  const uint32_t lbo_a = (uint32_t)HW_ * HH_ * 16u, sbo_a = (uint32_t)HW_ * 16u;
  const uint32_t lbo_b = (uint32_t)p.N * 16u, sbo_b = 128u;

  if (threadIdx.x == 0) {
    mbar_init(BAR(0), 1);
    mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);
    mbar_init(BAR(3), 1); mbar_init(BAR(4), 1);
    mbar_init(BAR(5), 1); mbar_init(BAR(6), 1);
    mbar_init(BAR(7), 4); mbar_init(BAR(8), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  const long long first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
      mbar_arrive_expect_tx(BAR(0), p.w_bytes);
      // bulk copies are limited in size only by the 20-bit tx-count per arrive: split in 64 KB pieces
      for (uint32_t off = 0; off < p.w_bytes; off += 65536u) {
        uint32_t n = p.w_bytes - off < 65536u ? p.w_bytes - off : 65536u;
        bulk_load(smem_u32(sW + off), reinterpret_cast<const unsigned char*>(p.w) + off, n, BAR(0));
      }
      uint32_t it = 0;
      for (long long t = first; t < p.ntiles; t += step, ++it) {
        const int b = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(BAR(3 + b), ph ^ 1);
        const int tx = (int)(t % p.tiles_x);
        const long long r = t / p.tiles_x;
        const int ty = (int)(r % p.tiles_y);
        const int n = (int)(r / p.tiles_y);
        mbar_arrive_expect_tx(BAR(1 + b), p.halo_bytes);
        tma_load_5d(smem_u32(sH[b]), &in_map, BAR(1 + b), 0, tx * TW - p.halo, ty * TH - p.halo, 0, n);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
      mbar_wait(BAR(0), 0);
      uint32_t it = 0;
      const int kc2n = p.Cin >> 4;
      for (long long t = first; t < p.ntiles; t += step, ++it) {
        const int b = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(BAR(1 + b), ph);            // halo landed
        mbar_wait(BAR(7 + b), ph ^ 1);        // accumulator drained by the epilogue
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_base = smem_u32(sH[b]);
        const uint32_t w_base = smem_u32(sW);
        const uint32_t d_tmem = tmem_base + (uint32_t)b * p.N;
        uint32_t acc = 0;
        for (int tp = 0; tp < p.ntaps; ++tp) {
          const uint32_t a_tap = a_base + (uint32_t)((p.halo + p.dy[tp]) * HW_ + (p.halo + p.dx[tp])) * 16u;
          const uint32_t w_tap = w_base + (uint32_t)tp * (uint32_t)(p.Cin >> 3) * lbo_b;
          for (int kc = 0; kc < kc2n; ++kc) {
            const uint64_t ad = make_desc(a_tap + (uint32_t)kc * 2u * lbo_a, lbo_a, sbo_a);
            const uint64_t bd = make_desc(w_tap + (uint32_t)kc * 2u * lbo_b, lbo_b, sbo_b);
            umma_f16(d_tmem, ad, bd, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(BAR(3 + b));              // halo buffer free once these MMAs retire
        umma_commit(BAR(5 + b));              // accumulator ready
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;            // MMA row = pixel inside the patch
    const int hy = row >> 3, wx = row & 7;
    uint32_t it = 0;
    for (long long t = first; t < p.ntiles; t += step, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int tx = (int)(t % p.tiles_x);
      const long long r = t / p.tiles_x;
      const int ty = (int)(r % p.tiles_y);
      const int n = (int)(r / p.tiles_y);
      const int y = ty * TH + hy, x = tx * TW + wx;
      const bool ok = y < p.H && x < p.W;
      mbar_wait(BAR(5 + b), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)b * p.N;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 32 >= p.N) {
          // all of this warp's TMEM reads for the tile are done: release the accumulator
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(7 + b));
        }
        if (!ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __ldg(p.bias + c0 + j);
        for (int qi = 0; qi < p.nres; ++qi) {
          const ResP& rr = p.res[qi];
          const bf16* rp = rr.p + ((size_t)((size_t)(rr.bs0 ? 0 : n) * rr.H + (y >> rr.shift)) * rr.W + (x >> rr.shift)) * rr.cs + rr.co + c0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            if (c0 + j4 * 8 < p.Cout) {
              uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + j4);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[j4 * 8 + 2 * e] += __bfloat162float(h2[e].x);
                f[j4 * 8 + 2 * e + 1] += __bfloat162float(h2[e].y);
              }
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        bf16* op = p.out + ((size_t)((size_t)n * p.H + y) * p.W + x) * p.out_cs + p.out_co + c0;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (c0 + j4 * 8 < p.Cout) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h2[e].x = __float2bfloat16_rn(f[j4 * 8 + 2 * e]);
              h2[e].y = __float2bfloat16_rn(f[j4 * 8 + 2 * e + 1]);
            }
            reinterpret_cast<uint4*>(op)[j4] = u;
          }
        }
      }
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

struct MapKey {
  const void* ptr; int cs, co, H, W, Cin, N, halo;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, cs, co, H, W, Cin, N, halo) < std::tie(o.ptr, o.cs, o.co, o.H, o.W, o.Cin, o.N, o.halo);
  }
};
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

int get_map(const ConvP& p, int halo, CUtensorMap* out) {
  MapKey k{p.in, p.in_cs, p.in_co, p.Hin, p.Win, p.Cin, p.N, halo};
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) { *out = it->second; return RSG_OK; }
  EncodeTiledFn enc = get_encode();
  RSG_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t es = 2;
  cuuint64_t dims[5] = {8, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)(p.Cin / 8), (cuuint64_t)p.N};
  cuuint64_t strides[4] = {(cuuint64_t)p.in_cs * es, (cuuint64_t)p.Win * p.in_cs * es, 16,
                           (cuuint64_t)p.Hin * p.Win * p.in_cs * es};
  cuuint32_t box[5] = {8, (cuuint32_t)(TW + 2 * halo), (cuuint32_t)(TH + 2 * halo), (cuuint32_t)(p.Cin / 8), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(p.in + p.in_co), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (Cin=%d H=%d W=%d cs=%d)", (int)r, p.Cin, p.Hin, p.Win, p.in_cs);
  g_maps[k] = m;
  *out = m;
  return RSG_OK;
}

}  // namespace

int conv_tc5_launch(const ConvP& p, cudaStream_t s, int* handled) {
  *handled = 0;
  static const bool disabled = getenv("RSG_DISABLE_TC5") != nullptr;   // A/B switch for debugging
  if (disabled) return RSG_OK;
  if (!p.w_tc5 || !p.out || p.out_f32) return RSG_OK;
  if (p.stride != 1 || p.omul != 1 || p.ooy != 0 || p.oox != 0) return RSG_OK;
  if (p.Hout != p.Hin || p.Wout != p.Win || p.oH != p.Hin || p.oW != p.Win) return RSG_OK;
  if (p.Cin % 16 != 0 || p.CoutPad % 32 != 0 || p.CoutPad > 256 || p.Cout % 8 != 0) return RSG_OK;
  if (p.ntaps > 9 || p.in_cs % 8 != 0 || p.in_co % 8 != 0 || p.out_cs % 8 != 0 || p.out_co % 8 != 0) return RSG_OK;
  int halo = 0;
  for (int t = 0; t < p.ntaps; ++t) {
    if (p.dy[t] < -1 || p.dy[t] > 1 || p.dx[t] < -1 || p.dx[t] > 1) return RSG_OK;
    if (p.dy[t] != 0 || p.dx[t] != 0) halo = 1;
  }
  for (int q = 0; q < p.nres; ++q)
    if (p.res[q].cs % 8 != 0 || p.res[q].co % 8 != 0) return RSG_OK;
  const uint32_t w_bytes = (uint32_t)p.ntaps * p.Cin * p.CoutPad * 2;
  const uint32_t halo_bytes = (uint32_t)(TH + 2 * halo) * (TW + 2 * halo) * p.Cin * 2;
  const size_t smem = (size_t)w_bytes + 2 * (size_t)halo_bytes;
  if (smem > 200 * 1024) return RSG_OK;          // weights do not fit: generic kernel
  if (p.M == 0) { *handled = 1; return RSG_OK; }

  Tc5P k;
  memset(&k, 0, sizeof(k));
  k.w = p.w_tc5; k.bias = p.bias; k.Cin = p.Cin; k.N = p.CoutPad; k.Cout = p.Cout;
  k.ntaps = p.ntaps;
  for (int t = 0; t < p.ntaps; ++t) { k.dy[t] = p.dy[t]; k.dx[t] = p.dx[t]; }
  k.halo = halo; k.H = p.Hin; k.W = p.Win; k.Nimg = p.N;
  k.tiles_x = (p.Win + TW - 1) / TW; k.tiles_y = (p.Hin + TH - 1) / TH;
  k.ntiles = (long long)k.tiles_x * k.tiles_y * p.N;
  k.out = p.out; k.out_cs = p.out_cs; k.out_co = p.out_co;
  k.nres = p.nres;
  for (int q = 0; q < p.nres; ++q) k.res[q] = p.res[q];
  k.relu = p.relu;
  k.w_bytes = w_bytes; k.halo_bytes = halo_bytes;
  uint32_t cols = 32;
  while (cols < 2u * p.CoutPad) cols <<= 1;
  k.tmem_cols = cols;

  CUtensorMap map;
  int rc = get_map(p, halo, &map);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    RSG_CUDA(cudaFuncSetAttribute(conv_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  long long grid = k.ntiles < rsg_num_sms() ? k.ntiles : rsg_num_sms();
  conv_tc5_kernel<<<(unsigned)grid, NTHREADS, smem, s>>>(map, k);
  RSG_LAUNCH_CHECK();
  *handled = 1;
  return RSG_OK;
}
