// Weight gradients of the 3x3 stride-1 convolutions (channel counts that are multiples of 8, more than 32 on one side) on the
// Blackwell tensor pipe
// (reference: the autograd of nn.Conv2d inside loss.backward(), function.py:299-301; layers pose_rsgnet.py:38-95, 194-249).
//
//     dW[tap][ci][co] += sum over pixels m of X[m + shift(tap)][ci] * dY[m][co]
//
// Both operands of this product are contiguous along their M / N dimension (channels) and strided along K (pixels): "MN-major"
// in tcgen05 terms.  kind::tf32 returns zeros for MN-major operands (profiles/r2_notes.md section 9), kind::f16 accepts them, so
// the product runs on bf16 SPLITS: x = hi + lo (two bf16 values, 16 mantissa bits together), and
//     x * y ~= hi_x hi_y + hi_x lo_y + lo_x hi_y            (three MMAs, fp32 accumulation; the dropped lo * lo term is 2^-16 relative)
// which is more accurate than one TF32 product (10 mantissa bits) and, at twice the TF32 MMA rate, costs 1.5 TF32 MMAs.
//
// Flat form as in train_tc5.cu: the pixels of all images are ONE array of pitch P = W + 1 with a zero column per row and a
// zero row per image (= the padding), so a tap is a constant shift of the position index.  Shared-memory planes
// [8-channel group][position][8 ch x bf16 = 16 B] are the canonical MN-major no-swizzle layout with the K (position) core
// stride = 128 B, i.e. LINEAR in the position: a tap's shift is a start-address offset of 16 B per position.  A CTA owns one
// (ci tile, co tile, kernel row dy) and a range of positions; per 64-position chunk its eight producer warps gather
// X (64 + 2 positions: the three dx shifts) and dY through registers, split them and store the hi / lo planes; one elected
// thread issues 3 taps x 3 products x 4 k16-steps MMAs into three TMEM accumulators [MT x 64] (one per dx); the producers
// finish with vector reductions (red.global.add.v4.f32) into the packed gradient.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "umma.cuh"
#include "../../include/rsg_b200.h"

namespace {
using namespace umma;

constexpr int W5_THREADS = 288, W5_PROD = 256, W5_KC = 64, W5_NT = 64;     // 8 producer / epilogue warps + the MMA warp
constexpr uint32_t W5_XP = (W5_KC + 3) * 16u;      // X plane pitch: 66 positions + 1 (odd number of 16-byte slots: conflict-free stores)
constexpr uint32_t W5_YP = (W5_KC + 1) * 16u;      // dY plane pitch

struct W5P {
  const float* X; const float* dY; float* dW;
  int Nimg, H, W, Ca, Nc, ldx, ldy, P, RPI, tiles_n;
  int total, per_split;                            // flat positions; per_split is a multiple of W5_KC
  uint32_t magicP, magicR;                         // ceil(2^32 / P), ceil(2^32 / RPI)
};

// flat position (may be negative: halo above the first image) -> pixel row of the NHWC array, or -1 for padding / outside
__device__ __forceinline__ int w5_pixel(const W5P& p, int flat) {
  const uint32_t gs = (uint32_t)(flat + p.RPI * p.P);             // shifted by one image block: image index -1 = outside
  const uint32_t R = __umulhi(gs, p.magicP), Xc = gs - R * (uint32_t)p.P;
  const uint32_t nn = __umulhi(R, p.magicR), yy = R - nn * (uint32_t)p.RPI;
  const int n = (int)nn - 1;
  const bool ok = flat + p.RPI * p.P >= 0 && n >= 0 && n < p.Nimg && Xc != 0 && yy != 0;
  return ok ? (n * p.H + ((int)yy - 1)) * p.W + ((int)Xc - 1) : -1;
}

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (lo = the rounding error of hi, itself rounded to bf16)
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * e] - __bfloat162float(hh.x), f[2 * e + 1] - __bfloat162float(hh.y));
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 8 consecutive floats of a pixel: one 32-byte load (a full sector per lane and instruction) when the row is 32-byte aligned
__device__ __forceinline__ void ld8(const float* p, bool v32, float4& a, float4& b) {
  if (v32) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
  } else {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
}

__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void red4(float* p, const uint32_t* v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(p), "f"(__uint_as_float(v[0])), "f"(__uint_as_float(v[1])), "f"(__uint_as_float(v[2])), "f"(__uint_as_float(v[3]))
               : "memory");
}

// MT = rows of the accumulator = input channels per CTA (64: M = 64 MMAs, accumulator row r in TMEM lane r % 16 + 32 (r / 16))
// two CTAs per SM (each other's gather latency, epilogue and prologue): three operand stages for MT = 64, two for MT = 128
template <int MT>
__global__ void __launch_bounds__(W5_THREADS, 2) wgrad_tc5_kernel(const W5P p) {
  constexpr int W5_STAGES = MT == 64 ? 3 : 2;
  constexpr int G = MT / 8;                                        // X planes per hi / lo set
  constexpr int XPASS = ((W5_KC + 2) * G + W5_PROD - 1) / W5_PROD; // (position, plane) items per producer thread
  constexpr int YPASS = W5_KC * (W5_NT / 8) / W5_PROD;
  constexpr uint32_t X_SET = G * W5_XP, Y_SET = (W5_NT / 8) * W5_YP;
  constexpr uint32_t STAGE = (2 * X_SET + 2 * Y_SET + 127u) & ~127u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * W5_STAGES + 1];        // full[S] | empty[S] | accumulators complete
  __shared__ uint32_t tmem_base_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tm = blockIdx.x / p.tiles_n, tn = blockIdx.x - tm * p.tiles_n;
  const int ci0 = tm * MT, co0 = tn * W5_NT, dy = blockIdx.y;
  const int fb = blockIdx.z * p.per_split;
  const int fe = min(fb + p.per_split, p.total);
  const int nchunk = fe > fb ? (fe - fb + W5_KC - 1) / W5_KC : 0;
  if (nchunk == 0) return;
  const int mvalid = min(MT, p.Ca - ci0), nvalid = min(W5_NT, p.Nc - co0);   // partial tiles: multiples of 8 (host check)
  if (tid == 0) {
    for (int i = 0; i < W5_STAGES; ++i) { mbar_init(BAR(i), W5_PROD); mbar_init(BAR(W5_STAGES + i), 1); }
    mbar_init(BAR(2 * W5_STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp < 8) {
    // ===================== producers: gather, split into bf16 hi / lo planes =====================
    const int gx = tid % G, hx = tid / G;                          // X item of pass j: plane gx, halo position hx + j (128 / G)
    const int gy = tid & 7, hy = tid >> 3;                         // dY item of pass j: plane gy, position hy + 32 j
    const float* Xc = p.X + ci0 + 8 * gx;
    const float* Yc = p.dY + co0 + 8 * gy;
    // channel groups past the end of a partial tile are neither loaded nor stored: their accumulator rows / columns hold
    // whatever the shared memory held and are never read (rows and columns of a product are independent)
    const bool xg = 8 * gx < mvalid, yg = 8 * gy < nvalid;
    const bool x32 = (p.ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 31u) == 0;
    const bool y32 = (p.ldy & 7) == 0 && (reinterpret_cast<uintptr_t>(p.dY) & 31u) == 0;
    uint32_t s = 0, ph = 0;
    for (int c = 0; c < nchunk; ++c) {
      const int f0 = fb + c * W5_KC;
      float4 xa[XPASS], xb[XPASS], ya[YPASS], yb[YPASS];
      const int fx = f0 + (dy - 1) * p.P - 1;                      // X halo position h <-> flat fx + h; tap dx reads h = r + dx
#pragma unroll
      for (int j = 0; j < XPASS; ++j) {
        const int h = hx + j * (W5_PROD / G);
        xa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        xb[j] = xa[j];
        if (h < W5_KC + 2 && xg) {
          const int px = w5_pixel(p, fx + h);
          if (px >= 0) {
            ld8(Xc + (long long)px * p.ldx, x32, xa[j], xb[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < YPASS; ++j) {
        const int r = hy + 32 * j;
        ya[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        yb[j] = ya[j];
        const int px = yg && f0 + r < fe ? w5_pixel(p, f0 + r) : -1;
        if (px >= 0) {
          ld8(Yc + (long long)px * p.ldy, y32, ya[j], yb[j]);
        }
      }
      mbar_wait(BAR(W5_STAGES + s), ph ^ 1u);                      // the MMAs that read this stage have completed
      const uint32_t st = sbase + s * STAGE;
#pragma unroll
      for (int j = 0; j < XPASS; ++j) {
        const int h = hx + j * (W5_PROD / G);
        if (h < W5_KC + 2 && xg) {
          uint4 hi, lo;
          split8(xa[j], xb[j], hi, lo);
          const uint32_t a = st + (uint32_t)gx * W5_XP + (uint32_t)h * 16u;
          sts16(a, hi);
          sts16(a + X_SET, lo);
        }
      }
#pragma unroll
      for (int j = 0; j < YPASS; ++j) {
        if (!yg) break;
        uint4 hi, lo;
        split8(ya[j], yb[j], hi, lo);
        const uint32_t a = st + 2 * X_SET + (uint32_t)gy * W5_YP + (uint32_t)(hy + 32 * j) * 16u;
        sts16(a, hi);
        sts16(a + Y_SET, lo);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(BAR(s));
      if (++s == W5_STAGES) { s = 0; ph ^= 1u; }
    }
    // ===================== epilogue: three [MT x 64] accumulators -> vector reductions into dW =====================
    mbar_wait(BAR(2 * W5_STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;                      // TMEM lane quarter of this warp | column half of each accumulator
    const int row = MT == 128 ? q * 32 + lane : q * 16 + lane;     // accumulator row = input channel ci0 + row
    const bool rok = (MT == 128 || lane < 16) && row < mvalid;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int dx = 0; dx < 3; ++dx) {
      float* wrow = p.dW + ((long long)(dy * 3 + dx) * p.Ca + ci0 + row) * p.Nc + co0;
#pragma unroll 1
      for (int cc = half * (W5_NT / 2); cc < (half + 1) * (W5_NT / 2) && cc < nvalid; cc += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)(dx * W5_NT + cc), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rok) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            if (cc + 4 * g4 < nvalid) red4(wrow + cc + 4 * g4, v + 4 * g4);
        }
      }
    }
  } else {
    // ===================== MMA issuer =====================
    // kind::f16: fp32 accumulators, bf16 x bf16, A and B MN-major (bits 15, 16)
    const uint32_t n_mma = (uint32_t)(nvalid + 15) & ~15u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n_mma >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
    const uint32_t hi_x = desc_hi(W5_XP), hi_y = desc_hi(W5_YP);   // SBO = one channel plane (MN direction)
    const uint32_t lbo = ((128u >> 4) & 0x3FFFu) << 16;            // LBO = 8 positions x 16 B (K direction)
    uint32_t s = 0, ph = 0;
    for (int c = 0; c < nchunk; ++c) {
      mbar_wait(BAR(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t st = sbase + s * STAGE;
        const uint32_t x16 = st >> 4, y16 = (st + 2 * X_SET) >> 4;
#pragma unroll 1
        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {                          // hi hi | hi lo | lo hi
            const uint32_t xa = x16 + (pr == 2 ? (X_SET >> 4) : 0u) + (uint32_t)dx;
            const uint32_t ya = y16 + (pr == 1 ? (Y_SET >> 4) : 0u);
#pragma unroll
            for (int k = 0; k < W5_KC / 16; ++k) {
              const uint64_t ad = ((uint64_t)hi_x << 32) | (((xa + 16u * k) & 0x3FFFu) | lbo);
              const uint64_t bd = ((uint64_t)hi_y << 32) | (((ya + 16u * k) & 0x3FFFu) | lbo);
              umma_f16(tmem_base + (uint32_t)(dx * W5_NT), ad, bd, idesc, (c | pr | k) ? 1u : 0u);
            }
          }
        }
        umma_commit(BAR(W5_STAGES + s));
        if (c == nchunk - 1) umma_commit(BAR(2 * W5_STAGES));
      }
      __syncwarp();
      if (++s == W5_STAGES) { s = 0; ph ^= 1u; }
    }
  }
#undef BAR
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

template <int MT>
constexpr size_t w5_smem() {
  return 128 + (size_t)(MT == 64 ? 3 : 2) * (((2u * (MT / 8) * W5_XP + 2u * (W5_NT / 8) * W5_YP) + 127u) & ~127u);
}

}  // namespace

// -1 when this kernel does not cover the call of rsg_train_wgrad (the caller takes the mma.sync kernels), else the launch status
int train_wgrad5_launch(cudaStream_t s, const float* X, const float* dY, float* dW, int M, int Ca, int Nc, int ldx, int ldy,
                        const int* geom) {
  if (!(geom[4] == 3 && geom[5] == 3 && geom[6] == 1 && geom[7] == 1 && geom[0] == geom[2] && geom[1] == geom[3])) return -1;
  if (Ca % 8 || Nc % 8 || (Ca <= 32 && Nc <= 32) || (ldx & 3) || (ldy & 3) || (reinterpret_cast<uintptr_t>(X) & 15u) || (reinterpret_cast<uintptr_t>(dY) & 15u) ||
      (reinterpret_cast<uintptr_t>(dW) & 15u))
    return -1;
  const int H = geom[0], W = geom[1];
  if (M % (H * W)) return -1;
  const long long nimg = M / (H * W), total = nimg * (H + 1) * (W + 1);
  if (total + (long long)(H + 1) * (W + 1) + 2 * (W + 1) + 256 >= (1ll << 31) || (long long)M * ldx >= (1ll << 31) ||
      (long long)M * ldy >= (1ll << 31))
    return -1;
  // the multiply-high position decode is exact for positions below 2^32 / P (two image blocks of slack: shift + halo)
  if (((nimg + 2) * (H + 1) * (W + 1) + 512) * (W + 1) >= (1ll << 32)) return -1;
  W5P p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.dY = dY; p.dW = dW; p.Nimg = (int)nimg; p.H = H; p.W = W; p.Ca = Ca; p.Nc = Nc; p.ldx = ldx; p.ldy = ldy;
  p.P = W + 1; p.RPI = H + 1; p.total = (int)total; p.tiles_n = (Nc + W5_NT - 1) / W5_NT;
  p.magicP = (uint32_t)(((1ull << 32) + p.P - 1) / p.P);
  p.magicR = (uint32_t)(((1ull << 32) + p.RPI - 1) / p.RPI);
  const int mt = Ca <= 64 ? 64 : 128;
  const int tiles = ((Ca + mt - 1) / mt) * p.tiles_n;
  const long long chunks = (total + W5_KC - 1) / W5_KC;
  long long want = 2ll * rsg_num_sms() / (3ll * tiles);            // the CTAs that are resident at once: two per SM
  if (want < 1) want = 1;
  if (want > chunks) want = chunks;
  if (want > 65535) want = 65535;
  const long long per = (chunks + want - 1) / want;
  p.per_split = (int)(per * W5_KC);
  const unsigned splits = (unsigned)((chunks + per - 1) / per);
  static DeviceOnce once;
  if (once.first()) {
    RSG_CUDA(cudaFuncSetAttribute(wgrad_tc5_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w5_smem<64>()));
    RSG_CUDA(cudaFuncSetAttribute(wgrad_tc5_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w5_smem<128>()));
    once.done();
  }
  const dim3 grid((unsigned)tiles, 3, splits);
  if (mt == 64) wgrad_tc5_kernel<64><<<grid, W5_THREADS, w5_smem<64>(), s>>>(p);
  else wgrad_tc5_kernel<128><<<grid, W5_THREADS, w5_smem<128>(), s>>>(p);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
