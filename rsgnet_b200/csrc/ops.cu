// HBM-bound helper kernels of the RSGNet forward: stem conv1 (Cin=3), fuse-layer sum with
// nearest up-sampling, 2x2 max-pool, GroupNorm, bilinear x2, relation_scores dump.
#include "ops.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// conv1: 3x3 s2 p1, 3 -> 64, fp32 NCHW in (optionally W-reversed), bf16 NHWC out, BN folded, ReLU.
// pose_rsgnet.py:612-613, 922-924 (+ input.flip(3), function.py:401).
//
// Warp-level tensor-core formulation (the fp32 FMA version ran at 16 TFLOP/s, 8x its HBM floor).  A CTA
// stages the (2*8+1) x (2*32+2) fp32 input patch of an 8 x 32 output tile in shared memory with cp.async,
// one tile ahead of the math (persistent CTAs, double buffer).  For kernel row ky the GEMM K index is
// k = kx*4 + c (kx = 0..3, c = 0..3; the kx = 3 and c = 3 entries have zero weights), so one k16 step of
// mma.sync.m16n8k16 is one kernel row and each A-fragment register is two LDS.32 + one bf16x2 convert:
// K = 3 x 16, N = 64 = 8 n-tiles, M = 16 consecutive output pixels.  All weights live in registers as
// B fragments (48 per thread).  Outputs go through a per-warp shared-memory transpose so that every
// global store is a fully used 16-byte piece of a contiguous 512-byte run.
// ---------------------------------------------------------------------------------------------
//
// Staging: the patch is an ALIGNED superset of the 66 needed columns ([64 tx - 4, 64 tx + 68): 18 sixteen-byte chunks per
// row), so that every copy is a 16-byte cp.async (the first version copied 3366 single floats per tile and was bound by
// the LSU, 2.4 TB/s).  The W flip of the second forward cannot be a reversed copy then; it is folded into the math
// instead: conv(flip(x))[ox] = sum_kx w[kx] x[2 ox' + 2 - kx] with ox' = Wo-1-ox, i.e. the UNFLIPPED patch read with a
// mirrored tap window and the output row written mirrored (FLIP instantiation) -- bit-identical to flipping the crop.
constexpr int ST_TY = 8, ST_TX = 32;
constexpr int ST_IR = 2 * ST_TY + 1, ST_IC = 2 * ST_TX + 8;      // 72 columns: patch column c is image column 64 tx - 4 + c
constexpr int ST_C4 = ST_IC / 4;
constexpr int ST_OPITCH = 144;                    // bytes per staged output pixel (128 + pad: conflict-free)

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t;
  t.x = __float2bfloat16_rn(a);
  t.y = __float2bfloat16_rn(b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool FLIP>
__global__ void __launch_bounds__(256, 2)
stem_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ w,
            const float* __restrict__ bias, bf16* __restrict__ out, int f0, int fl0, int n_crops,
            int tiles_x, int tiles_y, int ntiles) {
  // raw fp32 patches [buffer][c][row][col], filled by cp.async one tile ahead of the math
  __shared__ __align__(16) float sIn[2][3 * ST_IR * ST_IC];
  __shared__ __align__(16) unsigned char sOut[8][16 * ST_OPITCH];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int Ho = H >> 1, Wo = W >> 1;

  auto decode = [&](int tile, int& tx, int& ty, int& fl) {
    tx = tile % tiles_x; tile /= tiles_x;
    ty = tile % tiles_y;
    fl = tile / tiles_y;
  };
  // out-of-image chunks are the zero padding (cp.async with src-size 0 zero-fills); W is a multiple of 4, so an aligned
  // chunk is either inside the row or outside
  auto stage = [&](int tile, int buf) {
    int tx, ty, fl;
    decode(tile, tx, ty, fl);
    const int f = f0 + fl0 + fl;
    const float* xin = x + (size_t)(f % n_crops) * 3 * H * W;
    const int gy0 = 2 * ty * ST_TY - 1, gx0 = 2 * tx * ST_TX - 4;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&sIn[buf][0]);
    for (int i = tid; i < 3 * ST_IR * ST_C4; i += 256) {
      const int c = i / (ST_IR * ST_C4);
      const int rem = i - c * (ST_IR * ST_C4);
      const int r = rem / ST_C4, c4 = rem - r * ST_C4;
      const int gy = gy0 + r, gx = gx0 + 4 * c4;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* src = ok ? xin + ((size_t)c * H + gy) * W + gx : xin;
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sbase + 16u * (uint32_t)i), "l"(src), "r"(sz));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  int tile = blockIdx.x;
  if (tile < ntiles) stage(tile, 0);

  // ---- weights -> B fragments: b[ky][j][h] = {W[ky][kx][c], W[ky][kx][c+1]} for n = 8j+g, kx = tq/2 + 2h, c = 2(tq%2)
  uint32_t bfr[3][8][2];
  const int c0 = 2 * (tq & 1);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kx = (tq >> 1) + 2 * h;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = 8 * j + g;
        float w0 = 0.f, w1 = 0.f;
        if (kx < 3) {
          w0 = __ldg(w + ((c0 * 3 + ky) * 3 + kx) * 64 + n);
          if (c0 + 1 < 3) w1 = __ldg(w + (((c0 + 1) * 3 + ky) * 3 + kx) * 64 + n);
        }
        bfr[ky][j][h] = pack2(w0, w1);
      }
    }
  float bs[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { bs[j][0] = __ldg(bias + 8 * j + 2 * tq); bs[j][1] = __ldg(bias + 8 * j + 2 * tq + 1); }

  unsigned char* so = sOut[warp];
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tile + (int)gridDim.x < ntiles) {
      stage(tile + gridDim.x, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    int tx, ty, fl;
    decode(tile, tx, ty, fl);
    const int oy = ty * ST_TY + warp, x0 = tx * ST_TX;      // one output row per warp
    const float* p0 = &sIn[buf][c0 * ST_IR * ST_IC];        // plane of channel c0 (0 or 2)
#pragma unroll 1
    for (int xh = 0; xh < 2; ++xh) {
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j][0] = bs[j][0]; acc[j][1] = bs[j][1]; acc[j][2] = bs[j][0]; acc[j][3] = bs[j][1]; }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        // (row 2*warp+ky, col 2*(16*xh+g) + tq/2): channels c0, c0+1 of that pixel (channel 3 is zero)
        // K slot kx of the MMA holds w[kx] * x[image column 2 ox - 1 + kx] = patch column 2 ox + kx + 3.  FLIP: the same
        // slot holds w[kx] * flip(x)[2 ox - 1 + kx] = w[kx] * x[2 ox' + 2 - kx] with ox' = Wo - 1 - ox the column this
        // thread computes, i.e. patch column 2 ox' + 6 - kx: the slots (and with them the accumulation order) are those
        // of the unfused "flip the crop, then convolve", so both give the same bits.
        constexpr int KS = FLIP ? -1 : 1;
        const float* q0 = p0 + (2 * warp + ky) * ST_IC + 2 * (16 * xh + g) + (FLIP ? 6 - (tq >> 1) : 3 + (tq >> 1));
        const float* q1 = q0 + ST_IR * ST_IC;
        const bool two = c0 == 0;
        uint32_t a[4];
        a[0] = pack2(q0[0], two ? q1[0] : 0.f);                      // pixel g,    kx = tq/2
        a[1] = pack2(q0[16], two ? q1[16] : 0.f);                    // pixel g+8
        a[2] = pack2(q0[2 * KS], two ? q1[2 * KS] : 0.f);            // pixel g,    kx = tq/2 + 2
        a[3] = pack2(q0[16 + 2 * KS], two ? q1[16 + 2 * KS] : 0.f);  // pixel g+8
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
              : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
              : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfr[ky][j][0]), "r"(bfr[ky][j][1]));
        }
      }
      // ---- ReLU, bf16, transpose through shared memory: thread (g,tq) holds channels 8j+2tq(+1) of pixels g, g+8
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(so + g * ST_OPITCH + j * 16 + tq * 4) = pack2(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f));
        *reinterpret_cast<uint32_t*>(so + (g + 8) * ST_OPITCH + j * 16 + tq * 4) = pack2(fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
      }
      __syncwarp();
      if (oy < Ho) {
        bf16* orow = out + ((size_t)(fl0 + fl) * Ho + oy) * Wo * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int chunk = i * 32 + lane, px = chunk >> 3, part = chunk & 7;
          const int xo = x0 + 16 * xh + px;                  // computed column; the flipped forward stores it mirrored
          if (xo < Wo)
            *reinterpret_cast<uint4*>(orow + (size_t)(FLIP ? Wo - 1 - xo : xo) * 64 + part * 8) =
                *reinterpret_cast<const uint4*>(so + px * ST_OPITCH + part * 16);
        }
      }
      __syncwarp();
    }
    __syncthreads();          // every warp is done with sIn[buf] before the next iteration's prefetch overwrites it
  }
}

// ---------------------------------------------------------------------------------------------
// Fuse-layer sum: out = relu(sum_t term_t[y >> s_t, x >> s_t])  (pose_rsgnet.py:261-270 with
// nn.Upsample(nearest), :213).  One thread per (pixel, 8 channels): 16-byte loads and stores.
// ---------------------------------------------------------------------------------------------
struct FuseP {
  int nterms;
  ResP t[4];
  bf16* out;
  int out_cs, out_co, N, H, W, C, relu;
};

__device__ __forceinline__ void add8(float* a, const uint4& v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[2 * j] += __bfloat162float(h[j].x);
    a[2 * j + 1] += __bfloat162float(h[j].y);
  }
}
__device__ __forceinline__ uint4 pack8(const float* a) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j].x = __float2bfloat16_rn(a[2 * j]);
    h[j].y = __float2bfloat16_rn(a[2 * j + 1]);
  }
  return o;
}

// Flat element space (pixel, 8-channel group) with FUSE_E independent elements per thread: all loads of all terms are
// issued before the first add, so ~16 16-byte loads are in flight per thread (the row-per-block version kept one per
// term in flight and ran at 2.9 TB/s); divisions by runtime constants are multiplications by precomputed magics.
struct FuseIdx { uint32_t c8, W, H, magic_c8, magic_W, magic_H; };
template <int FUSE_E>
__global__ void __launch_bounds__(256) fuse_kernel(const FuseP p, const FuseIdx ix, uint32_t total) {
  const uint32_t e0 = blockIdx.x * (256u * FUSE_E) + threadIdx.x;
  uint4 v[FUSE_E][4];
  uint32_t ooff[FUSE_E];
  bool ok[FUSE_E];
#pragma unroll
  for (int k = 0; k < FUSE_E; ++k) {
    const uint32_t e = e0 + 256u * k;
    ok[k] = e < total;
    const uint32_t ee = ok[k] ? e : 0u;
    const uint32_t pix = ix.magic_c8 ? __umulhi(ee, ix.magic_c8) : ee, c = (ee - pix * ix.c8) * 8u;
    const uint32_t row = ix.magic_W ? __umulhi(pix, ix.magic_W) : pix, x = pix - row * ix.W;
    const uint32_t n = ix.magic_H ? __umulhi(row, ix.magic_H) : row, y = row - n * ix.H;
    ooff[k] = pix * (uint32_t)p.out_cs + c;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t < p.nterms) {
        const ResP& q = p.t[t];
        const bf16* src = q.p + ((size_t)((q.bs0 ? 0u : n) * (uint32_t)q.H + (y >> q.shift)) * (uint32_t)q.W + (x >> q.shift)) * q.cs + q.co + c;
        v[k][t] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < FUSE_E; ++k) {
    if (!ok[k]) continue;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < p.nterms) add8(a, v[k][t]);
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
    }
    *reinterpret_cast<uint4*>(p.out + p.out_co + ooff[k]) = pack8(a);
  }
}

__global__ void __launch_bounds__(256) fuse_rows_kernel(const FuseP p, int rows) {
  const int c8 = p.C >> 3;
  const int per_row = p.W * c8;
  const int row_end = min(rows, ((int)blockIdx.x + 1) * 8);
  for (int row = blockIdx.x * 8; row < row_end; ++row) {
    const int n = row / p.H, y = row - n * p.H;
    const bf16* base[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (t < p.nterms) {
        const ResP& q = p.t[t];
        base[t] = q.p + ((size_t)(q.bs0 ? 0 : n) * q.H + (y >> q.shift)) * q.W * q.cs + q.co;
      }
    }
    bf16* orow = p.out + ((size_t)n * p.H + y) * p.W * p.out_cs + p.out_co;
    for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
      const int x = e / c8, c = (e - x * c8) * 8;
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < p.nterms) {
          const ResP& q = p.t[t];
          add8(a, __ldg(reinterpret_cast<const uint4*>(base[t] + (x >> q.shift) * q.cs + c)));
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
      }
      *reinterpret_cast<uint4*>(orow + (size_t)x * p.out_cs + c) = pack8(a);
    }
  }
}

// 2x2 max-pool, stride 2 (association.py:254, 286-287)
__global__ void __launch_bounds__(256)
maxpool_kernel(const bf16* __restrict__ in, int cs, int co, int N, int H, int W, int C,
               bf16* __restrict__ out) {
  const int c8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * c8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % c8) * 8;
  long long r = i / c8;
  const int x = (int)(r % Wo);
  r /= Wo;
  const int y = (int)(r % Ho);
  const int n = (int)(r / Ho);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(
          in + ((size_t)((size_t)n * H + 2 * y + dy) * W + 2 * x + dx) * cs + co + c));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        m[2 * j] = fmaxf(m[2 * j], __bfloat162float(h[j].x));
        m[2 * j + 1] = fmaxf(m[2 * j + 1], __bfloat162float(h[j].y));
      }
    }
  *reinterpret_cast<uint4*>(out + ((size_t)((size_t)n * Ho + y) * Wo + x) * C + c) = pack8(m);
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(groups, C) over one crop's [S, C] map (association.py:243-245): one CTA per crop,
// pass 1 = per-channel sum / sum of squares (deterministic two-level reduction), pass 2 (L2-hot
// re-read) = normalise + affine, written into a channel slice of the next conv's input.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
groupnorm_kernel(const bf16* __restrict__ in, int in_cs, int in_co, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int groups, float eps, bf16* __restrict__ out,
                 int out_cs, int out_co, int S, int C) {
  extern __shared__ float sm[];
  const int c8 = C >> 3;
  const int nthr = (blockDim.x / c8) * c8;         // threads that own a fixed 8-channel chunk
  float* part = sm;                                // [nthr][16]
  float* chs = sm + (size_t)nthr * 16;             // [C] sums, [C] sums of squares
  float* gstat = chs + 2 * C;                      // [groups] mean, [groups] rstd
  const int n = blockIdx.x, tid = threadIdx.x;
  const bf16* src = in + (size_t)n * S * in_cs + in_co;
  const long long total = (long long)S * c8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (tid < nthr) {
    const int ch = tid % c8;
    for (long long i = tid; i < total; i += nthr) {
      const long long pix = i / c8;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(src + pix * in_cs + ch * 8));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = __bfloat162float(h[j].x), b = __bfloat162float(h[j].y);
        s[2 * j] += a; q[2 * j] = fmaf(a, a, q[2 * j]);
        s[2 * j + 1] += b; q[2 * j + 1] = fmaf(b, b, q[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { part[tid * 16 + j] = s[j]; part[tid * 16 + 8 + j] = q[j]; }
  }
  __syncthreads();
  for (int t = tid; t < 2 * C; t += blockDim.x) {
    const int c = t % C, which = t / C;
    const int ch = c >> 3, j = c & 7;
    float a = 0.f;
    for (int th = ch; th < nthr; th += c8) a += part[th * 16 + which * 8 + j];
    chs[t] = a;
  }
  __syncthreads();
  const int cpg = C / groups;
  if (tid < groups) {
    float a = 0.f, b = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { a += chs[c]; b += chs[C + c]; }
    const float cnt = (float)S * cpg;
    const float mean = a / cnt;
    const float var = fmaxf(b / cnt - mean * mean, 0.f);
    gstat[tid] = mean;
    gstat[groups + tid] = rsqrtf(var + eps);
  }
  __syncthreads();
  bf16* dst = out + (size_t)n * S * out_cs + out_co;
  for (long long i = tid; i < total; i += blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8);
    uint4 v = __ldg(reinterpret_cast<const uint4*>(src + pix * in_cs + ch * 8));
    const bf16* h = reinterpret_cast<const bf16*>(&v);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j, g = c / cpg;
      o[j] = (__bfloat162float(h[j]) - gstat[g]) * gstat[groups + g] * __ldg(gamma + c) + __ldg(beta + c);
    }
    *reinterpret_cast<uint4*>(dst + pix * out_cs + ch * 8) = pack8(o);
  }
}

// bilinear x2, align_corners=True, optional sigmoid (pose_rsgnet.py:1009-1013); fp32 NCHW
__global__ void __launch_bounds__(256)
bilinear2x_kernel(const float* __restrict__ in, float* __restrict__ out, long long NC, int H, int W,
                  int sig) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = NC * Ho * Wo;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int X = (int)(i % Wo);
  long long r = i / Wo;
  const int Y = (int)(r % Ho);
  const long long nc = r / Ho;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * Y, fx = sx * X;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float* p = in + nc * H * W;
  float v = hy * (hx * __ldg(p + y0 * W + x0) + lx * __ldg(p + y0 * W + x1)) +
            ly * (hx * __ldg(p + y1 * W + x0) + lx * __ldg(p + y1 * W + x1));
  if (sig) v = 1.f / (1.f + expf(-v));
  out[i] = v;
}

// relation_scores[n,i,j] = sigmoid(x_i . x_j) in fp32 from the bf16 TRP input (association.py:
// 294-295); only produced on request (it is S*S*4 bytes per crop and unused by the eval loop).
__global__ void __launch_bounds__(256)
relation_scores_kernel(const bf16* __restrict__ x, int cs, int co, int S, int C,
                       float* __restrict__ out) {
  extern __shared__ float sm[];
  const int pitch = C + 1;
  float* xi = sm;                 // [64][C+1]
  float* xj = sm + 64 * pitch;    // [64][C+1]
  const int n = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const bf16* base = x + (size_t)n * S * cs + co;
  for (int t = threadIdx.x; t < 64 * C; t += blockDim.x) {
    const int r = t / C, c = t - r * C;
    xi[r * pitch + c] = (i0 + r < S) ? __bfloat162float(base[(size_t)(i0 + r) * cs + c]) : 0.f;
    xj[r * pitch + c] = (j0 + r < S) ? __bfloat162float(base[(size_t)(j0 + r) * cs + c]) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int c = 0; c < C; ++c) {
    float a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = xi[(ty + 16 * u) * pitch + c]; b[u] = xj[(tx + 16 * u) * pitch + c]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
      if (i < S && j < S) out[((size_t)n * S + i) * S + j] = 1.f / (1.f + expf(-acc[u][v]));
    }
}

// ---------------------------------------------------------------------------------------------
// 1x1 heat-map head: bf16 NHWC [M, Cin] -> fp32 NCHW [N, Cout, H, W] (+bias), Cout <= 32, Cin <= 64
// (final_layer / multi_final_layer, pose_rsgnet.py:961, 1000).  HBM-bound: one thread per pixel, the
// Cin channels in registers, weights broadcast from shared memory, plane-coalesced fp32 stores.
// ---------------------------------------------------------------------------------------------
// Tensor-core version (mma.sync m16n8k16; the fp32-FMA version needed Cout * CIN FMAs per pixel and ran at 2.5 TB/s):
// a warp owns 16 consecutive pixels per step, A fragments come straight from global memory (lane (g, tq) reads the two
// bf16 pairs of pixel rows g and g+8 it needs: four lanes cover one 16-byte run), the weights live in registers as B
// fragments, and every accumulator register is one 32-byte run of a channel plane (8 consecutive pixels).
template <int CIN, int NT>
__global__ void __launch_bounds__(256)
head1x1_kernel(const bf16* __restrict__ in, int in_cs, int in_co, const bf16* __restrict__ w, int w_ld,
               const float* __restrict__ bias, int Cout, unsigned M, unsigned HW, float* __restrict__ out, int relu) {
  constexpr int KS = CIN / 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  // B fragments: b0 = W[n = 8j+g][k = 16ks + 2tq, +1], b1 = the same at k + 8
  uint32_t bfr[NT][KS][2];
  float bs[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int n = 8 * j + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t b0 = 0, b1 = 0;
      if (n < Cout) {
        b0 = *reinterpret_cast<const uint32_t*>(w + (size_t)n * w_ld + 16 * ks + 2 * tq);
        b1 = *reinterpret_cast<const uint32_t*>(w + (size_t)n * w_ld + 16 * ks + 2 * tq + 8);
      }
      bfr[j][ks][0] = b0; bfr[j][ks][1] = b1;
    }
    bs[j][0] = (8 * j + 2 * tq < Cout) ? __ldg(bias + 8 * j + 2 * tq) : 0.f;
    bs[j][1] = (8 * j + 2 * tq + 1 < Cout) ? __ldg(bias + 8 * j + 2 * tq + 1) : 0.f;
  }
  const unsigned ntiles = (M + 15u) >> 4;
  for (unsigned tile = warp_global; tile < ntiles; tile += nwarps) {
    const unsigned m0 = tile * 16u + g, m1 = m0 + 8u;
    const bool ok0 = m0 < M, ok1 = m1 < M;
    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(in + (size_t)(ok0 ? m0 : 0u) * in_cs + in_co) + tq;
    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(in + (size_t)(ok1 ? m1 : 0u) * in_cs + in_co) + tq;
    uint32_t a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      a[ks][0] = __ldg(r0 + 8 * ks); a[ks][1] = __ldg(r1 + 8 * ks);
      a[ks][2] = __ldg(r0 + 8 * ks + 4); a[ks][3] = __ldg(r1 + 8 * ks + 4);
    }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = bs[j][0]; acc[j][1] = bs[j][1]; acc[j][2] = bs[j][0]; acc[j][3] = bs[j][1]; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int j = 0; j < NT; ++j)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
            : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
            : "r"(a[ks][0]), "r"(a[ks][1]), "r"(a[ks][2]), "r"(a[ks][3]), "r"(bfr[j][ks][0]), "r"(bfr[j][ks][1]));
    const unsigned n0 = m0 / HW, n1 = m1 / HW;
    float* o0 = out + (size_t)n0 * Cout * HW + (m0 - n0 * HW);
    float* o1 = out + (size_t)n1 * Cout * HW + (m1 - n1 * HW);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * j + 2 * tq + e;
        if (c < Cout) {
          float v0 = acc[j][e], v1 = acc[j][2 + e];
          if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          if (ok0) o0[(size_t)c * HW] = v0;
          if (ok1) o1[(size_t)c * HW] = v1;
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// TRP tail (association.py:236-245, 300): z = W y + b (1x1 conv C -> C), out = GroupNorm(groups, C)(z), all in fp32.
// y = sum_j sigmoid(x_i . x_j) g_j has |mean| >> spread over the positions of a crop and GroupNorm subtracts the mean,
// so neither y, W nor z may be rounded to bf16 on the way (2^-9 |mean| of noise on a signal of size `spread`).
// One CTA per crop, one thread per position: pass 1 accumulates the per-group sum / sum of squares of z (fp32 per
// thread over a few dozen values, fp64 across threads), pass 2 recomputes z (y is L2-hot; cheaper than a round trip of
// z through HBM) and writes the normalised bf16 output.  W is read as 16-byte broadcast loads from shared memory.
template <int C, int G, int PPT>
__global__ void __launch_bounds__(256, 2)
trp_tail_kernel(const float* __restrict__ y32, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                bf16* __restrict__ out, int out_cs, int out_co, int S) {
  constexpr int groups = G, cpg = C / G;
  __shared__ __align__(16) float sW[C * C];
  __shared__ float sB[C], sG[C], sBe[C];
  __shared__ double red[8][2 * G];                  // [warp][sum g, sumsq g]
  __shared__ float gmean[G], grstd[G];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C * C; i += 256) sW[i] = w[i];
  for (int i = tid; i < C; i += 256) { sB[i] = bias[i]; sG[i] = gamma[i]; sBe[i] = beta[i]; }
  __syncthreads();
  const float* src = y32 + (size_t)n * S * C;
  float gs[G], gq[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { gs[g] = 0.f; gq[g] = 0.f; }

  // PPT positions at once (p, p + 256, ...): every 16-byte broadcast load of W feeds 4 * PPT FMAs; z is produced eight
  // channels at a time and consumed at once, so only y stays in registers.  Positions past the end of the crop are
  // clamped (computed, never used).
  auto load_y = [&](int p, float (&yv)[PPT][C]) {
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int pp = min(p + 256 * q, S - 1);
      const float4* yp = reinterpret_cast<const float4*>(src + (size_t)pp * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 v = __ldg(yp + k);
        yv[q][4 * k] = v.x; yv[q][4 * k + 1] = v.y; yv[q][4 * k + 2] = v.z; yv[q][4 * k + 3] = v.w;
      }
    }
  };
  auto z_chunk = [&](const float (&yv)[PPT][C], int c8, float (&z)[PPT][8]) {
    // (a packed fma.rn.f32x2 version with {w, w} weight pairs in shared memory was measured SLOWER: 411 vs 305 us)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
#pragma unroll
      for (int q = 0; q < PPT; ++q) z[q][j] = sB[c];
      const float4* wr = reinterpret_cast<const float4*>(sW + c * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 ww = wr[k];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          z[q][j] = fmaf(ww.x, yv[q][4 * k], z[q][j]); z[q][j] = fmaf(ww.y, yv[q][4 * k + 1], z[q][j]);
          z[q][j] = fmaf(ww.z, yv[q][4 * k + 2], z[q][j]); z[q][j] = fmaf(ww.w, yv[q][4 * k + 3], z[q][j]);
        }
      }
    }
  };

  for (int p = tid; p < S; p += 256 * PPT) {
    float yv[PPT][C];
    load_y(p, yv);
#pragma unroll
    for (int c8 = 0; c8 < C / 8; ++c8) {
      float z[PPT][8];
      z_chunk(yv, c8, z);
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        if (p + 256 * q < S) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int g = (c8 * 8 + j) / cpg;            // compile-time after unrolling: gs / gq stay in registers
            gs[g] += z[q][j];
            gq[g] = fmaf(z[q][j], z[q][j], gq[g]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    double a = (double)gs[g], b = (double)gq[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { red[warp][g] = a; red[warp][G + g] = b; }
  }
  __syncthreads();
  if (tid < groups) {
    double a = 0.0, b = 0.0;
    for (int wv = 0; wv < 8; ++wv) { a += red[wv][tid]; b += red[wv][G + tid]; }
    const double cnt = (double)S * cpg, mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    gmean[tid] = (float)mean;
    grstd[tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  bf16* dst = out + (size_t)n * S * out_cs + out_co;
  for (int p = tid; p < S; p += 256 * PPT) {
    float yv[PPT][C];
    load_y(p, yv);
#pragma unroll
    for (int c8 = 0; c8 < C / 8; ++c8) {
      float z[PPT][8];
      z_chunk(yv, c8, z);
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        const int pp = p + 256 * q;
        if (pp < S) {
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = c8 * 8 + 2 * j, g0 = c / cpg, g1 = (c + 1) / cpg;
            const float a = (z[q][2 * j] - gmean[g0]) * grstd[g0] * sG[c] + sBe[c];
            const float b = (z[q][2 * j + 1] - gmean[g1]) * grstd[g1] * sG[c + 1] + sBe[c + 1];
            h[j] = __floats2bfloat162_rn(a, b);
          }
          *reinterpret_cast<uint4*>(dst + (size_t)pp * out_cs + c8 * 8) = o;
        }
      }
    }
  }
}

__global__ void cvt_f32_kernel(const bf16* __restrict__ in, int cs, int co, float* __restrict__ out, long long rows, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  out[i] = __bfloat162float(in[r * cs + co + c]);
}

}  // namespace

int head1x1_launch(const ConvP& p, cudaStream_t s, int* handled) {
  *handled = 0;
  if (!p.out_f32 || p.out || p.ntaps != 1 || p.dy[0] != 0 || p.dx[0] != 0 || p.stride != 1 || p.omul != 1 ||
      p.nres != 0 || p.Cout > 32 || p.in_cs % 8 != 0 || p.in_co % 8 != 0 || p.M >= (1ll << 31))
    return RSG_OK;
  if (p.Cin != 16 && p.Cin != 32 && p.Cin != 48 && p.Cin != 64) return RSG_OK;
  *handled = 1;
  if (p.M == 0) return RSG_OK;
  const int HW = p.Hout * p.Wout;
  // persistent-ish: 8 CTAs of 8 warps per SM walk the 16-pixel tiles
  long long nblk = ceil_div(ceil_div(p.M, 16), 8);
  if (nblk > 8ll * rsg_num_sms()) nblk = 8ll * rsg_num_sms();
  dim3 grid((unsigned)nblk);
  RSG_REQUIRE(p.CinPad % 2 == 0 && ((uintptr_t)p.w % 4) == 0 && ((uintptr_t)p.in % 4) == 0, "head1x1: unaligned operand");
  const int nt = (p.Cout + 7) / 8;
#define RSG_HEAD2(C, T) head1x1_kernel<C, T><<<grid, 256, 0, s>>>(p.in, p.in_cs, p.in_co, p.w, p.CinPad, p.bias, p.Cout, (unsigned)p.M, (unsigned)HW, p.out_f32, p.relu)
#define RSG_HEAD(C) do { if (nt <= 2) RSG_HEAD2(C, 2); else if (nt == 3) RSG_HEAD2(C, 3); else RSG_HEAD2(C, 4); } while (0)
  switch (p.Cin) {
    case 16: RSG_HEAD(16); break;
    case 32: RSG_HEAD(32); break;
    case 48: RSG_HEAD(48); break;
    default: RSG_HEAD(64); break;
  }
#undef RSG_HEAD2
#undef RSG_HEAD
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int stem_launch(cudaStream_t s, const float* x, int H, int W, const float* w, const float* bias,
                bf16* out, int f0, int nb, int n_crops) {
  RSG_REQUIRE(H % 2 == 0 && W % 4 == 0, "stem: H must be even and W a multiple of 4");
  RSG_REQUIRE(((uintptr_t)x % 16) == 0, "stem: the input must be 16-byte aligned");
  if (nb == 0) return RSG_OK;
  const int tiles_x = ceil_div(W / 2, ST_TX), tiles_y = ceil_div(H / 2, ST_TY);
  // forwards [f0, f0 + nb) of the run: the first n_unf read their crop as it is, the rest W-flipped (f >= n_crops)
  int n_unf = n_crops - f0;
  if (n_unf < 0) n_unf = 0;
  if (n_unf > nb) n_unf = nb;
  for (int part = 0; part < 2; ++part) {
    const int cnt = part == 0 ? n_unf : nb - n_unf, fl0 = part == 0 ? 0 : n_unf;
    if (cnt == 0) continue;
    const long long nblk = (long long)tiles_x * tiles_y * cnt;
    RSG_REQUIRE(nblk < (1ll << 31), "stem: too many tiles");
    int grid = 2 * rsg_num_sms();                      // persistent: two CTAs per SM walk the tiles
    if (grid > nblk) grid = (int)nblk;
    if (part == 0) stem_kernel<false><<<grid, 256, 0, s>>>(x, H, W, w, bias, out, f0, fl0, n_crops, tiles_x, tiles_y, (int)nblk);
    else stem_kernel<true><<<grid, 256, 0, s>>>(x, H, W, w, bias, out, f0, fl0, n_crops, tiles_x, tiles_y, (int)nblk);
    RSG_LAUNCH_CHECK();
  }
  return RSG_OK;
}

int fuse_launch(cudaStream_t s, int nterms, const ResP* terms, bf16* out, int out_cs, int out_co,
                int N, int H, int W, int C, int relu) {
  RSG_REQUIRE(nterms >= 1 && nterms <= 4 && C % 8 == 0, "fuse: nterms=%d C=%d", nterms, C);
  FuseP p;
  p.nterms = nterms;
  for (int i = 0; i < nterms; ++i) p.t[i] = terms[i];
  p.out = out; p.out_cs = out_cs; p.out_co = out_co; p.N = N; p.H = H; p.W = W; p.C = C; p.relu = relu;
  const long long total = (long long)N * H * W * (C / 8);
  if (total == 0) return RSG_OK;
  RSG_REQUIRE(total < (1ll << 31) && (long long)N * H * W * out_cs < (1ll << 32), "fuse: tensor too large for 32-bit indexing");
  auto magic = [](uint32_t d) { return d > 1 ? (uint32_t)(((1ull << 32) + d - 1) / d) : 0u; };   // exact while n * d < 2^32
  FuseIdx ix{(uint32_t)(C / 8), (uint32_t)W, (uint32_t)H, magic((uint32_t)(C / 8)), magic((uint32_t)W), magic((uint32_t)H)};
  RSG_REQUIRE((unsigned long long)total * (unsigned long long)(C / 8 > W ? (C / 8 > H ? C / 8 : H) : (W > H ? W : H)) < (1ull << 32),
              "fuse: tensor too large for the division magics");
  const int fe = rsg_dbg_int("RSG_FUSE_E", 1);      // measured (tools/bench_ops.py fuse): 1 element per thread 51 us, 2: 59, 4: 95, row kernel: 72
  if (fe == 0) {
    const long long rows = (long long)N * H;
    const int per_row = W * (C / 8);
    const int threads = per_row >= 256 ? 256 : (per_row + 31) / 32 * 32;
    fuse_rows_kernel<<<(unsigned)((rows + 8 - 1) / 8), threads, 0, s>>>(p, (int)rows);
  } else if (fe == 1) fuse_kernel<1><<<(unsigned)((total + 256 - 1) / 256), 256, 0, s>>>(p, ix, (uint32_t)total);
  else if (fe == 2) fuse_kernel<2><<<(unsigned)((total + 512 - 1) / 512), 256, 0, s>>>(p, ix, (uint32_t)total);
  else fuse_kernel<4><<<(unsigned)((total + 1024 - 1) / 1024), 256, 0, s>>>(p, ix, (uint32_t)total);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int maxpool_launch(cudaStream_t s, const bf16* in, int cs, int co, int N, int H, int W, int C,
                   bf16* out) {
  RSG_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: bad shape");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return RSG_OK;
  maxpool_kernel<<<ceil_div(total, 256), 256, 0, s>>>(in, cs, co, N, H, W, C, out);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int groupnorm_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, const float* gamma,
                     const float* beta, int groups, float eps, bf16* out, int out_cs, int out_co,
                     int N, int S, int C) {
  RSG_REQUIRE(C % 8 == 0 && C % groups == 0 && C / 8 <= 256, "groupnorm: C=%d groups=%d", C, groups);
  if (N == 0) return RSG_OK;
  const int c8 = C / 8, nthr = (256 / c8) * c8;
  size_t smem = ((size_t)nthr * 16 + 2 * C + 2 * groups) * sizeof(float);
  groupnorm_kernel<<<N, 256, smem, s>>>(in, in_cs, in_co, gamma, beta, groups, eps, out, out_cs,
                                        out_co, S, C);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}


int trp_tail_launch(cudaStream_t s, const float* y32, const float* w, const float* bias, const float* gamma, const float* beta,
                    int groups, float eps, bf16* out, int out_cs, int out_co, int N, int S, int C) {
  RSG_REQUIRE(groups == 8 && C % 8 == 0, "trp_tail: GroupNorm(8, C) only (C=%d groups=%d)", C, groups);
  RSG_REQUIRE(out_cs % 8 == 0 && out_co % 8 == 0, "trp_tail: output channel stride/offset must be multiples of 8");
  if (N == 0) return RSG_OK;
  switch (C) {
    case 16: trp_tail_kernel<16, 8, 2><<<N, 256, 0, s>>>(y32, w, bias, gamma, beta, eps, out, out_cs, out_co, S); break;
    case 32: trp_tail_kernel<32, 8, 2><<<N, 256, 0, s>>>(y32, w, bias, gamma, beta, eps, out, out_cs, out_co, S); break;
    case 48: trp_tail_kernel<48, 8, 1><<<N, 256, 0, s>>>(y32, w, bias, gamma, beta, eps, out, out_cs, out_co, S); break;
    case 64: trp_tail_kernel<64, 8, 1><<<N, 256, 0, s>>>(y32, w, bias, gamma, beta, eps, out, out_cs, out_co, S); break;
    default: rsg_set_error("trp_tail: unsupported channel count %d (16/32/48/64)", C); return RSG_ERR_ARG;
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int cvt_f32_launch(cudaStream_t s, const bf16* in, int cs, int co, float* out, long long rows, int C) {
  if (rows == 0) return RSG_OK;
  cvt_f32_kernel<<<ceil_div(rows * C, 256), 256, 0, s>>>(in, cs, co, out, rows, C);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int bilinear2x_launch(cudaStream_t s, const float* in, float* out, int NC, int H, int W, int sig) {
  const long long total = (long long)NC * 4 * H * W;
  if (total == 0) return RSG_OK;
  bilinear2x_kernel<<<ceil_div(total, 256), 256, 0, s>>>(in, out, NC, H, W, sig);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

int relation_scores_launch(cudaStream_t s, const bf16* x, int cs, int co, int N, int S, int C,
                           float* out) {
  RSG_REQUIRE(C <= 96, "relation_scores: C=%d too large", C);
  if (N == 0) return RSG_OK;
  dim3 grid(ceil_div(S, 64), ceil_div(S, 64), N);
  size_t smem = (size_t)2 * 64 * (C + 1) * sizeof(float);
  if (smem > 48 * 1024)
    RSG_CUDA(cudaFuncSetAttribute(relation_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  relation_scores_kernel<<<grid, 256, smem, s>>>(x, cs, co, S, C, out);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
