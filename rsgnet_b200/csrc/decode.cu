// Flip-test averaging + heat-map decode: one HBM pass, one warp per (crop, joint) map.
//
// Replaces (paths under the reference checkout):
//   lib/core/function.py:417-427     flip_back -> 1-px shift -> (a+b)*0.5
//   lib/core/inference.py:21-49      get_max_preds (first arg-max, >0 mask)
//   lib/core/inference.py:52-82      get_final_preds (+-0.25 offsets, per-crop inverse affine)
//   lib/utils/transforms.py:23-37    flip_back
//   lib/utils/transforms.py:57-103   transform_preds / get_affine_transform(inv=1)
//
// Bit-exactness: every fp32/fp64 operation below uses an explicit round-to-nearest intrinsic so
// that nvcc cannot contract mul+add into FMA; the order of operations follows the reference.
// HBM traffic: N*K*H*W*4 B read per input tensor (once), 12 B written per map.
#include "common.cuh"
#include "../../include/rsg_b200.h"

namespace {

struct Best { float v; int i; };

__device__ __forceinline__ void take(Best& b, float v, int i) {
  // ascending scan inside a lane: strict '>' keeps the first maximum
  if (v > b.v) { b.v = v; b.i = i; }
}

__device__ __forceinline__ Best warp_argmax(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (ov > b.v || (ov == b.v && oi < b.i)) { b.v = ov; b.i = oi; }
  }
  return b;
}

// source column of the flipped map that lands on column x after flip_back (+ optional shift)
__device__ __forceinline__ int flip_src_x(int x, int W, int shift) {
  if (!shift) return W - 1 - x;
  return x == 0 ? W - 1 : W - x;
}

template <bool FLIP>
__device__ __forceinline__ float value_at(const float* __restrict__ a, const float* __restrict__ b,
                                          int y, int x, int W, int shift) {
  float va = __ldg(a + y * W + x);
  if (!FLIP) return va;
  float vb = __ldg(b + y * W + flip_src_x(x, W, shift));
  return __fmul_rn(__fadd_rn(va, vb), 0.5f);
}

template <bool FLIP>
__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ hm, const float* __restrict__ hmf,
              const int32_t* __restrict__ perm, int NK, int K, int H, int W,
              const float* __restrict__ center, const float* __restrict__ scale,
              int post_process, int shift, float* __restrict__ preds,
              float* __restrict__ maxvals, float* __restrict__ coords,
              float* __restrict__ avg_out) {
  const int lane = threadIdx.x & 31;
  const int map = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (map >= NK) return;
  const int n = map / K, k = map - n * K;
  const int HW = H * W;
  const float* a = hm + (size_t)map * HW;
  const float* b = nullptr;
  if (FLIP) b = hmf + ((size_t)n * K + __ldg(perm + k)) * HW;
  float* av = avg_out ? avg_out + (size_t)map * HW : nullptr;

  Best best; best.v = -INFINITY; best.i = 0x7fffffff;
  if ((W & 3) == 0 && (((size_t)a & 15) == 0)) {
    const int nv = HW >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
#pragma unroll 4
    for (int i4 = lane; i4 < nv; i4 += 32) {
      float4 va = __ldg(a4 + i4);
      float v[4] = {va.x, va.y, va.z, va.w};
      const int i = i4 << 2;
      if (FLIP) {
        const int y = i / W, x = i - y * W;
        const float* brow = b + y * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float vb = __ldg(brow + flip_src_x(x + j, W, shift));
          v[j] = __fmul_rn(__fadd_rn(v[j], vb), 0.5f);
        }
        if (av) *reinterpret_cast<float4*>(av + i) = make_float4(v[0], v[1], v[2], v[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) take(best, v[j], i + j);
    }
  } else {
    for (int i = lane; i < HW; i += 32) {
      const int y = i / W, x = i - y * W;
      float v = value_at<FLIP>(a, b, y, x, W, shift);
      if (FLIP && av) av[i] = v;
      take(best, v, i);
    }
  }
  if (best.i == 0x7fffffff) { best.i = 0; }   // all -inf / NaN maps: index 0 like np.argmax
  best = warp_argmax(best);
  if (lane != 0) return;

  const int py = best.i / W, px = best.i - py * W;
  const float mv = best.v;
  float cx = (float)px, cy = (float)py;
  if (!(mv > 0.0f)) { cx = 0.0f; cy = 0.0f; }          // preds *= (maxvals > 0)
  if (post_process) {
    const int qx = (int)floorf(cx + 0.5f), qy = (int)floorf(cy + 0.5f);
    if (qx > 1 && qx < W - 1 && qy > 1 && qy < H - 1) {
      float dx = __fsub_rn(value_at<FLIP>(a, b, qy, qx + 1, W, shift),
                           value_at<FLIP>(a, b, qy, qx - 1, W, shift));
      float dy = __fsub_rn(value_at<FLIP>(a, b, qy + 1, qx, W, shift),
                           value_at<FLIP>(a, b, qy - 1, qx, W, shift));
      float sx = dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
      float sy = dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
      cx = __fadd_rn(cx, sx);
      cy = __fadd_rn(cy, sy);
    }
  }
  maxvals[map] = mv;
  if (coords) { coords[2 * map] = cx; coords[2 * map + 1] = cy; }
  if (preds) {
    // get_affine_transform(center, scale, rot=0, [W,H], inv=1) in closed form, keeping the
    // reference's fp32 roundings of the three control points (transforms.py:65-97).
    const float ccx = __ldg(center + 2 * n), ccy = __ldg(center + 2 * n + 1);
    const float sw = __fmul_rn(__ldg(scale + 2 * n), 200.0f);
    const float d = __fmul_rn(sw, -0.5f);
    const float q1y = (float)__dadd_rn((double)ccy, (double)d);
    const float dd = __fsub_rn(ccy, q1y);
    const float q2x = __fsub_rn(ccx, dd);
    const double half_w = (double)W * 0.5, half_h = (double)H * 0.5;
    const double a11 = __ddiv_rn(__dsub_rn((double)ccx, (double)q2x), half_w);
    const double a22 = __ddiv_rn(__dsub_rn((double)ccy, (double)q1y), half_w);
    const double b1 = __dsub_rn((double)ccx, __dmul_rn(a11, half_w));
    const double b2 = __dsub_rn((double)ccy, __dmul_rn(a22, half_h));
    preds[2 * map] = (float)__dadd_rn(__dmul_rn(a11, (double)cx), b1);
    preds[2 * map + 1] = (float)__dadd_rn(__dmul_rn(a22, (double)cy), b2);
  }
}

__global__ void flip_back_kernel(const float* __restrict__ in, float* __restrict__ out,
                                 const int32_t* __restrict__ perm, int K, int H, int W,
                                 long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int x = (int)(i % W);
  long long r = i / W;
  int y = (int)(r % H);
  r /= H;
  int k = (int)(r % K);
  long long n = r / K;
  out[i] = __ldg(in + ((n * K + perm[k]) * H + y) * (long long)W + (W - 1 - x));
}

}  // namespace

extern "C" int rsg_flip_avg_decode(void* stream, const float* hm, const float* hm_flipped,
                                   const int32_t* flip_perm, int N, int K, int H, int W,
                                   const float* center, const float* scale, int post_process,
                                   int shift, float* preds, float* maxvals, float* coords,
                                   float* avg_out) {
  RSG_REQUIRE(N >= 0 && K > 0 && H > 0 && W > 0, "rsg_flip_avg_decode: bad shape N=%d K=%d H=%d W=%d", N, K, H, W);
  RSG_REQUIRE((long long)H * W < (1ll << 24), "rsg_flip_avg_decode: H*W must be < 2^24");
  if (N == 0) return RSG_OK;
  RSG_REQUIRE(hm && maxvals, "rsg_flip_avg_decode: hm and maxvals are required");
  RSG_REQUIRE(!preds || (center && scale), "rsg_flip_avg_decode: preds needs center and scale");
  RSG_REQUIRE(!hm_flipped || flip_perm, "rsg_flip_avg_decode: hm_flipped needs flip_perm");
  RSG_REQUIRE(!avg_out || hm_flipped, "rsg_flip_avg_decode: avg_out only with hm_flipped");
  const long long NK = (long long)N * K;
  RSG_REQUIRE(NK < (1ll << 31), "rsg_flip_avg_decode: N*K too large");
  const int warps = 8;
  dim3 grid(ceil_div(NK, warps)), block(warps * 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (hm_flipped)
    decode_kernel<true><<<grid, block, 0, s>>>(hm, hm_flipped, flip_perm, (int)NK, K, H, W, center,
                                               scale, post_process, shift, preds, maxvals, coords,
                                               avg_out);
  else
    decode_kernel<false><<<grid, block, 0, s>>>(hm, nullptr, nullptr, (int)NK, K, H, W, center,
                                                scale, post_process, shift, preds, maxvals, coords,
                                                nullptr);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_flip_back(void* stream, const float* in, float* out, const int32_t* flip_perm,
                             int N, int K, int H, int W) {
  RSG_REQUIRE(N >= 0 && K > 0 && H > 0 && W > 0, "rsg_flip_back: bad shape");
  if (N == 0) return RSG_OK;
  RSG_REQUIRE(in && out && flip_perm && in != out, "rsg_flip_back: null or aliased pointers");
  long long total = (long long)N * K * H * W;
  flip_back_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, flip_perm, K,
                                                                          H, W, total);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
