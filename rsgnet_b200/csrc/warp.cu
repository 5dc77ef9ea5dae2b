// Person-crop affine warp + normalisation on the device: what the reference's data loader does per
// crop on a CPU worker (lib/dataset/CPJointsDataset.py:1281-1290: cv2.warpAffine(img, trans, (W, H), INTER_LINEAR);
// tools/cp_test.py:107-115: ToTensor + Normalize), for a whole batch of crops in one launch.
//
// Bit-exact with OpenCV's 8-bit INTER_LINEAR / BORDER_CONSTANT(0) warpAffine (restated in oracle/warp_oracle.py, which
// is pinned to cv2 + the reference's crop()): the forward 2x3 matrix is inverted in fp64 with separate multiplies and
// adds (explicit _rn intrinsics: no FMA contraction), source coordinates are 10-bit fixed point rounded half-to-even,
// the 5-bit fractions select 15-bit bilinear weights ((0,0) -> (32767, 0, 0, 1), OpenCV's table quirk), taps outside
// the source read 0 and dst = (sum + 16384) >> 15.  Integer / byte work, HBM-bound: per crop it reads the source
// footprint once (L2 absorbs the 2x2 tap overlap) and writes C*H*W bytes (u8 HWC) and/or 4*C*H*W bytes (normalised
// f32 NCHW, the model's input layout) through a 256-entry-per-channel table built by the host from the reference's
// fp32 formula.
//
// Launch: one CTA of 64 x 4 threads per (64-pixel column strip x 32 rows, crop), walking its rows four at a time; the inverse
// matrix of the crop and the table are set up once per CTA (with one CTA per 64 x 4 pixels the serial fp64 inversion +
// barrier in front of 256 pixels of work was most of the kernel: 111 us per 256 crops).
#include "common.cuh"
#include "../../include/rsg_b200.h"

namespace {

constexpr int WARP_BX = 64, WARP_BY = 4, WARP_ROWS = 32;      // rows per CTA (WARP_BY at a time)

struct WarpP {
  const uint8_t* const* src;     // [N] device pointers, HWC uint8, 3 channels
  const int32_t* dims;           // [N][3] rows, cols, row stride in bytes
  const double* mats;            // [N][6] forward matrix (dst = M * src), as passed to cv2.warpAffine
  int N, H, W;
  uint8_t* out_u8;               // [N][H][W][3] or nullptr
  float* out_f32;                // [N][3][H][W] or nullptr
  const float* lut;              // [3][256], indexed by OUTPUT channel
  int reverse;                   // 1: output channel k = source channel 2-k (BGR -> RGB)
};

__global__ void __launch_bounds__(WARP_BX * WARP_BY) warp_affine_kernel(const WarpP p) {
  __shared__ double sM[6];
  __shared__ float sLut[3 * 256];
  const int n = blockIdx.z;
  const int tid = threadIdx.y * WARP_BX + threadIdx.x;
  if (tid == 0) {
    const double* m = p.mats + (size_t)n * 6;
    const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
    double d = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    d = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
    const double a11 = __dmul_rn(m4, d), a22 = __dmul_rn(m0, d);
    const double i0 = a11, i1 = __dmul_rn(m1, -d), i3 = __dmul_rn(m3, -d), i4 = a22;
    sM[0] = i0; sM[1] = i1; sM[3] = i3; sM[4] = i4;
    sM[2] = __dsub_rn(__dmul_rn(-i0, m2), __dmul_rn(i1, m5));
    sM[5] = __dsub_rn(__dmul_rn(-i3, m2), __dmul_rn(i4, m5));
  }
  if (p.out_f32)
    for (int i = tid; i < 3 * 256; i += WARP_BX * WARP_BY) sLut[i] = p.lut[i];
  __syncthreads();
  // per-row terms of the CTA's 32 rows, once per row instead of once per pixel (fp64 multiplies + conversions are
  // slow-pipe instructions): X0 = round((M1*y + M2)*1024) + 16, Y0 = round((M4*y + M5)*1024) + 16
  __shared__ int sXY[WARP_ROWS][2];
  if (tid < WARP_ROWS) {
    const double y = (double)(blockIdx.y * WARP_ROWS + tid);
    sXY[tid][0] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(sM[1], y), sM[2]), 1024.0)) + 16;
    sXY[tid][1] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(sM[4], y), sM[5]), 1024.0)) + 16;
  }
  __syncthreads();
  const int x = blockIdx.x * WARP_BX + threadIdx.x;
  if (x >= p.W) return;
  const int rows = p.dims[n * 3 + 0], cols = p.dims[n * 3 + 1], pitch = p.dims[n * 3 + 2];
  const uint8_t* __restrict__ src = p.src[n];
  // WarpAffineInvoker: adelta[x] = round(M0*x*1024), X0 = round((M1*y + M2)*1024) + 16  (AB_BITS = 10, INTER_BITS = 5)
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(sM[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(sM[3], (double)x), 1024.0));
  const size_t plane = (size_t)p.H * p.W;
  const int y_end = min(p.H, ((int)blockIdx.y + 1) * WARP_ROWS);
  // Two pixels (rows y and y + WARP_BY) per thread and iteration, branch-free: the taps are read from clamped
  // coordinates and an invalid tap gets weight 0, so the 24 byte loads of an iteration are independent and in flight
  // together (one pixel per thread kept ~15 KB per SM in flight: latency-bound at 1.8 TB/s).
  for (int yb = blockIdx.y * WARP_ROWS + threadIdx.y; yb < y_end; yb += 2 * WARP_BY) {
    int v[2][3];
    bool live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int y = yb + k * WARP_BY;
      live[k] = y < y_end;
      const int yr = live[k] ? y - blockIdx.y * WARP_ROWS : 0;
      const int X0 = sXY[yr][0], Y0 = sXY[yr][1];
      const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
      const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
      const int fx = X & 31, fy = Y & 31;
      int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
      if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
      const bool x0ok = sx >= 0 && sx < cols, x1ok = sx + 1 >= 0 && sx + 1 < cols;
      const bool y0ok = sy >= 0 && sy < rows, y1ok = sy + 1 >= 0 && sy + 1 < rows;
      w00 = (y0ok && x0ok) ? w00 : 0; w01 = (y0ok && x1ok) ? w01 : 0;
      w10 = (y1ok && x0ok) ? w10 : 0; w11 = (y1ok && x1ok) ? w11 : 0;
      const int cx0 = min(max(sx, 0), cols - 1), cx1 = min(max(sx + 1, 0), cols - 1);
      const int cy0 = min(max(sy, 0), rows - 1), cy1 = min(max(sy + 1, 0), rows - 1);
      const uint8_t* r0 = src + (size_t)cy0 * pitch;
      const uint8_t* r1 = src + (size_t)cy1 * pitch;
      const uint8_t* q00 = r0 + cx0 * 3; const uint8_t* q01 = r0 + cx1 * 3;
      const uint8_t* q10 = r1 + cx0 * 3; const uint8_t* q11 = r1 + cx1 * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[k][c] = w00 * (int)__ldg(q00 + c) + w01 * (int)__ldg(q01 + c) + w10 * (int)__ldg(q10 + c) + w11 * (int)__ldg(q11 + c);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!live[k]) continue;
      const int y = yb + k * WARP_BY;
      uint8_t u[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) u[c] = (uint8_t)min(255, (v[k][c] + (1 << 14)) >> 15);
      if (p.reverse) { const uint8_t t = u[0]; u[0] = u[2]; u[2] = t; }
      if (p.out_u8) {
        uint8_t* o = p.out_u8 + (((size_t)n * p.H + y) * p.W + x) * 3;
        o[0] = u[0]; o[1] = u[1]; o[2] = u[2];
      }
      if (p.out_f32) {
        float* o = p.out_f32 + (size_t)n * 3 * plane + (size_t)y * p.W + x;
        o[0] = sLut[u[0]]; o[plane] = sLut[256 + u[1]]; o[2 * plane] = sLut[512 + u[2]];
      }
    }
  }
}

}  // namespace

extern "C" int rsg_warp_affine(void* stream, const uint8_t* const* src, const int32_t* src_dims, const double* mats, int N,
                               int out_h, int out_w, int reverse_channels, uint8_t* out_u8, float* out_f32,
                               const float* lut) {
  RSG_REQUIRE(N >= 0 && out_h >= 1 && out_w >= 1, "rsg_warp_affine: bad sizes N=%d out=%dx%d", N, out_h, out_w);
  RSG_REQUIRE(N <= 65535, "rsg_warp_affine: at most 65535 crops per call (got %d)", N);
  if (N == 0) return RSG_OK;
  RSG_REQUIRE(src && src_dims && mats, "rsg_warp_affine: null input");
  RSG_REQUIRE(out_u8 || out_f32, "rsg_warp_affine: no output requested");
  RSG_REQUIRE(!out_f32 || lut, "rsg_warp_affine: the f32 output needs the normalisation table");
  WarpP p;
  p.src = src; p.dims = src_dims; p.mats = mats; p.N = N; p.H = out_h; p.W = out_w;
  p.out_u8 = out_u8; p.out_f32 = out_f32; p.lut = lut; p.reverse = reverse_channels ? 1 : 0;
  dim3 grid((unsigned)ceil_div(out_w, WARP_BX), (unsigned)ceil_div(out_h, WARP_ROWS), (unsigned)N);
  warp_affine_kernel<<<grid, dim3(WARP_BX, WARP_BY), 0, (cudaStream_t)stream>>>(p);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
