// Training-step element-wise, normalisation, re-sampling, loss and optimiser kernels.  All tensors are
// fp32; activations are NHWC = row-major [M pixels, C channels].  Reference semantics (paths under the reference):
//   BatchNorm2d / BatchNorm1d in train mode  torch batch_norm: biased variance for the output, unbiased for running_var,
//                                            running = (1 - momentum) running + momentum batch   (pose_rsgnet.py BN_MOMENTUM 0.1)
//   GroupNorm(8, C)                          association.py:243-246
//   nearest / bilinear(align_corners=True)   pose_rsgnet.py:207-219 (fuse layers), :1009-1012 (x2 up-sampling of the scores)
//   JointsMSELoss                            lib/core/loss.py:14-38
//   BCELoss, relation MSE                    lib/core/function.py:253, 307-311; pose_rsgnet.py:1014-1018
//   Adam                                     lib/utils/utils.py:70-74 (torch.optim.Adam defaults: betas .9/.999, eps 1e-8)
// These are HBM-bound streaming kernels: coalesced along the channel dimension, reductions in fp32 per thread and fp64
// across threads / blocks (atomicAdd on double), grids sized from the SM count.
#include "common.cuh"
#include "../../include/rsg_b200.h"

namespace {

constexpr int EW_THREADS = 256;

inline int ew_grid(long long n, int per_thread = 1) {
  long long b = (n + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
  const long long cap = 32ll * rsg_num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define GRID_STRIDE(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ------------------------------------------------------------------------------------------------------------------
// Per-channel reductions over the rows of [M, C]: out[0][c] += sum f0, out[1][c] += sum f1 (double).
// Thread layout: CT = min(C, 256) channel lanes, G = 256 / CT row groups; channel tiles loop when C > 256.
// KIND 0: f0 = x, f1 = x^2                       (batch statistics)
// KIND 1: f0 = dy', f1 = dy' * (x - mean) * invstd  with dy' = dy (or dy * [y > 0] when relu)   (BN backward sums)
// KIND 2: f0 = x only                              (column sum: bias gradients, repeat backward)
template <int KIND>
__global__ void __launch_bounds__(EW_THREADS) chan_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 const float* __restrict__ y, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, int relu, long long M, int C,
                                                                 int rows_per_block, double* __restrict__ out) {
  __shared__ double red0[EW_THREADS], red1[EW_THREADS];
  const int CT = C < EW_THREADS ? C : EW_THREADS, G = EW_THREADS / CT;
  const int tid = threadIdx.x, g = tid / CT, cl = tid - g * CT;
  const bool active = g < G;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  {
    const int cb = blockIdx.y * CT;              // one channel tile per blockIdx.y (C > 256: the batch-repeat backward has C = h w C0)
    const int c = cb + cl;
    // double accumulators: var = E[x^2] - mean^2 cancels (mean^2 / var) digits, and fp32 partial sums left a 1e-4-class
    // error in invstd for channels whose mean is ~10x their spread (visible as 3e-4 in the gradients of a 50-layer net)
    double s0 = 0.0, s1 = 0.0;
    if (active && c < C) {
      float mu = 0.f, is = 0.f;
      if (KIND == 1) { mu = mean[c]; is = invstd[c]; }
      for (long long r = r0 + g; r < r1; r += G) {
        const long long i = r * C + c;
        if (KIND == 0) { const double v = (double)x[i]; s0 += v; s1 += v * v; }
        else if (KIND == 1) {
          float d = dy[i];
          if (relu && !(y[i] > 0.f)) d = 0.f;
          s0 += (double)d;
          s1 += (double)d * (double)((x[i] - mu) * is);
        } else s0 += (double)x[i];
      }
    }
    red0[tid] = s0; red1[tid] = s1;
    __syncthreads();
    if (g == 0 && c < C) {
      double a0 = 0.0, a1 = 0.0;
      for (int k = 0; k < G; ++k) { a0 += red0[k * CT + cl]; a1 += red1[k * CT + cl]; }
      atomicAdd(out + c, a0);
      if (KIND != 2) atomicAdd(out + C + c, a1);
    }
    __syncthreads();
  }
}

// Row blocks of a channel reduction: every block ends in one fp64 atomicAdd per channel on the SAME 2 C addresses, so the
// block count is kept near 2 per SM (1184 blocks cost ~10 us of same-address atomics on a 12 MB tensor, 296 cost 2 us).
inline void chan_reduce_cfg(long long M, int& rows_per_block, int& blocks, int ytiles = 1) {
  long long b = (M + 63) / 64;                       // at least 64 rows per block (8 rows measured slower: more same-address atomics)
  long long cap = 2ll * rsg_num_sms() / ytiles;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  rows_per_block = (int)((M + b - 1) / b);
  blocks = (int)((M + rows_per_block - 1) / rows_per_block);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sums[c] / (double)M;
  double var = sums[C + c] / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) {
    const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void __launch_bounds__(EW_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int relu, long long n, int C,
                                                              float* __restrict__ y) {
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    float v = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
    if (relu) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

// dx = gamma * invstd * (dy' - sum_dy / M - xhat * sum_dy_xhat / M); block 0 also accumulates dgamma / dbeta
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  const float* __restrict__ y, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                  const double* __restrict__ sums, int relu, long long n, int C,
                                                                  long long M, float* __restrict__ dx, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta) {
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] += (float)sums[c];
      if (dgamma) dgamma[c] += (float)sums[C + c];
    }
  }
  if (!dx) return;
  const double invM = 1.0 / (double)M;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    float d = dy[i];
    if (relu && !(y[i] > 0.f)) d = 0.f;
    const float xh = (x[i] - mean[c]) * invstd[c];
    const float m0 = (float)(sums[c] * invM), m1 = (float)(sums[C + c] * invM);
    dx[i] = gamma[c] * invstd[c] * (d - m0 - xh * m1);
  }
}


// ---- float4 forms (C % 4 == 0: every BatchNorm of the network).  Same thread layout as chan_reduce_kernel on channel
// QUADS: a thread keeps its quad's statistics / affine terms in registers and streams rows, so the per-element work is one
// 16-byte load per operand and no index division (the scalar kernels spend a 64-bit modulo per element).
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); return F4{{t.x, t.y, t.z, t.w}}; }
__device__ __forceinline__ void st4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

// What the LAST block of a channel reduction does with the finished sums (ticket counter at out[2 C]): the BatchNorm
// finalize step (forward) or the means for the backward apply kernel + the dgamma / dbeta accumulation (backward); it then
// zeroes the sums and the ticket, so the scratch is zero on entry of the next reduction and no memset / finalize launch is
// needed (a BN is 2 launches instead of 4 forward and 3 backward: the step has 604 of them, each a dependent ~3 us launch).
struct BnFin {
  float eps, momentum;
  float* mean_out; float* invstd_out; float* running_mean; float* running_var;     // KIND 0
  float* fsum; float* dgamma; float* dbeta;                                        // KIND 1
};

template <int KIND>      // 0: sum x, sum x^2;  1: sum dy', sum dy' xhat
__global__ void __launch_bounds__(EW_THREADS) chan_reduce4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  const float* __restrict__ y, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, int relu, long long M, int C,
                                                                  int rows_per_block, double* out, const BnFin fin) {
  __shared__ double red[2][4][EW_THREADS];
  const int CQ = C >> 2, CT = CQ < EW_THREADS ? CQ : EW_THREADS, G = EW_THREADS / CT;
  const int tid = threadIdx.x, g = tid / CT, cl = tid - g * CT;
  const bool active = g < G;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  {
    const int cb = blockIdx.y * CT;
    const int cq = cb + cl;
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    if (active && cq < CQ) {
      F4 mu = {{0, 0, 0, 0}}, is = mu;
      if (KIND == 1) { mu = ld4(mean + 4 * cq); is = ld4(invstd + 4 * cq); }
      for (long long r = r0 + g; r < r1; r += G) {
        const long long i = r * C + 4 * cq;
        const F4 xv = ld4(x + i);
        if (KIND == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { const double v = (double)xv.v[e]; s0[e] += v; s1[e] += v * v; }
        } else {
          F4 d = ld4(dy + i);
          if (relu) {
            const F4 yv = ld4(y + i);
#pragma unroll
            for (int e = 0; e < 4; ++e) if (!(yv.v[e] > 0.f)) d.v[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) { s0[e] += (double)d.v[e]; s1[e] += (double)d.v[e] * (double)((xv.v[e] - mu.v[e]) * is.v[e]); }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[0][e][tid] = s0[e]; red[1][e][tid] = s1[e]; }
    __syncthreads();
    if (g == 0 && cq < CQ) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < G; ++k) { a0 += red[0][e][k * CT + cl]; a1 += red[1][e][k * CT + cl]; }
        atomicAdd(out + 4 * cq + e, a0);
        atomicAdd(out + C + 4 * cq + e, a1);
      }
    }
    __syncthreads();
  }
  // ---- ticket: the last block to arrive finalizes and cleans the scratch
  __shared__ int is_last;
  __threadfence();
  if (tid == 0) {
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(out + 2 * C), 1u);
    is_last = t == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = tid; c < C; c += EW_THREADS) {
    const double s0 = __ldcg(out + c), s1 = __ldcg(out + C + c);
    if (KIND == 0) {
      const double mu = s0 / (double)M;
      double var = s1 / (double)M - mu * mu;
      if (var < 0.0) var = 0.0;
      fin.mean_out[c] = (float)mu;
      fin.invstd_out[c] = (float)(1.0 / sqrt(var + (double)fin.eps));
      if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mu;
      if (fin.running_var) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unb;
      }
    } else {
      fin.fsum[c] = (float)(s0 / (double)M);
      fin.fsum[C + c] = (float)(s1 / (double)M);
      if (fin.dbeta) fin.dbeta[c] += (float)s0;
      if (fin.dgamma) fin.dgamma[c] += (float)s1;
    }
    out[c] = 0.0;
    out[C + c] = 0.0;
  }
  if (tid == 0) *reinterpret_cast<unsigned*>(out + 2 * C) = 0u;
}

__global__ void __launch_bounds__(EW_THREADS) bn_apply4_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ res, int relu,
                                                               long long M, int C, int rows_per_block, float* __restrict__ y) {
  const int CQ = C >> 2, CT = CQ < EW_THREADS ? CQ : EW_THREADS, G = EW_THREADS / CT;
  const int tid = threadIdx.x, g = tid / CT, cl = tid - g * CT;
  if (g >= G) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int cq = cl; cq < CQ; cq += CT) {
    const F4 mu = ld4(mean + 4 * cq), is = ld4(invstd + 4 * cq), ga = ld4(gamma + 4 * cq), be = ld4(beta + 4 * cq);
    for (long long r = r0 + g; r < r1; r += G) {
      const long long i = r * C + 4 * cq;
      F4 v = ld4(x + i);
      F4 rv = {{0.f, 0.f, 0.f, 0.f}};
      if (res) rv = ld4(res + i);                       // residual branch of a BasicBlock / Bottleneck: relu(bn(x) + res)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v.v[e] = (v.v[e] - mu.v[e]) * is.v[e] * ga.v[e] + be.v[e] + rv.v[e];
        if (relu) v.v[e] = fmaxf(v.v[e], 0.f);
      }
      st4(y + i, v);
    }
  }
}

__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   const float* __restrict__ y, const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ fsum, int relu, long long M, int C,
                                                                   int rows_per_block, float* __restrict__ dx, float* __restrict__ dres) {
  const int CQ = C >> 2, CT = CQ < EW_THREADS ? CQ : EW_THREADS, G = EW_THREADS / CT;
  const int tid = threadIdx.x, g = tid / CT, cl = tid - g * CT;
  if (g >= G) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int cq = cl; cq < CQ; cq += CT) {
    const F4 mu = ld4(mean + 4 * cq), is = ld4(invstd + 4 * cq), ga = ld4(gamma + 4 * cq);
    const F4 f0 = ld4(fsum + 4 * cq), f1 = ld4(fsum + C + 4 * cq);
    float m0[4], m1[4], sc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      m0[e] = f0.v[e];
      m1[e] = f1.v[e];
      sc[e] = ga.v[e] * is.v[e];
    }
    for (long long r = r0 + g; r < r1; r += G) {
      const long long i = r * C + 4 * cq;
      const F4 xv = ld4(x + i);
      F4 d = ld4(dy + i);
      if (relu) {
        const F4 yv = ld4(y + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) if (!(yv.v[e] > 0.f)) d.v[e] = 0.f;
      }
      if (dres) st4(dres + i, d);                      // gradient of the residual input: the ReLU-masked dy
      if (dx) {
        F4 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) o.v[e] = sc[e] * (d.v[e] - m0[e] - (xv.v[e] - mu.v[e]) * is.v[e] * m1[e]);
        st4(dx + i, o);
      }
    }
  }
}

inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline void rows_cfg(long long M, int per_sm, int& rows_per_block, int& blocks) {
  long long b = (M + 15) / 16;                       // at least 16 rows per block
  const long long cap = (long long)per_sm * rsg_num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  rows_per_block = (int)((M + b - 1) / b);
  blocks = (int)((M + rows_per_block - 1) / rows_per_block);
}

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm over [B, S, C] with G groups: one block per (b, group); two passes (mean, then centred variance).
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(EW_THREADS) gn_fwd_kernel(const float* __restrict__ x, int S, int C, int G, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ double sh[EW_THREADS];
  const int b = blockIdx.x / G, grp = blockIdx.x % G, cg = C / G, n = S * cg;
  const float* xb = x + (long long)b * S * C + grp * cg;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)xb[(long long)(i / cg) * C + i % cg];
  const double mu = block_sum(s, sh) / n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)xb[(long long)(i / cg) * C + i % cg] - mu; q += d * d; }
  const double var = block_sum(q, sh) / n;
  const float rs = (float)(1.0 / sqrt(var + (double)eps)), muf = (float)mu;
  if (threadIdx.x == 0) { mean[blockIdx.x] = muf; rstd[blockIdx.x] = rs; }
  float* yb = y + (long long)b * S * C + grp * cg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cg;
    const long long o = (long long)(i / cg) * C + c;
    yb[o] = (xb[o] - muf) * rs * gamma[grp * cg + c] + beta[grp * cg + c];
  }
}

__global__ void __launch_bounds__(EW_THREADS) gn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int S, int C, int G,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double sh[EW_THREADS];
  __shared__ float sdg[64], sdb[64];                 // per-channel partials of this (b, group); C / G <= 64
  const int b = blockIdx.x / G, grp = blockIdx.x % G, cg = C / G, n = S * cg;
  const float muf = mean[blockIdx.x], rs = rstd[blockIdx.x];
  const float* xb = x + (long long)b * S * C + grp * cg;
  const float* db = dy + (long long)b * S * C + grp * cg;
  for (int c = threadIdx.x; c < cg; c += blockDim.x) { sdg[c] = 0.f; sdb[c] = 0.f; }
  __syncthreads();
  double s1 = 0.0, s2 = 0.0;
  // thread-private per-channel sums are impractical for arbitrary cg: channel of element i is i % cg, and a thread's
  // elements i = tid + k * 256 keep the same channel only when 256 % cg == 0 -- shared atomics cover the general case
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cg;
    const long long o = (long long)(i / cg) * C + c;
    const float d = db[o], xh = (xb[o] - muf) * rs, dg = d * gamma[grp * cg + c];
    s1 += (double)dg;
    s2 += (double)(dg * xh);
    atomicAdd(&sdg[c], d * xh);
    atomicAdd(&sdb[c], d);
  }
  const double m1 = block_sum(s1, sh) / n, m2 = block_sum(s2, sh) / n;
  for (int c = threadIdx.x; c < cg; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + grp * cg + c, sdg[c]);
    if (dbeta) atomicAdd(dbeta + grp * cg + c, sdb[c]);
  }
  if (!dx) return;
  float* dxb = dx + (long long)b * S * C + grp * cg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % cg;
    const long long o = (long long)(i / cg) * C + c;
    const float xh = (xb[o] - muf) * rs, dg = db[o] * gamma[grp * cg + c];
    dxb[o] = rs * (dg - (float)m1 - xh * (float)m2);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// element-wise
struct Ptr4 { const float* p[4]; };

__global__ void __launch_bounds__(EW_THREADS) add_kernel(Ptr4 in, int nin, int relu, long long n, float* __restrict__ out) {
  GRID_STRIDE(i, n) {
    float v = in.p[0][i];
    for (int k = 1; k < nin; ++k) v += in.p[k][i];
    if (relu) v = fmaxf(v, 0.f);
    out[i] = v;
  }
}

__global__ void __launch_bounds__(EW_THREADS) add4_kernel(Ptr4 in, int nin, int relu, long long n4, float* __restrict__ out) {
  GRID_STRIDE(i, n4) {
    float4 v = reinterpret_cast<const float4*>(in.p[0])[i];
    for (int k = 1; k < nin; ++k) {
      const float4 w = reinterpret_cast<const float4*>(in.p[k])[i];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// OP 0: dx = dy * [y > 0]            (ReLU backward, y = the ReLU output)
// OP 1: y = sigmoid(x)               (a = x)
// OP 2: dx = dy * y * (1 - y)        (a = dy, b = y)
// OP 3: y = leaky(x, slope)          (a = x)
// OP 4: dx = dy * (x > 0 ? 1 : slope) (a = dy, b = x)
// OP 5: out = a * b
// OP 6: out += a
// OP 7: out = a * slope              (scale)
template <int OP>
__device__ __forceinline__ float ew_op(float a, float b, float o, float slope) {
  if (OP == 0) return b > 0.f ? a : 0.f;
  if (OP == 1) return 1.f / (1.f + expf(-a));
  if (OP == 2) return a * b * (1.f - b);
  if (OP == 3) return a > 0.f ? a : a * slope;
  if (OP == 4) return b > 0.f ? a : a * slope;
  if (OP == 5) return a * b;
  if (OP == 6) return o + a;
  return a * slope;
}

template <int OP>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(const float* __restrict__ a, const float* __restrict__ b, float slope, long long n,
                                                        float* __restrict__ out) {
  constexpr bool NB = OP == 0 || OP == 2 || OP == 4 || OP == 5;
  GRID_STRIDE(i, n) out[i] = ew_op<OP>(a[i], NB ? b[i] : 0.f, OP == 6 ? out[i] : 0.f, slope);
}

template <int OP>
__global__ void __launch_bounds__(EW_THREADS) ew4_kernel(const float* __restrict__ a, const float* __restrict__ b, float slope, long long n4,
                                                         float* __restrict__ out) {
  constexpr bool NB = OP == 0 || OP == 2 || OP == 4 || OP == 5;
  GRID_STRIDE(i, n4) {
    const float4 av = reinterpret_cast<const float4*>(a)[i];
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), ov = bv;
    if (NB) bv = reinterpret_cast<const float4*>(b)[i];
    if (OP == 6) ov = reinterpret_cast<const float4*>(out)[i];
    float4 r;
    r.x = ew_op<OP>(av.x, bv.x, ov.x, slope); r.y = ew_op<OP>(av.y, bv.y, ov.y, slope);
    r.z = ew_op<OP>(av.z, bv.z, ov.z, slope); r.w = ew_op<OP>(av.w, bv.w, ov.w, slope);
    reinterpret_cast<float4*>(out)[i] = r;
  }
}

// rows x cols copy between pitched matrices (pitch in elements; src_pitch 0 broadcasts one row); accumulate: dst += src
__global__ void __launch_bounds__(EW_THREADS) copy2d_kernel(const float* __restrict__ src, long long src_pitch, float* __restrict__ dst,
                                                            long long dst_pitch, long long rows, int cols, int accumulate) {
  const long long n = rows * cols;
  GRID_STRIDE(i, n) {
    const long long r = (long long)((unsigned)i / (unsigned)cols);
    const int c = (int)(i - r * cols);
    const float v = src[r * src_pitch + c];
    float* d = dst + r * dst_pitch + c;
    *d = accumulate ? *d + v : v;
  }
}

// dst[i0][i1][i2] (contiguous, dims D0 x D1 x D2) = src[i0 s0 + i1 s1 + i2 s2] where i1 < V1 and i2 < V2, else 0
struct Perm {
  const float* src; float* dst;
  int D0, D1, D2, V1, V2;
  long long s0, s1, s2;
  int accumulate;
};

__device__ __forceinline__ void perm_body(const Perm& q, long long lo, long long stride) {
  const long long n = (long long)q.D0 * q.D1 * q.D2;
  for (long long i = lo; i < n; i += stride) {
    const unsigned u = (unsigned)i;
    const int i2 = (int)(u % (unsigned)q.D2);
    const unsigned r = u / (unsigned)q.D2;
    const int i1 = (int)(r % (unsigned)q.D1);
    const long long i0 = r / (unsigned)q.D1;
    float v = 0.f;
    if (i1 < q.V1 && i2 < q.V2) v = q.src[i0 * q.s0 + i1 * q.s1 + i2 * q.s2];
    q.dst[i] = q.accumulate ? q.dst[i] + v : v;
  }
}

__global__ void __launch_bounds__(EW_THREADS) permute3_kernel(const Perm q) {
  perm_body(q, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// one launch for a whole table of permutations (the per-step weight packing / gradient unpacking): blockIdx.y = entry
__global__ void __launch_bounds__(EW_THREADS) permute3_batch_kernel(const Perm* __restrict__ table) {
  const Perm q = table[blockIdx.y];
  perm_body(q, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// nearest-neighbour up-sampling by an integer factor f: [N, h, w, C] -> [N, h f, w f, C]; backward = f x f block sums
__global__ void __launch_bounds__(EW_THREADS) nearest_fwd_kernel(const float* __restrict__ x, int h, int w, int C, int f, long long n,
                                                                 float* __restrict__ y) {
  const int W = w * f, H = h * f;
  GRID_STRIDE(i, n) {
    const unsigned u = (unsigned)i;                 // 32-bit index arithmetic (n < 2^32, checked on the host): a 64-bit
    const int c = (int)(u % (unsigned)C);           // division costs ~100 instructions, four of them per element
    unsigned r = u / (unsigned)C;
    const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
    const int Y = (int)(r % (unsigned)H);
    const long long nb = r / (unsigned)H;
    y[i] = x[((nb * h + Y / f) * w + X / f) * C + c];
  }
}

__global__ void __launch_bounds__(EW_THREADS) nearest_bwd_kernel(const float* __restrict__ dy, int h, int w, int C, int f, long long n,
                                                                 float* __restrict__ dx) {
  const int W = w * f, H = h * f;
  GRID_STRIDE(i, n) {                    // n = elements of dx
    const unsigned u = (unsigned)i;
    const int c = (int)(u % (unsigned)C);
    unsigned r = u / (unsigned)C;
    const int xx = (int)(r % (unsigned)w); r /= (unsigned)w;
    const int yy = (int)(r % (unsigned)h);
    const long long nb = r / (unsigned)h;
    float s = 0.f;
    for (int a = 0; a < f; ++a)
      for (int b2 = 0; b2 < f; ++b2) s += dy[((nb * H + yy * f + a) * W + xx * f + b2) * C + c];
    dx[i] = s;
  }
}

// bilinear x2, align_corners=True: src = dst * (in - 1) / (out - 1) (torch area_pixel_compute_source_index)
__device__ __forceinline__ void bil_src(int o, int in, int out, int& i0, int& i1, float& l1) {
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float s = scale * (float)o;
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(EW_THREADS) bilinear_fwd_kernel(const float* __restrict__ x, int h, int w, int C, long long n,
                                                                  float* __restrict__ y) {
  const int H = 2 * h, W = 2 * w;
  GRID_STRIDE(i, n) {
    const unsigned u = (unsigned)i;                 // 32-bit index arithmetic (n < 2^32, checked on the host): a 64-bit
    const int c = (int)(u % (unsigned)C);           // division costs ~100 instructions, four of them per element
    unsigned r = u / (unsigned)C;
    const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
    const int Y = (int)(r % (unsigned)H);
    const long long nb = r / (unsigned)H;
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(Y, h, H, y0, y1, ly);
    bil_src(X, w, W, x0, x1, lx);
    const float* xb = x + nb * h * w * C + c;
    const float v00 = xb[((long long)y0 * w + x0) * C], v01 = xb[((long long)y0 * w + x1) * C];
    const float v10 = xb[((long long)y1 * w + x0) * C], v11 = xb[((long long)y1 * w + x1) * C];
    const float hy = 1.f - ly, hx = 1.f - lx;
    y[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

__global__ void __launch_bounds__(EW_THREADS) bilinear_bwd_kernel(const float* __restrict__ dy, int h, int w, int C, long long n,
                                                                  float* __restrict__ dx) {   // dx zeroed by the caller
  const int H = 2 * h, W = 2 * w;
  GRID_STRIDE(i, n) {                    // n = elements of dy
    const unsigned u = (unsigned)i;                 // 32-bit index arithmetic (n < 2^32, checked on the host): a 64-bit
    const int c = (int)(u % (unsigned)C);           // division costs ~100 instructions, four of them per element
    unsigned r = u / (unsigned)C;
    const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
    const int Y = (int)(r % (unsigned)H);
    const long long nb = r / (unsigned)H;
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(Y, h, H, y0, y1, ly);
    bil_src(X, w, W, x0, x1, lx);
    float* xb = dx + nb * h * w * C + c;
    const float d = dy[i], hy = 1.f - ly, hx = 1.f - lx;
    atomicAdd(xb + ((long long)y0 * w + x0) * C, d * hy * hx);
    atomicAdd(xb + ((long long)y0 * w + x1) * C, d * hy * lx);
    atomicAdd(xb + ((long long)y1 * w + x0) * C, d * ly * hx);
    atomicAdd(xb + ((long long)y1 * w + x1) * C, d * ly * lx);
  }
}

// 2x2 max pooling (TRP sub_sample, association.py:221): idx = arg-max position 0..3 kept for the backward pass
__global__ void __launch_bounds__(EW_THREADS) maxpool_fwd_kernel(const float* __restrict__ x, int H, int W, int C, long long n,
                                                                 float* __restrict__ y, unsigned char* __restrict__ idx) {
  const int h = H / 2, w = W / 2;
  GRID_STRIDE(i, n) {
    const unsigned u = (unsigned)i;
    const int c = (int)(u % (unsigned)C);
    unsigned r = u / (unsigned)C;
    const int xx = (int)(r % (unsigned)w); r /= (unsigned)w;
    const int yy = (int)(r % (unsigned)h);
    const long long nb = r / (unsigned)h;
    float best = 0.f;
    int bi = 0;
    for (int k = 0; k < 4; ++k) {
      const float v = x[((nb * H + 2 * yy + (k >> 1)) * W + 2 * xx + (k & 1)) * C + c];
      if (k == 0 || v > best) { best = v; bi = k; }
    }
    y[i] = best;
    idx[i] = (unsigned char)bi;
  }
}

__global__ void __launch_bounds__(EW_THREADS) maxpool_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx, int H,
                                                                 int W, int C, long long n, float* __restrict__ dx) {
  const int h = H / 2, w = W / 2;
  GRID_STRIDE(i, n) {                    // n = elements of dx; odd trailing rows / columns get zero
    const unsigned u = (unsigned)i;                 // 32-bit index arithmetic (n < 2^32, checked on the host): a 64-bit
    const int c = (int)(u % (unsigned)C);           // division costs ~100 instructions, four of them per element
    unsigned r = u / (unsigned)C;
    const int X = (int)(r % (unsigned)W); r /= (unsigned)W;
    const int Y = (int)(r % (unsigned)H);
    const long long nb = r / (unsigned)H;
    float v = 0.f;
    if ((Y >> 1) < h && (X >> 1) < w) {
      const long long o = ((nb * h + (Y >> 1)) * w + (X >> 1)) * C + c;
      if (idx[o] == (unsigned char)(((Y & 1) << 1) | (X & 1))) v = dy[o];
    }
    dx[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// losses: out[0] += coef * sum; grad written per element
__device__ __forceinline__ void block_atomic_add(double v, double* out) {
  __shared__ double sh[EW_THREADS];
  const double s = block_sum(v, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// JointsMSELoss on NCHW [B, K, HW]: loss = 0.5 / (K B HW) sum (w (p - t))^2 ; dp = up / (K B HW) * w^2 (p - t)
__global__ void __launch_bounds__(EW_THREADS) mse_joints_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                                const float* __restrict__ tw, int HW, long long n, double coef, float up,
                                                                double* __restrict__ loss, float* __restrict__ grad) {
  double s = 0.0;
  const float gcoef = (float)(2.0 * coef) * up;
  GRID_STRIDE(i, n) {
    const float w = tw ? tw[i / HW] : 1.f;
    const float d = w * (p[i] - t[i]);
    s += (double)d * (double)d;
    if (grad) grad[i] = gcoef * w * d;
  }
  block_atomic_add(s * coef, loss);
}

// BCELoss(mean): loss = -coef sum (t log p + (1 - t) log(1 - p)), logs clamped at -100 (torch); dp = up * coef * (p - t) / max(p (1 - p), 1e-12)
__global__ void __launch_bounds__(EW_THREADS) bce_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n, double coef,
                                                         float up, double* __restrict__ loss, float* __restrict__ grad) {
  double s = 0.0;
  GRID_STRIDE(i, n) {
    const float pp = p[i], tt = t[i];
    const float lp = fmaxf(logf(pp), -100.f), lq = fmaxf(logf(1.f - pp), -100.f);
    s -= (double)(tt * lp + (1.f - tt) * lq);
    if (grad) grad[i] = up * (float)coef * (pp - tt) / fmaxf(pp * (1.f - pp), 1e-12f);
  }
  block_atomic_add(s * coef, loss);
}

// relation MSE (pose_rsgnet.py:1014-1018): out[b] = mean_ij (T[b,i,j] - P[b,i,j])^2, T given in full or as the rank-1
// factor v[b] (T = v v^T: lib/core/function.py:261-269 builds it from the person mask)
__global__ void __launch_bounds__(EW_THREADS) relation_mse_kernel(const float* __restrict__ P, const float* __restrict__ T,
                                                                  const float* __restrict__ v, int S, int blocks_per_b,
                                                                  double* __restrict__ out) {
  const int b = blockIdx.x / blocks_per_b, part = blockIdx.x % blocks_per_b;
  const long long n = (long long)S * S;
  const float* Pb = P + (long long)b * n;
  double s = 0.0;
  for (long long i = (long long)part * blockDim.x + threadIdx.x; i < n; i += (long long)blocks_per_b * blockDim.x) {
    const float tt = T ? T[(long long)b * n + i] : v[(long long)b * S + i / S] * v[(long long)b * S + i % S];
    const float d = tt - Pb[i];
    s += (double)d * (double)d;
  }
  block_atomic_add(s / (double)n, out + b);
}


// float4 form (S % 4 == 0): a block walks whole rows i of one sample, so v[i] is a scalar per row and no index is divided
// (the scalar kernel above spends two 64-bit divisions per element: 0.69 ms for 1.2 GB, a third of the HBM rate)
__global__ void __launch_bounds__(EW_THREADS) relation_mse4_kernel(const float* __restrict__ P, const float* __restrict__ T,
                                                                   const float* __restrict__ v, int S, int blocks_per_b,
                                                                   double* __restrict__ out) {
  const int b = blockIdx.x / blocks_per_b, part = blockIdx.x % blocks_per_b;
  const long long base = (long long)b * S * S;
  const float* vb = v ? v + (long long)b * S : nullptr;
  double s = 0.0;
  for (int i = part; i < S; i += blocks_per_b) {
    const float4* Pr = reinterpret_cast<const float4*>(P + base + (long long)i * S);
    const float4* Tr = T ? reinterpret_cast<const float4*>(T + base + (long long)i * S) : nullptr;
    const float vi = vb ? vb[i] : 0.f;
    float acc = 0.f;
    for (int j = threadIdx.x; j < S / 4; j += blockDim.x) {
      const float4 pp = Pr[j];
      float4 tt;
      if (Tr) tt = Tr[j];
      else { const float4 vj = reinterpret_cast<const float4*>(vb)[j]; tt = make_float4(vi * vj.x, vi * vj.y, vi * vj.z, vi * vj.w); }
      const float d0 = tt.x - pp.x, d1 = tt.y - pp.y, d2 = tt.z - pp.z, d3 = tt.w - pp.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    s += (double)acc;
  }
  block_atomic_add(s / ((double)S * (double)S), out + b);
}

// dA = (dP + coef[b] (P - T)) * P (1 - P):  the sigmoid backward of the TRP affinity with the relation-loss gradient folded in
// (dP may be NULL = zero, coef NULL = no relation term)
__global__ void __launch_bounds__(EW_THREADS) trp_dscore_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                                const float* __restrict__ T, const float* __restrict__ v,
                                                                const float* __restrict__ coef, int S, int blocks_per_b,
                                                                float* __restrict__ dA) {
  const int b = blockIdx.x / blocks_per_b, part = blockIdx.x % blocks_per_b;
  const long long n = (long long)S * S, base = (long long)b * n;
  const float cb = coef ? coef[b] : 0.f;
  for (long long i = (long long)part * blockDim.x + threadIdx.x; i < n; i += (long long)blocks_per_b * blockDim.x) {
    const float pp = P[base + i];
    float d = dP ? dP[base + i] : 0.f;
    if (coef) {
      const float tt = T ? T[base + i] : v[(long long)b * S + i / S] * v[(long long)b * S + i % S];
      d += cb * (pp - tt);
    }
    dA[base + i] = d * pp * (1.f - pp);
  }
}

// float4 / row-wise form of the same (S % 4 == 0): no index divisions, v[i] a scalar per row
__global__ void __launch_bounds__(EW_THREADS) trp_dscore4_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                                 const float* __restrict__ T, const float* __restrict__ v,
                                                                 const float* __restrict__ coef, int S, int blocks_per_b,
                                                                 float* __restrict__ dA) {
  const int b = blockIdx.x / blocks_per_b, part = blockIdx.x % blocks_per_b;
  const long long base = (long long)b * S * S;
  const float cb = coef ? coef[b] : 0.f;
  const float* vb = v ? v + (long long)b * S : nullptr;
  for (int i = part; i < S; i += blocks_per_b) {
    const long long ro = base + (long long)i * S;
    const float4* Pr = reinterpret_cast<const float4*>(P + ro);
    const float4* Dr = dP ? reinterpret_cast<const float4*>(dP + ro) : nullptr;
    const float4* Tr = (coef && T) ? reinterpret_cast<const float4*>(T + ro) : nullptr;
    float4* Ar = reinterpret_cast<float4*>(dA + ro);
    const float vi = (coef && vb) ? vb[i] : 0.f;
    for (int j = threadIdx.x; j < S / 4; j += blockDim.x) {
      const float4 pp = Pr[j];
      float4 d = Dr ? Dr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (coef) {
        float4 tt;
        if (Tr) tt = Tr[j];
        else { const float4 vj = reinterpret_cast<const float4*>(vb)[j]; tt = make_float4(vi * vj.x, vi * vj.y, vi * vj.z, vi * vj.w); }
        d.x += cb * (pp.x - tt.x); d.y += cb * (pp.y - tt.y); d.z += cb * (pp.z - tt.z); d.w += cb * (pp.w - tt.w);
      }
      Ar[j] = make_float4(d.x * pp.x * (1.f - pp.x), d.y * pp.y * (1.f - pp.y), d.z * pp.z * (1.f - pp.z), d.w * pp.w * (1.f - pp.w));
    }
  }
}

// lib/core/function.py:261-267: v[b, y, x] = bilinear(align_corners=True, size H/2 x W/2) of max_k target[b, k, :, :]
__global__ void __launch_bounds__(EW_THREADS) person_mask_kernel(const float* __restrict__ t, int K, int H, int W, long long n,
                                                                 float* __restrict__ v) {
  const int h = H / 2, w = W / 2;
  GRID_STRIDE(i, n) {
    const int x = (int)(i % w);
    long long r = i / w;
    const int y = (int)(r % h);
    const long long b = r / h;
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(y, H, h, y0, y1, ly);
    bil_src(x, W, w, x0, x1, lx);
    const float* tb = t + b * K * H * W;
    float m00 = -INFINITY, m01 = -INFINITY, m10 = -INFINITY, m11 = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float* tk = tb + (long long)k * H * W;
      m00 = fmaxf(m00, tk[y0 * W + x0]); m01 = fmaxf(m01, tk[y0 * W + x1]);
      m10 = fmaxf(m10, tk[y1 * W + x0]); m11 = fmaxf(m11, tk[y1 * W + x1]);
    }
    const float hy = 1.f - ly, hx = 1.f - lx;
    v[i] = hy * (hx * m00 + lx * m01) + ly * (hx * m10 + lx * m11);
  }
}

__global__ void d2f_kernel(const double* __restrict__ in, float scale, int n, int accumulate, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (accumulate ? out[i] : 0.f) + (float)in[i] * scale;
}

// Adam (torch.optim.Adam, no weight decay / amsgrad): one launch over the flat parameter buffer.  The step count comes from
// a kernel argument or, when `step_dev` is given, from device memory (incremented by step_inc_kernel), so that a captured
// CUDA graph of the whole training step replays with the right bias corrections.
__global__ void step_inc_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(EW_THREADS) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                          int step, const int* __restrict__ step_dev, float gscale) {
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    const int t = step_dev ? *step_dev : step;
    sh[0] = (float)(1.0 - pow((double)b1, (double)t));
    sh[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
  }
  __syncthreads();
  const float stepsz = lr / sh[0], bc2_sqrt = sh[1];
  GRID_STRIDE(i, n) {
    const float gg = g[i] * gscale;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vv;
    p[i] -= stepsz * mm / (sqrtf(vv) / bc2_sqrt + eps);
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// Scratch contract of the BatchNorm entry points: ws holds 3 C + 4 doubles and is ZERO on entry of the float4 path, which leaves
// it zero (the last reduction block cleans up); the scalar fallback and rsg_train_colsum clear what they use before and after.
extern "C" int rsg_train_bn_fwd_res(void* stream, const float* x, long long M, int C, const float* gamma, const float* beta, float eps,
                                    float momentum, float* running_mean, float* running_var, int relu, const float* res, float* y,
                                    float* save_mean, float* save_invstd, double* ws) {
  RSG_REQUIRE(x && y && gamma && beta && save_mean && save_invstd && ws && M > 0 && C > 0, "bn_fwd: bad arguments");
  const bool v4 = (C & 3) == 0 && al16p(x) && al16p(y) && al16p(gamma) && al16p(beta) && al16p(save_mean) && al16p(save_invstd) &&
                  al16p(res);
  RSG_REQUIRE(v4 || !res, "bn_fwd: the fused residual needs the float4 path (C %% 4 == 0, 16-byte aligned tensors)");
  int rpb, blocks;
  const int yt = v4 ? ceil_div(C / 4, EW_THREADS) : ceil_div(C, EW_THREADS);
  chan_reduce_cfg(M, rpb, blocks, yt);
  const long long n = M * C;
  if (v4) {
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    fin.eps = eps; fin.momentum = momentum; fin.mean_out = save_mean; fin.invstd_out = save_invstd;
    fin.running_mean = running_mean; fin.running_var = running_var;
    chan_reduce4_kernel<0><<<dim3(blocks, yt), EW_THREADS, 0, ST>>>(x, nullptr, nullptr, nullptr, nullptr, 0, M, C, rpb, ws, fin);
    rows_cfg(M, 16, rpb, blocks);
    bn_apply4_kernel<<<blocks, EW_THREADS, 0, ST>>>(x, save_mean, save_invstd, gamma, beta, res, relu, M, C, rpb, y);
  } else {
    RSG_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), ST));
    chan_reduce_kernel<0><<<dim3(blocks, yt), EW_THREADS, 0, ST>>>(x, nullptr, nullptr, nullptr, nullptr, 0, M, C, rpb, ws);
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, ST>>>(ws, M, C, eps, momentum, save_mean, save_invstd, running_mean, running_var);
    bn_apply_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(x, save_mean, save_invstd, gamma, beta, relu, n, C, y);
    RSG_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), ST));
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_bn_bwd_res(void* stream, const float* x, const float* y, const float* dy, long long M, int C,
                                    const float* gamma, const float* save_mean, const float* save_invstd, int relu, float* dx,
                                    float* dres, float* dgamma, float* dbeta, double* ws) {
  RSG_REQUIRE(x && dy && gamma && save_mean && save_invstd && ws && M > 0 && C > 0 && (!relu || y), "bn_bwd: bad arguments");
  const bool v4 = (C & 3) == 0 && al16p(x) && al16p(dy) && (!relu || al16p(y)) && (!dx || al16p(dx)) && al16p(gamma) &&
                  al16p(save_mean) && al16p(save_invstd) && al16p(dres);
  RSG_REQUIRE(v4 || !dres, "bn_bwd: the fused residual gradient needs the float4 path");
  int rpb, blocks;
  const int yt = v4 ? ceil_div(C / 4, EW_THREADS) : ceil_div(C, EW_THREADS);
  chan_reduce_cfg(M, rpb, blocks, yt);
  const long long n = M * C;
  if (v4) {
    BnFin fin;
    memset(&fin, 0, sizeof(fin));
    fin.fsum = reinterpret_cast<float*>(ws + 2 * (size_t)C + 2);              // 2 C floats behind the sums and the ticket
    fin.dgamma = dgamma; fin.dbeta = dbeta;
    chan_reduce4_kernel<1><<<dim3(blocks, yt), EW_THREADS, 0, ST>>>(x, dy, y, save_mean, save_invstd, relu, M, C, rpb, ws, fin);
    if (dx || dres) {
      rows_cfg(M, 16, rpb, blocks);
      bn_bwd_apply4_kernel<<<blocks, EW_THREADS, 0, ST>>>(x, dy, y, save_mean, save_invstd, gamma, fin.fsum, relu, M, C, rpb, dx, dres);
    }
  } else {
    RSG_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), ST));
    chan_reduce_kernel<1><<<dim3(blocks, yt), EW_THREADS, 0, ST>>>(x, dy, y, save_mean, save_invstd, relu, M, C, rpb, ws);
    bn_bwd_apply_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(x, dy, y, save_mean, save_invstd, gamma, ws, relu, n, C, M, dx, dgamma, dbeta);
    RSG_CUDA(cudaMemsetAsync(ws, 0, 2 * (size_t)C * sizeof(double), ST));
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_bn_fwd(void* stream, const float* x, long long M, int C, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, int relu, float* y, float* save_mean,
                                float* save_invstd, double* ws) {
  return rsg_train_bn_fwd_res(stream, x, M, C, gamma, beta, eps, momentum, running_mean, running_var, relu, nullptr, y, save_mean,
                              save_invstd, ws);
}

extern "C" int rsg_train_bn_bwd(void* stream, const float* x, const float* y, const float* dy, long long M, int C, const float* gamma,
                                const float* save_mean, const float* save_invstd, int relu, float* dx, float* dgamma, float* dbeta,
                                double* ws) {
  return rsg_train_bn_bwd_res(stream, x, y, dy, M, C, gamma, save_mean, save_invstd, relu, dx, nullptr, dgamma, dbeta, ws);
}

extern "C" int rsg_train_colsum(void* stream, const float* x, long long M, int C, float* out, int accumulate, double* ws) {
  RSG_REQUIRE(x && out && ws && M > 0 && C > 0, "colsum: bad arguments");
  RSG_CUDA(cudaMemsetAsync(ws, 0, (size_t)C * sizeof(double), ST));
  int rpb, blocks;
  const int yt = ceil_div(C, EW_THREADS);
  chan_reduce_cfg(M, rpb, blocks, yt);
  chan_reduce_kernel<2><<<dim3(blocks, yt), EW_THREADS, 0, ST>>>(x, nullptr, nullptr, nullptr, nullptr, 0, M, C, rpb, ws);
  d2f_kernel<<<ceil_div(C, 128), 128, 0, ST>>>(ws, 1.f, C, accumulate, out);
  RSG_CUDA(cudaMemsetAsync(ws, 0, (size_t)C * sizeof(double), ST));          // the BatchNorm float4 path expects a zeroed scratch
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_gn_fwd(void* stream, const float* x, int B, int S, int C, int G, const float* gamma, const float* beta, float eps,
                                float* y, float* mean, float* rstd) {
  RSG_REQUIRE(x && y && gamma && beta && mean && rstd && B > 0 && S > 0 && G > 0 && C % G == 0, "gn_fwd: bad arguments");
  gn_fwd_kernel<<<B * G, EW_THREADS, 0, ST>>>(x, S, C, G, gamma, beta, eps, y, mean, rstd);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_gn_bwd(void* stream, const float* x, const float* dy, int B, int S, int C, int G, const float* gamma,
                                const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta) {
  RSG_REQUIRE(x && dy && gamma && mean && rstd && B > 0 && S > 0 && G > 0 && C % G == 0 && C / G <= 64, "gn_bwd: bad arguments");
  gn_bwd_kernel<<<B * G, EW_THREADS, 0, ST>>>(x, dy, S, C, G, gamma, mean, rstd, dx, dgamma, dbeta);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_add(void* stream, int nin, const float* const* in, int relu, long long n, float* out) {
  RSG_REQUIRE(nin >= 1 && nin <= 4 && in && out, "add: 1..4 inputs");
  if (n == 0) return RSG_OK;
  Ptr4 q;
  bool v4 = (n & 3) == 0 && al16p(out);
  for (int k = 0; k < 4; ++k) { q.p[k] = k < nin ? in[k] : nullptr; v4 = v4 && al16p(q.p[k]); }
  if (v4) add4_kernel<<<ew_grid(n / 4, 2), EW_THREADS, 0, ST>>>(q, nin, relu, n / 4, out);
  else add_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(q, nin, relu, n, out);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

template <int OP>
static void ew_launch(cudaStream_t s, const float* a, const float* b, float slope, long long n, float* out) {
  if ((n & 3) == 0 && al16p(a) && al16p(b) && al16p(out)) ew4_kernel<OP><<<ew_grid(n / 4, 2), EW_THREADS, 0, s>>>(a, b, slope, n / 4, out);
  else ew_kernel<OP><<<ew_grid(n, 4), EW_THREADS, 0, s>>>(a, b, slope, n, out);
}

extern "C" int rsg_train_ew(void* stream, int op, const float* a, const float* b, float slope, long long n, float* out) {
  RSG_REQUIRE(a && out && op >= 0 && op <= 7, "ew: bad arguments");
  RSG_REQUIRE(b || !(op == 0 || op == 2 || op == 4 || op == 5), "ew: op %d needs a second operand", op);
  if (n == 0) return RSG_OK;
  switch (op) {
    case 0: ew_launch<0>(ST, a, b, slope, n, out); break;
    case 1: ew_launch<1>(ST, a, b, slope, n, out); break;
    case 2: ew_launch<2>(ST, a, b, slope, n, out); break;
    case 3: ew_launch<3>(ST, a, b, slope, n, out); break;
    case 4: ew_launch<4>(ST, a, b, slope, n, out); break;
    case 5: ew_launch<5>(ST, a, b, slope, n, out); break;
    case 6: ew_launch<6>(ST, a, b, slope, n, out); break;
    default: ew_launch<7>(ST, a, b, slope, n, out); break;
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_copy2d(void* stream, const float* src, long long src_pitch, float* dst, long long dst_pitch, long long rows,
                                int cols, int accumulate) {
  RSG_REQUIRE(src && dst && rows >= 0 && cols > 0, "copy2d: bad arguments");
  RSG_REQUIRE(rows * cols < (1ll << 32), "copy2d: more than 2^32 elements");
  if (rows == 0) return RSG_OK;
  copy2d_kernel<<<ew_grid(rows * cols, 4), EW_THREADS, 0, ST>>>(src, src_pitch, dst, dst_pitch, rows, cols, accumulate);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_permute3(void* stream, const float* src, float* dst, int D0, int D1, int D2, long long s0, long long s1,
                                  long long s2, int V1, int V2, int accumulate) {
  RSG_REQUIRE(src && dst && D0 > 0 && D1 > 0 && D2 > 0, "permute3: bad arguments");
  RSG_REQUIRE((long long)D0 * D1 * D2 < (1ll << 32), "permute3: more than 2^32 elements");
  Perm q;
  q.src = src; q.dst = dst; q.D0 = D0; q.D1 = D1; q.D2 = D2; q.V1 = V1; q.V2 = V2; q.s0 = s0; q.s1 = s1; q.s2 = s2;
  q.accumulate = accumulate;
  permute3_kernel<<<ew_grid((long long)D0 * D1 * D2, 4), EW_THREADS, 0, ST>>>(q);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

/* table: device array of `count` records laid out as struct Perm (see rsg_b200.h: rsg_perm_entry) */
extern "C" int rsg_train_permute3_batch(void* stream, const void* table, int count, int blocks_per_entry) {
  RSG_REQUIRE(table && count >= 0 && blocks_per_entry > 0, "permute3_batch: bad arguments");
  static_assert(sizeof(Perm) == sizeof(rsg_perm_entry), "rsg_perm_entry layout");
  if (count == 0) return RSG_OK;
  RSG_REQUIRE(count <= 65535, "permute3_batch: too many entries");
  permute3_batch_kernel<<<dim3((unsigned)blocks_per_entry, (unsigned)count), EW_THREADS, 0, ST>>>(reinterpret_cast<const Perm*>(table));
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_resample(void* stream, int kind, const float* in, int N, int h, int w, int C, int f, float* out) {
  // kind 0 nearest fwd (in [N,h,w,C] -> out [N,hf,wf,C]); 1 nearest bwd (in = dy [N,hf,wf,C] -> out = dx [N,h,w,C]);
  // 2 bilinear x2 fwd; 3 bilinear x2 bwd (out zeroed here)
  RSG_REQUIRE(in && out && N > 0 && h > 0 && w > 0 && C > 0 && kind >= 0 && kind <= 3, "resample: bad arguments");
  RSG_REQUIRE(kind >= 2 || f >= 1, "resample: factor");
  RSG_REQUIRE((long long)N * h * w * C * (kind < 2 ? (long long)f * f : 4) < (1ll << 32), "resample: more than 2^32 elements");
  const long long small = (long long)N * h * w * C;
  if (kind == 0) { const long long n = small * f * f; nearest_fwd_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(in, h, w, C, f, n, out); }
  else if (kind == 1) nearest_bwd_kernel<<<ew_grid(small, 2), EW_THREADS, 0, ST>>>(in, h, w, C, f, small, out);
  else if (kind == 2) { const long long n = small * 4; bilinear_fwd_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(in, h, w, C, n, out); }
  else {
    RSG_CUDA(cudaMemsetAsync(out, 0, (size_t)small * sizeof(float), ST));
    const long long n = small * 4;
    bilinear_bwd_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(in, h, w, C, n, out);
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_maxpool(void* stream, int backward, const float* in, unsigned char* idx, int N, int H, int W, int C, float* out) {
  RSG_REQUIRE(in && idx && out && N > 0 && H >= 2 && W >= 2 && C > 0, "maxpool: bad arguments");
  RSG_REQUIRE((long long)N * H * W * C < (1ll << 32), "maxpool: more than 2^32 elements");
  if (!backward) { const long long n = (long long)N * (H / 2) * (W / 2) * C; maxpool_fwd_kernel<<<ew_grid(n, 2), EW_THREADS, 0, ST>>>(in, H, W, C, n, out, idx); }
  else { const long long n = (long long)N * H * W * C; maxpool_bwd_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(in, idx, H, W, C, n, out); }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_mse_joints(void* stream, const float* pred, const float* target, const float* tw, int B, int K, int HW, float up,
                                    double* loss_acc, float* grad) {
  RSG_REQUIRE(pred && target && loss_acc && B > 0 && K > 0 && HW > 0, "mse_joints: bad arguments");
  const long long n = (long long)B * K * HW;
  mse_joints_kernel<<<ew_grid(n, 8), EW_THREADS, 0, ST>>>(pred, target, tw, HW, n, 0.5 / (double)n, up, loss_acc, grad);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_bce(void* stream, const float* p, const float* t, long long n, double weight, float up, double* loss_acc,
                             float* grad) {
  RSG_REQUIRE(p && t && loss_acc && n > 0, "bce: bad arguments");
  bce_kernel<<<ew_grid(n, 8), EW_THREADS, 0, ST>>>(p, t, n, weight / (double)n, up, loss_acc, grad);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_relation_mse(void* stream, const float* P, const float* T, const float* v, int B, int S, double* out_acc) {
  RSG_REQUIRE(P && (T || v) && out_acc && B > 0 && S > 0, "relation_mse: bad arguments");
  int bpb = ceil_div(8ll * rsg_num_sms(), B);
  const long long per = ((long long)S * S + EW_THREADS - 1) / EW_THREADS;
  if (bpb > per) bpb = (int)per;
  if (bpb < 1) bpb = 1;
  if ((S & 3) == 0 && al16p(P) && al16p(T) && al16p(v)) {
    int rb = ceil_div(8ll * rsg_num_sms(), B);
    if (rb > S) rb = S;
    relation_mse4_kernel<<<B * rb, EW_THREADS, 0, ST>>>(P, T, v, S, rb, out_acc);
  } else {
    relation_mse_kernel<<<B * bpb, EW_THREADS, 0, ST>>>(P, T, v, S, bpb, out_acc);
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_trp_dscore(void* stream, const float* P, const float* dP, const float* T, const float* v, const float* coef,
                                    int B, int S, float* dA) {
  RSG_REQUIRE(P && dA && B > 0 && S > 0 && (!coef || T || v), "trp_dscore: bad arguments");
  int bpb = ceil_div(16ll * rsg_num_sms(), B);
  const long long per = ((long long)S * S + EW_THREADS - 1) / EW_THREADS;
  if (bpb > per) bpb = (int)per;
  if (bpb < 1) bpb = 1;
  if ((S & 3) == 0 && al16p(P) && al16p(dP) && al16p(T) && al16p(v) && al16p(dA)) {
    int rb = ceil_div(16ll * rsg_num_sms(), B);
    if (rb > S) rb = S;
    trp_dscore4_kernel<<<B * rb, EW_THREADS, 0, ST>>>(P, dP, T, v, coef, S, rb, dA);
  } else {
    trp_dscore_kernel<<<B * bpb, EW_THREADS, 0, ST>>>(P, dP, T, v, coef, S, bpb, dA);
  }
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_person_mask(void* stream, const float* target, int B, int K, int H, int W, float* v) {
  RSG_REQUIRE(target && v && B > 0 && K > 0 && H >= 2 && W >= 2, "person_mask: bad arguments");
  const long long n = (long long)B * (H / 2) * (W / 2);
  person_mask_kernel<<<ew_grid(n), EW_THREADS, 0, ST>>>(target, K, H, W, n, v);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_zero(void* stream, void* p, size_t bytes) {
  RSG_REQUIRE(p || bytes == 0, "zero: null pointer");
  if (bytes) RSG_CUDA(cudaMemsetAsync(p, 0, bytes, ST));
  return RSG_OK;
}

extern "C" int rsg_train_d2f(void* stream, const double* in, float scale, int n, int accumulate, float* out) {
  RSG_REQUIRE(in && out && n > 0, "d2f: bad arguments");
  d2f_kernel<<<ceil_div(n, 128), 128, 0, ST>>>(in, scale, n, accumulate, out);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_adam(void* stream, float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, int step, float grad_scale) {
  RSG_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "adam: bad arguments");
  if (n == 0) return RSG_OK;
  adam_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, step, nullptr, grad_scale);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_train_adam_graph(void* stream, float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                                    float beta2, float eps, int* step_dev, float grad_scale) {
  RSG_REQUIRE(p && g && m && v && n >= 0 && step_dev, "adam_graph: bad arguments");
  step_inc_kernel<<<1, 1, 0, ST>>>(step_dev);
  if (n) adam_kernel<<<ew_grid(n, 4), EW_THREADS, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, 0, step_dev, grad_scale);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
