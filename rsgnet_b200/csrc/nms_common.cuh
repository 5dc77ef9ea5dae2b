// Device helpers shared by nms.cu and evaluate.cu: OKS of a detection pair in the reference's precisions
// (lib/nms/nms.py:75-94), NumPy's pairwise summation order, the score order of scores.argsort()[::-1], and the
// evaluate()-side rescoring (lib/dataset/crowdpose.py:1294-1306).
#pragma once
#include "common.cuh"

namespace rsgnms {

// np.add.reduce over a contiguous fp64 vector of length n <= 128 (numpy's pairwise_sum).
static __device__ double numpy_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

#define RSG_NMS_MAXK 64

// Strict total order "a (at index i) comes before b (at index j)" of scores.argsort()[::-1]: larger score first, equal
// scores: the later index first (what reversing a stable ascending sort gives; NumPy's default introsort is only
// stable for n <= 16, so the order of EXACT ties is unspecified against the reference beyond that).  NaN sorts like
// NumPy sorts it -- after everything in the ascending order, i.e. FIRST here --, so every rank is hit exactly once
// whatever the input.
__device__ __forceinline__ bool score_before(double a, int i, double b, int j) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return an == bn ? i > j : an;
  return a > b || (a == b && i > j);
}

// in_vis_thre (nms.py:85-90): the reference evaluates `list(vg > t) and list(vd > t)`, which is the SECOND list (a
// non-empty list is truthy): only key points of d with score > t count, compared in fp32 like NumPy does for a
// float32 array against a Python float; no visible key point -> 0.
static __device__ double oks_pair(const float* __restrict__ g, const float* __restrict__ d, double a_g,
                           double a_d, const double* __restrict__ vars, int K, int use_vis, float vis) {
  double ex[RSG_NMS_MAXK];
  const double denom = __dadd_rn(__ddiv_rn(__dadd_rn(a_g, a_d), 2.0), 2.220446049250313e-16);
  int m = 0;
  for (int k = 0; k < K; ++k) {
    float dx = __fsub_rn(d[3 * k], g[3 * k]);
    float dy = __fsub_rn(d[3 * k + 1], g[3 * k + 1]);
    float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    double e = __ddiv_rn(__ddiv_rn(__ddiv_rn((double)s, vars[k]), denom), 2.0);
    if (!use_vis || d[3 * k + 2] > vis) ex[m++] = exp(-e);
  }
  return m ? __ddiv_rn(numpy_sum(ex, m), (double)m) : 0.0;
}


// crowdpose.py:1294-1306 / coco.py:1249-1261 for one detection: np.float32 accumulation of the maxvals above the
// threshold (NumPy >= 2 / NEP 50 semantics: a float32 scalar against a Python float compares and adds in fp32; under
// NumPy 1.x the same lines promote to fp64 -- that variant is not reproduced), mean in fp32, times the fp64 box score.
// `mv` points at the first maxval, consecutive maxvals are `stride` floats apart.
__device__ __forceinline__ double rescore_one(const float* __restrict__ mv, int stride, int K, float thre32, double box) {
  float acc = 0.f;
  int valid = 0;
  for (int k = 0; k < K; ++k) {
    const float t = mv[(size_t)k * stride];
    if (t > thre32) { acc = __fadd_rn(acc, t); ++valid; }
  }
  const double s = valid != 0 ? (double)__fdiv_rn(acc, (float)valid) : 0.0;
  return __dmul_rn(s, box);
}

}  // namespace rsgnms
