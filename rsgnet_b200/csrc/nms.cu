// Segmented greedy OKS-NMS (one image per CTA) and evaluate()-side rescoring.
//
// Replaces (paths under the reference checkout):
//   lib/nms/nms.py:75-94    oks_iou  -- dx,dy and their squares in fp32, everything after the
//                                       division by vars in fp64, np.sum's pairwise order
//   lib/nms/nms.py:97-124   oks_nms  -- descending-score greedy sweep, suppress when oks > thresh
//   lib/dataset/crowdpose.py:1294-1306, lib/dataset/coco.py:1249-1261   rescoring
// Like the reference, OKS is evaluated lazily: only between a kept detection and the detections
// still alive behind it in the score order.
#include "nms_common.cuh"
#include "../../include/rsg_b200.h"

namespace {
using namespace rsgnms;

__global__ void __launch_bounds__(128)
oks_nms_kernel(const float* __restrict__ kpts, const double* __restrict__ scores,
               const double* __restrict__ areas, const int32_t* __restrict__ offs,
               const double* __restrict__ sigmas, int K, double thresh, int use_vis, float vis, int max_per_img,
               int32_t* __restrict__ keep, int32_t* __restrict__ keep_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double vars[RSG_NMS_MAXK];
  const int img = blockIdx.x;
  const int beg = offs[img], n = offs[img + 1] - beg;
  int* order = reinterpret_cast<int*>(smem_raw);          // [n] local indices, best first
  int* dead = order + n;                                  // [n] by position in `order`
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n < 0 || n > max_per_img) {                         // shared memory was sized for max_per_img: refuse, loudly
    if (tid == 0) keep_counts[img] = -1;
    return;
  }
  if (tid < K) {
    double s2 = __dmul_rn(sigmas[tid], 2.0);
    vars[tid] = __dmul_rn(s2, s2);
  }
  // rank = position in scores.argsort()[::-1]; ties: the later index first
  for (int i = tid; i < n; i += nt) {
    const double si = scores[beg + i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (j != i) && score_before(scores[beg + j], j, si, i);
    order[r] = i;
    dead[i] = 0;
  }
  __syncthreads();
  int nkeep = 0;
  for (int p = 0; p < n; ++p) {
    if (dead[p]) continue;                                 // uniform: smem value
    const int i = order[p];
    if (tid == 0) keep[beg + nkeep] = i;
    ++nkeep;
    const float* g = kpts + (size_t)(beg + i) * K * 3;
    const double a_g = areas[beg + i];
    for (int q = p + 1 + tid; q < n; q += nt) {
      if (dead[q]) continue;
      const int j = order[q];
      double oks = oks_pair(g, kpts + (size_t)(beg + j) * K * 3, a_g, areas[beg + j], vars, K, use_vis, vis);
      if (oks > thresh) dead[q] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) keep_counts[img] = nkeep;
}

// soft_oks_nms (lib/nms/nms.py:138-180): up to max_dets rounds per image; every round keeps the head of the score-ordered
// list and multiplies the scores of the others by exp(-oks^2 / thresh) (rescore(), nms.py:127-135, 'gaussian'), then
// re-sorts what is left with scores.argsort()[::-1] -- reproduced with an explicit order list and a STABLE sort (equal
// scores come out in the reverse of their current relative order; NumPy's own order of exactly equal scores is
// implementation-defined, so ties are unspecified against the reference).  One CTA per image.
__global__ void __launch_bounds__(128)
soft_oks_nms_kernel(const float* __restrict__ kpts, const double* __restrict__ scores,
                    const double* __restrict__ areas, const int32_t* __restrict__ offs,
                    const double* __restrict__ sigmas, int K, double thresh, int max_dets, int use_vis, float vis,
                    int max_per_img, int32_t* __restrict__ keep, int32_t* __restrict__ keep_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double vars[RSG_NMS_MAXK];
  const int img = blockIdx.x;
  const int beg = offs[img], n = offs[img + 1] - beg;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n < 0 || n > max_per_img) {
    if (tid == 0) keep_counts[img] = -1;
    return;
  }
  double* cur = reinterpret_cast<double*>(smem_raw);      // [n] current scores, by detection
  int* ord = reinterpret_cast<int*>(cur + max_per_img);   // [n] the remaining detections, best first
  int* tmp = ord + max_per_img;                           // [n]
  if (tid < K) {
    double s2 = __dmul_rn(sigmas[tid], 2.0);
    vars[tid] = __dmul_rn(s2, s2);
  }
  for (int i = tid; i < n; i += nt) cur[i] = scores[beg + i];
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    const double si = cur[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (j != i) && score_before(cur[j], j, si, i);
    ord[r] = i;
  }
  __syncthreads();
  int cnt = 0, L = n;
  while (L > 0 && cnt < max_dets) {
    const int i = ord[0];
    if (tid == 0) keep[(size_t)img * max_dets + cnt] = i;
    ++cnt;
    const float* g = kpts + (size_t)(beg + i) * K * 3;
    const double a_g = areas[beg + i];
    for (int u = 1 + tid; u < L; u += nt) {
      const int j = ord[u];
      const double oks = oks_pair(g, kpts + (size_t)(beg + j) * K * 3, a_g, areas[beg + j], vars, K, use_vis, vis);
      cur[j] = __dmul_rn(cur[j], exp(__ddiv_rn(-__dmul_rn(oks, oks), thresh)));
    }
    __syncthreads();
    --L;
    for (int u = tid; u < L; u += nt) {
      const int e = ord[u + 1];
      const double se = cur[e];
      int r = 0;
      for (int v = 0; v < L; ++v) r += (v != u) && score_before(cur[ord[v + 1]], v, se, u);
      tmp[r] = e;
    }
    __syncthreads();
    for (int u = tid; u < L; u += nt) ord[u] = tmp[u];
    __syncthreads();
  }
  if (tid == 0) keep_counts[img] = cnt;
}

__global__ void oks_iou_kernel(const float* __restrict__ g, const float* __restrict__ d, double a_g,
                               const double* __restrict__ a_d, const double* __restrict__ sigmas, int K,
                               int M, int use_vis, float vis, double* __restrict__ out) {
  __shared__ double vars[RSG_NMS_MAXK];
  if (threadIdx.x < K) {
    double s2 = __dmul_rn(sigmas[threadIdx.x], 2.0);
    vars[threadIdx.x] = __dmul_rn(s2, s2);
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) out[i] = oks_pair(g, d + (size_t)i * K * 3, a_g, a_d[i], vars, K, use_vis, vis);
}

__global__ void rescore_kernel(const float* __restrict__ maxvals,
                               const double* __restrict__ box, int n, int K, double thre,
                               double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = rescore_one(maxvals + (size_t)i * K, 1, K, (float)thre, box[i]);
}

}  // namespace

extern "C" int rsg_oks_nms(void* stream, const float* kpts, const double* scores,
                           const double* areas, const int32_t* img_offsets, int n_imgs,
                           int max_per_img, const double* sigmas, int K, double thresh,
                           int32_t* keep, int32_t* keep_counts, int use_in_vis_thre, double in_vis_thre) {
  RSG_REQUIRE(n_imgs >= 0 && K > 0 && K <= RSG_NMS_MAXK, "rsg_oks_nms: bad n_imgs=%d or K=%d", n_imgs, K);
  if (n_imgs == 0) return RSG_OK;
  RSG_REQUIRE(kpts && scores && areas && img_offsets && sigmas && keep && keep_counts,
              "rsg_oks_nms: null pointer");
  RSG_REQUIRE(max_per_img >= 0, "rsg_oks_nms: max_per_img < 0");
  size_t smem = (size_t)max_per_img * 2 * sizeof(int);
  RSG_REQUIRE(smem <= 200 * 1024, "rsg_oks_nms: more than %d detections in one image", 25600);
  if (smem > 48 * 1024)
    RSG_CUDA(cudaFuncSetAttribute(oks_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  oks_nms_kernel<<<n_imgs, 128, smem, (cudaStream_t)stream>>>(kpts, scores, areas, img_offsets, sigmas, K, thresh,
                                                             use_in_vis_thre ? 1 : 0, (float)in_vis_thre, max_per_img, keep, keep_counts);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_soft_oks_nms(void* stream, const float* kpts, const double* scores, const double* areas,
                                const int32_t* img_offsets, int n_imgs, int max_per_img, const double* sigmas, int K,
                                double thresh, int max_dets, int32_t* keep, int32_t* keep_counts, int use_in_vis_thre,
                                double in_vis_thre) {
  RSG_REQUIRE(n_imgs >= 0 && K > 0 && K <= RSG_NMS_MAXK, "rsg_soft_oks_nms: bad n_imgs=%d or K=%d", n_imgs, K);
  RSG_REQUIRE(max_dets >= 1, "rsg_soft_oks_nms: max_dets=%d", max_dets);
  if (n_imgs == 0) return RSG_OK;
  RSG_REQUIRE(kpts && scores && areas && img_offsets && sigmas && keep && keep_counts, "rsg_soft_oks_nms: null pointer");
  RSG_REQUIRE(max_per_img >= 0, "rsg_soft_oks_nms: max_per_img < 0");
  size_t smem = (size_t)max_per_img * (sizeof(double) + 2 * sizeof(int));
  RSG_REQUIRE(smem <= 200 * 1024, "rsg_soft_oks_nms: more than %d detections in one image", 12800);
  if (smem > 48 * 1024)
    RSG_CUDA(cudaFuncSetAttribute(soft_oks_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  soft_oks_nms_kernel<<<n_imgs, 128, smem, (cudaStream_t)stream>>>(kpts, scores, areas, img_offsets, sigmas, K, thresh,
                                                                  max_dets, use_in_vis_thre ? 1 : 0, (float)in_vis_thre,
                                                                  max_per_img, keep, keep_counts);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_oks_iou(void* stream, const float* g, const float* d, double a_g, const double* a_d,
                           const double* sigmas, int K, int M, double* out, int use_in_vis_thre, double in_vis_thre) {
  RSG_REQUIRE(K > 0 && K <= RSG_NMS_MAXK && M >= 0, "rsg_oks_iou: bad K=%d or M=%d", K, M);
  if (M == 0) return RSG_OK;
  RSG_REQUIRE(g && d && a_d && sigmas && out, "rsg_oks_iou: null pointer");
  oks_iou_kernel<<<ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(g, d, a_g, a_d, sigmas, K, M, use_in_vis_thre ? 1 : 0,
                                                                     (float)in_vis_thre, out);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}

extern "C" int rsg_rescore(void* stream, const float* maxvals, const double* box_scores, int n,
                           int K, double in_vis_thre, double* scores) {
  RSG_REQUIRE(n >= 0 && K > 0, "rsg_rescore: bad shape");
  if (n == 0) return RSG_OK;
  RSG_REQUIRE(maxvals && box_scores && scores, "rsg_rescore: null pointer");
  rescore_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(maxvals, box_scores, n, K,
                                                                     in_vis_thre, scores);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
