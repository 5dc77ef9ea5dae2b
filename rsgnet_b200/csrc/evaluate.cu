// Device-resident evaluate(): per-image grouping -> rescoring -> segmented (soft-)OKS-NMS -> compacted keep lists,
// with no host round trip between the steps.
//
// Replaces lib/dataset/crowdpose.py:1272-1324 and lib/dataset/coco.py:1227-1277 (paths under the reference checkout):
//   _kpts / kpts = defaultdict(list)       group the detections by image id, images in FIRST-APPEARANCE order, the
//                                          detections of an image in their original order
//   rescoring (:1294-1306)                 score = box_score * mean(maxval > in_vis_thre)
//   oks_nms / soft_oks_nms (:1308-1319)    per image, WITHOUT in_vis_thre (no caller of the reference passes it)
//   len(keep) == 0 -> keep all (:1321-1324)
//
// Five small launches, all integer / fp64 scalar work (latency-bound; 100 k detections are 22 MB):
//   1 hash      open-addressing table keyed by image id: slot of every detection, first index and count per image
//   2 rank      single CTA: exclusive scan of "is the first detection of its image" -> image rank (first-appearance
//               order), then exclusive scan of the per-image counts -> img_offsets; n_imgs stays on the device
//   3 scatter   detections into their image's segment (arbitrary order inside the segment)
//   4 nms       grid-stride over images: restore the original order inside the segment (the keys are unique, so the
//               result is deterministic whatever order step 3 produced), rescore, score order, greedy / soft sweep
// Every output is indexed by GLOBAL detection number, so one D2H copy at the end is all a caller needs.
#include "nms_common.cuh"
#include "../../include/rsg_b200.h"

namespace {
using namespace rsgnms;

// empty-slot sentinel of the hash table = what cudaMemsetAsync(.., 0x80, ..) writes; an image id equal to it is refused
constexpr long long EMPTY_KEY = (long long)0x8080808080808080ull;

__device__ __forceinline__ uint32_t hash_id(long long v) {
  unsigned long long x = (unsigned long long)v;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

__global__ void eval_hash_kernel(const long long* __restrict__ ids, int n, unsigned long long* keys, int* first, int* cnt,
                                 uint32_t mask, int* __restrict__ slot_of, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = ids[i];
  if (id == EMPTY_KEY) { atomicExch(err, 1); slot_of[i] = 0; return; }
  uint32_t h = hash_id(id) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&keys[h], (unsigned long long)EMPTY_KEY, (unsigned long long)id);
    if (prev == (unsigned long long)EMPTY_KEY || prev == (unsigned long long)id) break;
    h = (h + 1) & mask;
  }
  atomicMin(&first[h], i);
  atomicAdd(&cnt[h], 1);
  slot_of[i] = (int)h;
}

// exclusive scan of v[0..n) by ONE CTA (each thread owns a contiguous chunk); returns the total to every thread
__device__ int block_exclusive_scan(const int* __restrict__ in, int* __restrict__ out, int n, int* sh /*[blockDim+1]*/) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt, b = tid * per, e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += in[i];
  sh[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int t = 0; t < nt; ++t) { const int v = sh[t]; sh[t] = acc; acc += v; }
    sh[nt] = acc;
  }
  __syncthreads();
  int acc = sh[tid];
  for (int i = b; i < e; ++i) { const int v = in[i]; out[i] = acc; acc += v; }
  const int total = sh[nt];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(1024)
eval_rank_kernel(const long long* __restrict__ ids, int n, const int* __restrict__ slot_of, const int* __restrict__ first,
                 const int* __restrict__ cnt, int* flag /*[n]*/, int* rank /*[n]*/, int* rank_of_slot, int* counts_by_rank /*[n]*/,
                 long long* __restrict__ images, int* __restrict__ img_offsets /*[n+1]*/, int* __restrict__ n_imgs_out,
                 const int* __restrict__ err) {
  __shared__ int sh[1025];
  if (*err) {                                             // a reserved image id: report it instead of a wrong grouping
    if (threadIdx.x == 0) { *n_imgs_out = -1; img_offsets[0] = 0; }
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) flag[i] = first[slot_of[i]] == i;
  __syncthreads();
  const int n_imgs = block_exclusive_scan(flag, rank, n, sh);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (flag[i]) {
      const int r = rank[i], sl = slot_of[i];
      rank_of_slot[sl] = r;
      images[r] = ids[i];
      counts_by_rank[r] = cnt[sl];
    }
  __syncthreads();
  block_exclusive_scan(counts_by_rank, img_offsets, n_imgs, sh);
  if (threadIdx.x == 0) { img_offsets[n_imgs] = n; *n_imgs_out = n_imgs; }
}

__global__ void eval_scatter_kernel(int n, const int* __restrict__ slot_of, const int* __restrict__ rank_of_slot,
                                    const int* __restrict__ img_offsets, int* cursor, int* __restrict__ seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rank_of_slot[slot_of[i]];
  seg[img_offsets[r] + atomicAdd(&cursor[r], 1)] = i;
}

// One CTA per image (grid-stride).  Per-image scratch lives in global workspace arrays indexed like the segment, so an
// image may hold any number of detections.
__global__ void __launch_bounds__(128)
eval_nms_kernel(const float* __restrict__ preds /*[n,K,3]*/, const double* __restrict__ boxes /*[n,6]*/,
                const int* __restrict__ n_imgs_p, const int* __restrict__ img_offsets, int* seg /*[n]*/, int* tmp /*[n]*/,
                int* ord /*[n]*/, int* dead /*[n]*/, double* cur /*[n]*/, const double* __restrict__ sigmas, int K,
                float vis_thre32, double oks_thre, int soft, int max_dets, double* __restrict__ scores /*[n] by detection*/,
                int32_t* __restrict__ keep /*[n]*/, int32_t* __restrict__ keep_counts) {
  __shared__ double vars[RSG_NMS_MAXK];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < K) {
    const double s2 = __dmul_rn(sigmas[tid], 2.0);
    vars[tid] = __dmul_rn(s2, s2);
  }
  __syncthreads();
  const int n_imgs = *n_imgs_p;                           // -1: refused input (see eval_rank_kernel)
  for (int img = blockIdx.x; img < n_imgs; img += gridDim.x) {
    const int beg = img_offsets[img], m = img_offsets[img + 1] - beg;
    // original order inside the image: rank by counting on the (unique) detection indices
    for (int a = tid; a < m; a += nt) {
      const int ia = seg[beg + a];
      int r = 0;
      for (int b = 0; b < m; ++b) r += seg[beg + b] < ia;
      tmp[beg + r] = ia;
    }
    __syncthreads();
    for (int a = tid; a < m; a += nt) {
      const int det = tmp[beg + a];
      seg[beg + a] = det;
      const double sc = rescore_one(preds + (size_t)det * K * 3 + 2, 3, K, vis_thre32, boxes[(size_t)det * 6 + 5]);
      scores[det] = sc;
      cur[beg + a] = sc;
      dead[beg + a] = 0;
    }
    __syncthreads();
    int nkeep = 0;
    if (!soft) {
      // position in scores.argsort()[::-1] over the image's list
      for (int a = tid; a < m; a += nt) {
        const double sa = cur[beg + a];
        int r = 0;
        for (int b = 0; b < m; ++b) r += (b != a) && score_before(cur[beg + b], b, sa, a);
        ord[beg + r] = a;
      }
      __syncthreads();
      for (int p = 0; p < m; ++p) {
        if (dead[beg + p]) continue;                       // uniform: written before the last barrier
        const int a = ord[beg + p], det = seg[beg + a];
        if (tid == 0) keep[beg + nkeep] = det;
        ++nkeep;
        const float* g = preds + (size_t)det * K * 3;
        const double a_g = boxes[(size_t)det * 6 + 4];
        for (int q = p + 1 + tid; q < m; q += nt) {
          if (dead[beg + q]) continue;
          const int dj = seg[beg + ord[beg + q]];
          const double oks = oks_pair(g, preds + (size_t)dj * K * 3, a_g, boxes[(size_t)dj * 6 + 4], vars, K, 0, 0.f);
          if (oks > oks_thre) dead[beg + q] = 1;
        }
        __syncthreads();
      }
    } else {
      // soft_oks_nms (nms.py:138-180) with the reference's own bookkeeping: an explicit `order` list that loses its
      // head every round and is re-sorted by the decayed scores with scores.argsort()[::-1].  The re-sort is modelled
      // as a STABLE sort (equal scores come out in the reverse of their current relative order); what NumPy does with
      // exactly equal scores is implementation-defined, so ties are unspecified against the reference.
      int L = m;
      for (int a = tid; a < m; a += nt) {
        const double sa = cur[beg + a];
        int r = 0;
        for (int b = 0; b < m; ++b) r += (b != a) && score_before(cur[beg + b], b, sa, a);
        ord[beg + r] = a;
      }
      __syncthreads();
      while (L > 0 && nkeep < max_dets) {
        const int a = ord[beg], det = seg[beg + a];
        if (tid == 0) keep[beg + nkeep] = det;
        ++nkeep;
        const float* g = preds + (size_t)det * K * 3;
        const double a_g = boxes[(size_t)det * 6 + 4];
        for (int u = 1 + tid; u < L; u += nt) {
          const int e = ord[beg + u], dj = seg[beg + e];
          const double oks = oks_pair(g, preds + (size_t)dj * K * 3, a_g, boxes[(size_t)dj * 6 + 4], vars, K, 0, 0.f);
          cur[beg + e] = __dmul_rn(cur[beg + e], exp(__ddiv_rn(-__dmul_rn(oks, oks), oks_thre)));
        }
        __syncthreads();
        --L;                                               // positions u' = 0 .. L-1 of the remaining list = ord[u' + 1]
        for (int u = tid; u < L; u += nt) {
          const int e = ord[beg + u + 1];
          const double se = cur[beg + e];
          int r = 0;
          for (int v = 0; v < L; ++v) r += (v != u) && score_before(cur[beg + ord[beg + v + 1]], v, se, u);
          tmp[beg + r] = e;
        }
        __syncthreads();
        for (int u = tid; u < L; u += nt) ord[beg + u] = tmp[beg + u];
        __syncthreads();
      }
    }
    if (nkeep == 0) {                                      // crowdpose.py:1321-1322 (only an empty image gets here)
      for (int a = tid; a < m; a += nt) keep[beg + a] = seg[beg + a];
      nkeep = m;
    }
    if (tid == 0) keep_counts[img] = nkeep;
    __syncthreads();
  }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
inline uint32_t table_size(int n) {
  uint32_t t = 64;
  while (t < 2u * (uint32_t)n) t <<= 1;
  return t;
}

}  // namespace

extern "C" int rsg_evaluate_workspace_bytes(int n, size_t* bytes) {
  RSG_REQUIRE(bytes && n >= 0, "rsg_evaluate_workspace_bytes: bad arguments");
  const size_t T = table_size(n), N = (size_t)(n > 0 ? n : 1);
  *bytes = align256(T * 8) + 3 * align256(T * 4) + 9 * align256(N * 4 + 4) + align256(N * 8) + 256;
  return RSG_OK;
}

extern "C" int rsg_evaluate(void* stream, const float* preds, const double* boxes, const int64_t* image_ids, int n, int K,
                            const double* sigmas, double in_vis_thre, double oks_thre, int soft_nms, int max_dets,
                            void* workspace, size_t workspace_bytes, int32_t* n_imgs, int64_t* images,
                            int32_t* img_offsets, double* scores, int32_t* keep, int32_t* keep_counts) {
  RSG_REQUIRE(n >= 0 && K > 0 && K <= RSG_NMS_MAXK, "rsg_evaluate: bad n=%d or K=%d", n, K);
  RSG_REQUIRE(n_imgs && images && img_offsets && scores && keep && keep_counts, "rsg_evaluate: null output pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    RSG_CUDA(cudaMemsetAsync(n_imgs, 0, sizeof(int32_t), s));
    RSG_CUDA(cudaMemsetAsync(img_offsets, 0, sizeof(int32_t), s));
    return RSG_OK;
  }
  RSG_REQUIRE(preds && boxes && image_ids && sigmas && workspace, "rsg_evaluate: null input pointer");
  RSG_REQUIRE(max_dets >= 1, "rsg_evaluate: max_dets=%d", max_dets);
  size_t need = 0;
  rsg_evaluate_workspace_bytes(n, &need);
  RSG_REQUIRE(workspace_bytes >= need, "rsg_evaluate: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  const uint32_t T = table_size(n);
  char* w = (char*)workspace;
  auto take = [&](size_t b) { char* p = w; w += align256(b); return p; };
  unsigned long long* keys = (unsigned long long*)take((size_t)T * 8);
  int* first = (int*)take((size_t)T * 4);
  int* cnt = (int*)take((size_t)T * 4);
  int* rank_of_slot = (int*)take((size_t)T * 4);
  const size_t nb = (size_t)n * 4 + 4;
  int* slot_of = (int*)take(nb);
  int* flag = (int*)take(nb);
  int* rank = (int*)take(nb);
  int* counts_by_rank = (int*)take(nb);
  int* cursor = (int*)take(nb);
  int* seg = (int*)take(nb);
  int* tmp = (int*)take(nb);
  int* ord = (int*)take(nb);
  int* dead = (int*)take(nb);
  double* cur = (double*)take((size_t)n * 8);
  int* err = (int*)take(4);
  RSG_CUDA(cudaMemsetAsync(keys, 0x80, (size_t)T * 8, s));
  RSG_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)T * 4, s));
  RSG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)T * 4, s));
  RSG_CUDA(cudaMemsetAsync(cursor, 0, nb, s));
  RSG_CUDA(cudaMemsetAsync(err, 0, 4, s));
  const int tb = 256, gb = (n + tb - 1) / tb;
  eval_hash_kernel<<<gb, tb, 0, s>>>((const long long*)image_ids, n, keys, first, cnt, T - 1, slot_of, err);
  RSG_LAUNCH_CHECK();
  eval_rank_kernel<<<1, 1024, 0, s>>>((const long long*)image_ids, n, slot_of, first, cnt, flag, rank, rank_of_slot,
                                      counts_by_rank, (long long*)images, img_offsets, n_imgs, err);
  RSG_LAUNCH_CHECK();
  eval_scatter_kernel<<<gb, tb, 0, s>>>(n, slot_of, rank_of_slot, img_offsets, cursor, seg);
  RSG_LAUNCH_CHECK();
  int grid = rsg_num_sms() * 8;
  if (grid > n) grid = n;
  eval_nms_kernel<<<grid, 128, 0, s>>>(preds, boxes, n_imgs, img_offsets, seg, tmp, ord, dead, cur, sigmas, K,
                                       (float)in_vis_thre, oks_thre, soft_nms ? 1 : 0, max_dets, scores, keep, keep_counts);
  RSG_LAUNCH_CHECK();
  return RSG_OK;
}
