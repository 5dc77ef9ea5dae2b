/* librsg_b200 -- C ABI of the B200-native (sm_100a) RSGNet per-crop inference hot path.
 *
 * Every entry point takes plain pointers and sizes; `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  Unless an entry point says "host", every data pointer is a
 * DEVICE pointer and the call is asynchronous on `stream`.  All functions return 0 on success;
 * otherwise a non-zero code, with a thread-local message available from rsg_last_error().
 * There is no CPU fallback anywhere in this library.
 *
 * Each group cites the reference interface it replaces (paths under the reference checkout).
 */
#ifndef RSG_B200_H_
#define RSG_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSG_ABI_VERSION 1

int rsg_abi_version(void);
const char* rsg_last_error(void);
/* Device properties the host side sizes work against: out[0]=SM count, out[1]=cc major,
 * out[2]=cc minor, out[3]=max dynamic smem per block (bytes). */
int rsg_device_info(int* out4);

/* ------------------------------------------------------------------------------------------
 * Post-processing.
 * ---------------------------------------------------------------------------------------- */

/* Fused flip-test average + argmax decode + quarter-pixel offset + inverse crop affine.
 * Replaces lib/core/function.py:417-427 (flip_back, 1-px shift, (a+b)*0.5),
 * lib/core/inference.py:21-82 (get_max_preds, get_final_preds) and
 * lib/utils/transforms.py:57-103 (transform_preds / get_affine_transform(inv=1)).
 *   hm          f32 [N,K,H,W]  heat-maps of the un-flipped crops
 *   hm_flipped  f32 [N,K,H,W]  RAW heat-maps of the W-flipped crops, or NULL (no flip test)
 *   flip_perm   i32 [K]        channel permutation pi(k) of flip_back (ignored when hm_flipped==NULL)
 *   center,scale f32 [N,2]     crop centre / scale (scale[:,1] is unused, as in the reference)
 *   post_process            TEST.POST_PROCESS  (quarter-pixel offset)
 *   shift                   TEST.SHIFT_HEATMAP (1-px shift of the flipped map)
 *   preds   f32 [N,K,2]  image-space coordinates          (NULL = skip)
 *   maxvals f32 [N,K]    value at the arg-max
 *   coords  f32 [N,K,2]  heat-map-space coordinates incl. the +-0.25 offset (NULL = skip)
 *   avg_out f32 [N,K,H,W] averaged heat-maps (NULL = never materialised) */
int rsg_flip_avg_decode(void* stream, const float* hm, const float* hm_flipped,
                        const int32_t* flip_perm, int N, int K, int H, int W,
                        const float* center, const float* scale, int post_process, int shift,
                        float* preds, float* maxvals, float* coords, float* avg_out);

/* lib/utils/transforms.py:23-37 flip_back: out[n,k,y,x] = in[n,pi(k),y,W-1-x]. */
int rsg_flip_back(void* stream, const float* in, float* out, const int32_t* flip_perm,
                  int N, int K, int H, int W);

/* Segmented greedy OKS-NMS, one image per CTA.  Replaces lib/nms/nms.py:75-124
 * (oks_iou + oks_nms) for all images of an evaluate() call at once
 * (lib/dataset/crowdpose.py:1315, lib/dataset/coco.py:1269).
 *   kpts    f32 [n,K,3] (x,y,score);  scores f64 [n];  areas f64 [n]
 *   img_offsets i32 [n_imgs+1]  detections of image i are [off[i], off[i+1])
 *   sigmas  f64 [K];  thresh: suppress when oks > thresh
 *   keep    i32 [n]   per image, at keep[off[i] ...]: kept indices RELATIVE to the image, in
 *                     greedy selection order;  keep_counts i32 [n_imgs] */
int rsg_oks_nms(void* stream, const float* kpts, const double* scores, const double* areas,
                const int32_t* img_offsets, int n_imgs, int max_per_img, const double* sigmas,
                int K, double thresh, int32_t* keep, int32_t* keep_counts, int use_in_vis_thre, double in_vis_thre);
/* use_in_vis_thre != 0 (the reference's in_vis_thre is not None, nms.py:85-90; all three OKS entry points): only key
 * points of the compared detection d with score > in_vis_thre enter the mean (the reference's `list(a) and list(b)` is the
 * second list), compared in fp32; a detection without visible key points has OKS 0. */

/* lib/nms/nms.py:75-94 oks_iou: OKS of one detection g (f32 [K,3]) against M detections d (f32 [M,K,3]);
 * a_g, a_d f64 areas; out f64 [M]. */
int rsg_oks_iou(void* stream, const float* g, const float* d, double a_g, const double* a_d,
                const double* sigmas, int K, int M, double* out, int use_in_vis_thre, double in_vis_thre);

/* Segmented soft OKS-NMS, one image per CTA.  Replaces lib/nms/nms.py:127-180 (rescore 'gaussian' + soft_oks_nms;
 * imported by lib/dataset/coco.py:24, used when TEST.SOFT_NMS is set): up to max_dets rounds (the reference hard-codes
 * 20), each keeping the best remaining detection and multiplying the other scores by exp(-oks^2 / thresh).
 *   keep i32 [n_imgs][max_dets]: kept indices RELATIVE to the image in selection order;  keep_counts i32 [n_imgs] */
int rsg_soft_oks_nms(void* stream, const float* kpts, const double* scores, const double* areas,
                     const int32_t* img_offsets, int n_imgs, int max_per_img, const double* sigmas, int K,
                     double thresh, int max_dets, int32_t* keep, int32_t* keep_counts, int use_in_vis_thre,
                     double in_vis_thre);

/* evaluate()-side rescoring, lib/dataset/crowdpose.py:1294-1306 / coco.py:1249-1261:
 * score[i] = box_score[i] * mean(maxvals[i,k] for maxvals[i,k] > in_vis_thre). */
int rsg_rescore(void* stream, const float* maxvals, const double* box_scores, int n, int K,
                double in_vis_thre, double* scores);
/* NumPy >= 2 (NEP 50) arithmetic is what is reproduced: the np.float32 maxvals are compared with the Python-float
 * threshold and accumulated in fp32; the mean is multiplied with the fp64 box score in fp64.  (Under NumPy 1.x the
 * same reference lines promote to fp64; that variant differs at ~1e-8 and is not reproduced.)
 * Order of EXACTLY equal scores in the NMS entry points: the later index first (= reversing a stable ascending sort,
 * what NumPy's argsort does for n <= 16); NumPy's order of exact ties is unspecified for longer lists.  NaN scores sort
 * first, as NumPy's argsort()[::-1] places them.  An image with more than max_per_img detections is refused: its keep_counts entry is set to -1. */

/* Device-resident evaluate(): group the detections by image, rescore, (soft-)OKS-NMS per image, compacted keep lists --
 * lib/dataset/crowdpose.py:1272-1324 and lib/dataset/coco.py:1227-1277 in one call with no host round trip (n_imgs
 * included: it stays on the device).  Inputs are what lib/core/function.py:376-380,452-458 accumulates:
 *   preds     f32 [n,K,3]  image-space x, y and the maxval of every joint (all_preds)
 *   boxes     f64 [n,6]    centre, scale, area (index 4), box score (index 5) (all_boxes)
 *   image_ids i64 [n]      the image of every detection, in ANY order (the reference parses it from the file name);
 *                          0x8080808080808080 is reserved (n_imgs = -1 is reported if it occurs)
 *   sigmas    f64 [K];  in_vis_thre = TEST.IN_VIS_THRE (rescoring only, like the reference), oks_thre = TEST.OKS_THRE,
 *   soft_nms = TEST.SOFT_NMS (max_dets rounds; the reference hard-codes 20)
 * Outputs (device):
 *   n_imgs      i32 [1]
 *   images      i64 [n]    the distinct image ids in first-appearance order (entries [0, n_imgs))
 *   img_offsets i32 [n+1]  kept detections of image r are keep[img_offsets[r] ... + keep_counts[r])
 *   scores      f64 [n]    rescored score of EVERY detection, indexed like the inputs
 *   keep        i32 [n]    GLOBAL detection indices, per image in selection order
 *   keep_counts i32 [n]
 * workspace: rsg_evaluate_workspace_bytes(n) bytes of device memory (no alignment requirement beyond 256). */
int rsg_evaluate_workspace_bytes(int n, size_t* bytes);
int rsg_evaluate(void* stream, const float* preds, const double* boxes, const int64_t* image_ids, int n, int K,
                 const double* sigmas, double in_vis_thre, double oks_thre, int soft_nms, int max_dets, void* workspace,
                 size_t workspace_bytes, int32_t* n_imgs, int64_t* images, int32_t* img_offsets, double* scores,
                 int32_t* keep, int32_t* keep_counts);

/* Person-crop affine warp + normalisation for a batch of crops (the loader's input path).
 * Replaces, per crop, cv2.warpAffine(img, trans, (out_w, out_h), flags=cv2.INTER_LINEAR) as called by
 * lib/dataset/CPJointsDataset.py:1281-1285 and lib/utils/transforms.py:121-129 (crop), and
 * transforms.Compose([ToTensor(), Normalize(mean, std)]) of tools/cp_test.py:107-115.  Bit-exact with OpenCV's 8-bit
 * fixed-point bilinear warp (BORDER_CONSTANT 0).  All pointers are DEVICE pointers.
 *   src      [N] pointers to uint8 HWC images with 3 channels (several crops may share one image)
 *   src_dims i32 [N][3]  rows, cols, row stride in bytes of each source image
 *   mats     f64 [N][6]  the forward 2x3 matrix `trans` (get_affine_transform, lib/utils/transforms.py:65-97)
 *   reverse_channels     1: output channel k = source channel 2-k (cv2.cvtColor BGR2RGB before the warp,
 *                        CPJointsDataset.py:1231-1232)
 *   out_u8   u8 [N][out_h][out_w][3] or NULL: the warped crop as cv2.warpAffine returns it
 *   out_f32  f32 [N][3][out_h][out_w] or NULL: the normalised model input; lut f32 [3][256] = ((u/255 - mean[k])/std[k])
 *            evaluated in fp32 by the host for every byte value (indexed by OUTPUT channel) */
int rsg_warp_affine(void* stream, const uint8_t* const* src, const int32_t* src_dims, const double* mats, int N,
                    int out_h, int out_w, int reverse_channels, uint8_t* out_u8, float* out_f32, const float* lut);

/* ------------------------------------------------------------------------------------------
 * Model plan: a flat list of device ops (the folded / packed network) executed per chunk of crops.
 * Replaces the nn.Module forward of lib/models/pose_rsgnet.py:955-1021 and
 * lib/models/pose_hrnet.py:428-463.  The host side (rsgnet_b200/_engine.py, or any other
 * host language) folds BatchNorm etc., packs weights into device memory, and describes the
 * network once with the rsg_plan_add_* calls; rsg_plan_run then executes it.
 * ---------------------------------------------------------------------------------------- */
typedef struct rsg_plan rsg_plan;

#define RSG_MAX_TAPS 16
#define RSG_MAX_RES 4

/* A pointer inside a plan is either absolute, or relative to an external per-run base:
 * address = (ext_slot < 0 ? ptr : ext_base[ext_slot]) + offset + first_crop * crop_stride. */
typedef struct rsg_ref {
  void* ptr;             /* absolute device pointer (ext_slot < 0) */
  int32_t ext_slot;      /* >= 0: index into the ext[] array given to rsg_plan_run */
  int64_t offset;        /* bytes */
  int64_t crop_stride;   /* bytes per crop, applied with the chunk's first crop (ext only) */
} rsg_ref;

typedef struct rsg_res {       /* residual / fuse term added in a conv epilogue */
  rsg_ref src;                 /* bf16 NHWC */
  int32_t cs, co;              /* channel stride / offset of the source buffer (elements) */
  int32_t H, W;                /* source spatial size; read at (y >> shift, x >> shift) */
  int32_t shift;
  int32_t batch_stride0;       /* 1: the same map for every crop (broadcast) */
} rsg_res;

typedef struct rsg_conv_desc {
  rsg_ref in;                  /* bf16 NHWC [N,Hin,Win,in_cs], channels [in_co, in_co+Cin) */
  int32_t in_cs, in_co, Hin, Win, Cin;
  rsg_ref w;                   /* bf16 [ntaps][CoutPad][CinPad32] (BN folded), generic kernel */
  rsg_ref w_tc5;               /* bf16 [CoutPad/NS][ntaps][Cin/8][NS][8] (same weights in the UMMA
                                  K-major core-matrix order, NS from rsg_conv_tc5_config) or NULL:
                                  enables the tcgen05 kernel */
  rsg_ref bias;                /* f32 [CoutPad] */
  int32_t Cout, CoutPad;
  int32_t ntaps;
  int8_t tap_dy[RSG_MAX_TAPS], tap_dx[RSG_MAX_TAPS];   /* input offset of each tap */
  int32_t stride;              /* input pixel = out * stride + tap */
  int32_t Hout, Wout;          /* iteration space */
  rsg_ref out;                 /* bf16 NHWC or NULL */
  int32_t out_cs, out_co, oH, oW, omul, ooy, oox;      /* out pixel = (y*omul+ooy, x*omul+oox) */
  rsg_ref out_f32;             /* f32 NCHW [N,Cout,oH,oW] or NULL */
  int32_t nres;
  rsg_res res[RSG_MAX_RES];
  int32_t relu;
  int32_t engine;              /* 0 = auto, 1 = force mma.sync path, 2 = force tcgen05 path,
                                  3 = weight-streaming tcgen05 path (w_tc5 packed with the NS of
                                  rsg_conv_ws_config), 4 = the same on CTA pairs (w_tc5 packed as
                                  rsg_conv_ws2_config describes) */
  int32_t pixel_shuffle_c;     /* > 0 (tcgen05 path only, omul = 2): the Cout = 4 * pixel_shuffle_c output
                                  columns are 4 sub-pixel phases of pixel_shuffle_c channels: column c goes
                                  to out pixel (2y + (c / psc) / 2, 2x + (c / psc) % 2), channel c % psc.
                                  ConvTranspose2d(4, 2, 1) as ONE 3x3 conv (pose_rsgnet.py:733-744) */
} rsg_conv_desc;

int rsg_plan_create(rsg_plan** out, int chunk);
void rsg_plan_destroy(rsg_plan*);
int rsg_plan_num_ops(const rsg_plan*);

/* conv1 of the stem (pose_rsgnet.py:612-613,922-924): fp32 NCHW [*,3,H,W] -> bf16 NHWC
 * [N,H/2,W/2,64], 3x3 s2 p1, BN folded, ReLU.  Forward f of a run reads crop f % n_crops,
 * W-reversed when f >= n_crops (the flip-test second forward, function.py:401). */
int rsg_plan_add_stem(rsg_plan*, rsg_ref x, int H, int W, rsg_ref w /*f32 [27][64]*/,
                      rsg_ref bias /*f32[64]*/, rsg_ref out);
int rsg_plan_add_conv(rsg_plan*, const rsg_conv_desc*);
/* out = relu(sum_i up_i(term_i)): HighResolutionModule fuse (pose_rsgnet.py:261-270). */
int rsg_plan_add_fuse(rsg_plan*, int nterms, const rsg_res* terms, rsg_ref out, int out_cs,
                      int out_co, int H, int W, int C, int relu);
/* 2x2 max-pool (association.py:286-287). */
int rsg_plan_add_maxpool(rsg_plan*, rsg_ref in, int cs, int co, int H, int W, int C, rsg_ref out);
/* TRP core (association.py:288-299): y[n,i,:] = sum_j sigmoid(x_i . x_j) g[n,j,:];
 * x, g, y are bf16 [N,S,C] views with channel strides/offsets. */
int rsg_plan_add_attention(rsg_plan*, rsg_ref x, int x_cs, int x_co, rsg_ref g, int g_cs,
                           int g_co, rsg_ref y, int y_cs, int y_co, int S, int C);
/* The same with y ALSO / INSTEAD as dense fp32 [N,S,C] (y_f32 non-null: the tcgen05 kernel writes fp32 only; y_bf16 is
 * then just the scratch of the mma.sync fallback and may be null when the views are 16-byte aligned), and the fp32 tail
 * of the TRP (association.py:236-245, 300): out = GroupNorm(8, C)(W y + b) with W f32 [C][C], b / gamma / beta f32 [C].
 * y is a sum of S sigmoid-weighted terms whose mean dwarfs its spread over the positions and GroupNorm removes the
 * mean, so this stretch is kept in fp32 end to end (bf16 storage of y or z costs ~10 % of the normalised signal). */
int rsg_plan_add_attention_f32(rsg_plan*, rsg_ref x, int x_cs, int x_co, rsg_ref g, int g_cs, int g_co, rsg_ref y_bf16,
                               int y_cs, int y_co, rsg_ref y_f32, int S, int C);
int rsg_plan_add_trp_tail(rsg_plan*, rsg_ref y_f32, rsg_ref w, rsg_ref bias, rsg_ref gamma, rsg_ref beta, int groups,
                          float eps, rsg_ref out, int out_cs, int out_co, int S, int C);
/* relation_scores f32 [N,S,S] = sigmoid(x x^T) (4th output of RSGNet.forward). */
int rsg_plan_add_relation_scores(rsg_plan*, rsg_ref x, int x_cs, int x_co, int S, int C,
                                 rsg_ref out);
/* GroupNorm(groups, C) over [S,C] per crop (association.py:243-245). */
int rsg_plan_add_groupnorm(rsg_plan*, rsg_ref in, int in_cs, int in_co, rsg_ref gamma,
                           rsg_ref beta, int groups, float eps, rsg_ref out, int out_cs,
                           int out_co, int S, int C);
/* Fused BasicBlock (pose_rsgnet.py:25-54): out = relu(bn2(conv2(relu(bn1(conv1(x))))) + x), two 3x3 stride-1
 * C -> C convs with BN folded into w1/b1, w2/b2 (bf16 [9][C/8][C][8] = the tcgen05 packing with NS = C, f32 [C]).
 * The intermediate never leaves the SM.  rsg_basic_block_supported says whether the fused kernel covers (C, H, W);
 * otherwise describe the block as two rsg_plan_add_conv ops. */
int rsg_basic_block_supported(int C, int H, int W);
int rsg_plan_add_basic_block(rsg_plan*, rsg_ref in, int in_cs, int in_co, int H, int W, int C, rsg_ref w1, rsg_ref b1,
                             rsg_ref w2, rsg_ref b2, rsg_ref out, int out_cs, int out_co);
/* Fused Bottleneck (pose_rsgnet.py:57-95; layer1, :619): out = relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x)))))))) + res)
 * with conv1 = 1x1 Cin -> 64, conv2 = 3x3 64 -> 64, conv3 = 1x1 64 -> 256, BN folded (weights bf16 in the tcgen05 packing
 * [tap][K/8][N][8] with N = the conv's own output channels, biases f32).  `res` is a bf16 NHWC tensor with 256 channels:
 * x itself for an identity block, the output of the block's 1x1 downsample conv otherwise.  Both intermediates stay in
 * shared memory.  rsg_bottleneck_supported says whether the fused kernel covers the shape; otherwise describe the block
 * as three rsg_plan_add_conv ops. */
int rsg_bottleneck_supported(int Cin, int planes, int Cout, int H, int W);
int rsg_plan_add_bottleneck(rsg_plan*, rsg_ref in, int in_cs, int in_co, int H, int W, int Cin, rsg_ref w1, rsg_ref b1,
                            rsg_ref w2, rsg_ref b2, rsg_ref w3, rsg_ref b3, rsg_ref res, int res_cs, int res_co,
                            rsg_ref out, int out_cs, int out_co);
/* bilinear x2, align_corners=True (+ optional sigmoid) on f32 NCHW (pose_rsgnet.py:1009-1013). */
int rsg_plan_add_bilinear2x(rsg_plan*, rsg_ref in, rsg_ref out, int C, int H, int W, int sigmoid);
/* Mark ops added after this call as "aux": skipped by rsg_plan_run unless with_aux != 0. */
int rsg_plan_begin_aux(rsg_plan*);

/* Run `n_fwd` forwards in chunks.  ext[] resolves rsg_ref.ext_slot.  n_crops: see add_stem
 * (n_fwd == n_crops: plain forward; n_fwd == 2*n_crops: flip-test batch).  use_graph: capture the
 * whole run into a CUDA graph keyed by (n_fwd, n_crops, with_aux, ext pointers) and replay it. */
int rsg_plan_run(rsg_plan*, void* stream, void* const* ext, int n_ext, int n_fwd, int n_crops,
                 int with_aux, int use_graph);
/* Measurement aid: run ONE chunk of `nb` forwards eagerly with a CUDA event pair around every op.
 * ms[i] = device time of op i, kind[i]: 0 stem, 1 conv (generic mma.sync kernel), 2 conv (tcgen05
 * kernel), 3 fuse, 4 maxpool, 5 attention, 6 relation_scores, 7 groupnorm / fp32 TRP tail, 8 bilinear, 9 conv
 * (weight-streaming tcgen05 kernel), 10 fused basic block, 11 fused bottleneck, 12 conv (weight-streaming CTA-pair kernel);
 * flops[i] = MAC*2 the op executes for nb forwards (0 for the element-wise ops).  Arrays must hold
 * rsg_plan_num_ops entries; ops skipped (aux) get ms = -1. */
int rsg_plan_profile(rsg_plan*, void* stream, void* const* ext, int n_ext, int nb, int n_crops,
                     int with_aux, float* ms, int32_t* kind, double* flops);
/* Kernel launches issued by the last rsg_plan_run (graph nodes when replayed). */
int rsg_plan_last_launches(const rsg_plan*);

/* mode: 0 = 1x1 stride 1, 1 = taps inside the 3x3 neighbourhood, stride 1, 2 = the same at stride 2.
 * Tiling the tcgen05 conv kernel uses for a shape: output channels per CTA (NS), channels per TMA
 * halo stage (KC) and ring depth (S).  Returns 0 when the shape is not covered (the generic kernel
 * runs instead).  The host packer needs NS to lay w_tc5 out as [CoutPad/NS][ntaps][Cin/8][NS][8]. */
int rsg_conv_tc5_config(int Cin, int CoutPad, int ntaps, int mode, int* NS, int* KC, int* S);

/* Weight-streaming tcgen05 conv kernel (many channels, small maps: the stage-3/4 low-resolution
 * branches): returns 1 and the output channels per CTA (NS) when it covers a stride-1 conv of this
 * shape on H x W maps, else 0.  The host packs w_tc5 as [CoutPad/NS][Cin/16][ntaps][2][NS][8] (the weights of a
 * 16-channel chunk are ONE contiguous block: one bulk copy per pipeline stage) and sets engine = 3. */
int rsg_conv_ws_config(int Cin, int CoutPad, int ntaps, int H, int W, int* NS);
/* The same for stride 1 or 2 (Hin x Win = the INPUT map).  Stride 2 (the fuse-layer / transition down-paths onto the
 * 16x12 and 8x6 maps, pose_rsgnet.py:219-246, 841-853) keeps the four pixel-parity phases of the input as separate
 * plane sets per stage (TMA boxes with an element stride of 2), so no MMA row or tap is wasted. */
int rsg_conv_ws_config2(int Cin, int CoutPad, int ntaps, int Hin, int Win, int stride, int* NS);

/* The same layers on CTA pairs (tcgen05.mma.cta_group::2, conv_ws2.cu): returns 1 when the pair kernel covers a stride-1
 * conv of this shape (Cout a multiple of 128, a 256-pixel supertile at least 70 % full).  The host then packs w_tc5 as
 * [CoutPad/128][2 halves][Cin/16][ntaps][2][64][8] (each CTA of a pair streams the 64 output channels of its half,
 * one contiguous block per 16-channel chunk; inside every group of 64 output channels accumulator column 8g + 2j + e
 * holds channel 16j + 2g + e, so that a lane quad of the epilogue owns one 128-byte line) and sets engine = 4. */
int rsg_conv_ws2_config(int Cin, int CoutPad, int ntaps, int H, int W);

/* Stand-alone conv launch (unit tests / micro-benchmarks): desc refs must be absolute. */
int rsg_conv_run(void* stream, const rsg_conv_desc*, int N);

/* ------------------------------------------------------------------------------------------
 * Training step (BASELINE.json configs[4]).  Replaces what torch autograd + cuDNN / cuBLAS + torch.optim
 * execute for lib/core/function.py:240-363 (rsgnet_train): train-mode forward of lib/models/pose_rsgnet.py:955-1021 (batch-
 * statistics BatchNorm), the losses (lib/core/loss.py:14-38 JointsMSELoss, BCELoss, the relation MSE), the backward pass
 * and Adam (lib/utils/utils.py:70-74).  Every tensor is fp32; activations are NHWC, i.e. row-major [pixels, channels]
 * matrices.  The host side (rsgnet_b200/train) sequences these calls with its own tape.
 * ---------------------------------------------------------------------------------------- */

/* C[b] (M x Nc, row pitch ldc) = (beta ? C[b] : 0) + gather(A[b]) . B[b] (+ bias[n]),  b < batch (strides sA/sB/sC elements).
 *   mode 0  plain GEMM: A is M x Ca (row pitch lda), or Ca x M when transA
 *   mode 1  convolution forward gather: C rows are the pixels (n, y, x) of the Hc x Wc grid, A rows the pixels of the
 *           Ha x Wa grid; tap (dy, dx) reads A pixel (y*stride - pad + dy, x*stride - pad + dx), zero outside
 *   mode 2  convolution input-gradient gather (= ConvTranspose forward): tap (dy, dx) reads A pixel
 *           ((y + pad - dy) / stride, (x + pad - dx) / stride) when both divisions are exact and inside the grid
 *   B       [taps][Ca][Nc] (row pitch ldb), or with transB [taps][Nc][Ca]: element (c, n) at B[tap*Nc*ldb + n*ldb + c]
 *           (the input-gradient pass reads the forward weights [tap][Cin][Cout] transposed)
 *   geom    {Ha, Wa, Hc, Wc, kh, kw, stride, pad} for modes 1 / 2 (taps = kh*kw), ignored for mode 0
 *   precise 0: TF32 mma.sync with fp32 accumulation; 1: 3xTF32 (fp32-class products), for parity tests */
int rsg_train_gemm(void* stream, const float* A, const float* B, float* C, const float* bias, int M, int Nc, int Ca,
                   int lda, int ldb, int ldc, int batch, long long sA, long long sB, long long sC, int mode,
                   int transA, int transB, int beta, const int* geom, int precise);

/* Weight gradient: dW[tap][ci][co] += sum over the M pixels m of dY's grid of X[src(m, tap)][ci] * dY[m][co]  (fp32 atomics
 * over pixel splits: dW must hold zeros or the gradient accumulated so far).  mode / geom as for rsg_train_gemm, the
 * gather applies to X.  With mode 0 this is dW[ci][co] += X^T dY (Linear / 1x1 weight gradients, either orientation).
 * precise: 0 = tensor pipe where a kernel covers the shape (3x3 stride-1, channel counts multiples of 8: tcgen05 on bf16
 * hi / lo splits, ~5e-6 of max; else TF32 mma.sync), 1 = 3xTF32 (~2e-7), 2 = TF32 mma.sync (~3e-4). */
int rsg_train_wgrad(void* stream, const float* X, const float* dY, float* dW, int M, int Ca, int Nc, int ldx, int ldy,
                    int mode, const int* geom, int precise);

/* BatchNorm over the rows of x [M, C] with batch statistics (torch batch_norm, training=True): y = (x - mean) * invstd *
 * gamma + beta (+ ReLU); running_mean / running_var (may be NULL) are updated with `momentum`, running_var with the
 * unbiased variance.  save_mean / save_invstd [C] feed the backward pass.  ws = 3*C + 4 doubles of scratch that must be ZERO on entry and is
 * left zero (the last reduction block finalizes the statistics and cleans up: no memset / finalize launches). */
int rsg_train_bn_fwd(void* stream, const float* x, long long M, int C, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, int relu, float* y, float* save_mean,
                     float* save_invstd, double* ws);
/* dx (may be NULL) = gamma * invstd * (dy' - mean(dy') - xhat * mean(dy' * xhat)), dgamma += sum dy' * xhat, dbeta += sum dy',
 * with dy' = dy * [y > 0] when relu (y = the forward output).  ws as for rsg_train_bn_fwd. */
int rsg_train_bn_bwd(void* stream, const float* x, const float* y, const float* dy, long long M, int C, const float* gamma,
                     const float* save_mean, const float* save_invstd, int relu, float* dx, float* dgamma, float* dbeta,
                     double* ws);
/* The same with the residual branch of a BasicBlock / Bottleneck fused in (pose_rsgnet.py:47-52, 88-93): y = relu(bn(x) + res)
 * (res NULL = plain), and in the backward pass dres = dy * [y > 0], the gradient of the residual input (dres NULL = none). */
int rsg_train_bn_fwd_res(void* stream, const float* x, long long M, int C, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, int relu, const float* res, float* y,
                         float* save_mean, float* save_invstd, double* ws);
int rsg_train_bn_bwd_res(void* stream, const float* x, const float* y, const float* dy, long long M, int C, const float* gamma,
                         const float* save_mean, const float* save_invstd, int relu, float* dx, float* dres, float* dgamma,
                         float* dbeta, double* ws);
/* out[c] (+)= sum over the M rows of x[m][c]  (bias gradients; backward of a batch broadcast).  ws = C doubles (cleared
 * before and after use). */
int rsg_train_colsum(void* stream, const float* x, long long M, int C, float* out, int accumulate, double* ws);

/* GroupNorm(G, C) over x [B, S, C] (association.py:243-246); mean / rstd are [B*G]. */
int rsg_train_gn_fwd(void* stream, const float* x, int B, int S, int C, int G, const float* gamma, const float* beta, float eps,
                     float* y, float* mean, float* rstd);
int rsg_train_gn_bwd(void* stream, const float* x, const float* dy, int B, int S, int C, int G, const float* gamma,
                     const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta);

/* out = in[0] + ... + in[nin-1] (+ ReLU); `in` is a HOST array of 1..4 device pointers. */
int rsg_train_add(void* stream, int nin, const float* const* in, int relu, long long n, float* out);
/* Element-wise: op 0 out = a * [b > 0] (ReLU backward, b = the ReLU output); 1 out = sigmoid(a); 2 out = a * b * (1 - b)
 * (sigmoid backward, b = the sigmoid output); 3 out = leaky_relu(a, slope); 4 out = a * (b > 0 ? 1 : slope); 5 out = a * b;
 * 6 out += a; 7 out = a * slope. */
int rsg_train_ew(void* stream, int op, const float* a, const float* b, float slope, long long n, float* out);
/* rows x cols copy between pitched matrices (pitches in elements; src_pitch 0 broadcasts one row; accumulate: dst += src):
 * channel concatenation / slicing and the batch repeat of the location branch. */
int rsg_train_copy2d(void* stream, const float* src, long long src_pitch, float* dst, long long dst_pitch, long long rows,
                     int cols, int accumulate);
/* dst[i0][i1][i2] (contiguous D0 x D1 x D2) (+)= src[i0*s0 + i1*s1 + i2*s2] for i1 < V1 and i2 < V2, else 0: NCHW <-> NHWC and
 * the weight packing OIHW -> [tap][Cin][Cout] / gradient unpacking. */
int rsg_train_permute3(void* stream, const float* src, float* dst, int D0, int D1, int D2, long long s0, long long s1,
                       long long s2, int V1, int V2, int accumulate);
typedef struct rsg_perm_entry {
  const float* src; float* dst;
  int32_t D0, D1, D2, V1, V2;
  long long s0, s1, s2;
  int32_t accumulate;
} rsg_perm_entry;
/* The same for a DEVICE table of `count` entries in one launch (all conv weights of a step). */
int rsg_train_permute3_batch(void* stream, const void* table, int count, int blocks_per_entry);
/* kind 0 nearest up-sampling x f ([N,h,w,C] -> [N,hf,wf,C]); 1 its backward (in = dy [N,hf,wf,C], out = dx); 2 bilinear x2
 * with align_corners=True; 3 its backward (out is zeroed here).  pose_rsgnet.py:207-219, 1009-1012. */
int rsg_train_resample(void* stream, int kind, const float* in, int N, int h, int w, int C, int f, float* out);
/* 2x2 max pooling of [N,H,W,C] (association.py:221); idx [N,H/2,W/2,C] keeps the arg-max for backward (in = dy, out = dx). */
int rsg_train_maxpool(void* stream, int backward, const float* in, unsigned char* idx, int N, int H, int W, int C, float* out);

/* JointsMSELoss (lib/core/loss.py:14-38) on NCHW [B,K,HW] with target_weight tw [B,K] (NULL = 1): *loss_acc += 0.5/(K B HW) *
 * sum (tw (pred - target))^2; grad (may be NULL) = up * d loss / d pred. */
int rsg_train_mse_joints(void* stream, const float* pred, const float* target, const float* tw, int B, int K, int HW, float up,
                         double* loss_acc, float* grad);
/* weight * BCELoss(mean) (function.py:253, 307): logs clamped at -100 as torch does. */
int rsg_train_bce(void* stream, const float* p, const float* t, long long n, double weight, float up, double* loss_acc,
                  float* grad);
/* pose_rsgnet.py:1014-1018: out_acc[b] += mean_ij (T[b,i,j] - P[b,i,j])^2; T [B,S,S] or NULL with T = v v^T, v [B,S]
 * (function.py:261-269 builds exactly that outer product). */
int rsg_train_relation_mse(void* stream, const float* P, const float* T, const float* v, int B, int S, double* out_acc);
/* The rank-1 factor of the relation target (lib/core/function.py:261-269): v [B, (H/2)*(W/2)] = bilinear(align_corners=True,
 * scale 1/2) of max_k target[b,k]; relation_target = v v^T is never materialised on this path. */
int rsg_train_person_mask(void* stream, const float* target, int B, int K, int H, int W, float* v);
/* dA = (dP + coef[b] * (P - T)) * P * (1 - P): backward of relation_score = sigmoid(A) with the relation-loss term folded
 * in (dP NULL = 0; coef NULL = no loss term; coef[b] = 2 / S^2 * d loss / d out[b]). */
int rsg_train_trp_dscore(void* stream, const float* P, const float* dP, const float* T, const float* v, const float* coef,
                         int B, int S, float* dA);
/* cudaMemsetAsync(p, 0, bytes) on `stream` (gradient buffers, loss accumulators). */
int rsg_train_zero(void* stream, void* p, size_t bytes);
/* out[i] (+)= (float)in[i] * scale  (loss accumulators are doubles). */
int rsg_train_d2f(void* stream, const double* in, float scale, int n, int accumulate, float* out);
/* torch.optim.Adam (no weight decay, no amsgrad) on flat buffers; step >= 1; the gradient is read as g * grad_scale
 * (1 / world_size after a summing all-reduce). */
int rsg_train_adam(void* stream, float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale);
/* The same with the step count in DEVICE memory: *step_dev is incremented, then used -- a captured CUDA graph of the whole
 * training step replays with the right bias corrections. */
int rsg_train_adam_graph(void* stream, float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                         float beta2, float eps, int* step_dev, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* RSG_B200_H_ */
