// Plan executor + C ABI glue: a model is a flat list of device ops (built once by the host side
// from the folded / packed network), run per chunk of crops so that activations stay L2-sized,
// optionally replayed as one CUDA graph.
#include <stdarg.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/rsg_b200.h"
#include "ops.cuh"

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void rsg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int rsg_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  rsg_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return RSG_ERR_CUDA;
}
int rsg_num_sms() {
  static int cache[DeviceOnce::MAX_DEV] = {};      // per device; a racing first call writes the same value twice
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= DeviceOnce::MAX_DEV) return 148;
  int n = cache[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev] = n;
  }
  return n;
}

extern "C" int rsg_abi_version(void) { return RSG_ABI_VERSION; }
extern "C" const char* rsg_last_error(void) { return g_err; }
extern "C" int rsg_device_info(int* out4) {
  RSG_REQUIRE(out4, "rsg_device_info: null");
  int dev = 0;
  RSG_CUDA(cudaGetDevice(&dev));
  RSG_CUDA(cudaDeviceGetAttribute(&out4[0], cudaDevAttrMultiProcessorCount, dev));
  RSG_CUDA(cudaDeviceGetAttribute(&out4[1], cudaDevAttrComputeCapabilityMajor, dev));
  RSG_CUDA(cudaDeviceGetAttribute(&out4[2], cudaDevAttrComputeCapabilityMinor, dev));
  RSG_CUDA(cudaDeviceGetAttribute(&out4[3], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------
enum OpKind { OP_STEM, OP_CONV, OP_FUSE, OP_MAXPOOL, OP_ATTN, OP_RELSCORES, OP_GROUPNORM, OP_BILINEAR, OP_BBLOCK, OP_BNECK, OP_TRPTAIL };

struct Op {
  OpKind kind;
  int aux;
  // generic slots
  rsg_ref r[10];
  int i[16];
  float f[2];
  rsg_conv_desc conv;
  rsg_res terms[RSG_MAX_RES];
};

struct GraphKey {
  int n_fwd, n_crops, with_aux;
  std::vector<void*> ext;
  bool operator<(const GraphKey& o) const {
    if (n_fwd != o.n_fwd) return n_fwd < o.n_fwd;
    if (n_crops != o.n_crops) return n_crops < o.n_crops;
    if (with_aux != o.with_aux) return with_aux < o.with_aux;
    return ext < o.ext;
  }
};
struct GraphEntry {
  int seen = 0;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
};

struct rsg_plan {
  int chunk;
  int aux_mode = 0;
  std::vector<Op> ops;
  std::map<GraphKey, GraphEntry> graphs;
  int last_launches = 0;
};

namespace {

struct RunCtx {
  void* const* ext;
  int n_ext;
  int f0;       // first forward of this chunk
};

inline void* resolve(const rsg_ref& r, const RunCtx& c) {
  if (r.ext_slot < 0) return r.ptr ? (char*)r.ptr + r.offset : nullptr;
  if (r.ext_slot >= c.n_ext || c.ext[r.ext_slot] == nullptr) return nullptr;
  return (char*)c.ext[r.ext_slot] + r.offset + (int64_t)c.f0 * r.crop_stride;
}
inline bool is_null(const rsg_ref& r) { return r.ext_slot < 0 && r.ptr == nullptr; }

ResP resolve_res(const rsg_res& s, const RunCtx& c) {
  ResP o;
  o.p = (const bf16*)resolve(s.src, c);
  o.cs = s.cs; o.co = s.co; o.H = s.H; o.W = s.W; o.shift = s.shift; o.bs0 = s.batch_stride0;
  return o;
}

int fill_conv(const rsg_conv_desc& d, const RunCtx& c, int N, ConvP* p) {
  RSG_REQUIRE(d.ntaps >= 1 && d.ntaps <= RSG_MAX_TAPS, "conv: ntaps=%d", d.ntaps);
  RSG_REQUIRE(d.nres >= 0 && d.nres <= RSG_MAX_RES, "conv: nres=%d", d.nres);
  RSG_REQUIRE(d.stride >= 1 && d.omul >= 1, "conv: stride/omul must be >= 1");
  memset(p, 0, sizeof(*p));
  p->in = (const bf16*)resolve(d.in, c);
  p->in_cs = d.in_cs; p->in_co = d.in_co; p->Hin = d.Hin; p->Win = d.Win; p->Cin = d.Cin;
  p->CinPad = (d.Cin + 31) / 32 * 32;
  p->w = (const bf16*)resolve(d.w, c);
  p->w_tc5 = is_null(d.w_tc5) ? nullptr : (const bf16*)resolve(d.w_tc5, c);
  p->bias = (const float*)resolve(d.bias, c);
  p->Cout = d.Cout; p->CoutPad = d.CoutPad;
  p->ntaps = d.ntaps;
  for (int t = 0; t < d.ntaps; ++t) { p->dy[t] = d.tap_dy[t]; p->dx[t] = d.tap_dx[t]; }
  p->stride = d.stride; p->Hout = d.Hout; p->Wout = d.Wout;
  p->out = is_null(d.out) ? nullptr : (bf16*)resolve(d.out, c);
  p->out_cs = d.out_cs; p->out_co = d.out_co; p->oH = d.oH; p->oW = d.oW;
  p->omul = d.omul; p->ooy = d.ooy; p->oox = d.oox; p->psC = d.pixel_shuffle_c;
  p->out_f32 = is_null(d.out_f32) ? nullptr : (float*)resolve(d.out_f32, c);
  p->nres = d.nres;
  for (int q = 0; q < d.nres; ++q) p->res[q] = resolve_res(d.res[q], c);
  p->relu = d.relu;
  p->force = d.engine == 2;
  p->N = N;
  p->M = (long long)N * d.Hout * d.Wout;
  RSG_REQUIRE(p->in && p->w && p->bias, "conv: unresolved in/w/bias pointer");
  RSG_REQUIRE(p->out || p->out_f32, "conv: no output");
  return RSG_OK;
}

int run_conv(const rsg_conv_desc& d, const RunCtx& c, int N, cudaStream_t s, int* used_tc5 = nullptr) {
  ConvP p;
  int rc = fill_conv(d, c, N, &p);
  if (rc) return rc;
  if (used_tc5) *used_tc5 = 0;
  if (d.engine == 0) {
    int handled = 0;
    rc = head1x1_launch(p, s, &handled);
    if (rc) return rc;
    if (handled) return RSG_OK;
  }
  if (d.engine == 4) {
    // w_tc5 holds the CTA-pair kernel's packing (rsg_conv_ws2_config): no other kernel can read it
    int handled = 0;
    rc = conv_ws2_launch(p, s, &handled);
    if (rc) return rc;
    RSG_REQUIRE(handled, "conv: shape not covered by the CTA-pair weight-streaming kernel (engine=4)");
    if (used_tc5) *used_tc5 = 3;
    return RSG_OK;
  }
  if (d.engine == 3) {
    // w_tc5 holds the weight-streaming kernel's packing (NS from rsg_conv_ws_config)
    int handled = 0;
    rc = conv_ws_launch(p, s, &handled);
    if (rc) return rc;
    if (handled) {
      if (used_tc5) *used_tc5 = 2;
      return RSG_OK;
    }
    return conv_mma_launch(p, s);
  }
  if (d.engine != 1) {
    int handled = 0;
    rc = conv_tc5_launch(p, s, &handled);
    if (rc) return rc;
    if (handled) {
      if (used_tc5) *used_tc5 = 1;
      return RSG_OK;
    }
    RSG_REQUIRE(d.engine != 2, "conv: shape not supported by the tcgen05 kernel (engine=2 forced)");
  }
  RSG_REQUIRE(p.psC == 0, "conv: pixel-shuffle output needs the tcgen05 kernel (shape not covered)");
  return conv_mma_launch(p, s);
}

int run_op(const Op& op, const RunCtx& c, int nb, int n_crops, cudaStream_t s, int* used_tc5 = nullptr) {
  switch (op.kind) {
    case OP_STEM:
      return stem_launch(s, (const float*)resolve(op.r[0], c), op.i[0], op.i[1],
                         (const float*)resolve(op.r[1], c), (const float*)resolve(op.r[2], c),
                         (bf16*)resolve(op.r[3], c), c.f0, nb, n_crops);
    case OP_CONV:
      return run_conv(op.conv, c, nb, s, used_tc5);
    case OP_FUSE: {
      ResP t[RSG_MAX_RES];
      for (int q = 0; q < op.i[0]; ++q) t[q] = resolve_res(op.terms[q], c);
      return fuse_launch(s, op.i[0], t, (bf16*)resolve(op.r[0], c), op.i[1], op.i[2], nb, op.i[3],
                         op.i[4], op.i[5], op.i[6]);
    }
    case OP_MAXPOOL:
      return maxpool_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1], nb, op.i[2],
                            op.i[3], op.i[4], (bf16*)resolve(op.r[1], c));
    case OP_ATTN: {
      int handled = 0;
      float* y32 = is_null(op.r[3]) ? nullptr : (float*)resolve(op.r[3], c);
      bf16* y16 = is_null(op.r[2]) ? nullptr : (bf16*)resolve(op.r[2], c);
      int rc = attention_tc5_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1],
                                    (const bf16*)resolve(op.r[1], c), op.i[2], op.i[3],
                                    y32 ? nullptr : y16, op.i[4], op.i[5], y32, nb, op.i[6], op.i[7], &handled);
      if (rc || handled) return rc;
      // shapes the tcgen05 kernel leaves out (unaligned views): the mma.sync kernel writes bf16, converted if fp32 is wanted
      RSG_REQUIRE(y16, "attention: this shape needs the bf16 output buffer (mma.sync fallback)");
      rc = attention_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1],
                            (const bf16*)resolve(op.r[1], c), op.i[2], op.i[3], y16, op.i[4], op.i[5], nb, op.i[6], op.i[7]);
      if (rc || !y32) return rc;
      return cvt_f32_launch(s, y16, op.i[4], op.i[5], y32, (long long)nb * op.i[6], op.i[7]);
    }
    case OP_TRPTAIL:
      return trp_tail_launch(s, (const float*)resolve(op.r[0], c), (const float*)resolve(op.r[1], c),
                             (const float*)resolve(op.r[2], c), (const float*)resolve(op.r[3], c),
                             (const float*)resolve(op.r[4], c), op.i[0], op.f[0], (bf16*)resolve(op.r[5], c), op.i[1], op.i[2],
                             nb, op.i[3], op.i[4]);
    case OP_RELSCORES:
      return relation_scores_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1], nb,
                                    op.i[2], op.i[3], (float*)resolve(op.r[1], c));
    case OP_GROUPNORM:
      return groupnorm_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1],
                              (const float*)resolve(op.r[1], c), (const float*)resolve(op.r[2], c),
                              op.i[2], op.f[0], (bf16*)resolve(op.r[3], c), op.i[3], op.i[4], nb,
                              op.i[5], op.i[6]);
    case OP_BBLOCK:
      return conv_bb_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1], nb, op.i[2], op.i[3], op.i[4],
                            (const bf16*)resolve(op.r[1], c), (const float*)resolve(op.r[2], c),
                            (const bf16*)resolve(op.r[3], c), (const float*)resolve(op.r[4], c),
                            (bf16*)resolve(op.r[5], c), op.i[5], op.i[6]);
    case OP_BNECK:
      return conv_bneck_launch(s, (const bf16*)resolve(op.r[0], c), op.i[0], op.i[1], nb, op.i[2], op.i[3], op.i[4],
                               (const bf16*)resolve(op.r[1], c), (const float*)resolve(op.r[2], c),
                               (const bf16*)resolve(op.r[3], c), (const float*)resolve(op.r[4], c),
                               (const bf16*)resolve(op.r[5], c), (const float*)resolve(op.r[6], c),
                               (const bf16*)resolve(op.r[7], c), op.i[5], op.i[6],
                               (bf16*)resolve(op.r[8], c), op.i[7], op.i[8]);
    case OP_BILINEAR:
      return bilinear2x_launch(s, (const float*)resolve(op.r[0], c), (float*)resolve(op.r[1], c),
                               nb * op.i[0], op.i[1], op.i[2], op.i[3]);
  }
  rsg_set_error("plan: unknown op kind");
  return RSG_ERR_STATE;
}

int run_all(rsg_plan* p, cudaStream_t s, void* const* ext, int n_ext, int n_fwd, int n_crops,
            int with_aux, int* launches) {
  int count = 0;
  for (int f0 = 0; f0 < n_fwd; f0 += p->chunk) {
    const int nb = n_fwd - f0 < p->chunk ? n_fwd - f0 : p->chunk;
    RunCtx c{ext, n_ext, f0};
    for (const Op& op : p->ops) {
      if (op.aux && !with_aux) continue;
      int rc = run_op(op, c, nb, n_crops, s);
      if (rc) return rc;
      ++count;
    }
  }
  *launches = count;
  return RSG_OK;
}

}  // namespace

extern "C" int rsg_plan_create(rsg_plan** out, int chunk) {
  RSG_REQUIRE(out && chunk >= 1, "rsg_plan_create: bad arguments");
  *out = new rsg_plan();
  (*out)->chunk = chunk;
  return RSG_OK;
}
extern "C" void rsg_plan_destroy(rsg_plan* p) {
  if (!p) return;
  for (auto& kv : p->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  delete p;
}
extern "C" int rsg_plan_num_ops(const rsg_plan* p) { return p ? (int)p->ops.size() : 0; }
extern "C" int rsg_plan_last_launches(const rsg_plan* p) { return p ? p->last_launches : 0; }
extern "C" int rsg_plan_begin_aux(rsg_plan* p) {
  RSG_REQUIRE(p, "null plan");
  p->aux_mode = 1;
  return RSG_OK;
}

static Op& new_op(rsg_plan* p, OpKind k) {
  p->ops.emplace_back();
  Op& op = p->ops.back();
  memset(&op, 0, sizeof(Op));
  for (rsg_ref& r : op.r) r.ext_slot = -1;          // unset refs are null, not "external slot 0"
  op.kind = k;
  op.aux = p->aux_mode;
  return op;
}

extern "C" int rsg_plan_add_stem(rsg_plan* p, rsg_ref x, int H, int W, rsg_ref w, rsg_ref bias,
                                 rsg_ref out) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_STEM);
  op.r[0] = x; op.r[1] = w; op.r[2] = bias; op.r[3] = out;
  op.i[0] = H; op.i[1] = W;
  return RSG_OK;
}
extern "C" int rsg_plan_add_conv(rsg_plan* p, const rsg_conv_desc* d) {
  RSG_REQUIRE(p && d, "null plan/desc");
  RSG_REQUIRE(d->CoutPad % 32 == 0 && d->CoutPad >= d->Cout, "conv: CoutPad=%d Cout=%d", d->CoutPad, d->Cout);
  Op& op = new_op(p, OP_CONV);
  op.conv = *d;
  return RSG_OK;
}
extern "C" int rsg_plan_add_fuse(rsg_plan* p, int nterms, const rsg_res* terms, rsg_ref out,
                                 int out_cs, int out_co, int H, int W, int C, int relu) {
  RSG_REQUIRE(p && terms && nterms >= 1 && nterms <= RSG_MAX_RES, "fuse: bad arguments");
  Op& op = new_op(p, OP_FUSE);
  for (int q = 0; q < nterms; ++q) op.terms[q] = terms[q];
  op.r[0] = out;
  op.i[0] = nterms; op.i[1] = out_cs; op.i[2] = out_co; op.i[3] = H; op.i[4] = W; op.i[5] = C; op.i[6] = relu;
  return RSG_OK;
}
extern "C" int rsg_plan_add_maxpool(rsg_plan* p, rsg_ref in, int cs, int co, int H, int W, int C,
                                    rsg_ref out) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_MAXPOOL);
  op.r[0] = in; op.r[1] = out;
  op.i[0] = cs; op.i[1] = co; op.i[2] = H; op.i[3] = W; op.i[4] = C;
  return RSG_OK;
}
extern "C" int rsg_plan_add_attention(rsg_plan* p, rsg_ref x, int x_cs, int x_co, rsg_ref g,
                                      int g_cs, int g_co, rsg_ref y, int y_cs, int y_co, int S,
                                      int C) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_ATTN);
  op.r[0] = x; op.r[1] = g; op.r[2] = y;
  op.i[0] = x_cs; op.i[1] = x_co; op.i[2] = g_cs; op.i[3] = g_co; op.i[4] = y_cs; op.i[5] = y_co;
  op.i[6] = S; op.i[7] = C;
  return RSG_OK;
}
extern "C" int rsg_plan_add_attention_f32(rsg_plan* p, rsg_ref x, int x_cs, int x_co, rsg_ref g, int g_cs, int g_co,
                                          rsg_ref y_bf16, int y_cs, int y_co, rsg_ref y_f32, int S, int C) {
  int rc = rsg_plan_add_attention(p, x, x_cs, x_co, g, g_cs, g_co, y_bf16, y_cs, y_co, S, C);
  if (rc) return rc;
  p->ops.back().r[3] = y_f32;
  return RSG_OK;
}
extern "C" int rsg_plan_add_trp_tail(rsg_plan* p, rsg_ref y_f32, rsg_ref w, rsg_ref bias, rsg_ref gamma, rsg_ref beta, int groups,
                                     float eps, rsg_ref out, int out_cs, int out_co, int S, int C) {
  RSG_REQUIRE(p, "null plan");
  RSG_REQUIRE(groups == 8 && (C == 16 || C == 32 || C == 48 || C == 64), "trp_tail: GroupNorm(8, C) with C in 16/32/48/64 (C=%d groups=%d)", C, groups);
  Op& op = new_op(p, OP_TRPTAIL);
  op.r[0] = y_f32; op.r[1] = w; op.r[2] = bias; op.r[3] = gamma; op.r[4] = beta; op.r[5] = out;
  op.i[0] = groups; op.i[1] = out_cs; op.i[2] = out_co; op.i[3] = S; op.i[4] = C;
  op.f[0] = eps;
  return RSG_OK;
}
extern "C" int rsg_plan_add_relation_scores(rsg_plan* p, rsg_ref x, int x_cs, int x_co, int S,
                                            int C, rsg_ref out) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_RELSCORES);
  op.r[0] = x; op.r[1] = out;
  op.i[0] = x_cs; op.i[1] = x_co; op.i[2] = S; op.i[3] = C;
  return RSG_OK;
}
extern "C" int rsg_plan_add_groupnorm(rsg_plan* p, rsg_ref in, int in_cs, int in_co, rsg_ref gamma,
                                      rsg_ref beta, int groups, float eps, rsg_ref out, int out_cs,
                                      int out_co, int S, int C) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_GROUPNORM);
  op.r[0] = in; op.r[1] = gamma; op.r[2] = beta; op.r[3] = out;
  op.i[0] = in_cs; op.i[1] = in_co; op.i[2] = groups; op.i[3] = out_cs; op.i[4] = out_co;
  op.i[5] = S; op.i[6] = C;
  op.f[0] = eps;
  return RSG_OK;
}
extern "C" int rsg_plan_add_basic_block(rsg_plan* p, rsg_ref in, int in_cs, int in_co, int H, int W, int C, rsg_ref w1,
                                        rsg_ref b1, rsg_ref w2, rsg_ref b2, rsg_ref out, int out_cs, int out_co) {
  RSG_REQUIRE(p, "null plan");
  RSG_REQUIRE(rsg_basic_block_supported(C, H, W), "basic block: C=%d on %dx%d is not covered by the fused kernel", C, H, W);
  Op& op = new_op(p, OP_BBLOCK);
  op.r[0] = in; op.r[1] = w1; op.r[2] = b1; op.r[3] = w2; op.r[4] = b2; op.r[5] = out;
  op.i[0] = in_cs; op.i[1] = in_co; op.i[2] = H; op.i[3] = W; op.i[4] = C; op.i[5] = out_cs; op.i[6] = out_co;
  return RSG_OK;
}
extern "C" int rsg_plan_add_bottleneck(rsg_plan* p, rsg_ref in, int in_cs, int in_co, int H, int W, int Cin, rsg_ref w1,
                                       rsg_ref b1, rsg_ref w2, rsg_ref b2, rsg_ref w3, rsg_ref b3, rsg_ref res, int res_cs,
                                       int res_co, rsg_ref out, int out_cs, int out_co) {
  RSG_REQUIRE(p, "null plan");
  RSG_REQUIRE(rsg_bottleneck_supported(Cin, 64, 256, H, W), "bottleneck: Cin=%d on %dx%d is not covered by the fused kernel", Cin, H, W);
  Op& op = new_op(p, OP_BNECK);
  op.r[0] = in; op.r[1] = w1; op.r[2] = b1; op.r[3] = w2; op.r[4] = b2; op.r[5] = w3; op.r[6] = b3; op.r[7] = res; op.r[8] = out;
  op.i[0] = in_cs; op.i[1] = in_co; op.i[2] = H; op.i[3] = W; op.i[4] = Cin; op.i[5] = res_cs; op.i[6] = res_co;
  op.i[7] = out_cs; op.i[8] = out_co;
  return RSG_OK;
}
extern "C" int rsg_plan_add_bilinear2x(rsg_plan* p, rsg_ref in, rsg_ref out, int C, int H, int W,
                                       int sigmoid) {
  RSG_REQUIRE(p, "null plan");
  Op& op = new_op(p, OP_BILINEAR);
  op.r[0] = in; op.r[1] = out;
  op.i[0] = C; op.i[1] = H; op.i[2] = W; op.i[3] = sigmoid;
  return RSG_OK;
}

extern "C" int rsg_plan_run(rsg_plan* p, void* stream, void* const* ext, int n_ext, int n_fwd,
                            int n_crops, int with_aux, int use_graph) {
  RSG_REQUIRE(p, "null plan");
  RSG_REQUIRE(n_fwd >= 0 && n_crops >= 1 && (n_fwd == n_crops || n_fwd == 2 * n_crops || n_fwd == 0),
              "rsg_plan_run: n_fwd=%d must be n_crops=%d or twice that", n_fwd, n_crops);
  cudaStream_t s = (cudaStream_t)stream;
  if (n_fwd == 0) return RSG_OK;
  if (!use_graph) return run_all(p, s, ext, n_ext, n_fwd, n_crops, with_aux, &p->last_launches);

  GraphKey key{n_fwd, n_crops, with_aux, std::vector<void*>(ext, ext + n_ext)};
  GraphEntry& e = p->graphs[key];
  if (e.exec) {
    RSG_CUDA(cudaGraphLaunch(e.exec, s));
    p->last_launches = e.launches;
    return RSG_OK;
  }
  if (e.seen++ == 0)   // first sight: run eagerly (sets function attributes, warms the module)
    return run_all(p, s, ext, n_ext, n_fwd, n_crops, with_aux, &p->last_launches);
  RSG_REQUIRE(s != nullptr, "rsg_plan_run: graph capture needs a non-default stream");
  RSG_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int launches = 0;
  int rc = run_all(p, s, ext, n_ext, n_fwd, n_crops, with_aux, &launches);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(s, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess) return rsg_cuda_fail(ce, "cudaStreamEndCapture", __FILE__, __LINE__);
  ce = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { e.exec = nullptr; return rsg_cuda_fail(ce, "cudaGraphInstantiate", __FILE__, __LINE__); }
  e.launches = launches;
  p->last_launches = launches;
  RSG_CUDA(cudaGraphLaunch(e.exec, s));
  return RSG_OK;
}

extern "C" int rsg_plan_profile(rsg_plan* p, void* stream, void* const* ext, int n_ext, int nb,
                                int n_crops, int with_aux, float* ms, int32_t* kind, double* flops) {
  RSG_REQUIRE(p && ms && kind && flops, "rsg_plan_profile: null argument");
  RSG_REQUIRE(nb >= 1 && nb <= p->chunk, "rsg_plan_profile: nb=%d must be in [1, chunk=%d]", nb, p->chunk);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = p->ops.size();
  std::vector<cudaEvent_t> ev(2 * n, nullptr);
  RunCtx c{ext, n_ext, 0};
  int rc = RSG_OK;
  for (size_t i = 0; i < n && rc == RSG_OK; ++i) {
    const Op& op = p->ops[i];
    ms[i] = -1.f; flops[i] = 0.0;
    static const int kmap[] = {0, 1, 3, 4, 5, 6, 7, 8, 10, 11, 7};
    kind[i] = kmap[op.kind];
    if (op.aux && !with_aux) continue;
    cudaEventCreate(&ev[2 * i]); cudaEventCreate(&ev[2 * i + 1]);
    cudaEventRecord(ev[2 * i], s);
    int tc5 = 0;
    rc = run_op(op, c, nb, n_crops, s, &tc5);
    cudaEventRecord(ev[2 * i + 1], s);
    if (op.kind == OP_CONV) {
      if (tc5) kind[i] = tc5 == 3 ? 12 : (tc5 == 2 ? 9 : 2);
      flops[i] = 2.0 * op.conv.ntaps * op.conv.Cin * op.conv.Cout * (double)op.conv.Hout * op.conv.Wout * nb;
      if (op.conv.pixel_shuffle_c) flops[i] *= 16.0 / 36.0;      // the zero taps of the fused deconv are not credited
    } else if (op.kind == OP_BBLOCK) {
      flops[i] = 2.0 * 2.0 * 9 * op.i[4] * op.i[4] * (double)op.i[2] * op.i[3] * nb;
    } else if (op.kind == OP_BNECK) {
      flops[i] = 2.0 * (64.0 * op.i[4] + 9.0 * 64 * 64 + 64.0 * 256) * (double)op.i[2] * op.i[3] * nb;
    } else if (op.kind == OP_ATTN) {
      flops[i] = 4.0 * (double)op.i[6] * op.i[6] * op.i[7] * nb;
    } else if (op.kind == OP_STEM) {
      flops[i] = 2.0 * 27 * 64 * (double)(op.i[0] / 2) * (op.i[1] / 2) * nb;
    }
  }
  cudaError_t ce = cudaStreamSynchronize(s);
  for (size_t i = 0; i < n; ++i) {
    if (ev[2 * i]) {
      if (ce == cudaSuccess && rc == RSG_OK) cudaEventElapsedTime(&ms[i], ev[2 * i], ev[2 * i + 1]);
      cudaEventDestroy(ev[2 * i]); cudaEventDestroy(ev[2 * i + 1]);
    }
  }
  if (rc) return rc;
  if (ce != cudaSuccess) return rsg_cuda_fail(ce, "cudaStreamSynchronize", __FILE__, __LINE__);
  return RSG_OK;
}

extern "C" int rsg_conv_run(void* stream, const rsg_conv_desc* d, int N) {
  RSG_REQUIRE(d, "null desc");
  RunCtx c{nullptr, 0, 0};
  return run_conv(*d, c, N, (cudaStream_t)stream);
}
