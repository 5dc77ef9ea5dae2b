// Launchers of the non-GEMM kernels (ops.cu) and the TRP attention kernel (attention.cu).
#pragma once
#include "conv_params.cuh"

// returns *handled = 1 when the fp32-NCHW 1x1 head kernel covers this conv
int head1x1_launch(const ConvP& p, cudaStream_t s, int* handled);
int stem_launch(cudaStream_t s, const float* x, int H, int W, const float* w, const float* bias,
                bf16* out, int f0, int nb, int n_crops);
int fuse_launch(cudaStream_t s, int nterms, const ResP* terms, bf16* out, int out_cs, int out_co,
                int N, int H, int W, int C, int relu);
int maxpool_launch(cudaStream_t s, const bf16* in, int cs, int co, int N, int H, int W, int C,
                   bf16* out);
int groupnorm_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, const float* gamma,
                     const float* beta, int groups, float eps, bf16* out, int out_cs, int out_co,
                     int N, int S, int C);
int bilinear2x_launch(cudaStream_t s, const float* in, float* out, int NC, int H, int W,
                      int sigmoid);
int relation_scores_launch(cudaStream_t s, const bf16* x, int cs, int co, int N, int S, int C,
                           float* out);
int attention_launch(cudaStream_t s, const bf16* x, int x_cs, int x_co, const bf16* g, int g_cs,
                     int g_co, bf16* y, int y_cs, int y_co, int N, int S, int C);
// tcgen05 version of the TRP core (attention_tc5.cu); *handled = 0 when the shape is left to attention_launch
int attention_tc5_launch(cudaStream_t s, const bf16* x, int x_cs, int x_co, const bf16* g, int g_cs, int g_co,
                         bf16* y, int y_cs, int y_co, float* y32, int N, int S, int C, int* handled);
// TRP tail in fp32 (association.py:236-245, 300): out = GroupNorm(groups, C)(W y + b), y fp32 [N,S,C] dense, W fp32 [C][C]
int trp_tail_launch(cudaStream_t s, const float* y32, const float* w, const float* bias, const float* gamma, const float* beta,
                    int groups, float eps, bf16* out, int out_cs, int out_co, int N, int S, int C);
// bf16 [N,S,C] view -> dense fp32 [N,S,C]
int cvt_f32_launch(cudaStream_t s, const bf16* in, int cs, int co, float* out, long long rows, int C);
// fused BasicBlock (conv_bb.cu): out = relu(conv2(relu(conv1(x))) + x), both 3x3 s1, BN folded, C = 32 / 48
int conv_bb_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, int N, int H, int W, int C, const bf16* w1,
                   const float* b1, const bf16* w2, const float* b2, bf16* out, int out_cs, int out_co);
extern "C" int rsg_basic_block_supported(int C, int H, int W);
// fused Bottleneck (conv_bneck.cu): out = relu(conv3(relu(conv2(relu(conv1(x))))) + res), 1x1 Cin->64, 3x3 64->64, 1x1 64->256
int conv_bneck_launch(cudaStream_t s, const bf16* in, int in_cs, int in_co, int N, int H, int W, int Cin, const bf16* w1,
                      const float* b1, const bf16* w2, const float* b2, const bf16* w3, const float* b3, const bf16* res,
                      int res_cs, int res_co, bf16* out, int out_cs, int out_co);
extern "C" int rsg_bottleneck_supported(int Cin, int planes, int Cout, int H, int W);
