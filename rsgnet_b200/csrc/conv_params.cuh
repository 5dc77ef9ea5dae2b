// Device-side (resolved) parameter blocks shared by the conv kernels and the plan executor.
#pragma once
#include "common.cuh"

struct ResP {
  const bf16* p;
  int cs, co, H, W, shift, bs0;
};

struct ConvP {
  const bf16* in;
  int in_cs, in_co, Hin, Win, Cin, CinPad;
  const bf16* w;        // [ntaps][CoutPad][CinPad]
  const bf16* w_tc5;    // [CoutPad/NS][ntaps][Cin/8][NS][8] or nullptr (tcgen05 kernel layout)
  const float* bias;    // [CoutPad]
  int Cout, CoutPad;
  int ntaps;
  int8_t dy[16], dx[16];
  int stride;
  int Hout, Wout;
  bf16* out;
  int out_cs, out_co, oH, oW, omul, ooy, oox;
  int psC;              // > 0: pixel-shuffle output (tcgen05 kernel only): column c -> phase c / psC, channel c % psC
  float* out_f32;
  int nres;
  ResP res[4];
  int relu;
  int force;            // engine == 2: the caller insists on the tcgen05 kernel (unit tests): no size-based routing to the generic kernel
  int N;
  long long M;          // N * Hout * Wout
};

int conv_mma_launch(const ConvP& p, cudaStream_t s);
// returns RSG_OK and sets *handled=1 when the tcgen05 kernel covers this shape
int conv_tc5_launch(const ConvP& p, cudaStream_t s, int* handled);
// weight-streaming tcgen05 kernel (conv_ws.cu): many-channel stride-1 convs on small maps; w_tc5 must be
// packed with the NS of rsg_conv_ws_config
int conv_ws_launch(const ConvP& p, cudaStream_t s, int* handled);
// the same on CTA pairs (conv_ws2.cu, tcgen05.mma.cta_group::2): w_tc5 packed as [CoutPad/128][2][ntaps][Cin/8][64][8]
int conv_ws2_launch(const ConvP& p, cudaStream_t s, int* handled);
