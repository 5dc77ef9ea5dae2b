// Shared helpers for librsg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef __nv_bfloat16 bf16;

#define RSG_OK 0
#define RSG_ERR_ARG 1
#define RSG_ERR_CUDA 2
#define RSG_ERR_STATE 3

void rsg_set_error(const char* fmt, ...);
int rsg_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define RSG_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return rsg_cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define RSG_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      rsg_set_error(__VA_ARGS__);              \
      return RSG_ERR_ARG;                      \
    }                                          \
  } while (0)

#define RSG_LAUNCH_CHECK() RSG_CUDA(cudaGetLastError())

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int rsg_num_sms();
