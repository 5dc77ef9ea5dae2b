// Shared helpers for librsg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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

int rsg_num_sms();          // SM count of the CURRENT device (cached per device)

// Per-device one-time initialisation (cudaFuncSetAttribute is per device, and DataParallel-style callers drive
// several devices from several threads of one process): `static DeviceOnce once; if (once.first()) {...}`.
// first() returns true exactly until done() has been called for the current device.
#include <mutex>
struct DeviceOnce {
  static constexpr int MAX_DEV = 64;
  std::mutex mu;
  bool flag[MAX_DEV] = {};
  bool first(int* dev_out = nullptr) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_out) *dev_out = dev;
    if (dev < 0 || dev >= MAX_DEV) return true;
    std::lock_guard<std::mutex> g(mu);
    return !flag[dev];
  }
  void done() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEV) return;
    std::lock_guard<std::mutex> g(mu);
    flag[dev] = true;
  }
};

// Debug / tuning switches (RSG_*_SKIP, RSG_TC5_KC, ...) exist only in builds made with -DRSG_DEBUG_SWITCHES
// (`make DEBUG_SWITCHES=1`, what the tools/ micro-benchmarks use): the product library never reads the environment
// on a launch path, so no variable can silently change its results.
#ifdef RSG_DEBUG_SWITCHES
static inline const char* rsg_dbg_env(const char* name) { return getenv(name); }
#else
static inline const char* rsg_dbg_env(const char*) { return nullptr; }
#endif
static inline int rsg_dbg_int(const char* name, int dflt) {
  const char* e = rsg_dbg_env(name);
  return e ? atoi(e) : dflt;
}

// Programmatic dependent launch (PDL): the kernel may start while its stream predecessor is still draining;
// it must execute pdl_wait() before its first access to memory the predecessor produces or still reads.
// Hides the launch latency and the prologue (barrier init, TMEM allocation, weight loads) of the ~300 kernels
// of a step.  RSG_NO_PDL=1 restores plain stream-ordered launches.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  static const bool off = rsg_dbg_env("RSG_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif
