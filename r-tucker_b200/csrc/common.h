// Shared helpers for librtucker_b200 (sm_100a only).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/rtucker.h"
#include "../../include/rtucker_debug.h"

namespace rt {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;  // kernels launched by this library in this process
extern double g_small_flops;           // fp64 flops of the GEMM operations RECORDED by the small stage (host side, eager launches)

#define RT_CHECK_CUDA(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      rt::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                    cudaGetErrorString(_e));                                           \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define RT_REQUIRE(cond, ...)                                                          \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      rt::set_error(__VA_ARGS__);                                                      \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define RT_LAUNCH_CHECK()                 \
  do {                                    \
    ++rt::g_launches;                     \
    RT_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (kernel, device) instead of on every launch
inline cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
  constexpr int kSlots = 64, kDevs = 16;
  static const void* funcs[kSlots];
  static size_t done[kSlots][kDevs];
  int dev = 0;
  cudaGetDevice(&dev);
  int slot = 0;
  while (slot < kSlots && funcs[slot] && funcs[slot] != func) ++slot;
  if (slot < kSlots && dev < kDevs) {
    if (funcs[slot] == func && done[slot][dev] >= bytes) return cudaSuccess;
    funcs[slot] = func;
  }
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && slot < kSlots && dev < kDevs) done[slot][dev] = bytes;
  return e;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rt
