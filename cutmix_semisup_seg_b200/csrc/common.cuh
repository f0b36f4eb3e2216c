// Shared helpers for libb200seg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200seg.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200seg is written for sm_100a (Blackwell B200) only"
#endif

extern thread_local char g_b2_err[512];

static inline int b2_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_b2_err, sizeof(g_b2_err), fmt, ap);
  va_end(ap);
  return code;
}

#define B2_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return b2_fail(B2_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

#define B2_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return b2_fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                                    \
  } while (0)

#define B2_LAUNCH_CHECK(name)                                                                \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return b2_fail(B2_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));  \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int b2_sm_count_cached();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide fixed-order sum of doubles (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ double block_sum_d(double v, double* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.0;
    r = warp_sum_d(r);
  }
  return r;
}
