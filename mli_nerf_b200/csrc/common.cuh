// common.cuh -- error plumbing + launch helpers shared by the .cu files of libmli_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "mli_math.h"

void mli_set_error(const char* fmt, ...);
int mli_check_device();

#define MLI_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      mli_set_error(__VA_ARGS__);       \
      return MLI_EINVAL;                \
    }                                   \
  } while (0)

#define MLI_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      mli_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return MLI_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

#define MLI_LAUNCH_OK() MLI_CUDA_OK(cudaGetLastError())

#define MLI_ENTRY()                      \
  do {                                   \
    int _d = mli_check_device();         \
    if (_d != MLI_OK) return _d;         \
  } while (0)

static inline unsigned mli_cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

#define MLI_NUM_SMS 148
// SMs the persistent (one CTA per SM, full shared memory) kernels may occupy; the rest stay free for concurrently
// running communication kernels (mli_set_sm_limit, capi.cu)
int mli_sm_limit();

// block-wide sum of one float (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ float mli_block_sum(float v, float* smem32) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? smem32[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}
