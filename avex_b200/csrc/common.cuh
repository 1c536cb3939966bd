// common.cuh -- shared helpers for libavexk (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avexk.h"

namespace avexk {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define AVEXK_CHECK_ARG(cond, ...)        \
  do {                                    \
    if (!(cond)) {                        \
      ::avexk::set_error(__VA_ARGS__);    \
      return AVEXK_EINVAL;                \
    }                                     \
  } while (0)

#define AVEXK_CUDA(call)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::avexk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AVEXK_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define AVEXK_LAUNCH_CHECK()                                                                  \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::avexk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AVEXK_ECUDA;                                                                     \
    }                                                                                         \
    ::avexk::count_launch();                                                                  \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int num_sms();

// kernel ids of the optional event profiler (avexk_profile_*)
enum { KID_FBANK = 0, KID_GEMM = 1, KID_ATTN = 2, KID_LAYERNORM = 3, KID_POSCONV = 4, KID_OTHER = 5 };
bool prof_enabled();
void prof_begin(cudaStream_t st, int kid, double work);
void prof_end(cudaStream_t st);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace avexk
