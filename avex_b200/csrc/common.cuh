// common.cuh -- shared helpers for libavexk (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
int current_device();  // clamped to [0, 64)

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

// Exact (erf) GELU for two values with packed fp32x2 arithmetic and one special-function op per value:
// gelu(v) = relu(v) - |v| * (erfc(|v| / sqrt2) / 2) with log2(erfc(|v| / sqrt2) / 2) as a degree-5 polynomial (|err| <= 6.5e-7
// over the real line; derivation in gemm_tc.cu).
__device__ __forceinline__ float2 gelu_erf_fast2(float2 v) {
  const float2 n = make_float2(-fabsf(v.x), -fabsf(v.y));
  const float2 r = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  float2 q = __ffma2_rn(make_float2(4.712853462e-4f, 4.712853462e-4f), n, make_float2(7.064215249e-3f, 7.064215249e-3f));
  q = __ffma2_rn(q, n, make_float2(5.175841926e-2f, 5.175841926e-2f));
  q = __ffma2_rn(q, n, make_float2(-4.600918231e-1f, -4.600918231e-1f));
  q = __ffma2_rn(q, n, make_float2(1.150727814f, 1.150727814f));
  q = __ffma2_rn(q, n, make_float2(-1.000049354f, -1.000049354f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  return __ffma2_rn(n, e, r);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp16 storage of bounded activations (EfficientNet path, pos-conv operands): three more mantissa bits than bf16 at the same
// tensor-core rate; values beyond the fp16 range saturate instead of becoming inf
__device__ __forceinline__ uint32_t pack_h16(float lo, float hi) {
  uint32_t r;  // one F2FP.SATFINITE: |x| > 65504 -> +-65504, NaN stays NaN
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h16(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

}  // namespace avexk
