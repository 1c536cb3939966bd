// elementwise.cu -- the memory-bound glue of the BEATs path: LayerNorm, pos-conv operand packing, masked mean-pool, weight packing.  All HBM-bound: one pass, 16-byte
// vector accesses, a warp per row with shuffle reductions.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace avexk {
namespace {

// residual of a bf16 rounding, itself rounded to bf16: v ~= hi + lo with ~16 mantissa bits
__device__ __forceinline__ uint2 split_lo4(const float4& v, const uint2& hi) {
  const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&hi.y);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  return make_uint2(pack_bf16(v.x - f0.x, v.y - f0.y), pack_bf16(v.z - f1.x, v.w - f1.y));
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm (beats.py:353, backbone.py:176-177, :362, :373).  One warp per row, row held in registers.
// ---------------------------------------------------------------------------------------------------------
template <int VEC>  // VEC float4 per lane: C = 128 * VEC
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int split3) {
  constexpr int C = 128 * VEC;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
    q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = fmaf(v[i].x * rstd, g.x, b.x); y.y = fmaf(v[i].y * rstd, g.y, b.y);
    y.z = fmaf(v[i].z * rstd, g.z, b.z); y.w = fmaf(v[i].w * rstd, g.w, b.w);
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * C)[c4] = y;
    if (out_bf16 && !split3) reinterpret_cast<uint2*>(out_bf16 + (size_t)row * C)[c4] = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
    if (out_bf16 && split3) {  // [hi | lo | hi] operand of a 3-term split-bf16 GEMM (row pitch 3C)
      const uint2 hi = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      const uint2 lo = split_lo4(y, hi);
      uint2* o = reinterpret_cast<uint2*>(out_bf16 + (size_t)row * 3 * C);
      o[c4] = hi;
      o[C / 4 + c4] = lo;
      o[2 * (C / 4) + c4] = hi;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// pos-conv operand: x0 [M, C] fp32 -> xg [M, G*64] fp16 (16-bit storage typed __nv_bfloat16 in the signatures) (each group's C/G = 48 channels padded to 64 so that one
// tap of one group is a 128-byte TMA row).  Rows of padded tokens are zeroed (backbone.py:169-170) in xg AND in x0.
// ---------------------------------------------------------------------------------------------------------
// The pos-conv operands are FP16, not bf16: the projection output is bounded (LayerNorm(512) -> Linear), the tensor core runs
// kind::f16 at the same rate for both formats, and the three extra mantissa bits matter -- the 6144-term pos-conv sum was the
// largest single contributor to the end-to-end error of the bf16 path (tools/emulate_bf16.py: final max-abs 0.0196 -> 0.0039
// on the perturbed-weights golden).  Values beyond the fp16 range saturate instead of becoming inf.
__global__ void __launch_bounds__(256)
group_pad_kernel(float* __restrict__ x0, const uint8_t* __restrict__ key_pad, long long M, int G, int cg,
                 __nv_bfloat16* __restrict__ xg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, group, 8-channel chunk)
  const long long total = M * G * 8;
  if (idx >= total) return;
  const int ch = idx & 7, g = (idx >> 3) % G;
  const long long row = idx / (8 * G);
  const bool dead = key_pad != nullptr && key_pad[row] != 0;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (ch * 8 < cg) {
    float4* src = reinterpret_cast<float4*>(x0 + (size_t)row * (G * cg) + g * cg + ch * 8);
    if (dead) {
      src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float4 a = src[0], b = src[1];
      o = make_uint4(pack_h16(a.x, a.y), pack_h16(a.z, a.w), pack_h16(b.x, b.y), pack_h16(b.z, b.w));
    }
  }
  reinterpret_cast<uint4*>(xg + (size_t)row * (G * 64) + g * 64)[ch] = o;
}

// ---------------------------------------------------------------------------------------------------------
// mean over tokens: x [B,N,C] -> [B,C]; masked mean when the clip has padded tokens (beats_model.py:269-275)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mean_pool_kernel(const float* __restrict__ x, const uint8_t* __restrict__ key_pad, int any_pad, int N, int C,
                 float* __restrict__ out) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = x + (size_t)b * N * C + c;
  float s = 0.f;
  int cnt = 0;
  if (key_pad != nullptr && any_pad) {
    const uint8_t* kp = key_pad + (size_t)b * N;
    for (int n = 0; n < N; ++n)
      if (!kp[n]) { s += p[(size_t)n * C]; ++cnt; }
    out[(size_t)b * C + c] = s / (float)(cnt > 0 ? cnt : 1);
  } else {
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int n = 0;
    for (; n + 3 < N; n += 4) {
      s += p[(size_t)n * C]; s1 += p[(size_t)(n + 1) * C]; s2 += p[(size_t)(n + 2) * C]; s3 += p[(size_t)(n + 3) * C];
    }
    for (; n < N; ++n) s += p[(size_t)n * C];
    out[(size_t)b * C + c] = ((s + s1) + (s2 + s3)) / (float)N;
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight packing (run once at load time)
// ---------------------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// W [N,K] fp32 -> [N, 3K] bf16 = [hi | hi | lo]: with an [hi | lo | hi] activation the GEMM computes
// a_hi w_hi + a_lo w_hi + a_hi w_lo (the three significant terms of (a_hi + a_lo)(w_hi + w_lo)).
__global__ void f32_to_bf16_split3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K) {
  const long long total = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / K, k = i % K;
    const float w = src[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    __nv_bfloat16* row = dst + n * 3 * K;
    row[k] = hi;
    row[K + k] = hi;
    row[2 * K + k] = lo;
  }
}

// per-tap norm of v over (out-channel, in-channel): nrm[t] = sqrt(sum_{co,ci} v[co,ci,t]^2)   (weight_norm dim=2)
__global__ void posconv_norm_kernel(const float* __restrict__ v, int CoCi, int K, float* __restrict__ nrm) {
  const int t = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < CoCi; i += blockDim.x) {
    const double x = v[(size_t)i * K + t];
    s += x * x;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) nrm[t] = (float)sqrt(sh[0]);
}

// Wpc [G][K][cg][64] bf16: Wpc[grp, t, co, ci] = g[t] * v[grp*cg + co, ci, t] / nrm[t] for ci < cg, 0 for the pad channels
// (consecutive taps of one group are consecutive 128-byte rows: the pos-conv kernel fetches four taps per TMA box).
__global__ void posconv_pack_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ nrm,
                                    int C, int cg, int K, __nv_bfloat16* __restrict__ W) {
  const long long total = (long long)C * K * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = i & 63;
    const long long q = i >> 6;  // (grp, t, co)
    const int co = q % cg, t = (q / cg) % K, grp = q / ((long long)cg * K);
    float w = 0.f;
    if (ci < cg) w = g[t] * v[((size_t)(grp * cg + co) * cg + ci) * K + t] / nrm[t];
    const __half hw = __float2half_rn(w);  // weight-normalised taps, |w| << 1: fp16 keeps three more bits than bf16
    W[i] = *reinterpret_cast<const __nv_bfloat16*>(&hw);
  }
}

// gate_w[2,64] / gate_b[2]: grep_linear rows summed in groups of four (view(..,2,4).sum(-1), backbone.py:547)
__global__ void gate_pack_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ gw,
                                 float* __restrict__ gb) {
  const int t = threadIdx.x;  // 128 threads
  const int half = t >> 6, d = t & 63;
  gw[t] = (w[(half * 4 + 0) * 64 + d] + w[(half * 4 + 1) * 64 + d]) + (w[(half * 4 + 2) * 64 + d] + w[(half * 4 + 3) * 64 + d]);
  if (t < 2) gb[t] = (b[t * 4] + b[t * 4 + 1]) + (b[t * 4 + 2] + b[t * 4 + 3]);
}

}  // namespace

int launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, float* out_f32,
                     void* out_bf16, cudaStream_t st, int split3) {
  AVEXK_CHECK_ARG(C % 128 == 0 && C >= 128 && C <= 1024, "layernorm: C=%d must be a multiple of 128 up to 1024", C);
  if (M == 0) return AVEXK_OK;
  const int grid = ceil_div(M, 8);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  prof_begin(st, KID_LAYERNORM, (double)M * C * (4.0 + (out_f32 ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)));
  switch (C / 128) {
    case 1: layernorm_kernel<1><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 2: layernorm_kernel<2><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 3: layernorm_kernel<3><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 4: layernorm_kernel<4><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 5: layernorm_kernel<5><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 6: layernorm_kernel<6><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    case 7: layernorm_kernel<7><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
    default: layernorm_kernel<8><<<grid, 256, 0, st>>>(x, M, gamma, beta, eps, out_f32, ob, split3); break;
  }
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_group_pad(float* x0, const uint8_t* key_pad, long long M, int G, int cg, __nv_bfloat16* xg, cudaStream_t st) {
  AVEXK_CHECK_ARG(cg % 8 == 0 && cg <= 64, "group_pad: channels per group %d unsupported", cg);
  const long long total = M * G * 8;
  if (total == 0) return AVEXK_OK;
  group_pad_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x0, key_pad, M, G, cg, xg);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

__global__ void pool_finalize_kernel(const long long* __restrict__ acc, long long n, float scale, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)((double)acc[i] * (1.0 / 16777216.0)) * scale;
}

int launch_pool_finalize(const long long* acc, int B, int C, float scale, float* out, cudaStream_t st) {
  const long long n = (long long)B * C;
  if (n == 0) return AVEXK_OK;
  pool_finalize_kernel<<<ceil_div(n, 256), 256, 0, st>>>(acc, n, scale, out);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_mean_pool(const float* x, const uint8_t* key_pad, int any_pad, int B, int N, int C, float* out, cudaStream_t st) {
  if (B == 0) return AVEXK_OK;
  dim3 grid(ceil_div(C, 256), B);
  mean_pool_kernel<<<grid, 256, 0, st>>>(x, key_pad, any_pad, N, C, out);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(fminf(fmaxf(src[i], -65504.f), 65504.f));
}

int launch_f32_to_f16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n == 0) return AVEXK_OK;
  int grid = ceil_div(n, 256);
  if (grid > 4096) grid = 4096;
  f32_to_f16_kernel<<<grid, 256, 0, st>>>(src, reinterpret_cast<__half*>(dst), n);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st) {
  if (n == 0) return AVEXK_OK;
  int grid = ceil_div(n, 256);
  if (grid > 4096) grid = 4096;
  f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, dst, n);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_f32_to_bf16_split3(const float* src, __nv_bfloat16* dst, int N, int K, cudaStream_t st) {
  f32_to_bf16_split3_kernel<<<1024, 256, 0, st>>>(src, dst, N, K);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_posconv_pack(const float* v, const float* g, int C, int cg, int K, float* nrm_ws, __nv_bfloat16* W, cudaStream_t st) {
  posconv_norm_kernel<<<K, 256, 0, st>>>(v, C * cg, K, nrm_ws);
  AVEXK_LAUNCH_CHECK();
  posconv_pack_kernel<<<2048, 256, 0, st>>>(v, g, nrm_ws, C, cg, K, W);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_gate_pack(const float* w, const float* b, float* gw, float* gb, cudaStream_t st) {
  gate_pack_kernel<<<1, 128, 0, st>>>(w, b, gw, gb);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

}  // namespace avexk

extern "C" int avexk_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, float* out_f32,
                               void* out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(x && gamma && beta && (out_f32 || out_bf16) && M >= 0, "avexk_layernorm: null argument");
  return launch_layernorm(x, M, C, gamma, beta, eps, out_f32, out_bf16, reinterpret_cast<cudaStream_t>(stream), 0);
}
