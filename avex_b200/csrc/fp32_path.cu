// fp32_path.cu -- the pieces of the fp32 mode of the BEATs path (north_star: max-abs <= 1e-3 against the fp32 reference).
//
// In fp32 mode every nn.Linear still runs on the tcgen05 GEMM, but as a 3-term split-bf16 product
//   x W^T ~= x_hi W_hi^T + x_lo W_hi^T + x_hi W_lo^T      (operands [hi|lo|hi] x [hi|hi|lo], K tripled, fp32 accumulation)
// which carries ~16 mantissa bits per operand (tools/emulate_bf16.py).  What cannot be expressed that way runs here in plain
// fp32 on the CUDA cores: the gated relative-position-bias attention with q, k, v, P kept in fp32
// (avex/models/beats/backbone.py:526-571) and the convolutional position embedding (backbone.py:52-68, :172-174).  This is a
// validation / high-precision mode: ~10x slower than the bf16 path, same C ABI, no library calls.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace avexk {
namespace {

// ---------------------------------------------------------------------------------------------------------
// fp32 row [M, K] -> bf16 [M, 3K] = [hi | lo | hi]: the A operand of a 3-term split GEMM
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ src, long long M, int K, __nv_bfloat16* __restrict__ dst) {
  const long long total = M * (K / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (K / 4);
    const int c4 = (int)(i % (K / 4));
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + row * K) + c4);
    const uint2 hi = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.x));
    const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.y));
    const uint2 lo = make_uint2(pack_bf16(v.x - h0.x, v.y - h0.y), pack_bf16(v.z - h1.x, v.w - h1.y));
    uint2* o = reinterpret_cast<uint2*>(dst + row * 3 * K);
    o[c4] = hi;
    o[K / 4 + c4] = lo;
    o[2 * (K / 4) + c4] = hi;
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 attention with the gated relative-position bias.  One thread per query row (q and the output accumulator in
// registers), 128 rows per CTA, keys / values streamed through shared memory 32 at a time, exact online softmax.
//   qkv [B*N, 3*H*64] fp32 (q | k | v, token-major)   out [B*N, H*64] fp32
// ---------------------------------------------------------------------------------------------------------
constexpr int AQ = 128, AK = 32, HD = 64;
__global__ void __launch_bounds__(AQ)
attention_fp32_kernel(const float* __restrict__ qkv, int B, int N, int H, const float* __restrict__ gate_w,
                      const float* __restrict__ gate_b, const float* __restrict__ grep_a, const float* __restrict__ bias_vec,
                      const uint8_t* __restrict__ key_pad, float* __restrict__ out) {
  __shared__ __align__(16) float sK[AK][HD];
  __shared__ __align__(16) float sV[AK][HD];
  __shared__ float sDead[AK];
  const int b = blockIdx.z, h = blockIdx.y, i = blockIdx.x * AQ + threadIdx.x;
  const int C3 = 3 * H * HD;
  const bool row_ok = i < N;
  float q[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + ((size_t)b * N + (row_ok ? i : 0)) * C3 + h * HD);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      const float4 v = __ldg(qp + d);
      q[4 * d] = v.x; q[4 * d + 1] = v.y; q[4 * d + 2] = v.z; q[4 * d + 3] = v.w;
    }
  }
  // gate from the UNscaled q (backbone.py:544-550; rows of grep_linear pre-summed in groups of four)
  float ga = __ldg(gate_b), gb = __ldg(gate_b + 1);
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    ga = fmaf(q[d], __ldg(gate_w + d), ga);
    gb = fmaf(q[d], __ldg(gate_w + HD + d), gb);
  }
  ga = 1.0f / (1.0f + expf(-ga));
  gb = 1.0f / (1.0f + expf(-gb));
  const float gate = ga * (gb * __ldg(grep_a + h) - 1.0f) + 2.0f;
  const float* bias_row = bias_vec + (size_t)h * (2 * N - 1) + (N - 1) - i;  // + j
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < N; j0 += AK) {
    __syncthreads();
    for (int e = threadIdx.x; e < AK * (HD / 4); e += AQ) {
      const int jj = e / (HD / 4), d4 = e % (HD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + jj < N) {
        const float* base = qkv + ((size_t)b * N + j0 + jj) * C3 + h * HD;
        kv = __ldg(reinterpret_cast<const float4*>(base + H * HD) + d4);
        vv = __ldg(reinterpret_cast<const float4*>(base + 2 * H * HD) + d4);
      }
      reinterpret_cast<float4*>(&sK[jj][0])[d4] = kv;
      reinterpret_cast<float4*>(&sV[jj][0])[d4] = vv;
    }
    if (threadIdx.x < AK) {
      const int j = j0 + threadIdx.x;
      sDead[threadIdx.x] = (j >= N || (key_pad != nullptr && key_pad[(size_t)b * N + j] != 0)) ? 1.f : 0.f;
    }
    __syncthreads();
    if (!row_ok) continue;
#pragma unroll 1
    for (int jj = 0; jj < AK; ++jj) {
      if (sDead[jj] != 0.f) continue;  // masked key: contributes exp(-inf) = 0 (backbone.py:554-559)
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&sK[jj][d]);
        s0 = fmaf(q[d], kk.x, s0); s1 = fmaf(q[d + 1], kk.y, s1);
        s0 = fmaf(q[d + 2], kk.z, s0); s1 = fmaf(q[d + 3], kk.w, s1);
      }
      const float s = (s0 + s1) * 0.125f + gate * __ldg(bias_row + j0 + jj);
      if (s > m) {
        const float c = expf(m - s);  // 0 when m = -inf
        l *= c;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] *= c;
        m = s;
      }
      const float p = expf(s - m);
      l += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&sV[jj][d]);
        o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
  }
  if (row_ok) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    float4* op = reinterpret_cast<float4*>(out + ((size_t)b * N + i) * (H * HD) + h * HD);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) op[d] = make_float4(o[4 * d] * inv, o[4 * d + 1] * inv, o[4 * d + 2] * inv, o[4 * d + 3] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 pos-conv.  Wf [G][taps][cg][cg] fp32: Wf[g, t, ci, co] = g[t] * v[g*cg + co, ci, t] / nrm[t]
// ---------------------------------------------------------------------------------------------------------
__global__ void posconv_pack_f32_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ nrm,
                                        int G, int cg, int K, float* __restrict__ W) {
  const long long total = (long long)G * K * cg * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = i % cg, ci = (i / cg) % cg, t = (i / ((long long)cg * cg)) % K, grp = i / ((long long)cg * cg * K);
    W[i] = g[t] * v[((size_t)(grp * cg + co) * cg + ci) * K + t] / nrm[t];
  }
}

constexpr int PT = 64, PCG = 48, PTAPS = 128, PCHUNK = 4;  // tokens per CTA, channels per group, taps, taps per weight chunk
__global__ void __launch_bounds__(256)
posconv_fp32_kernel(const float* __restrict__ x0, const float* __restrict__ Wf, const float* __restrict__ bias, int B, int N, int G,
                    float* __restrict__ out) {
  extern __shared__ __align__(16) float psm[];
  float* xs = psm;                                // [PT + PTAPS - 1][PCG]
  float* ws = psm + (PT + PTAPS - 1) * PCG;       // [PCHUNK][PCG][PCG]
  const int b = blockIdx.z, grp = blockIdx.y, n0 = blockIdx.x * PT, C = G * PCG;
  for (int e = threadIdx.x; e < (PT + PTAPS - 1) * (PCG / 4); e += blockDim.x) {
    const int r = e / (PCG / 4), c4 = e % (PCG / 4);
    const int n = n0 + r - PTAPS / 2;  // zero padding outside the clip (padded tokens were zeroed by group_pad)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= 0 && n < N) v = __ldg(reinterpret_cast<const float4*>(x0 + ((size_t)b * N + n) * C + grp * PCG) + c4);
    reinterpret_cast<float4*>(xs + r * PCG)[c4] = v;
  }
  const int nl = threadIdx.x >> 2, cq = threadIdx.x & 3;  // token inside the tile, group of 12 output channels
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int t0 = 0; t0 < PTAPS; t0 += PCHUNK) {
    __syncthreads();
    const float4* wsrc = reinterpret_cast<const float4*>(Wf + ((size_t)grp * PTAPS + t0) * PCG * PCG);
    for (int e = threadIdx.x; e < PCHUNK * PCG * PCG / 4; e += blockDim.x) reinterpret_cast<float4*>(ws)[e] = __ldg(wsrc + e);
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < PCHUNK; ++tt) {
      const float* xr = xs + (nl + t0 + tt) * PCG;
      const float* wr = ws + tt * PCG * PCG + cq * 12;
#pragma unroll 4
      for (int ci = 0; ci < PCG; ++ci) {
        const float xv = xr[ci];
        const float4 w0 = *reinterpret_cast<const float4*>(wr + ci * PCG), w1 = *reinterpret_cast<const float4*>(wr + ci * PCG + 4);
        const float4 w2 = *reinterpret_cast<const float4*>(wr + ci * PCG + 8);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        acc[8] = fmaf(xv, w2.x, acc[8]); acc[9] = fmaf(xv, w2.y, acc[9]); acc[10] = fmaf(xv, w2.z, acc[10]); acc[11] = fmaf(xv, w2.w, acc[11]);
      }
    }
  }
  const int n = n0 + nl;
  if (n < N) {
    const size_t off = ((size_t)b * N + n) * C + grp * PCG + cq * 12;
#pragma unroll
    for (int k = 0; k < 12; ++k) out[off + k] = __ldg(x0 + off + k) + gelu_erf(acc[k] + __ldg(bias + grp * PCG + cq * 12 + k));
  }
}

}  // namespace

int launch_split3_rows(const float* src, long long M, int K, __nv_bfloat16* dst, cudaStream_t st) {
  AVEXK_CHECK_ARG(K % 4 == 0, "split3_rows: K=%d must be a multiple of 4", K);
  if (M == 0) return AVEXK_OK;
  const long long total = M * (K / 4);
  const int grid = (int)(total / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  split3_rows_kernel<<<grid, 256, 0, st>>>(src, M, K, dst);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_attention_fp32(const float* qkv, int B, int N, int H, const float* gate_w, const float* gate_b, const float* grep_a,
                          const float* bias_vec, const uint8_t* key_pad, float* out, cudaStream_t st) {
  AVEXK_CHECK_ARG(B <= 65535 && H <= 65535, "attention_fp32: grid limits (B=%d H=%d)", B, H);
  if (B == 0 || N == 0) return AVEXK_OK;
  dim3 grid(ceil_div(N, AQ), H, B);
  prof_begin(st, KID_ATTN, 4.0 * B * H * (double)N * N * HD);
  attention_fp32_kernel<<<grid, AQ, 0, st>>>(qkv, B, N, H, gate_w, gate_b, grep_a, bias_vec, key_pad, out);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

// nrm_ws [K] must already hold the per-tap norms (launch_posconv_pack computes them)
int launch_posconv_pack_f32(const float* v, const float* g, const float* nrm, int G, int cg, int K, float* W, cudaStream_t st) {
  posconv_pack_f32_kernel<<<2048, 256, 0, st>>>(v, g, nrm, G, cg, K, W);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_posconv_fp32(const float* x0, const float* Wf, const float* bias, float* out, int B, int N, int G, int cg, int taps,
                        cudaStream_t st) {
  AVEXK_CHECK_ARG(cg == PCG && taps == PTAPS, "posconv_fp32 is specialised to 48 channels/group and 128 taps (got %d, %d)", cg, taps);
  AVEXK_CHECK_ARG(B <= 65535, "posconv_fp32: B=%d exceeds grid.z", B);
  if (B == 0 || N == 0) return AVEXK_OK;
  const int smem = ((PT + PTAPS - 1) * PCG + PCHUNK * PCG * PCG) * sizeof(float);
  static bool attr_set[64] = {};  // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    AVEXK_CUDA(cudaFuncSetAttribute(posconv_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev_] = true;
  }
  dim3 grid(ceil_div(N, PT), G, B);
  prof_begin(st, KID_POSCONV, 2.0 * B * N * (double)(G * cg) * cg * taps);
  posconv_fp32_kernel<<<grid, 256, smem, st>>>(x0, Wf, bias, B, N, G, out);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

}  // namespace avexk
