// attention.cu -- fused multi-head attention with BEATs' gated relative-position bias applied in-tile.
//
// Replaces avex/models/beats/backbone.py:526-571: compute_bias expansion to [B,H,N,N], the gate
// (grep_linear -> view(..,2,4).sum -> sigmoid -> gate_a*(gate_b*grep_a-1)+2), the materialised
// `gate * position_bias` mask (3 GB fp32 per layer at B=256, N=496), the key-padding -inf mask, SDPA and the
// permute/contiguous that follows.  Nothing of size N^2 touches HBM: the Toeplitz bias is a [H, 2N-1] vector
// indexed as bias[h, j-i+N-1] inside the score tile, scaled by the per-row gate.
//
// v1 compute path: mma.sync m16n8k16 bf16 (fp32 accumulate), flash-style online softmax in fp32 (exp2 domain).
//   grid = (ceil(N/64), H, B), CTA = 128 threads: 4 warps x 16 query rows; K/V streamed in 64-key tiles through a
//   2-stage cp.async ring, 128-byte rows XOR-swizzled for conflict-free ldmatrix.
// Roofline: tensor pipe (4*N*N*64 FLOP per (b,h)); HBM traffic = qkv read once per q-tile (L2 resident) + out.
#include <math.h>

#include "common.cuh"

namespace avexk {
namespace {

constexpr int BQ = 64, BKV = 64, HD = 64, ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` of row `r` in a [rows][64] bf16 tile with XOR swizzle
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

struct AttnArgs {
  const __nv_bfloat16* qkv;  // [B*N, 3*H*64]
  int B, N, H;
  const float* gate_w;  // [2,64]
  const float* gate_b;  // [2]
  const float* grep_a;  // [H]
  const float* bias_vec;  // [H, 2N-1]
  const uint8_t* key_pad;  // [B,N] or null
  __nv_bfloat16* out;  // [B*N, H*64]
};

__global__ void __launch_bounds__(ATT_THREADS)
attention_gated_kernel(const AttnArgs a) {
  __shared__ __align__(128) unsigned char sQ[BQ * 128];
  __shared__ __align__(128) unsigned char sK[2][BKV * 128];
  __shared__ __align__(128) unsigned char sV[2][BKV * 128];
  __shared__ float sBias[2][128];   // bias window of the current kv tile, pre-multiplied by log2(e)
  __shared__ float sMask[2][BKV];   // 0 or -inf per key
  __shared__ float sGate[BQ];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int N = a.N, C3 = 3 * a.H * HD;
  const __nv_bfloat16* base = a.qkv + (size_t)b * N * C3;
  const __nv_bfloat16* qptr = base + h * HD;
  const __nv_bfloat16* kptr = base + a.H * HD + h * HD;
  const __nv_bfloat16* vptr = base + 2 * a.H * HD + h * HD;
  const float* bvec = a.bias_vec + (size_t)h * (2 * N - 1);
  const int n_kv = (N + BKV - 1) / BKV;

  auto load_kv = [&](int t, int buf) {
    const int kv0 = t * BKV;
    for (int i = tid; i < BKV * 8; i += ATT_THREADS) {
      const int r = i >> 3, c = i & 7;
      const bool ok = kv0 + r < N;
      const size_t g = (size_t)(ok ? kv0 + r : 0) * C3 + c * 8;
      cp_async16(sm_u32(sK[buf]) + tile_off(r, c), kptr + g, ok);
      cp_async16(sm_u32(sV[buf]) + tile_off(r, c), vptr + g, ok);
    }
    // bias window: x in [0,127) <-> (j - kv0) - (i - q0) + 63
    if (tid < 127) {
      const int idx = (kv0 - q0 - 63 + tid) + (N - 1);
      sBias[buf][tid] = (idx >= 0 && idx <= 2 * N - 2) ? __ldg(bvec + idx) * LOG2E : 0.f;
    }
    if (tid < BKV) {
      const int j = kv0 + tid;
      bool dead = j >= N;
      if (!dead && a.key_pad != nullptr) dead = a.key_pad[(size_t)b * N + j] != 0;
      sMask[buf][tid] = dead ? -INFINITY : 0.f;
    }
  };

  // ---- prologue: Q tile + first K/V tile ----------------------------------------------------------------------
  for (int i = tid; i < BQ * 8; i += ATT_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bool ok = q0 + r < N;
    cp_async16(sm_u32(sQ) + tile_off(r, c), qptr + (size_t)(ok ? q0 + r : 0) * C3 + c * 8, ok);
  }
  load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // gate per query row from UNscaled q (backbone.py:544-550): 2 threads per row, one per gate half
  {
    const int r = tid >> 1, half = tid & 1;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 raw = *reinterpret_cast<const uint4*>(sQ + tile_off(r, c));
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        acc = fmaf(f.x, __ldg(a.gate_w + half * 64 + c * 8 + 2 * e), acc);
        acc = fmaf(f.y, __ldg(a.gate_w + half * 64 + c * 8 + 2 * e + 1), acc);
      }
    }
    acc += __ldg(a.gate_b + half);
    const float sg = 1.0f / (1.0f + __expf(-acc));
    const float other = __shfl_xor_sync(0xffffffffu, sg, 1);
    if (half == 0) sGate[r] = sg * (other * __ldg(a.grep_a + h) - 1.0f) + 2.0f;  // gate_a*(gate_b*grep_a-1)+2
  }

  // Q fragments (held for the whole kernel): 4 k-steps of 16 along d
  uint32_t qf[4][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(sm_u32(sQ) + tile_off(row, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
  __syncthreads();  // sGate visible
  const int r_lo = warp * 16 + (lane >> 2), r_hi = r_lo + 8;  // rows of this thread inside the q tile
  const float gate_lo = sGate[r_lo], gate_hi = sGate[r_hi];
  const float qk_scale = 0.125f * LOG2E;  // head_dim^-0.5, backbone.py:403; exp2 domain

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int t = 0; t < n_kv; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_kv) {  // prefetch next tile into the other buffer (its previous readers passed the barrier below)
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
    }
    // ---- S = Q K^T (16 x 64 per warp) -----------------------------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    const uint32_t kb = sm_u32(sK[buf]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        uint32_t b0, b1, b2, b3;
        const int krow = np * 16 + (lane & 7) + (lane >> 4) * 8;
        ldsm_x4(kb + tile_off(krow, ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // ---- scores: scale, gated Toeplitz bias, key mask; online softmax ------------------------------------------
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int jc = nt * 8 + (lane & 3) * 2;  // key column inside the tile
      const float mk0 = sMask[buf][jc], mk1 = sMask[buf][jc + 1];
      const float* bw_lo = &sBias[buf][jc - r_lo + 63];
      const float* bw_hi = &sBias[buf][jc - r_hi + 63];
      s[nt][0] = fmaf(s[nt][0], qk_scale, fmaf(gate_lo, bw_lo[0], mk0));
      s[nt][1] = fmaf(s[nt][1], qk_scale, fmaf(gate_lo, bw_lo[1], mk1));
      s[nt][2] = fmaf(s[nt][2], qk_scale, fmaf(gate_hi, bw_hi[0], mk0));
      s[nt][3] = fmaf(s[nt][3], qk_scale, fmaf(gate_hi, bw_hi[1], mk1));
      mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float ms_lo = mn_lo == -INFINITY ? 0.f : mn_lo, ms_hi = mn_hi == -INFINITY ? 0.f : mn_hi;
    const float corr_lo = exp2f(m_lo - ms_lo), corr_hi = exp2f(m_hi - ms_hi);
    m_lo = mn_lo;
    m_hi = mn_hi;
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - ms_lo);
      s[nt][1] = exp2f(s[nt][1] - ms_lo);
      s[nt][2] = exp2f(s[nt][2] - ms_hi);
      s[nt][3] = exp2f(s[nt][3] - ms_hi);
      sum_lo += s[nt][0] + s[nt][1];
      sum_hi += s[nt][2] + s[nt][3];
    }
    l_lo = l_lo * corr_lo + sum_lo;
    l_hi = l_hi * corr_hi + sum_hi;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= corr_lo; o[dt][1] *= corr_lo;
      o[dt][2] *= corr_hi; o[dt][3] *= corr_hi;
    }
    // ---- O += P V  (P rounded to bf16, fp32 accumulate) ---------------------------------------------------------
    const uint32_t vb = sm_u32(sV[buf]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
      const uint32_t a0 = pack_bf16(s[2 * ks][0], s[2 * ks][1]), a1 = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      const uint32_t a2 = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]), a3 = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d tiles
        uint32_t b0, b1, b2, b3;
        const int vrow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(vb + tile_off(vrow, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    if (t + 1 < n_kv) cp_async_wait<0>();
    __syncthreads();  // next tile landed and everyone is done with `buf`
  }

  // ---- finalise: O / l -> bf16, staged through sQ for 16-byte coalesced stores ----------------------------------
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.0f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.0f / l_hi : 0.f;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + (lane & 3) * 2;  // element column; chunk = dt, byte offset inside chunk = (lane&3)*4
    *reinterpret_cast<uint32_t*>(sQ + tile_off(r_lo, dt) + (lane & 3) * 4) = pack_bf16(o[dt][0] * inv_lo, o[dt][1] * inv_lo);
    *reinterpret_cast<uint32_t*>(sQ + tile_off(r_hi, dt) + (lane & 3) * 4) = pack_bf16(o[dt][2] * inv_hi, o[dt][3] * inv_hi);
    (void)col;
  }
  __syncthreads();
  const int CO = a.H * HD;
  for (int i = tid; i < BQ * 8; i += ATT_THREADS) {
    const int r = i >> 3, c = i & 7;
    if (q0 + r < N)
      *reinterpret_cast<uint4*>(a.out + ((size_t)b * N + q0 + r) * CO + h * HD + c * 8) = *reinterpret_cast<const uint4*>(sQ + tile_off(r, c));
  }
}

}  // namespace
}  // namespace avexk

extern "C" int avexk_attention_gated(const void* qkv, int B, int N, int H, const float* gate_w, const float* gate_b,
                                     const float* grep_a, const float* bias_vec, const uint8_t* key_pad, void* out,
                                     void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(qkv && gate_w && gate_b && grep_a && bias_vec && out, "avexk_attention_gated: null argument");
  AVEXK_CHECK_ARG(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "avexk_attention_gated: bad shape B=%d N=%d H=%d", B, N, H);
  AVEXK_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "avexk_attention_gated: qkv/out must be 16-byte aligned");
  if (B == 0) return AVEXK_OK;
  AttnArgs a{reinterpret_cast<const __nv_bfloat16*>(qkv), B, N, H, gate_w, gate_b, grep_a, bias_vec, key_pad,
             reinterpret_cast<__nv_bfloat16*>(out)};
  dim3 grid(ceil_div(N, BQ), H, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prof_begin(st, KID_ATTN, 4.0 * B * H * (double)N * N * HD);
  attention_gated_kernel<<<grid, ATT_THREADS, 0, st>>>(a);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}
