// gemm_tc.cu -- persistent, warp-specialised bf16 GEMM on the sm_100a tensor cores.
//
//   D[M,N] = epilogue(A[M,K] @ W[N,K]^T)      (A, W bf16 K-major; fp32 accumulation in TMEM)
//
// Replaces every nn.Linear on the BEATs path (backbone.py:531-533 q/k/v, :572 out_proj, :365-370 fc1/fc2,
// beats.py:350 patch-embed as im2col GEMM, :359 post_extract_proj) and the elementwise ops that follow them
// (bias, exact GELU, DeepNorm residual `residual * alpha + x`, backbone.py:360,:372), fused in the epilogue.
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer : cp.async.bulk.tensor 128x64 (A) + 256x64 (W) bf16 tiles, 128B swizzle, 4-stage ring
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma 128x256x16 (cta_group::1), commits to mbarriers
//   warps 2..9  epilogue     : tcgen05.ld 32x32b.x32 from TMEM -> registers -> per-warp smem transposition ->
//                              coalesced 16-byte global stores; the fp32 residual is register-prefetched one chunk ahead
//   TMEM: 512 columns = two 128x256 fp32 accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Roofline: dense bf16 tensor pipe; algorithmic FLOPs = 2*M*N*K.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int STG_BYTES_PER_WARP = 32 * 32 * 4;                // 32 rows x 32 fp32, XOR-swizzled (no padding)
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 256;
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;

struct GemmArgs {
  int M, N, K;
  const float* bias;
  int gelu;
  float* raw_out;
  const float* residual;
  float res_scale;
  void* out;
  long long ldo;
  int out_bf16;
  // conv mode (EfficientNet 1x1 convolutions): y = act(acc * scale[n] + bias[n]) (+ res_bf16), raw_out = acc
  const float* scale;
  int silu;
  const __nv_bfloat16* res_bf16;
};

// Exact GELU (modules.py:191-200) for two values per call with packed fp32x2 arithmetic.  erf via Abramowitz-Stegun
// 7.1.26 (|err| < 1.5e-7, far inside the bf16 rounding of the output).  With a = |v|, z = a / sqrt(2) and
// E(z) = erf(z) = 1 - P(t) e^{-z^2}, t = 1 / (1 + p z):   gelu(v) = 0.5 v + 0.5 a E(z)   (v erf(v/sqrt2) = a E(z)).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 gelu_fast2(float2 v) {
  const float2 a = make_float2(fabsf(v.x), fabsf(v.y));
  const float2 d = __ffma2_rn(a, make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f),
                              make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  p = __fmul2_rn(p, t);
  // e^{-z^2} = 2^{-(a k)^2}, k = sqrt(log2(e) / 2)
  constexpr float k = 0.84932180028801904272f;
  const float2 zk = __fmul2_rn(a, make_float2(k, k)), nzk = __fmul2_rn(a, make_float2(-k, -k));
  const float2 q = __fmul2_rn(zk, nzk);
  const float2 pe = __fmul2_rn(p, make_float2(ex2_approx(q.x), ex2_approx(q.y)));
  const float2 ha = __fmul2_rn(a, make_float2(0.5f, 0.5f)), nha = __fmul2_rn(a, make_float2(-0.5f, -0.5f));
  const float2 s = __ffma2_rn(nha, pe, ha);  // 0.5 a E(z)
  return __ffma2_rn(v, make_float2(0.5f, 0.5f), s);
}

// CONV = false: BEATs epilogue (bias, GELU, raw store, fp32 residual * alpha).  CONV = true: folded-BatchNorm epilogue of
// a 1x1 convolution in NHWC (raw pre-BN store, per-channel scale + shift, SiLU, bf16 residual); K and N need only be
// multiples of 8: the TMA maps zero-fill the out-of-range part of the last K block / N tile.
template <bool CONV>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle pattern shared by TMA and the UMMA descriptors
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space (LDS/STS, not generic LD/ST)
  unsigned char* stage_base = smem;
  float* stg_base = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (g.M + BM - 1) / BM, n_blocks = (g.N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks, num_kb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = stage_base + stage * STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          ptx::tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);
          ptx::tma_load_2d(sa + A_BYTES, &map_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(stage_base + stage * STAGE_BYTES);
          const uint64_t da = ptx::make_sw128_desc(sa), db = ptx::make_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
            ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[stage]);                   // smem slot free once these MMAs retire
          if (kb == num_kb - 1) ptx::umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter; each owns four of the tile's eight 32-column chunks.  The fp32 residual of
    // the NEXT chunk is prefetched into registers (8 independent 16-byte loads per lane) before the current chunk
    // is processed, so HBM latency is overlapped instead of serialised behind the accumulator read.
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    float* stg = stg_base + ew * (32 * 32);
    const int rsub = lane >> 3, csub = lane & 7;  // row-in-group-of-4, 16-byte column slot inside a 128-byte row segment
    const bool has_res = !CONV && g.residual != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    float4 res_next[8];
    auto load_res = [&](int tile, int c, float4 (&dst)[8]) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      const int cc = n_blk * BN + c * 32 + csub * 4;
      const int row0 = m_blk * BM + quarter * 32;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int grow = row0 + it * 4 + rsub;
        dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res && tile < num_tiles && grow < g.M && cc < g.N)
          dst[it] = __ldg(reinterpret_cast<const float4*>(g.residual + (size_t)grow * g.N + cc));
      }
    };
    if (has_res) load_res(blockIdx.x, half * 4, res_next);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      const int row0 = m_blk * BM + quarter * 32;
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int c = half * 4 + ci;
        const int col0 = n_blk * BN + c * 32;
        float4 res_cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) res_cur[it] = res_next[it];
        if (has_res) {
          if (ci < 3) load_res(tile, c + 1, res_next);
          else load_res(tile + gridDim.x, half * 4, res_next);
        }
        if (col0 < g.N) {  // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_addr + c * 32, r);
          ptx::tmem_ld_wait();
          // lane owns one row: park its 32 columns (XOR-swizzled 16-byte slots), re-read so 8 lanes cover a 128 B row segment
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          __syncwarp();
          const int cc = col0 + csub * 4;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), scale4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (g.bias != nullptr && cc < g.N) bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + cc));
          if (CONV && g.scale != nullptr && cc < g.N) scale4 = __ldg(reinterpret_cast<const float4*>(g.scale + cc));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            const int grow = row0 + rr;
            if (grow < g.M && cc < g.N) {
              float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((csub ^ (rr & 7)) << 2));
              const size_t off = (size_t)grow * g.N + cc;
              if (CONV) {
                if (g.raw_out != nullptr) *reinterpret_cast<float4*>(g.raw_out + off) = v;  // pre-BN conv output (hook)
                v.x = fmaf(v.x, scale4.x, bias4.x); v.y = fmaf(v.y, scale4.y, bias4.y);
                v.z = fmaf(v.z, scale4.z, bias4.z); v.w = fmaf(v.w, scale4.w, bias4.w);
                if (g.silu) {
                  v.x *= rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v.x));
                  v.y *= rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v.y));
                  v.z *= rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v.z));
                  v.w *= rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v.w));
                }
                if (g.res_bf16 != nullptr) {
                  const uint2 rb = __ldg(reinterpret_cast<const uint2*>(g.res_bf16 + off));
                  const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rb.x));
                  const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rb.y));
                  v.x += r0.x; v.y += r0.y; v.z += r1.x; v.w += r1.y;
                }
              } else {
                v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                if (g.gelu) {
                  const float2 g0 = gelu_fast2(make_float2(v.x, v.y)), g1 = gelu_fast2(make_float2(v.z, v.w));
                  v = make_float4(g0.x, g0.y, g1.x, g1.y);
                }
                if (g.raw_out != nullptr) *reinterpret_cast<float4*>(g.raw_out + off) = v;
                if (has_res) {
                  v.x = fmaf(g.res_scale, res_cur[it].x, v.x); v.y = fmaf(g.res_scale, res_cur[it].y, v.y);
                  v.z = fmaf(g.res_scale, res_cur[it].z, v.z); v.w = fmaf(g.res_scale, res_cur[it].w, v.w);
                }
              }
              if (g.out != nullptr) {
                const size_t oo = (size_t)grow * g.ldo + cc;
                if (g.out_bf16) {
                  uint2 pk = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + oo) = pk;
                } else {
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + oo) = v;
                }
              }
            }
          }
          __syncwarp();
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int gemm_bf16_launch(const CUtensorMap& map_a, const CUtensorMap& map_b, int M, int N, int K, const float* bias, int gelu,
                     float* raw_out, const float* residual, float res_scale, void* out, long long ldo, int out_bf16,
                     cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    AVEXK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  GemmArgs g{M, N, K, bias, gelu, raw_out, residual, res_scale, out, ldo, out_bf16, nullptr, 0, nullptr};
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  prof_begin(st, KID_GEMM, 2.0 * M * N * K);
  gemm_bf16_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(map_a, map_b, g);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int gemm_make_maps(CUtensorMap* map_a, CUtensorMap* map_b, const void* A, long long lda, const void* W, long long ldw, int M,
                   int N, int K);

// 1x1 convolution in NHWC: out[M, N] = act((A[M, K] @ W[N, K]^T) * scale + shift) (+ res); K, N multiples of 8.
int conv1x1_launch(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                   const __nv_bfloat16* res, float* raw_out, void* out, int out_bf16, cudaStream_t st) {
  AVEXK_CHECK_ARG(K % 8 == 0 && N % 8 == 0 && K > 0 && N > 0, "conv1x1: channel counts must be multiples of 8 (K=%d N=%d)", K, N);
  if (M == 0) return AVEXK_OK;
  static bool attr_set = false;
  if (!attr_set) {
    AVEXK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ma, mb;
  int rc = gemm_make_maps(&ma, &mb, A, K, W, K, M, N, K);
  if (rc) return rc;
  GemmArgs g{M, N, K, shift, 0, raw_out, nullptr, 0.f, out, N, out_bf16, scale, silu, res};
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  prof_begin(st, KID_GEMM, 2.0 * M * N * K);
  gemm_bf16_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(ma, mb, g);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int gemm_make_maps(CUtensorMap* map_a, CUtensorMap* map_b, const void* A, long long lda, const void* W, long long ldw, int M,
                   int N, int K) {
  int rc = make_tmap_2d_bf16(map_a, A, M, K, lda, BM, BK);
  if (rc) return rc;
  return make_tmap_2d_bf16(map_b, W, N, K, ldw, BN, BK);
}

}  // namespace avexk

extern "C" int avexk_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                               const float* bias, int gelu, float* raw_out, const float* residual, float res_scale, void* out,
                               long long ldo, int out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && (out || raw_out), "avexk_gemm_bf16: null operand");
  AVEXK_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, "avexk_gemm_bf16: unsupported shape M=%d N=%d K=%d", M, N, K);
  AVEXK_CHECK_ARG(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0 && (out == nullptr || (ldo >= N && ldo % 8 == 0)),
                  "avexk_gemm_bf16: bad leading dimensions");
  if (M == 0) return AVEXK_OK;
  CUtensorMap ma, mb;
  int rc = gemm_make_maps(&ma, &mb, A, lda, W, ldw, M, N, K);
  if (rc) return rc;
  return gemm_bf16_launch(ma, mb, M, N, K, bias, gelu, raw_out, residual, res_scale, out, ldo, out_bf16,
                          reinterpret_cast<cudaStream_t>(stream));
}
