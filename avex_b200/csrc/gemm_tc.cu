// gemm_tc.cu -- persistent, warp-specialised bf16 GEMM on the sm_100a tensor cores.
//
//   D[M,N] = epilogue(A[M,K] @ W[N,K]^T)      (A, W bf16 K-major; fp32 accumulation in TMEM)
//
// Replaces every nn.Linear on the BEATs path (backbone.py:531-533 q/k/v, :572 out_proj, :365-370 fc1/fc2,
// beats.py:350 patch-embed as im2col GEMM, :359 post_extract_proj) and the elementwise ops that follow them
// (bias, exact GELU, DeepNorm residual `residual * alpha + x`, backbone.py:360,:372, and the post-LN LayerNorm of
// backbone.py:362,:373), fused in the epilogue.  Also the 1x1 convolutions of EfficientNet (MODE_CONV).
//
// Structure (one CTA per SM; PAIR = two CTAs of a cluster drive one 256x256 tile with cta_group::2):
//   warp 0      TMA producer : cp.async.bulk.tensor 128x64 (A) + 256x64 (W; 128x64 per CTA of a pair) bf16 tiles,
//                              128B swizzle, mbarrier ring
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma 128x256x16 (cta_group::1) or 256x256x16
//                              (cta_group::2, leader CTA only; each CTA holds half of W's rows, so the shared-memory
//                              operand traffic per SM drops from 96 to 64 bytes/clk), commits to mbarriers
//   warps 2-3   idle register donors (setmaxnreg works per warpgroup)
//   warps 4..   epilogue, two flavours:
//     bias / GELU / conv (16 warps): tcgen05.ld 32x32b.x32 (next chunk in flight) -> per-warp smem transposition ->
//                              coalesced 16-byte global stores
//     residual / LayerNorm (8 warps, MODE_RES): everything stays in the TMEM layout (lane = output row).  Each warp
//                              streams its 32x32 fp32 residual boxes in by TMA (3 deep), adds bias + alpha * residual, and
//                              either TMA-stores the result, or -- fused LayerNorm -- accumulates the row's (sum, sum of
//                              squares) in registers, parks the pre-LN values back in TMEM (tcgen05.st), swaps the row
//                              statistics with the two CTAs holding the other column tiles of the same rows through L2,
//                              re-reads TMEM, normalises and writes fp32 (TMA store) + bf16: the pre-LN [M,768] tensor
//                              never exists in memory and the separate LayerNorm launch disappears.
//   TMEM: 512 columns = two 128x256 fp32 accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Roofline: dense bf16 tensor pipe; algorithmic FLOPs = 2*M*N*K.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int CTRL_WARPS = 4;
constexpr int STG_BYTES_PER_WARP = 32 * 32 * 4;  // 32 rows x 32 fp32 (128-byte rows), XOR-swizzled
constexpr int LN_C = 768, LN_NB = LN_C / BN;     // the fused LayerNorm epilogue is built for rows of 3 tiles

enum { MODE_PLAIN = 0, MODE_GELU = 1, MODE_RES = 2, MODE_CONV = 3 };

// DEEPK (MODE_RES with a long K loop, fc2): the tile's MMA time hides the epilogue, so shared memory goes to a fourth
// operand stage instead of a third residual box per warp.
template <bool PAIR, int EW, int MODE, bool DEEPK>
struct Cfg {
  static constexpr int EPI_WARPS = EW, EPI_THREADS = EW * 32, NTHREADS = 32 * (CTRL_WARPS + EW);
  static constexpr int CHUNKS = 32 / EW;  // 32-column chunks of the 256-column tile per epilogue warp
  // setmaxnreg moves registers inside the CTA's launch allocation (threads x compiled count), it cannot grow it:
  // EW = 8: 384 x 168 = 64512 >= 128*64 + 256*216 = 63488;  EW = 16: 640 x 96 = 61440 >= 128*56 + 512*104 = 60416
  static constexpr int CTRL_REGS = EW == 8 ? 64 : 56, EPI_REGS = EW == 8 ? 216 : 104;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;  // rows of W each CTA stages per k-block
  // a pipeline stage holds KSUB k-blocks of 64 (one 128-byte swizzle atom wide each): fewer barrier round trips per MMA
  // (measured: qkv 1262 -> 1276, fc1+GELU 1216 -> 1291 TFLOP/s; with only two such stages fc2 lost 7 %, so MODE_RES keeps KSUB = 1)
  static constexpr int KSUB = (MODE == MODE_PLAIN || MODE == MODE_GELU) ? 2 : 1;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = B_ROWS * BK * 2, SUB_BYTES = A_BYTES + B_BYTES, STAGE_BYTES = KSUB * SUB_BYTES;
  // MODE_RES spends 96 KB of shared memory on the residual ring, so it keeps fewer operand stages
  // residual boxes in flight per epilogue warp (MODE_RES).  DEEPK: a tile's 48 k-blocks hide the whole epilogue, so the
  // residual is fetched box by box without prefetch and the ring's shared memory goes to a FIFTH operand stage (fc2+LN:
  // 3 stages 0.585 ms, 4 stages 0.549 ms -- the DRAM-sourced A rows need the depth)
  static constexpr int RES_DEPTH = (PAIR && DEEPK) ? 1 : 3;
  // per-warp output staging: 32 rows x 64 bytes for the bias / GELU epilogue, 32 x 128 bytes for the others
  static constexpr int STG_BYTES = (MODE == MODE_PLAIN || MODE == MODE_GELU) ? STG_BYTES_PER_WARP / 2 : STG_BYTES_PER_WARP;
  static constexpr int STAGES = MODE == MODE_RES ? (PAIR ? (DEEPK ? 5 : 3) : 2) : MODE == MODE_CONV ? (PAIR ? 5 : 3) : (PAIR ? 3 : 2);
  static constexpr int TX_SUB = SUB_BYTES * (PAIR ? 2 : 1);  // bytes landing on the (leader's) full barrier per k-block
  static constexpr int UM = PAIR ? 2 * BM : BM;                  // output rows per scheduling unit
  static constexpr int RES_BYTES = MODE == MODE_RES ? EW * RES_DEPTH * STG_BYTES_PER_WARP : 0;
  static constexpr int OFF_STG = STAGES * STAGE_BYTES, OFF_RES = OFF_STG + EW * STG_BYTES, OFF_BAR = OFF_RES + RES_BYTES;
  static constexpr int SMEM_BYTES = 1024 + OFF_BAR + 512;
  static_assert(SMEM_BYTES <= 232448, "gemm: shared memory budget");
};

struct GemmArgs {
  int M, N, K;
  const float* bias;
  float* raw_out;
  const float* residual;
  float res_scale;
  void* out;
  long long ldo;
  int out_bf16;
  // MODE_CONV (EfficientNet 1x1 convolutions): y = act(acc * scale[n] + bias[n]) (+ res_bf16), raw_out = acc
  const float* scale;
  int silu;
  const __nv_bfloat16* res_bf16;
  // fused LayerNorm (MODE_RES, N == 768)
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  float* ln_out_f32;  // written through map_y
  __nv_bfloat16* ln_out_bf16;
  float2* ln_stats;  // [ceil(M/256)*256][2*LN_NB] (sum, sum of squares) over 128 columns of a row
  int* ln_count;     // [2][2*ceil(M/256)]: arrivals / departures per 128-row block; zero before and after every launch
  // column sums over the token rows of every clip (mean-pooled embeddings without a second pass over [M,768]):
  // 40.24 fixed-point accumulators [M / pool_rows][768], so the atomics are order-independent and the result reproducible
  // fused QKV projection: the gate of BEATs' relative-position bias from the q part of the output (columns < gate_heads * 64),
  // gate[m, h] = sig_a (sig_b grep_a[h] - 1) + 2, (sig_a, sig_b) = sigmoid(q_h . gate_w[0|1] + gate_b[0|1]) (backbone.py:544-550),
  // formed on the fp32 accumulators (+ bias) before they are rounded to bf16
  const float* gate_w;   // [2, 64]
  const float* gate_b;   // [2]
  const float* grep_a;   // [heads]
  float* gate_out;       // [M, heads] or null
  int gate_heads;
  long long* pool_raw;  // of v = A @ W^T + bias (the tensor a forward hook on the Linear sees), or null
  long long* pool_y;    // of the LayerNorm output, or null
  int pool_rows;        // token rows per clip (>= 32)
};

constexpr float POOL_FIX = 16777216.0f;  // 2^24

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

// Exact (erf) GELU (modules.py:191-200) for two values per call with packed fp32x2 arithmetic and ONE special-function
// op per value.  With a = |v|:  gelu(v) = v Phi(v) = relu(v) - a * (erfc(a / sqrt2) / 2), and log2(erfc(a / sqrt2) / 2)
// is smooth, so a degree-5 polynomial q (least-squares fit of a 2^q(a) on [0, 6], leading coefficient negative so the
// tail decays) reproduces GELU to |err| <= 6.5e-7 over the whole real line -- three decimal orders inside the bf16
// rounding of the stored activation.  Evaluated in n = -a so the final product needs no negation.
__device__ __forceinline__ float2 gelu_fast2(float2 v) {
  const float2 n = make_float2(-fabsf(v.x), -fabsf(v.y));
  const float2 r = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  float2 q = __ffma2_rn(make_float2(4.712853462e-4f, 4.712853462e-4f), n, make_float2(7.064215249e-3f, 7.064215249e-3f));
  q = __ffma2_rn(q, n, make_float2(5.175841926e-2f, 5.175841926e-2f));
  q = __ffma2_rn(q, n, make_float2(-4.600918231e-1f, -4.600918231e-1f));
  q = __ffma2_rn(q, n, make_float2(1.150727814f, 1.150727814f));
  q = __ffma2_rn(q, n, make_float2(-1.000049354f, -1.000049354f));
  const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));
  return __ffma2_rn(n, e, r);
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Column sums of a 32 x 32 fp32 box staged in shared memory (128-byte rows, 16-byte chunk j of row r at chunk j ^ (r & 7)):
// lane c reads column c of every row (one row = 32 banks: conflict-free) and adds the sum over the rows of each clip to that
// clip's accumulator.  row0: first global row of the box; rows >= M are excluded; a box straddles at most one clip boundary
// (pool_rows >= 32).
__device__ __forceinline__ void pool_box(uint32_t box, int lane, int row0, int M, int pool_rows, long long* acc, int col) {
  const int b0 = row0 / pool_rows;
  const int lim = (b0 + 1) * pool_rows - row0;        // rows of clip b0 in this box (warp-uniform)
  const int nval = M - row0 < 32 ? M - row0 : 32;     // valid rows (warp-uniform)
  const uint32_t base = box + (lane & 3) * 4;
  const int cj = lane >> 2;
  float sa = 0.f, sb = 0.f;
  if (nval == 32 && lim >= 32) {
#pragma unroll
    for (int r = 0; r < 32; ++r) sa += lds32f(base + r * 128 + ((cj ^ (r & 7)) << 4));
  } else {
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float x = lds32f(base + r * 128 + ((cj ^ (r & 7)) << 4));
      if (r < nval) {
        if (r < lim) sa += x;
        else sb += x;
      }
    }
  }
  if (nval > 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc + (size_t)b0 * LN_C + col + lane),
                          static_cast<unsigned long long>(__float2ll_rn(sa * POOL_FIX)));
  if (nval > lim) atomicAdd(reinterpret_cast<unsigned long long*>(acc + (size_t)(b0 + 1) * LN_C + col + lane),
                            static_cast<unsigned long long>(__float2ll_rn(sb * POOL_FIX)));
}

template <int MODE, bool PAIR, int EW, bool DEEPK>
__global__ void __launch_bounds__(32 * (CTRL_WARPS + EW), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_y,
                 const __grid_constant__ CUtensorMap map_xb, const GemmArgs g) {
  using C = Cfg<PAIR, EW, MODE, DEEPK>;
  constexpr int RES_DEPTH = C::RES_DEPTH;
  constexpr int STAGES = C::STAGES, EPI_WARPS = C::EPI_WARPS, EPI_THREADS = C::EPI_THREADS, CHUNKS = C::CHUNKS;
  static_assert(MODE != MODE_RES || EW == 8, "the residual / LayerNorm epilogue is written for 8 warps");
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle pattern shared by TMA and the UMMA descriptors.  The offset is the
  // same in both CTAs of a pair (same kernel, same dynamic-smem base), which the pair MMA and multicast commits rely on.
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stage_base = smem;
  float* stg_base = reinterpret_cast<float*>(smem + C::OFF_STG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full_bar = bars;                      // [STAGES]  (pair: only the leader's are waited on)
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // [2]       (pair: the leader's collect both CTAs' epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* res_bar = bars + 32;                  // [EPI_WARPS][RES_DEPTH] residual boxes landed (MODE_RES)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_units = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_units = (g.M + C::UM - 1) / C::UM, n_blocks = (g.N + BN - 1) / BN;
  const int num_tiles = m_units * n_blocks, num_kb = (g.K + BK - 1) / BK;
  // iteration -> tile of this scheduling unit: tiles round-robin over units, n fastest, so the units working on the n tiles
  // of one row block run side by side (they share the A rows in L2 and, with the fused LayerNorm, exchange row statistics)
  auto tile_of = [&](int it, int& mu, int& nb) -> bool {
    const int t = unit + it * num_units;
    mu = t / n_blocks;
    nb = t % n_blocks;
    return t < num_tiles;
  };

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    if (MODE == MODE_RES) ptx::prefetch_tensormap(&map_res);
    if (MODE != MODE_CONV) ptx::prefetch_tensormap(&map_y);
    if (MODE == MODE_RES && g.ln_out_bf16 != nullptr) ptx::prefetch_tensormap(&map_xb);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], EPI_WARPS * (PAIR ? 2 : 1));  // one arrive per epilogue warp (of both CTAs)
    }
    if (MODE == MODE_RES)
      for (int i = 0; i < EPI_WARPS * RES_DEPTH; ++i) ptx::mbar_init(&res_bar[i], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) {
    if (PAIR) {
      ptx::tmem_alloc_pair(tmem_slot, 512);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < CTRL_WARPS) {
    ptx::setmaxnreg_dec<C::CTRL_REGS>();
    if (warp == 0) {
      // ===================== TMA producer (every CTA loads its own A rows and its share of W) =====================
      if (lane == 0) {
        const uint32_t full0 = PAIR ? ptx::mapa(ptx::smem_u32(&full_bar[0]), 0) : 0u;  // the leader's full barriers
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0;; ++it) {
          int mu, nb;
          if (!tile_of(it, mu, nb)) break;
          const int m0 = mu * C::UM + static_cast<int>(rank) * BM;
          const int n0 = nb * BN + static_cast<int>(rank) * C::B_ROWS;
          for (int kb = 0; kb < num_kb; kb += C::KSUB) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            const int nsub = num_kb - kb < C::KSUB ? num_kb - kb : C::KSUB;
            if (!PAIR || rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], nsub * C::TX_SUB);
#pragma unroll
            for (int sb = 0; sb < C::KSUB; ++sb) {
              if (sb >= nsub) break;
              unsigned char* sa = stage_base + stage * C::STAGE_BYTES + sb * C::SUB_BYTES;
              if (PAIR) {
                ptx::tma_load_2d_pair(sa, &map_a, full0 + stage * 8, (kb + sb) * BK, m0);
                ptx::tma_load_2d_pair(sa + C::A_BYTES, &map_b, full0 + stage * 8, (kb + sb) * BK, n0);
              } else {
                ptx::tma_load_2d(sa, &map_a, &full_bar[stage], (kb + sb) * BK, m0);
                ptx::tma_load_2d(sa + C::A_BYTES, &map_b, &full_bar[stage], (kb + sb) * BK, n0);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        // tail: every fill has been consumed and its release has landed here (no commit may target an exited CTA)
        for (int s = 0; s < STAGES; ++s) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && (!PAIR || rank == 0)) {
      // ===================== MMA issuer (leader CTA of a pair) =====================
      constexpr uint32_t idesc = MODE == MODE_CONV ? ptx::make_idesc_f16(C::UM, BN) : ptx::make_idesc_bf16(C::UM, BN);  // EfficientNet: fp16
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0;; ++it) {
        int mu, nb;
        if (!tile_of(it, mu, nb)) break;
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; kb += C::KSUB) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (lane == 0) {
            const int nsub = num_kb - kb < C::KSUB ? num_kb - kb : C::KSUB;
#pragma unroll
            for (int sb = 0; sb < C::KSUB; ++sb) {
              if (sb >= nsub) break;
              const uint32_t sa = ptx::smem_u32(stage_base + stage * C::STAGE_BYTES + sb * C::SUB_BYTES);
              const uint64_t da = ptx::make_sw128_desc(sa), db = ptx::make_sw128_desc(sa + C::A_BYTES);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {  // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
                if (PAIR) ptx::umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | sb | k) != 0 ? 1u : 0u);
                else ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | sb | k) != 0 ? 1u : 0u);
              }
            }
            const bool last = kb + C::KSUB >= num_kb;
            if (PAIR) {
              ptx::umma_commit_pair(&empty_bar[stage], 3);              // both CTAs' smem slots are free
              if (last) ptx::umma_commit_pair(&tfull_bar[acc], 3);      // both halves of the accumulator complete
            } else {
              ptx::umma_commit(&empty_bar[stage]);
              if (last) ptx::umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if constexpr (MODE == MODE_RES) {
    // ===================== residual / LayerNorm epilogue (warps 4..11): lane = output row =====================
    // Warp (quarter q, half h) owns rows 32q..32q+31 and the four 32-column chunks 4h..4h+3 of every tile.  Its work is a
    // flat sequence of 32x32 fp32 boxes (4 per tile); the residuals of the next RES_DEPTH-1 boxes are in flight (TMA) while box n is processed.
    ptx::setmaxnreg_inc<C::EPI_REGS>();
    const int ew = warp - CTRL_WARPS, quarter = warp & 3, half = ew >> 2;
    const bool ln = g.ln_gamma != nullptr;
    const bool has_res = g.residual != nullptr;
    const uint32_t ybuf = ptx::smem_u32(stg_base) + ew * STG_BYTES_PER_WARP;                       // staging of one output box
    const uint32_t rbuf = ptx::smem_u32(smem + C::OFF_RES) + ew * (RES_DEPTH * STG_BYTES_PER_WARP);  // residual ring
    uint64_t* rbar = res_bar + ew * RES_DEPTH;
    const uint32_t tempty0 = PAIR ? ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    const uint32_t rowoff = lane * 128, sw = lane & 7;  // 128-byte rows; 16-byte chunk j of row r is stored at chunk j ^ (r & 7)
    const int nrows_blk = 2 * ((g.M + 2 * BM - 1) / (2 * BM));  // counters per 128-row block, pair-padded (host: ln_scratch_layout)
    int acc = 0;
    uint32_t acc_phase = 0;

    // residual box n of this warp (tile n / 4, chunk 4h + n % 4) -> ring slot n % RES_DEPTH.  The caller has made sure
    // (__syncwarp) that every lane is done reading the slot being overwritten.
    auto issue_res = [&](int n) {
      int mu, nb;
      if (!has_res || lane != 0 || !tile_of(n >> 2, mu, nb)) return;
      const int s = n % RES_DEPTH;
      // the slot of a tile's last box doubles as the bf16 staging of its LayerNorm pass: that store must have left it
      if (ln && (n & 3) == RES_DEPTH - 1) ptx::tma_store_wait_read();
      ptx::fence_proxy_async();  // the generic-proxy reads of this slot are ordered before the async-proxy overwrite
      ptx::mbar_arrive_expect_tx(&rbar[s], STG_BYTES_PER_WARP);
      ptx::tma_load_2d_s(rbuf + s * STG_BYTES_PER_WARP, &map_res, &rbar[s], nb * BN + (half * 4 + (n & 3)) * 32,
                         mu * C::UM + static_cast<int>(rank) * BM + quarter * 32);
    };
    // stage a finished 32x32 fp32 box (this lane's row in v) and hand it to the TMA engine
    auto store_box = [&](const float (&v)[32], int col, int row) {
      if (lane == 0) ptx::tma_store_wait_read();  // the previous box has left the staging buffer
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128f(ybuf + rowoff + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      ptx::fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d_s(&map_y, ybuf, col, row);  // rows >= M / columns >= N are clipped by the tensor map
        ptx::tma_store_commit();
      }
    };

#pragma unroll
    for (int n = 0; n < RES_DEPTH - 1; ++n) issue_res(n);
    for (int it = 0;; ++it) {
      int mu, nb;
      if (!tile_of(it, mu, nb)) break;
      const int m0 = mu * C::UM + static_cast<int>(rank) * BM, row0 = m0 + quarter * 32, grow = row0 + lane;
      const bool row_ok = grow < g.M;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * 128;
      float sum = 0.f, sq = 0.f;
      // ---- pass 1: v = acc + bias + alpha * residual ------------------------------------------------------------
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int n = it * 4 + ci, s = n % RES_DEPTH;
        const int col0 = nb * BN + (half * 4 + ci) * 32;
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_addr + ci * 32, r);
        issue_res(n + RES_DEPTH - 1);  // its slot held box n - 1, released by the __syncwarp that ended the previous iteration
        if (has_res) ptx::mbar_wait(&rbar[s], (n / RES_DEPTH) & 1);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + j);  // warp-uniform address
          v[4 * j] = __uint_as_float(r[4 * j]) + b4.x;
          v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
          v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
        }
        if (g.raw_out != nullptr && row_ok) {  // hook on the Linear (fc2): rare, so plain per-row stores
          float4* dst = reinterpret_cast<float4*>(g.raw_out + (size_t)grow * g.N + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (ln && g.pool_raw != nullptr) {  // mean-pooled hook: column sums of the raw Linear output, via the (idle) output staging
          if (lane == 0) ptx::tma_store_wait_read();  // pass 2 of the previous tile may still be reading the staging buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) sts128f(ybuf + rowoff + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          pool_box(ybuf, lane, row0, g.M, g.pool_rows, g.pool_raw, col0);
          __syncwarp();
        }
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = lds128f(rbuf + s * STG_BYTES_PER_WARP + rowoff + ((j ^ sw) << 4));
            v[4 * j] = fmaf(g.res_scale, q.x, v[4 * j]);
            v[4 * j + 1] = fmaf(g.res_scale, q.y, v[4 * j + 1]);
            v[4 * j + 2] = fmaf(g.res_scale, q.z, v[4 * j + 2]);
            v[4 * j + 3] = fmaf(g.res_scale, q.w, v[4 * j + 3]);
          }
        }
        if (ln) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sum += v[j];
            sq = fmaf(v[j], v[j], sq);
            r[j] = __float_as_uint(v[j]);
          }
          ptx::tmem_st_32x32(t_addr + ci * 32, r);  // park the pre-LN values where they came from
        } else if (g.out != nullptr) {
          store_box(v, col0, row0);
        }
        __syncwarp();  // every lane is done with residual slot s
      }
      if (!ln) {
        // all TMEM reads of this accumulator have completed (wait::ld in the last iteration)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
          else ptx::mbar_arrive(&tempty_bar[acc]);
        }
      } else {
        // ---- fused LayerNorm (backbone.py:362,:373): the LN_NB CTAs holding the column tiles of these 128 rows swap
        // per-row (sum, sum of squares) through L2.  Requires all CTAs of the grid to be co-resident (grid <= #SMs, one
        // CTA per SM): a CTA publishes its own statistics before it waits for its neighbours'.
        const int mb = m0 / BM;
        if (row_ok) g.ln_stats[(size_t)grow * (2 * LN_NB) + nb * 2 + half] = make_float2(sum, sq);
        ptx::tmem_st_wait();
        // CTA barrier + release at gpu scope by the announcing thread: the other warps' statistics are ordered before the
        // arrival by cumulativity (no per-thread __threadfence())
        ptx::named_bar_sync(1, EPI_THREADS);
        if (ew == 0 && lane == 0) {
          red_release_gpu_add(&g.ln_count[mb], 1);
          uint32_t spins = 0;
          while (ld_acquire_gpu(&g.ln_count[mb]) < LN_NB) {
            __nanosleep(32);
            if (++spins > (1u << 24)) __trap();  // a missing neighbour fails the launch instead of hanging the GPU
          }
        }
        ptx::named_bar_sync(1, EPI_THREADS);
        float mean = 0.f, rstd = 0.f;
        {
          float s1 = 0.f, s2 = 0.f;
          if (row_ok) {
            const float4* sp = reinterpret_cast<const float4*>(g.ln_stats + (size_t)grow * (2 * LN_NB));
#pragma unroll
            for (int i = 0; i < LN_NB; ++i) {
              const float4 p = __ldcg(sp + i);  // two (sum, sq) pairs; written by other SMs: read at L2
              s1 += p.x + p.z;
              s2 += p.y + p.w;
            }
          }
          mean = s1 * (1.0f / LN_C);
          rstd = rsqrtf(fmaxf(s2 * (1.0f / LN_C) - mean * mean, 0.f) + g.ln_eps);
        }
        ptx::named_bar_sync(1, EPI_THREADS);  // every statistic of this row block has been read
        if (ew == 0 && lane == 0) {
          // the last of the LN_NB CTAs to get here re-arms the counters for the next launch
          if (atomicAdd(&g.ln_count[nrows_blk + mb], 1) == LN_NB - 1) {
            g.ln_count[mb] = 0;
            g.ln_count[nrows_blk + mb] = 0;
          }
        }
        // ---- pass 2: y = (v - mean) * rstd * gamma + beta ----------------------------------------------------
        const uint32_t xbuf = rbuf + ((it * 4 + 3) % RES_DEPTH) * STG_BYTES_PER_WARP;  // free until the next tile's first box
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int col0 = nb * BN + (half * 4 + ci) * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_addr + ci * 32, r);
          ptx::tmem_ld_wait();
          if (ci == 3) {  // the accumulator is drained: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
              else ptx::mbar_arrive(&tempty_bar[acc]);
            }
          }
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 ga = __ldg(reinterpret_cast<const float4*>(g.ln_gamma + col0) + j);  // warp-uniform addresses
            const float4 be = __ldg(reinterpret_cast<const float4*>(g.ln_beta + col0) + j);
            y[4 * j] = fmaf((__uint_as_float(r[4 * j]) - mean) * rstd, ga.x, be.x);
            y[4 * j + 1] = fmaf((__uint_as_float(r[4 * j + 1]) - mean) * rstd, ga.y, be.y);
            y[4 * j + 2] = fmaf((__uint_as_float(r[4 * j + 2]) - mean) * rstd, ga.z, be.z);
            y[4 * j + 3] = fmaf((__uint_as_float(r[4 * j + 3]) - mean) * rstd, ga.w, be.w);
          }
          // Both outputs leave through TMA: the fp32 box from the warp's staging buffer, the bf16 box (32 rows x 64 bytes,
          // 64-byte swizzle) from the residual-ring slot that the tile's last box has just vacated.  (Per-thread 16-byte
          // stores of the bf16 rows -- 32 scattered lines per instruction -- cost 0.09 ms of the 0.31 ms out_proj launch.)
          if (lane == 0) ptx::tma_store_wait_read();  // the previous boxes have left both staging buffers
          __syncwarp();
          if (g.ln_out_bf16 != nullptr) {
            const uint32_t sw64 = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128u(xbuf + lane * 64 + ((j ^ sw64) << 4), pack_bf16(y[8 * j], y[8 * j + 1]), pack_bf16(y[8 * j + 2], y[8 * j + 3]),
                      pack_bf16(y[8 * j + 4], y[8 * j + 5]), pack_bf16(y[8 * j + 6], y[8 * j + 7]));
          }
          if (g.ln_out_f32 != nullptr || g.pool_y != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128f(ybuf + rowoff + ((j ^ sw) << 4), y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
          }
          ptx::fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
          __syncwarp();
          if (g.pool_y != nullptr) pool_box(ybuf, lane, row0, g.M, g.pool_rows, g.pool_y, col0);  // mean-pool fused: no pass over y
          if (lane == 0) {  // rows >= M are clipped by the tensor maps
            if (g.ln_out_bf16 != nullptr) ptx::tma_store_2d_s(&map_xb, xbuf, col0, row0);
            if (g.ln_out_f32 != nullptr) ptx::tma_store_2d_s(&map_y, ybuf, col0, row0);
            ptx::tma_store_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) ptx::tma_store_wait_all();  // the bulk stores must have been performed before the CTA exits
  } else if constexpr (MODE != MODE_CONV) {
    // ===================== bias / GELU epilogue (warps 4..19): lane = output row =====================
    // Warp (quarter q, group h) owns rows 32q..32q+31 and the 64 columns 64h..64h+63 of every tile.  Both 32-column TMEM
    // chunks are read up front, so the accumulator goes back to the MMA warp before any arithmetic; the finished 32x64
    // bf16 box (or two 32x32 fp32 boxes) is staged in 128-byte-swizzled shared memory and written by TMA: no transposition,
    // no per-element address arithmetic, full 128-byte lines on the way out.
    ptx::setmaxnreg_inc<C::EPI_REGS>();
    static_assert(CHUNKS == 2, "the bias / GELU epilogue is written for 16 warps");
    const int ew = warp - CTRL_WARPS, quarter = warp & 3, half = ew >> 2;
    const uint32_t ybuf = ptx::smem_u32(stg_base) + ew * C::STG_BYTES;  // 32 rows x 64 bytes, 64-byte swizzle
    const uint32_t tempty0 = PAIR ? ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    const uint32_t rowoff = lane * 64, sw = (lane >> 1) & 3;  // 16-byte chunk j of row r is stored at chunk j ^ ((r >> 1) & 3)
    int acc = 0;
    uint32_t acc_phase = 0;
    // hand a staged 32-row x 64-byte box to the TMA engine (rows >= M / columns >= N are clipped by the tensor map)
    auto flush_box = [&](int col, int row) {
      ptx::fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d_s(&map_y, ybuf, col, row);
        ptx::tma_store_commit();
      }
    };
    auto claim_box = [&]() {  // the previous box has left the staging buffer
      if (lane == 0) ptx::tma_store_wait_read();
      __syncwarp();
    };
    for (int it = 0;; ++it) {
      int mu, nb;
      if (!tile_of(it, mu, nb)) break;
      const int row0 = mu * C::UM + static_cast<int>(rank) * BM + quarter * 32, grow = row0 + lane;
      const int colb = nb * BN + half * 64;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * 64;
      uint32_t r[2][32];
      ptx::tmem_ld_32x32(t_addr, r[0]);
      ptx::tmem_ld_32x32(t_addr + 32, r[1]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
        else ptx::mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // the warp's 64 columns are one attention head of q: its two gate logits are dot products over them (warp-uniform test)
      const bool do_gate = MODE == MODE_PLAIN && g.gate_out != nullptr && colb < g.gate_heads * 64;
      float2 za2 = make_float2(0.f, 0.f), zb2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = colb + c * 32;
        if (col0 >= g.N) continue;  // warp-uniform (N tail)
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias != nullptr && col0 + 4 * j < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + j);  // warp-uniform
          v[4 * j] = __uint_as_float(r[c][4 * j]) + b4.x;
          v[4 * j + 1] = __uint_as_float(r[c][4 * j + 1]) + b4.y;
          v[4 * j + 2] = __uint_as_float(r[c][4 * j + 2]) + b4.z;
          v[4 * j + 3] = __uint_as_float(r[c][4 * j + 3]) + b4.w;
        }
        if (do_gate) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 wa = __ldg(reinterpret_cast<const float4*>(g.gate_w + c * 32) + j);       // warp-uniform, L1-resident
            const float4 wb = __ldg(reinterpret_cast<const float4*>(g.gate_w + 64 + c * 32) + j);
            za2 = __ffma2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(wa.x, wa.y), za2);
            za2 = __ffma2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(wa.z, wa.w), za2);
            zb2 = __ffma2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(wb.x, wb.y), zb2);
            zb2 = __ffma2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(wb.z, wb.w), zb2);
          }
        }
        if (MODE == MODE_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 gl = gelu_fast2(make_float2(v[2 * j], v[2 * j + 1]));
            v[2 * j] = gl.x;
            v[2 * j + 1] = gl.y;
          }
        }
        if (g.raw_out != nullptr && grow < g.M) {  // hook on the Linear: rare, plain per-row stores
          float4* dst = reinterpret_cast<float4*>(g.raw_out + (size_t)grow * g.N + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col0 + 4 * j < g.N) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (g.out == nullptr) continue;
        if (g.out_bf16) {  // one 32 x 32 bf16 box
          claim_box();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = ybuf + rowoff + ((j ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(v[8 * j], v[8 * j + 1])),
                         "r"(pack_bf16(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16(v[8 * j + 4], v[8 * j + 5])),
                         "r"(pack_bf16(v[8 * j + 6], v[8 * j + 7]))
                         : "memory");
          }
          flush_box(col0, row0);
        } else {  // two 32 x 16 fp32 boxes
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (col0 + 16 * h2 >= g.N) continue;
            claim_box();
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128f(ybuf + rowoff + ((j ^ sw) << 4), v[16 * h2 + 4 * j], v[16 * h2 + 4 * j + 1], v[16 * h2 + 4 * j + 2], v[16 * h2 + 4 * j + 3]);
            flush_box(col0 + 16 * h2, row0);
          }
        }
      }
      if (do_gate && grow < g.M) {
        const int hh = colb >> 6;
        const float za = za2.x + za2.y + __ldg(g.gate_b), zb = zb2.x + zb2.y + __ldg(g.gate_b + 1);
        const float ga = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * za)), gb = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * zb));
        g.gate_out[(size_t)grow * g.gate_heads + hh] = ga * (gb * __ldg(g.grep_a + hh) - 1.0f) + 2.0f;
      }
    }
    if (lane == 0) ptx::tma_store_wait_all();  // the bulk stores must have been performed before the CTA exits
  } else {
    // ===================== conv epilogue (warps 4..19) =====================
    // Four warps per TMEM lane quarter; each owns CHUNKS of the tile's eight 32-column chunks.  The TMEM load of the next
    // chunk is in flight while the current one is transposed through shared memory and stored with 16-byte coalesced writes.
    ptx::setmaxnreg_inc<C::EPI_REGS>();
    const int ew = warp - CTRL_WARPS, quarter = warp & 3, half = ew >> 2;  // half: which group of CHUNKS chunks
    float* stg = stg_base + ew * (32 * 32);
    static_assert(MODE != MODE_CONV || C::STG_BYTES == STG_BYTES_PER_WARP, "conv epilogue staging");
    const int rsub = lane >> 3, csub = lane & 7;  // row-in-group-of-4, 16-byte column slot inside a 128-byte row segment
    const uint32_t tempty0 = PAIR ? ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0) : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;

    auto process = [&](const uint32_t (&r)[32], int mu, int nb, int c, const float4 bias4, const float4 scale4) {
      const int col0 = nb * BN + c * 32;
      if (col0 >= g.N) return;  // warp-uniform
      const int row0 = mu * C::UM + static_cast<int>(rank) * BM + quarter * 32;
      // lane owns one row: park its 32 columns (XOR-swizzled 16-byte slots), re-read so 8 lanes cover a 128 B row segment
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      __syncwarp();
      const int cc = col0 + csub * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + rsub;
        const int grow = row0 + rr;
        if (grow < g.M && cc < g.N) {
          float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((csub ^ (rr & 7)) << 2));
          const size_t off = (size_t)grow * g.N + cc;
          if (MODE == MODE_CONV) {
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
              const float2 r0 = unpack_h16(rb.x), r1 = unpack_h16(rb.y);  // the skip connection is stored as fp16
              v.x += r0.x; v.y += r0.y; v.z += r1.x; v.w += r1.y;
            }
          } else {
            v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
            if (MODE == MODE_GELU) {
              const float2 g0 = gelu_fast2(make_float2(v.x, v.y)), g1 = gelu_fast2(make_float2(v.z, v.w));
              v = make_float4(g0.x, g0.y, g1.x, g1.y);
            }
            if (g.raw_out != nullptr) *reinterpret_cast<float4*>(g.raw_out + off) = v;
          }
          if (g.out != nullptr) {
            const size_t oo = (size_t)grow * g.ldo + cc;
            if (g.out_bf16) {
              uint2 pk = MODE == MODE_CONV ? make_uint2(pack_h16(v.x, v.y), pack_h16(v.z, v.w)) : make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + oo) = pk;
            } else {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + oo) = v;
            }
          }
        }
      }
      __syncwarp();
    };

    float4 bias_t[CHUNKS], scale_t[CHUNKS];  // this lane's bias (and BN scale) columns, loaded ahead of the accumulator
    for (int it = 0;; ++it) {
      int mu, nb;
      if (!tile_of(it, mu, nb)) break;
#pragma unroll
      for (int ci = 0; ci < CHUNKS; ++ci) {
        const int cc = nb * BN + (half * CHUNKS + ci) * 32 + csub * 4;
        bias_t[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
        scale_t[ci] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (g.bias != nullptr && cc < g.N) bias_t[ci] = __ldg(reinterpret_cast<const float4*>(g.bias + cc));
        if (MODE == MODE_CONV && g.scale != nullptr && cc < g.N) scale_t[ci] = __ldg(reinterpret_cast<const float4*>(g.scale + cc));
      }
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (CHUNKS * 32);
      const int c0 = half * CHUNKS;
      uint32_t ra[32], rb[32];
      ptx::tmem_ld_32x32(t_addr, ra);
      if (CHUNKS == 4) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_addr + 32, rb);
        process(ra, mu, nb, c0, bias_t[0], scale_t[0]);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_addr + 64, ra);
        process(rb, mu, nb, c0 + 1, bias_t[1], scale_t[1]);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_addr + 96, rb);
        process(ra, mu, nb, c0 + 2, bias_t[CHUNKS - 2], scale_t[CHUNKS - 2]);
      } else {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_addr + 32, rb);
        process(ra, mu, nb, c0, bias_t[0], scale_t[0]);
      }
      // last chunk: every TMEM read of this accumulator has completed -> hand it back to the MMA warp first
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
        else ptx::mbar_arrive(&tempty_bar[acc]);
      }
      process(rb, mu, nb, c0 + CHUNKS - 1, bias_t[CHUNKS - 1], scale_t[CHUNKS - 1]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// CTA pairs (cta_group::2) by default; AVEXK_GEMM_PAIR=0 or avexk_gemm_config(0) selects the single-CTA kernel.
int g_pair = -1;
bool use_pair() {
  if (g_pair < 0) {
    const char* e = getenv("AVEXK_GEMM_PAIR");
    g_pair = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pair != 0;
}

template <int MODE, bool PAIR, bool DEEPK>
int launch_mode(const void* A, long long lda, const void* W, long long ldw, const GemmArgs& g, cudaStream_t st) {
  constexpr int EW = MODE == MODE_RES ? 8 : 16;
  using C = Cfg<PAIR, EW, MODE, DEEPK>;
  constexpr int NTHREADS = C::NTHREADS;
  static bool attr_set[64] = {};  // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    AVEXK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<MODE, PAIR, EW, DEEPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[dev_] = true;
  }
  CUtensorMap ma, mb;
  CUtensorMap mxb{};  // bf16 output of the fused LayerNorm epilogue (MODE_RES)
  int rc = make_tmap_2d_bf16(&ma, A, g.M, g.K, lda, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&mb, W, g.N, g.K, ldw, C::B_ROWS, BK);
  if (rc) return rc;
  CUtensorMap mres = ma, my = ma;  // mres: residual / LayerNorm epilogue; my: output of every epilogue but the conv one
  if ((MODE == MODE_PLAIN || MODE == MODE_GELU) && g.out != nullptr) {
    rc = make_tmap_2d_64B(&my, g.out, g.M, g.N, g.ldo, g.out_bf16 ? 2 : 4, 32);
    if (rc) return rc;
  }
  if (MODE == MODE_RES) {
    if (g.residual != nullptr) {
      rc = make_tmap_2d_f32(&mres, g.residual, g.M, g.N, g.N, 32);
      if (rc) return rc;
    }
    if (g.ln_gamma != nullptr && g.ln_out_bf16 != nullptr) {
      rc = make_tmap_2d_64B(&mxb, g.ln_out_bf16, g.M, g.N, g.N, 2, 32);
      if (rc) return rc;
    }
    float* y = g.ln_gamma != nullptr ? g.ln_out_f32 : reinterpret_cast<float*>(g.out);
    if (y != nullptr) {
      rc = make_tmap_2d_f32(&my, y, g.M, g.N, g.ln_gamma != nullptr ? (long long)g.N : g.ldo, 32);
      if (rc) return rc;
    }
  }
  const int m_units = ceil_div(g.M, C::UM), n_blocks = ceil_div(g.N, BN);
  const int work = m_units * n_blocks;
  const int max_units = PAIR ? num_sms() / 2 : num_sms();
  int units = work < max_units ? work : max_units;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The persistent grid must be CO-RESIDENT: the fused LayerNorm epilogue spins on counters that CTAs of the same row block
  // publish, so an unscheduled peer would never arrive.  Ask the runtime how many clusters of this configuration fit (MPS
  // thread limits, green contexts or a concurrent kernel can shrink it below #SMs) and size the grid to that; if not even the
  // LN_NB tiles of one row block fit, refuse -- the caller falls back to the unfused GEMM + LayerNorm launches.
  {
    static int resident[64] = {};
    if (resident[dev_] == 0) {
      int ncl = 0;
      cfg.gridDim = dim3(PAIR ? 2 * max_units : max_units);
      if (cudaOccupancyMaxActiveClusters(&ncl, gemm_bf16_kernel<MODE, PAIR, EW, DEEPK>, &cfg) != cudaSuccess || ncl <= 0) {
        cudaGetLastError();
        ncl = -1;  // cannot be queried: keep the #SM-based grid
      }
      resident[dev_] = ncl;
    }
    if (resident[dev_] > 0 && units > resident[dev_]) units = resident[dev_];
    if (MODE == MODE_RES && g.ln_gamma != nullptr && units < LN_NB && work >= LN_NB) {
      set_error("gemm+LN: only %d co-resident tile units (need %d): fused LayerNorm epilogue unavailable", units, LN_NB);
      return AVEXK_ENOMEM;
    }
  }
  cfg.gridDim = dim3(PAIR ? 2 * units : units);
  prof_begin(st, KID_GEMM, 2.0 * g.M * g.N * g.K);
  AVEXK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<MODE, PAIR, EW, DEEPK>, ma, mb, mres, my, mxb, g));
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

template <int MODE>
int launch_any(const void* A, long long lda, const void* W, long long ldw, const GemmArgs& g, cudaStream_t st) {
  if (MODE == MODE_RES && use_pair() && g.K >= 1536) return launch_mode<MODE, true, MODE == MODE_RES>(A, lda, W, ldw, g, st);
  return use_pair() ? launch_mode<MODE, true, false>(A, lda, W, ldw, g, st) : launch_mode<MODE, false, false>(A, lda, W, ldw, g, st);
}

}  // namespace

int gemm_bf16_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias, int gelu,
                     float* raw_out, const float* residual, float res_scale, void* out, long long ldo, int out_bf16,
                     cudaStream_t st, const GemmGate* gate) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.raw_out = raw_out; g.residual = residual; g.res_scale = res_scale;
  g.out = out; g.ldo = ldo; g.out_bf16 = out_bf16;
  if (gate != nullptr) {
    AVEXK_CHECK_ARG(!gelu && residual == nullptr && gate->heads * 64 <= N, "gemm: the gate epilogue belongs to the plain (QKV) flavour");
    g.gate_w = gate->w; g.gate_b = gate->b; g.grep_a = gate->grep_a; g.gate_out = gate->out; g.gate_heads = gate->heads;
  }
  if (gelu) {
    AVEXK_CHECK_ARG(residual == nullptr, "gemm: GELU and residual epilogues are exclusive");
    return launch_any<MODE_GELU>(A, lda, W, ldw, g, st);
  }
  if (residual != nullptr) {
    AVEXK_CHECK_ARG(!out_bf16, "gemm: the residual epilogue writes fp32 (bf16 output + residual is not built)");
    AVEXK_CHECK_ARG(N % 4 == 0 && (out == nullptr || ldo % 4 == 0), "gemm: residual epilogue needs 16-byte aligned rows");
    return launch_any<MODE_RES>(A, lda, W, ldw, g, st);
  }
  return launch_any<MODE_PLAIN>(A, lda, W, ldw, g, st);
}

// scratch of the fused LayerNorm epilogue: [row statistics][arrival / departure counters] (48 bytes per row)
struct LnScratch {
  size_t tiles, stats, count, total;
};
LnScratch ln_scratch_layout(int M) {
  const size_t blocks = 2 * (size_t)ceil_div(M, 2 * BM);  // 128-row blocks, padded to whole CTA pairs
  LnScratch L;
  L.tiles = 0;
  L.stats = 0;
  L.count = L.stats + blocks * BM * 2 * LN_NB * sizeof(float2);
  L.total = L.count + ((2 * blocks * sizeof(int) + 255) & ~size_t(255));
  return L;
}
size_t gemm_ln_scratch_bytes(int M) { return ln_scratch_layout(M).total; }

// y = LayerNorm(A @ W^T + bias + res_scale * residual) in one launch (N must be 768).  raw_out (optional) receives
// A @ W^T + bias (the fc2 hook); ln_out_f32 may alias `residual` (a CTA reads a residual element before it rewrites it).
// zero_counters: the counters at the end of `scratch` must be zero on entry; every launch leaves them zero, so only the
// first use of a scratch buffer needs the memset.
int gemm_bf16_ln_launch(const void* A, long long lda, const void* W, long long ldw, int M, int K, const float* bias, float* raw_out,
                        const float* residual, float res_scale, const float* gamma, const float* beta, float eps,
                        float* ln_out_f32, __nv_bfloat16* ln_out_bf16, void* scratch, size_t scratch_bytes, int zero_counters,
                        cudaStream_t st, long long* pool_raw, long long* pool_y, int pool_rows) {
  AVEXK_CHECK_ARG((pool_raw == nullptr && pool_y == nullptr) || (pool_rows >= 32 && M % pool_rows == 0),
                  "gemm+LN: fused pooling needs >= 32 token rows per clip and M a multiple of it (M=%d rows=%d)", M, pool_rows);
  const LnScratch L = ln_scratch_layout(M);
  AVEXK_CHECK_ARG(scratch != nullptr && scratch_bytes >= L.total && (reinterpret_cast<uintptr_t>(scratch) & 255) == 0,
                  "gemm+LN: scratch too small or misaligned");
  char* base = reinterpret_cast<char*>(scratch);
  if (zero_counters) AVEXK_CUDA(cudaMemsetAsync(base + L.count, 0, L.total - L.count, st));
  GemmArgs g{};
  g.M = M; g.N = LN_C; g.K = K;
  g.bias = bias; g.raw_out = raw_out; g.residual = residual; g.res_scale = res_scale;
  g.ln_gamma = gamma; g.ln_beta = beta; g.ln_eps = eps; g.ln_out_f32 = ln_out_f32; g.ln_out_bf16 = ln_out_bf16;
  g.ln_stats = reinterpret_cast<float2*>(base + L.stats);
  g.ln_count = reinterpret_cast<int*>(base + L.count);
  g.pool_raw = pool_raw; g.pool_y = pool_y; g.pool_rows = pool_rows;
  return launch_any<MODE_RES>(A, lda, W, ldw, g, st);
}

// 1x1 convolution in NHWC: out[M, N] = act((A[M, K] @ W[N, K]^T) * scale + shift) (+ res); K, N multiples of 8: the TMA
// maps zero-fill the out-of-range part of the last K block / N tile.
int conv1x1_launch(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                   const __nv_bfloat16* res, float* raw_out, void* out, int out_bf16, cudaStream_t st) {
  AVEXK_CHECK_ARG(K % 8 == 0 && N % 8 == 0 && K > 0 && N > 0, "conv1x1: channel counts must be multiples of 8 (K=%d N=%d)", K, N);
  if (M == 0) return AVEXK_OK;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.bias = shift; g.raw_out = raw_out; g.out = out; g.ldo = N; g.out_bf16 = out_bf16;
  g.scale = scale; g.silu = silu; g.res_bf16 = res;
  return launch_any<MODE_CONV>(A, K, W, K, g, st);
}

}  // namespace avexk

extern "C" int avexk_gemm_config(int pair) {
  const int prev = avexk::use_pair() ? 1 : 0;
  if (pair == 0 || pair == 1) avexk::g_pair = pair;
  return prev;
}

extern "C" int avexk_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                               const float* bias, int gelu, float* raw_out, const float* residual, float res_scale, void* out,
                               long long ldo, int out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && (out || raw_out), "avexk_gemm_bf16: null operand");
  AVEXK_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0, "avexk_gemm_bf16: unsupported shape M=%d N=%d K=%d", M, N, K);
  AVEXK_CHECK_ARG(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0 && (out == nullptr || (ldo >= N && ldo % 8 == 0)),
                  "avexk_gemm_bf16: bad leading dimensions");
  if (M == 0) return AVEXK_OK;
  return gemm_bf16_launch(A, lda, W, ldw, M, N, K, bias, gelu, raw_out, residual, res_scale, out, ldo, out_bf16,
                          reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t avexk_gemm_ln_scratch_bytes(int M) { return M > 0 ? avexk::gemm_ln_scratch_bytes(M) : 0; }

extern "C" int avexk_gemm_bf16_ln_pooled(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                                         const float* bias, float* raw_out, const float* residual, float res_scale,
                                         const float* gamma, const float* beta, float eps, float* out_f32, void* out_bf16,
                                         void* scratch, size_t scratch_bytes, int rows_per_clip, float* pooled_raw,
                                         float* pooled_y, void* pool_ws, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && gamma && beta && (out_f32 || out_bf16 || pooled_raw || pooled_y), "avexk_gemm_bf16_ln_pooled: null operand");
  AVEXK_CHECK_ARG(N == LN_C, "avexk_gemm_bf16_ln_pooled: the fused LayerNorm epilogue is built for N = %d (got %d)", LN_C, N);
  AVEXK_CHECK_ARG(M >= 0 && K > 0 && K % 8 == 0 && lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0,
                  "avexk_gemm_bf16_ln_pooled: unsupported shape / leading dimensions");
  AVEXK_CHECK_ARG((!pooled_raw && !pooled_y) || (pool_ws && rows_per_clip >= 32 && M % rows_per_clip == 0),
                  "avexk_gemm_bf16_ln_pooled: pooling needs pool_ws and rows_per_clip >= 32 dividing M");
  if (M == 0) return AVEXK_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nclips = rows_per_clip > 0 ? M / rows_per_clip : 0;
  long long* acc_raw = pooled_raw ? reinterpret_cast<long long*>(pool_ws) : nullptr;
  long long* acc_y = pooled_y ? reinterpret_cast<long long*>(pool_ws) + (size_t)nclips * LN_C : nullptr;
  if (acc_raw || acc_y) AVEXK_CUDA(cudaMemsetAsync(pool_ws, 0, (size_t)2 * nclips * LN_C * sizeof(long long), st));
  int rc = gemm_bf16_ln_launch(A, lda, W, ldw, M, K, bias, raw_out, residual, res_scale, gamma, beta, eps, out_f32,
                               reinterpret_cast<__nv_bfloat16*>(out_bf16), scratch, scratch_bytes, 1, st, acc_raw, acc_y, rows_per_clip);
  if (rc) return rc;
  if (acc_raw) rc = launch_pool_finalize(acc_raw, nclips, LN_C, 1.0f / rows_per_clip, pooled_raw, st);
  if (rc) return rc;
  if (acc_y) rc = launch_pool_finalize(acc_y, nclips, LN_C, 1.0f / rows_per_clip, pooled_y, st);
  return rc;
}

extern "C" int avexk_gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias,
                                  float* raw_out, const float* residual, float res_scale, const float* gamma, const float* beta,
                                  float eps, float* out_f32, void* out_bf16, void* scratch, size_t scratch_bytes, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && gamma && beta && (out_f32 || out_bf16), "avexk_gemm_bf16_ln: null operand");
  AVEXK_CHECK_ARG(N == LN_C, "avexk_gemm_bf16_ln: the fused LayerNorm epilogue is built for N = %d (got %d)", LN_C, N);
  AVEXK_CHECK_ARG(M >= 0 && K > 0 && K % 8 == 0 && lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0,
                  "avexk_gemm_bf16_ln: unsupported shape / leading dimensions");
  if (M == 0) return AVEXK_OK;
  return gemm_bf16_ln_launch(A, lda, W, ldw, M, K, bias, raw_out, residual, res_scale, gamma, beta, eps, out_f32,
                             reinterpret_cast<__nv_bfloat16*>(out_bf16), scratch, scratch_bytes, 1,
                             reinterpret_cast<cudaStream_t>(stream), nullptr, nullptr, 0);
}
