// posconv.cu -- BEATs convolutional position embedding on the sm_100a tensor cores.
//
// Replaces avex/models/beats/backbone.py:52-68,172-174 + modules.py:67-94:
//   y = x + GELU(Conv1d(C, C, k=128, pad=64, groups=16)(x^T)[:, :, :N]^T)      (weight-norm resolved at load)
// as 16 implicit GEMMs, one per channel group:
//   out[m, co] = sum_{t<128} sum_{ci<48} x[m + t - 64, g*48 + ci] * w[g*48 + co, ci, t]
//
// A grouped conv has only 48 output columns per GEMM, so a tap-by-tap formulation (UMMA 128 x 48 x 16) is bound by the
// issue rate of tcgen05.mma from its single issuing thread (768 instructions per 256 tokens; measured 2.9 ms per
// 256 x 10 s batch, tensor pipe 25 % busy).  This kernel widens N to 192 by taking FOUR taps per instruction:
//   D_j[n', co] += x_win[4u + n', :] . w[:, co, 4u + j]        j = 0..3, one UMMA with B = [w_4u ; w_4u+1 ; w_4u+2 ; w_4u+3]
// The window row that output row m needs for tap 4u+j is m + 4u + j = (m + j) + 4u, so D_j holds the contribution to
// output row n' - j:   out[m] = sum_j D_j[m + j].   The epilogue undoes the shift with warp shuffles (TMEM lane = row)
// and a 3-row halo between lane quarters through shared memory; rows 125..127 of a 128-row tile have no complete sum,
// so tiles advance by 125 tokens.  Two CTAs form a pair (tcgen05 cta_group::2): M = 256 (each CTA its own 125-token tile
// and its own activation window), each CTA stages half of the 192 weight rows.  32 x 3 = 96 UMMAs of 256 x 192 x 16 per
// 250 tokens instead of 768, and the A operand is read from shared memory once per four taps.
//
// The activation window (rows n0-64 .. n0+191 of the group-padded bf16 copy xg [B, N, 16*64]) is loaded ONCE per tile by
// TMA (3-D tensor map: rows outside the clip are zero-filled by hardware -- exactly the conv's zero padding).  Step u
// reads it shifted down by 4u rows: the UMMA shared-memory descriptor simply starts 4u*128 bytes later (the hardware
// swizzles on absolute address bits, so the phase still matches what TMA wrote).  K = 48 per tap (the 16 pad channels are
// never multiplied); accumulators in TMEM (2 buffers x 192 columns).  Epilogue: shifted sum, + bias, GELU, + x0, fp32 store.
// Roofline: tensor pipe (2 * 48 * 48 * 128 FLOP per token and group).
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int BM = 128, CG = 48, TAPS = 128, J = 4, STEPS = TAPS / J, NCOL = J * CG;  // 192 accumulator columns
constexpr int VALID = BM - (J - 1), PAIR_TOKENS = 2 * VALID;                           // 125 complete rows per tile
constexpr int WSTAGES = 8;
constexpr int WIN_ROWS = 256, WIN_BYTES = WIN_ROWS * 128;  // rows n0-64 .. n0+191 (251 needed)
constexpr int W_ROWS = NCOL / 2, W_BYTES = W_ROWS * 128;   // each CTA of the pair stages 96 of the 192 weight rows
constexpr int HALO_FLOATS = 5 * (J - 1) * (J - 1) * CG;    // [quarter (+1 never-read pad)][j-1][row][co]
constexpr int OFF_W = 2 * WIN_BYTES, OFF_HALO = OFF_W + WSTAGES * W_BYTES, OFF_BAR = OFF_HALO + HALO_FLOATS * 4;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 512;
constexpr int NTHREADS = 192;
static_assert(OFF_HALO % 16 == 0 && OFF_BAR % 8 == 0 && SMEM_BYTES <= 232448, "posconv: shared memory layout");

struct PcArgs {
  int B, N, G, pairs_per_clip;
  const float* bias;   // [C]
  const float* x0;     // [B*N, C] fp32 residual
  float* out;          // [B*N, C] fp32
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major SWIZZLE_128B operand that starts `row` 128-byte rows below a 1024-byte aligned base.  The tensor core applies
// the 128-byte swizzle to ABSOLUTE shared-memory address bits (as TMA does when it writes the tile), so a start address
// that is only 128-byte aligned reads the shifted rows correctly with the descriptor's base-offset field left at 0
// (measured on B200: base_offset = row mod 8 gives wrong data, 0 is bit-correct against the oracle).
__device__ __forceinline__ uint64_t make_sw128_desc_rows(uint32_t base_addr, int row) {
  return ptx::make_sw128_desc(base_addr + row * 128);
}

__global__ void __launch_bounds__(NTHREADS, 1)
posconv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const PcArgs g) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  float* halo = reinterpret_cast<float*>(smem + OFF_HALO);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* win_full = bars;                      // [2]        (leader's are waited on)
  uint64_t* win_empty = bars + 2;                 // [2]
  uint64_t* w_full = bars + 4;                    // [WSTAGES]  (leader's are waited on)
  uint64_t* w_empty = bars + 4 + WSTAGES;         // [WSTAGES]
  uint64_t* tfull_bar = bars + 4 + 2 * WSTAGES;   // [2]
  uint64_t* tempty_bar = bars + 6 + 2 * WSTAGES;  // [2]        (leader's collect both CTAs' epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * WSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int unit0 = static_cast<int>(blockIdx.x >> 1), num_units = static_cast<int>(gridDim.x >> 1);
  const int C = g.G * CG;
  const int num_tiles = g.B * g.pairs_per_clip * g.G;  // a tile = (clip, 250-token span, group), one per CTA pair

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_w);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&win_full[i], 1);
      ptx::mbar_init(&win_empty[i], 1);
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 8);  // four epilogue warps of each CTA
    }
    for (int i = 0; i < WSTAGES; ++i) {
      ptx::mbar_init(&w_full[i], 1);
      ptx::mbar_init(&w_empty[i], 1);
    }
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (clip b, span pt, group grp); groups fastest so concurrent pairs share the activation rows in L2
  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own window, own half of the weight rows) =====================
    if (lane == 0) {
      const uint32_t win_full0 = ptx::mapa(ptx::smem_u32(&win_full[0]), 0), w_full0 = ptx::mapa(ptx::smem_u32(&w_full[0]), 0);
      auto load_window = [&](int tile, int it) {
        const int grp = tile % g.G, mt = tile / g.G;
        const int pt = mt % g.pairs_per_clip, b = mt / g.pairs_per_clip;
        const int wb = it & 1;
        const int n0 = pt * PAIR_TOKENS + static_cast<int>(rank) * VALID;  // first output token of this CTA
        ptx::mbar_wait(&win_empty[wb], ((it >> 1) & 1) ^ 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&win_full[wb], 2 * WIN_BYTES);
        for (int bx = 0; bx < WIN_ROWS / 128; ++bx)
          tma_load_3d_pair(smem + wb * WIN_BYTES + bx * 128 * 128, &map_x, win_full0 + wb * 8, grp * 64, n0 - TAPS / 2 + bx * 128, b);
      };
      int stage = 0, it = 0;
      uint32_t phase = 0;
      if (unit0 < num_tiles) load_window(unit0, 0);
      for (int tile = unit0; tile < num_tiles; tile += num_units, ++it) {
        if (tile + num_units < num_tiles) load_window(tile + num_units, it + 1);  // one tile ahead
        const int grp = tile % g.G;
        for (int u = 0; u < STEPS; ++u) {
          ptx::mbar_wait(&w_empty[stage], phase ^ 1);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&w_full[stage], 2 * W_BYTES);
          ptx::tma_load_2d_pair(smem + OFF_W + stage * W_BYTES, &map_w, w_full0 + stage * 8, 0,
                                (grp * TAPS + u * J) * CG + static_cast<int>(rank) * W_ROWS);
          if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
        }
      }
      // tail: every release has landed here before the CTA may exit
      for (int s = 0; s < WSTAGES; ++s) {
        ptx::mbar_wait(&w_empty[stage], phase ^ 1);
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
      for (int k = 0; k < 2; ++k, ++it) ptx::mbar_wait(&win_empty[it & 1], ((it >> 1) & 1) ^ 1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * BM, NCOL);  // fp16 operands: see elementwise.cu group_pad_kernel
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += num_units, ++it) {
        const int acc = it & 1, wb = it & 1;
        ptx::mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        ptx::mbar_wait(&win_full[wb], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t win_addr = ptx::smem_u32(smem + wb * WIN_BYTES);
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int u = 0; u < STEPS; ++u) {
          ptx::mbar_wait(&w_full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t db = ptx::make_sw128_desc(ptx::smem_u32(smem + OFF_W + stage * W_BYTES));
          const uint64_t da = make_sw128_desc_rows(win_addr, u * J);
#pragma unroll
          for (int k = 0; k < CG / 16; ++k) ptx::umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (u | k) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(&w_empty[stage], 3);
          if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tfull_bar[acc], 3);
        ptx::umma_commit_pair(&win_empty[wb], 3);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs): TMEM lane = window-aligned row n' =====================
    const int quarter = warp & 3;
    const uint32_t tempty0 = ptx::mapa(ptx::smem_u32(&tempty_bar[0]), 0);
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += num_units, ++it) {
      const int acc = it & 1;
      const int grp = tile % g.G, mt = tile / g.G;
      const int pt = mt % g.pairs_per_clip, b = mt / g.pairs_per_clip;
      const int m = quarter * 32 + lane;                                             // output row inside the tile
      const int n = pt * PAIR_TOKENS + static_cast<int>(rank) * VALID + m;             // token inside the clip
      const bool ok = m < VALID && n < g.N;
      const size_t row_off = ((size_t)b * g.N + (ok ? n : 0)) * C + grp * CG;
      float4 rs[CG / 4];  // the residual row, in flight while the tensor core finishes the tile
#pragma unroll
      for (int i = 0; i < CG / 4; ++i) rs[i] = ok ? __ldcs(reinterpret_cast<const float4*>(g.x0 + row_off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      ptx::mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256;
      // halo: rows 0..j-1 of this quarter's D_j are the tail of the previous quarter's shifted sums
      if (quarter > 0) {
#pragma unroll
        for (int j = 1; j < J; ++j) {
#pragma unroll
          for (int c = 0; c < CG / 16; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(t_addr + j * CG + c * 16, r);
            ptx::tmem_ld_wait();
            if (lane < j) {
              float4* dst = reinterpret_cast<float4*>(halo + ((quarter * (J - 1) + (j - 1)) * (J - 1) + lane) * CG + c * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
          }
        }
      }
      ptx::named_bar_sync(1, 128);  // the four epilogue warps: halos are visible
#pragma unroll
      for (int c = 0; c < CG / 16; ++c) {
        float s[16];
        {
          uint32_t r[16];
          tmem_ld_32x16(t_addr + c * 16, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i] = __uint_as_float(r[i]);
        }
#pragma unroll
        for (int j = 1; j < J; ++j) {
          uint32_t r[16];
          tmem_ld_32x16(t_addr + j * CG + c * 16, r);
          ptx::tmem_ld_wait();
          const bool from_halo = lane + j >= 32;  // row m + j lives in the next quarter
          const float* hp = halo + (((quarter + 1) * (J - 1) + (j - 1)) * (J - 1) + ((lane + j) & 31)) * CG + c * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float up = __shfl_down_sync(0xffffffffu, __uint_as_float(r[i]), j);
            s[i] += (from_halo && quarter < 3) ? hp[i] : up;  // quarter 3's last rows are never stored (m >= VALID)
          }
        }
        if (c == CG / 16 - 1) {  // the accumulator is drained: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
        }
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = c * 16 + i * 4;
            const float4 bs = __ldg(reinterpret_cast<const float4*>(g.bias + grp * CG + col));  // warp-uniform address
            const float2 g0 = gelu_erf_fast2(make_float2(s[4 * i + 0] + bs.x, s[4 * i + 1] + bs.y));
            const float2 g1 = gelu_erf_fast2(make_float2(s[4 * i + 2] + bs.z, s[4 * i + 3] + bs.w));
            const float4 r4 = rs[c * 4 + i];
            __stcs(reinterpret_cast<float4*>(g.out + row_off + col), make_float4(g0.x + r4.x, g0.y + r4.y, g1.x + r4.z, g1.y + r4.w));
          }
        }
      }
      ptx::named_bar_sync(1, 128);  // the halos may be overwritten by the next tile
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

// xg [B, N, G*64] bf16 (group-padded activations), Wpc [G][128][48][64] bf16, out = x0 + gelu(conv + bias), fp32 [B*N, G*48]
int launch_posconv(const __nv_bfloat16* xg, const __nv_bfloat16* Wpc, const float* bias, const float* x0, float* out, int B,
                   int N, int G, int cg, int taps, cudaStream_t st) {
  AVEXK_CHECK_ARG(cg == CG && taps == TAPS, "posconv kernel is specialised to 48 channels/group and 128 taps (got %d, %d)", cg, taps);
  if (B == 0 || N == 0) return AVEXK_OK;
  static bool attr_set[64] = {};  // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    AVEXK_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[dev_] = true;
  }
  CUtensorMap mx, mw;
  int rc = make_tmap_3d_bf16(&mx, xg, (long long)G * 64, N, B, (long long)G * 64, (long long)N * G * 64, 64, 128, 1, true);
  if (rc) return rc;
  // Wpc [G][TAPS][48][64] bf16 (launch_posconv_pack): four consecutive taps of a group are 192 consecutive 128-byte rows
  rc = make_tmap_2d_bf16(&mw, Wpc, (long long)G * TAPS * CG, 64, 64, W_ROWS, 64);
  if (rc) return rc;
  PcArgs a{B, N, G, ceil_div(N, PAIR_TOKENS), bias, x0, out};
  const int tiles = B * a.pairs_per_clip * G;
  const int units = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  prof_begin(st, KID_POSCONV, 2.0 * B * N * (double)(G * CG) * CG * TAPS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * units);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AVEXK_CUDA(cudaLaunchKernelEx(&cfg, posconv_kernel, mx, mw, a));
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

}  // namespace avexk

extern "C" size_t avexk_posconv_workspace_bytes(int B, int N, int C, int groups, int taps) {
  if (B <= 0 || N <= 0 || C <= 0 || groups <= 0 || taps <= 0) return 0;
  const size_t M = (size_t)B * N;
  return ((M * groups * 64 * 2 + 255) & ~size_t(255)) + (((size_t)C * taps * 64 * 2 + 255) & ~size_t(255)) + 4096;
}

extern "C" int avexk_posconv(float* x0, int B, int N, int C, int groups, int taps, const float* weight_g, const float* weight_v,
                             const float* bias, const uint8_t* key_pad, float* out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(x0 && weight_g && weight_v && bias && out && workspace, "avexk_posconv: null argument");
  AVEXK_CHECK_ARG(B > 0 && N > 0 && groups > 0 && C % groups == 0, "avexk_posconv: bad shape");
  AVEXK_CHECK_ARG(workspace_bytes >= avexk_posconv_workspace_bytes(B, N, C, groups, taps), "avexk_posconv: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t M = (size_t)B * N;
  char* p = reinterpret_cast<char*>(workspace);
  __nv_bfloat16* xg = reinterpret_cast<__nv_bfloat16*>(p);
  p += (M * groups * 64 * 2 + 255) & ~size_t(255);
  __nv_bfloat16* W = reinterpret_cast<__nv_bfloat16*>(p);
  p += ((size_t)C * taps * 64 * 2 + 255) & ~size_t(255);
  float* nrm = reinterpret_cast<float*>(p);
  int rc = launch_posconv_pack(weight_v, weight_g, C, C / groups, taps, nrm, W, st);
  if (rc) return rc;
  rc = launch_group_pad(x0, key_pad, (long long)M, groups, C / groups, xg, st);
  if (rc) return rc;
  return launch_posconv(xg, W, bias, x0, out, B, N, groups, C / groups, taps, st);
}
