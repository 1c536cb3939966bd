// attention_tc.cu -- fused multi-head attention on the sm_100a tensor cores (tcgen05 + TMEM), with BEATs' gated
// relative-position bias applied in-tile.
//
// Replaces avex/models/beats/backbone.py:526-571: the [B,H,N,N] expansion of compute_bias, the gate
// (grep_linear -> view(..,2,4).sum(-1) -> sigmoid -> gate_a*(gate_b*grep_a-1)+2, from UNscaled q), the materialised
// `gate * position_bias` mask (3.0 GB fp32 per layer at B=256, N=496), the key-padding -inf mask, SDPA, and the
// permute/contiguous that follows.  Nothing of size N^2 touches HBM: the Toeplitz bias is the [H, 2N-1] vector
// bias[h, j-i+N-1], indexed inside the score tile and scaled by the per-row gate.
//
// Persistent kernel, one CTA per SM, 20 warps (16 softmax, loader, MMA issuer, 2 register donors: setmaxnreg gives the
// softmax warps 104 registers).  A work item is (clip, head, pair of 128-query tiles); both tiles of the pair share every K/V tile that streams through a 3-stage TMA ring, and their softmax phases interleave so the tensor
// pipe, the MUFU pipe and the TMA engine are all busy:
//   warps 0..7   softmax group A, warps 8..15 softmax group B.  thread = (query row r = 32*(warp&3)+lane, column half
//                ch = (warp>>2)&1).  Per 128-key tile a thread reads its 64 scores from TMEM (tcgen05.ld), adds scale /
//                gated bias / mask, joins the row max with its partner through shared memory, exponentiates
//                (ex2.approx) and writes P as bf16 pairs back into TMEM over its own score columns, where the PV MMA reads
//                it as its A operand (P never touches shared memory).  The running max is lazy (FA4-style): O in TMEM is only rescaled when the
//                max grows by more than 2^8, which is exact after the final 1/l.
//   warp 16      loaders: lane 0 streams 128-key K and V tiles (3-D tensor map: rows past the clip end are
//                zero-filled by hardware), running ahead across work items; lane 1 loads the Q tiles.
//   warp 17      MMA issuer: S_g = Q_g K^T (128x128x64, K-major operands from shared memory), the gate logits Q_g Wg^T, and
//                O_g += P_g V (128x64x128; A = P from TMEM, B = V consumed MN-major exactly as it lies in the qkv buffer --
//                no transposed copy), tcgen05.commit -> mbarriers.  PV_g(t) is issued before S_g(t+1), which overwrites P.
//   TMEM (512 columns): S_A @0, S_B @128 (fp32 128x128 each; P_g as bf16 pairs over columns 0..31 and 64..95 of S_g),
//                O_A @256, O_B @320 (fp32 128x64 each), gate logits @384 / @400.
//   Shared memory: Q (2 tiles), a 5-stage K/V ring, the bias / mask tables -- the measured limiter before P moved to TMEM
//                was the shared-memory pipe (64 % busy: bias-window loads 31 %, P stores 11 %, MMA operand reads 21 %).
// Roofline: MUFU (one ex2 per score: B*H*N^2 per layer) and FP32 issue, not the tensor pipe; see DESIGN.md section 4.
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int GROUP_WARPS = 8, GROUP_THREADS = 32 * GROUP_WARPS;
constexpr int WARP_LOAD = 16, WARP_MMA = 17, NTHREADS = 32 * 20;  // warps 18-19: idle register donors (setmaxnreg is per warpgroup)
// setmaxnreg moves registers inside the launch allocation: 640 x 96 = 61440 = 128*64 + 512*104
constexpr int CTRL_REGS = 64, SOFTMAX_REGS = 104;
constexpr int KV_STAGES = 4;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 domain: P stays below 2^8, exact after normalisation
constexpr float P_CLAMP = 96.0f;           // a score more than 2^96 above the reference is clamped (row sums stay finite)

constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
// Four copies of the 255-entry bias window per tile, copy s shifted right by s floats so that every row can read
// 16-byte aligned float4s.  Copy offsets (floats) are chosen so that the 8 lanes of each LDS.128 phase hit 8 distinct
// 16-byte bank groups (offsets / 4 mod 8 = 0, 1, 3, 5).
constexpr int WIN_FLOATS = 1056;
constexpr int OFF_Q = 0;                                 // [2] tiles
constexpr int OFF_KV = OFF_Q + 2 * TILE_BYTES;           // [KV_STAGES] x (K tile, V tile)
constexpr int OFF_OSTG = OFF_KV + KV_STAGES * 2 * TILE_BYTES;  // [16 softmax warps] x (32 rows x 64 bytes, 64-byte swizzle): O on its way out (TMA store)
constexpr int OFF_WG = OFF_OSTG + 16 * 2048;  // gate weights as a 16 x 64 bf16 UMMA operand (K-major, 128-byte swizzle)
constexpr int OFF_TAB = OFF_WG + 2048;
constexpr int TAB_WIN = 0;                               // [2][WIN_FLOATS] float
constexpr int TAB_MASK = TAB_WIN + 2 * WIN_FLOATS * 4;  // [2][128] float
constexpr int TAB_PMAX = TAB_MASK + 2 * 128 * 4;         // exchange buffers: [2 parity][2][128] tile max, [2][128] row sums l, [128] shared estimate
constexpr int TAB_BYTES = TAB_PMAX + 1024 * 4 + 16;  // + one int: first tile with a valid key
constexpr int OFF_GATEB = OFF_TAB + 2 * TAB_BYTES;  // [2] float: gate bias
constexpr int OFF_BAR = OFF_GATEB + 16;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "attention_tc: shared memory budget");

struct AttnTcArgs {
  int B, N, H;
  int n_items;             // B * H * ceil(N / 256): read from the constant bank where needed (a register-resident copy spilled)
  const float* gate_w;     // [2,64]
  const float* gate_b;     // [2]
  const float* grep_a;     // [H]
  const float* bias_vec;   // [H, 2N-1]
  const uint8_t* key_pad;  // [B,N] or null
  __nv_bfloat16* out;      // [B*N, H*64]
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// the same barrier, returning the OR of `pred` over its threads
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(out)
      : "r"((uint32_t)pred), "r"(id), "r"(nthreads)
      : "memory");
  return out != 0;
}
// explicit shared-space accesses (32-bit shared addresses): the carve-up below goes through an aligned byte offset, and
// generic LD/ST on those pointers costs a long-scoreboard round trip
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns (small chunks keep the softmax threads inside their 104 registers)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int win_copy_offset(int s) { return s == 0 ? 0 : (s == 1 ? 260 : (s == 2 ? 524 : 788)); }

// x = S * scale + gate * bias (+ mask): two scores per packed fp32x2 instruction.  Software pipeline per 16-column
// chunk: the TMEM load of chunk c+1 and the window loads of chunk c+1 are in flight while chunk c is processed
// (tcgen05.wait::ld waits for every outstanding load, so the next load is issued right after the wait).
#define AVEXK_SCORE_QUAD(R, lo, hi)                                                                                       \
  float2 lo = __ffma2_rn(make_float2(__uint_as_float(R[4 * k + 0]), __uint_as_float(R[4 * k + 1])), s2,                    \
                         __fmul2_rn(g2, make_float2(w[k].x, w[k].y)));                                                    \
  float2 hi = __ffma2_rn(make_float2(__uint_as_float(R[4 * k + 2]), __uint_as_float(R[4 * k + 3])), s2,                    \
                         __fmul2_rn(g2, make_float2(w[k].z, w[k].w)));                                                    \
  if (chunk < 3) w[k] = lds128(win_addr + ((chunk + 1) * 16 + k * 4) * 4);                                               \
  if (MASKED) {                                                                                                           \
    const float4 mk = lds128(mask_addr + (chunk * 16 + k * 4) * 4);                                                       \
    lo = __fadd2_rn(lo, make_float2(mk.x, mk.y));                                                                         \
    hi = __fadd2_rn(hi, make_float2(mk.z, mk.w));                                                                         \
  }

// Reference estimate (only on the first tile with a valid key): the max of ONE 16-key chunk of the row, the chunk that holds
// the clip's first valid key.  Both column halves of a row read the same chunk, so they agree without an exchange.  Any
// reference within ~2^100 of the true row max is exact after the final 1/l (P is bf16: the precision is relative, the
// reference only positions the exponent range); from the next tile on the reference follows the true running max.
template <bool MASKED>
__device__ __forceinline__ float score_est(uint32_t taddr, uint32_t win_addr, uint32_t mask_addr, float gate, float qk_scale) {
  float mx = -INFINITY;
  uint32_t ra[16];
  tmem_ld_32x16(taddr, ra);
  float4 w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = lds128(win_addr + k * 16);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float x0 = fmaf(__uint_as_float(ra[4 * k + 0]), qk_scale, gate * w[k].x);
    float x1 = fmaf(__uint_as_float(ra[4 * k + 1]), qk_scale, gate * w[k].y);
    float x2 = fmaf(__uint_as_float(ra[4 * k + 2]), qk_scale, gate * w[k].z);
    float x3 = fmaf(__uint_as_float(ra[4 * k + 3]), qk_scale, gate * w[k].w);
    if (MASKED) {
      const float4 mk = lds128(mask_addr + k * 16);
      x0 += mk.x; x1 += mk.y; x2 += mk.z; x3 += mk.w;
    }
    mx = fmaxf(fmaxf(mx, fmaxf(x0, x1)), fmaxf(x2, x3));
  }
  return mx;
}

// Stream the 64 scores once: p = 2^(min(x - m, P_CLAMP)) -> bf16 pairs -> back into TMEM, over the first half of the
// thread's own score columns (chunk c of 16 fp32 scores becomes 8 packed columns at 8c: always columns this thread has
// already read), where the PV MMA takes them as its A operand -- P never touches shared memory.  Accumulates the row sum
// and tracks the tile max (used to move the reference max for the NEXT tile).  Nothing is kept in registers.
template <bool MASKED>
__device__ __forceinline__ void stream_tile(uint32_t taddr, uint32_t win_addr, uint32_t mask_addr, float gate, float qk_scale,
                                            float m_ref, float& mx_out, float& sum_out) {
  float mx = -INFINITY;
  const float2 g2 = make_float2(gate, gate), s2 = make_float2(qk_scale, qk_scale), nm2 = make_float2(-m_ref, -m_ref);
  float2 sum2 = make_float2(0.f, 0.f);
  uint32_t ra[16], rb[16];
  float4 w[4];
  tmem_ld_32x16(taddr, ra);
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = lds128(win_addr + k * 16);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int chunk = 0; chunk < 4; ++chunk) {
    if (chunk < 3) tmem_ld_32x16(taddr + (chunk + 1) * 16, (chunk & 1) ? ra : rb);
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 lo_, hi_;
      if (chunk & 1) {
        AVEXK_SCORE_QUAD(rb, lo, hi)
        lo_ = lo; hi_ = hi;
      } else {
        AVEXK_SCORE_QUAD(ra, lo, hi)
        lo_ = lo; hi_ = hi;
      }
      mx = fmaxf(fmaxf(mx, fmaxf(lo_.x, lo_.y)), fmaxf(hi_.x, hi_.y));
      lo_ = __fadd2_rn(lo_, nm2);
      hi_ = __fadd2_rn(hi_, nm2);
      // the reference max may lag the true max (it moves between tiles): clamp so bf16 P cannot overflow
      const float2 p0 = make_float2(ex2(fminf(lo_.x, P_CLAMP)), ex2(fminf(lo_.y, P_CLAMP)));
      const float2 p1 = make_float2(ex2(fminf(hi_.x, P_CLAMP)), ex2(fminf(hi_.y, P_CLAMP)));
      sum2 = __fadd2_rn(sum2, __fadd2_rn(p0, p1));
      pk[2 * k] = pack_bf16(p0.x, p0.y);
      pk[2 * k + 1] = pack_bf16(p1.x, p1.y);
    }
    if (chunk < 3) ptx::tmem_ld_wait();  // also orders the store below after the loads of the columns it overwrites
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr + chunk * 8),
                 "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                 : "memory");
  }
  tmem_st_wait();
  mx_out = mx;
  sum_out = sum2.x + sum2.y;
}

struct Item {
  int b, h, q0;
  bool has_b;  // the pair's second tile holds at least one valid query row
};
__device__ __forceinline__ Item decode_item(int item, int npairs, int H, int N) {
  Item it;
  const int pair = item % npairs, bh = item / npairs;
  it.h = bh % H;
  it.b = bh / H;
  it.q0 = pair * 2 * BQ;
  it.has_b = it.q0 + BQ < N;
  return it;
}

__global__ void __launch_bounds__(NTHREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out, const AttnTcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment (128B swizzle atoms); pointer arithmetic on the __shared__ array keeps the address space
  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_a = ptx::smem_u32(smem);
  unsigned char* sQ = smem + OFF_Q;
  unsigned char* sKV = smem + OFF_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;     // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [KV_STAGES <= 8]
  uint64_t* kv_empty = bars + 12;  // [KV_STAGES <= 8]
  uint64_t* s_full = bars + 20;    // [2]
  uint64_t* p_full = bars + 22;    // [2]
  uint64_t* o_full = bars + 24;    // [2]
  uint64_t* o_free = bars + 26;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  static_assert(KV_STAGES <= 8, "barrier layout");

  int tid;  // read once through a volatile asm: otherwise the compiler re-reads %tid.x (S2R, behind the MUFU queue) per item
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5, lane = tid & 31;
  const int N = a.N;
  const int n_kv = (N + BKV - 1) / BKV;
  const int npairs = (N + 2 * BQ - 1) / (2 * BQ);
  const int n_items = a.n_items;

  // Gate weights -> a 16-row UMMA B operand, once per CTA: rows 0,1 = bf16(w), rows 2,3 = bf16(w - hi) (the two halves
  // are summed after the MMA, so the gate logits carry ~16 mantissa bits of w), rows 4..15 = 0.
  for (int e = tid; e < 16 * HD; e += NTHREADS) {
    const int row = e >> 6, col = e & 63;
    const float w = row < 4 ? __ldg(a.gate_w + (row & 1) * HD + col) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 v = row < 2 ? hi : __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(smem_a + OFF_WG + off), "h"(__bfloat16_as_ushort(v)) : "memory");
  }
  if (tid < 2) sts32(smem_a + OFF_GATEB + tid * 4, __ldg(a.gate_b + tid));
  ptx::fence_proxy_async();  // the gate operand is read by the tensor core (async proxy)
  if (warp == WARP_MMA) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&map_qkv);
      ptx::prefetch_tensormap(&map_out);
      for (int g = 0; g < 2; ++g) {
        ptx::mbar_init(&q_full[g], 1);
        ptx::mbar_init(&q_empty[g], 1);  // the commit of the item's last S MMA
        ptx::mbar_init(&s_full[g], 1);
        ptx::mbar_init(&p_full[g], GROUP_WARPS);
        ptx::mbar_init(&o_full[g], 1);
        ptx::mbar_init(&o_free[g], GROUP_WARPS);
      }
      for (int s = 0; s < KV_STAGES; ++s) {
        ptx::mbar_init(&kv_full[s], 1);
        ptx::mbar_init(&kv_empty[s], 1);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_LOAD) {
    ptx::setmaxnreg_dec<CTRL_REGS>();
    if (lane == 0) {
      // ===================== K/V loader =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(item, npairs, a.H, N);
        const int col_k = (a.H + it.h) * HD, col_v = (2 * a.H + it.h) * HD;
        for (int t = 0; t < n_kv; ++t) {
          ptx::mbar_wait(&kv_empty[stage], phase ^ 1);
          unsigned char* dst = sKV + stage * 2 * TILE_BYTES;
          ptx::mbar_arrive_expect_tx(&kv_full[stage], 2 * TILE_BYTES);
          ptx::tma_load_3d(dst, &map_qkv, &kv_full[stage], col_k, t * BKV, it.b);
          ptx::tma_load_3d(dst + TILE_BYTES, &map_qkv, &kv_full[stage], col_v, t * BKV, it.b);
          if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 1) {
      // ===================== Q loader =====================
      uint32_t cnt[2] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item(item, npairs, a.H, N);
        for (int g = 0; g < 2; ++g) {
          if (g == 1 && !it.has_b) break;
          ptx::mbar_wait(&q_empty[g], (cnt[g] & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&q_full[g], TILE_BYTES);
          ptx::tma_load_3d(sQ + g * TILE_BYTES, &map_qkv, &q_full[g], it.h * HD, it.q0 + g * BQ, it.b);
          ++cnt[g];
        }
      }
    }
  } else if (warp > WARP_MMA) {
    ptx::setmaxnreg_dec<CTRL_REGS>();  // register donors
  } else if (warp == WARP_MMA) {
    ptx::setmaxnreg_dec<CTRL_REGS>();
    // ===================== MMA issuer =====================
    // Event loop over the two query tiles of the item: whichever group has delivered P_g(t) gets S_g(t+1) issued first
    // (it is on that group's critical path: its softmax warps are idle until it lands), then O_g += P_g(t) V(t).
    // State is packed (done counters 16 bits per group, parities one bit per group) and barriers / descriptors are 32-bit
    // shared addresses: the issuer thread lives in the 64 registers left after the softmax warps took theirs.
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BQ, HD) | (1u << 16);  // B operand (V) is MN-major
      constexpr uint32_t idesc_g = ptx::make_idesc_bf16(BQ, 16);
      const uint32_t lo_wg = ptx::sw128_desc_lo(smem_a + OFF_WG);
      const uint32_t bar_a = smem_a + OFF_BAR;
      constexpr uint32_t B_QFULL = 0, B_QEMPTY = 16, B_KVFULL = 32, B_KVEMPTY = 96, B_SFULL = 160, B_PFULL = 176, B_OFULL = 192, B_OFREE = 208;
      const uint32_t lo_q = ptx::sw128_desc_lo(smem_a + OFF_Q), lo_kv = ptx::sw128_desc_lo(smem_a + OFF_KV);
      uint32_t gt0 = 0;   // global index (over this CTA's items) of the item's first K/V tile
      uint32_t ipar = 0;  // bit g: parity of the items group g has processed
      uint32_t tpar = 0;  // bit g: parity of the tiles group g has processed
      auto issue_s = [&](int g, uint32_t gt) {
        const uint32_t st = gt % KV_STAGES;
        ptx::mbar_wait_a(bar_a + B_KVFULL + st * 8, (gt / KV_STAGES) & 1);  // returns at once if this phase was already observed
        ptx::tc_fence_after();
        const uint32_t dq = lo_q + g * (TILE_BYTES >> 4), dk = lo_kv + st * (2 * TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::umma_bf16(tmem_base + g * 128, ptx::sw128_desc_from_lo(dq + 2 * k), ptx::sw128_desc_from_lo(dk + 2 * k), idesc_s,
                         k != 0 ? 1u : 0u);
        ptx::umma_commit_a(bar_a + B_SFULL + g * 8);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ng = (item % npairs) * 2 * BQ + BQ < N ? 2 : 1;  // Item::has_b
        for (int g = 0; g < ng; ++g) {
          ptx::mbar_wait_a(bar_a + B_QFULL + g * 8, (ipar >> g) & 1);
          ptx::tc_fence_after();
          // gate logits of the 128 query rows: G_g = Q_g Wg^T (128 x 16 x 64) -> TMEM columns 384 + 16 g; the commit of
          // S_g(0) below covers them
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::umma_bf16(tmem_base + 384 + g * 16, ptx::sw128_desc_from_lo(lo_q + g * (TILE_BYTES >> 4) + 2 * k),
                           ptx::sw128_desc_from_lo(lo_wg + 2 * k), idesc_g, k != 0 ? 1u : 0u);
          issue_s(g, gt0);
          if (n_kv == 1) ptx::umma_commit_a(bar_a + B_QEMPTY + g * 8);
        }
        uint32_t done = ng == 2 ? 0u : (uint32_t)n_kv << 16;  // tiles finished: group 0 in the low half, group 1 in the high half
        while ((int)(done & 0xffffu) < n_kv || (int)(done >> 16) < n_kv) {
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            const int t = (done >> (16 * g)) & 0xffffu;
            if (t >= n_kv) continue;
            if (!ptx::mbar_try_wait_a(bar_a + B_PFULL + g * 8, (tpar >> g) & 1)) continue;  // P_g(t) stored, S_g(t) read out of TMEM
            if (t + 1 < n_kv) {
              // never block here: the stage of tile t+1 may only free up after the OTHER group's PV, issued by this thread
              const uint32_t gn = gt0 + t + 1;
              if (!ptx::mbar_try_wait_a(bar_a + B_KVFULL + (gn % KV_STAGES) * 8, (gn / KV_STAGES) & 1)) continue;
            }
            ptx::tc_fence_after();
            if (t == 0) {
              ptx::mbar_wait_a(bar_a + B_OFREE + g * 8, ((ipar >> g) & 1) ^ 1);  // previous item's O_g has been drained
              ptx::tc_fence_after();
            }
            // O_g += P_g(t) V(t).  A: P as bf16 pairs in TMEM, inside the columns of S_g that its writer owned (keys
            // 0..63 at columns 0..31, keys 64..127 at columns 64..95); B: V rows = keys (MN-major), 16 keys = 2048 bytes
            const uint32_t st = (gt0 + t) % KV_STAGES;
            const uint32_t dv = lo_kv + st * (2 * TILE_BYTES >> 4) + (TILE_BYTES >> 4);
#pragma unroll
            for (int ks = 0; ks < BKV / 16; ++ks)
              ptx::umma_bf16_ts(tmem_base + 256 + g * 64, tmem_base + g * 128 + (ks >> 2) * 64 + (ks & 3) * 8,
                                ptx::sw128_desc_from_lo(dv + ks * (2048 >> 4)), idesc_o, (t | ks) != 0 ? 1u : 0u);
            if (t + 1 == n_kv) ptx::umma_commit_a(bar_a + B_OFULL + g * 8);  // the item's O_g is complete (one phase per item)
            // S_g(t+1) overwrites S_g(t) / P_g(t): the tensor pipe executes it after the PV above (same issuing thread)
            if (t + 1 < n_kv) {
              issue_s(g, gt0 + t + 1);
              if (t + 2 == n_kv) ptx::umma_commit_a(bar_a + B_QEMPTY + g * 8);  // last S of the item: Q_g may be overwritten
            }
            done += 1u << (16 * g);
            tpar ^= 1u << g;
            // the stage of tile t is free once both groups' MMAs on it retire; the group that issues last commits
            if (ng == 1 || (int)((done >> (16 * (g ^ 1))) & 0xffffu) > t) ptx::umma_commit_a(bar_a + B_KVEMPTY + st * 8);
          }
        }
        gt0 += n_kv;
        ipar ^= ng == 2 ? 3u : 1u;
      }
    }
  } else {
    // ===================== softmax groups =====================
    ptx::setmaxnreg_inc<SOFTMAX_REGS>();
    const int g = warp >> 3;
    const int quarter = warp & 3, ch = (warp >> 2) & 1;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int stid = tid & (GROUP_THREADS - 1);
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 128, tmem_o = tmem_base + 256 + g * 64;
    const uint32_t tab_a = smem_a + OFF_TAB + g * TAB_BYTES;
    const uint32_t win_a = tab_a + TAB_WIN, mask_a = tab_a + TAB_MASK, pmax_a = tab_a + TAB_PMAX;
    const bool has_pad = a.key_pad != nullptr;
    const float qk_scale = 0.125f * LOG2E;  // head_dim^-0.5 (backbone.py:403), exp2 domain
    const int shift = (r + 1) & 3;          // which shifted window copy makes (c - r + 127 + shift) a multiple of 4
    const uint32_t win_row = win_a + (win_copy_offset(shift) + ch * 64 - r + 127 + shift) * 4;
    const int bar_id = 1 + g;
    // barrier addresses are derived from bar_a (32-bit shared address); par: bit 0 = item parity, bit 1 = tile parity
    const uint32_t bar_a = smem_a + OFF_BAR + g * 8;
    constexpr uint32_t B_SFULL = 160, B_PFULL = 176, B_OFULL = 192, B_OFREE = 208;
    uint32_t par = 0;

    // table entries of a tile: thread stid < 255 owns entry stid of the bias window, thread stid < 128 one mask entry
    auto fetch_bias = [&](const Item& it, int t) -> float {
      const int q0 = it.q0 + g * BQ;
      const int rel = t * BKV - q0 - 127 + stid;  // j - i
      return (stid < 255 && rel > -N && rel < N) ? __ldg(a.bias_vec + (it.h * (2 * N - 1) + (N - 1) + rel)) : 0.f;
    };
    auto fetch_dead = [&](const Item& it, int t) -> bool {
      const int j = t * BKV + stid;
      bool dead = j >= N;
      if (stid < BKV && !dead && has_pad) dead = a.key_pad[(size_t)it.b * N + j] != 0;
      return dead;
    };
    auto store_tables = [&](uint32_t slot, float bv, bool dead) {
      const uint32_t wbuf = win_a + slot * WIN_FLOATS * 4;
      if (stid < 255) {
        bv *= LOG2E;  // exp2 domain; scaled here so that the global load stays in flight across the tile
#pragma unroll
        for (int s = 0; s < 4; ++s) sts32(wbuf + (win_copy_offset(s) + stid + s) * 4, bv);
      }
      if (stid < BKV) sts32(mask_a + (slot * BKV + stid) * 4, dead ? -INFINITY : 0.f);
    };
    // The table slot of a tile is the parity of the group's running tile count (par bit 1), so slots alternate across item
    // boundaries too: the tables of the NEXT item's first tile are fetched and stored during the last tile of the current
    // one (have_tab) and published by the barrier of the finalisation.  Every barrier that publishes a tile's tables also
    // ORs its `dead` flags: a tile without a padded / out-of-range key runs the unmasked fast path even when a padding
    // mask was passed (the production pipeline always passes one; it is mostly false).
    bool have_tab = false;
    bool tile_dead = false;  // the tile about to be processed holds a dead key (group-uniform)

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode_item(item, npairs, a.H, N);
      if (g == 1 && !it.has_b) continue;
      const int q0 = it.q0 + g * BQ, b = it.b, h = it.h;
      if (!have_tab) {
        const bool dead0 = fetch_dead(it, 0);
        store_tables((par >> 1) & 1, fetch_bias(it, 0), dead0);
        tile_dead = named_bar_or(bar_id, GROUP_THREADS, dead0 && stid < BKV);  // tables of the first tile are visible
      }
      const float grep_a = __ldg(a.grep_a + h);  // 12 floats: an L1/L2 hit, consumed after the first S wait

      // first tile that holds a valid key (0 unless the clip starts with >= 128 padded keys): group-uniform
      int jc_first = 0;  // first valid key rounded down to its 16-key chunk: tile = jc_first / 128, chunk column = jc_first % 128
      if (has_pad && a.key_pad[(size_t)b * N] != 0) {  // the clip STARTS with padding: find its first valid key
        const uint8_t* kpad = a.key_pad + (size_t)b * N;
        int jmin = N;
        for (int j = stid; j < N; j += GROUP_THREADS)
          if (kpad[j] == 0) { jmin = j; break; }
        if (stid == 0) sts32(pmax_a + 1024 * 4, __int_as_float(N));
        named_bar_sync(bar_id, GROUP_THREADS);
        if (jmin < N) atomicMin(reinterpret_cast<int*>(smem + OFF_TAB + g * TAB_BYTES + TAB_PMAX) + 1024, jmin);
        named_bar_sync(bar_id, GROUP_THREADS);
        jc_first = __float_as_int(lds32(pmax_a + 1024 * 4)) & ~15;  // tile == n_kv when every key is padded: no estimate
      }

      have_tab = false;
      float gate = 0.f;  // set on the item's first tile

      // m_run: reference max of the row (log2 domain).  An estimate on tile t_first (score_est); afterwards it only moves
      // (between tiles) when a tile's max exceeds it by more than RESCALE_THRESHOLD -- O and l are rescaled by
      // `pending` at the start of the next tile.  The result is exact after the final 1/l for any reference.
      float m_run = -INFINITY, l_run = 0.f, pending = 1.0f;
      bool tab_dead = false;  // this thread's `dead` flag of the next item's first tile (stored on the last tile)
      for (int t = 0; t < n_kv; ++t) {
        const bool masked = tile_dead;
        const bool more = t + 1 < n_kv;
        float nbias = 0.f;
        bool ndead = false;
        if (more) {  // global loads in flight during the tile
          nbias = fetch_bias(it, t + 1);
          ndead = fetch_dead(it, t + 1);
        } else if (item + (int)gridDim.x < n_items) {  // decoded here, not held in registers across the tiles
          const Item nit = decode_item(item + gridDim.x, npairs, a.H, N);
          if (!(g == 1 && !nit.has_b)) {
            nbias = fetch_bias(nit, 0);
            ndead = fetch_dead(nit, 0);
            have_tab = true;
          }
        }
        const uint32_t slot = (par >> 1) & 1;
        const uint32_t taddr = tmem_s + lane_addr + ch * 64;
        const uint32_t wrow = win_row + slot * WIN_FLOATS * 4;
        const uint32_t mrow = mask_a + (slot * BKV + ch * 64) * 4;
        ptx::mbar_wait_a(bar_a + B_SFULL, slot);
        ptx::tc_fence_after();
        if (t == 0) {
          // gate of this query row from UNscaled q (backbone.py:544-550): the logits q.w_a, q.w_b were formed by the
          // tensor core next to S_g(0) (hi and lo halves of w in columns 0,1 and 2,3); both column halves of the row
          // read the same values: gate = sig_a * (sig_b * grep_a - 1) + 2
          uint32_t z[4];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(z[0]), "=r"(z[1]), "=r"(z[2]), "=r"(z[3])
                       : "r"(tmem_base + 384 + g * 16 + lane_addr)
                       : "memory");
          ptx::tmem_ld_wait();
          const float za = __uint_as_float(z[0]) + __uint_as_float(z[2]) + lds32(smem_a + OFF_GATEB);
          const float zb = __uint_as_float(z[1]) + __uint_as_float(z[3]) + lds32(smem_a + OFF_GATEB + 4);
          const float ga = 1.0f / (1.0f + __expf(-za)), gb = 1.0f / (1.0f + __expf(-zb));
          gate = ga * (gb * grep_a - 1.0f) + 2.0f;
        }
        if (t == jc_first / BKV) {
          // Reference estimate from a 16-key chunk that holds a valid key.  Both column halves of the row read the SAME
          // chunk, and the other half's warp may already be writing P over its first 32 columns: only chunks in columns
          // 32..63 / 96..127 of S (never covered by P) can be read by both; otherwise the owning half reads and shares.
          int c = jc_first % BKV;
          bool shared_est = false;
          if (!(c & 32)) {
            if (!masked || lds32(mask_a + (slot * BKV + 32) * 4) == 0.f) c = 32;
            else if (lds32(mask_a + (slot * BKV + 96) * 4) == 0.f) c = 96;
            else shared_est = true;
          }
          const uint32_t ta = tmem_s + lane_addr + c, wa = wrow + (c - ch * 64) * 4, ma = mask_a + (slot * BKV + c) * 4;
          if (!shared_est || ch == (c >> 6))
            m_run = masked ? score_est<true>(ta, wa, ma, gate, qk_scale) : score_est<false>(ta, wa, ma, gate, qk_scale);
          if (shared_est) {  // group-uniform (the mask is per key)
            if (ch == (c >> 6)) sts32(pmax_a + (768 + r) * 4, m_run);
            named_bar_sync(bar_id, GROUP_THREADS);
            m_run = lds32(pmax_a + (768 + r) * 4);
          }
        }
        // O is only touched here when the reference moved (rare).  No barrier is needed: S_g(t) was issued after PV_g(t-1)
        // and its commit (s_full, waited for above) covers every earlier MMA of the issuing thread, so O is consistent.
        if (t > 0 && __any_sync(0xffffffffu, pending != 1.0f)) {
          {
            l_run *= pending;
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {
              uint32_t o[16];
              const uint32_t oaddr = tmem_o + lane_addr + ch * 32 + hc * 16;
              tmem_ld_32x16(oaddr, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * pending);
              tmem_st_32x16(oaddr, o);
            }
            tmem_st_wait();
          }
        }
        const float m_eff = m_run == -INFINITY ? 0.f : m_run;  // -inf only while every key so far was masked (p = 0)
        float mx, sum;
        if (masked) stream_tile<true>(taddr, wrow, mrow, gate, qk_scale, m_eff, mx, sum);
        else stream_tile<false>(taddr, wrow, mrow, gate, qk_scale, m_eff, mx, sum);
        l_run += sum;
        ptx::tc_fence_before();  // P is in TMEM (tcgen05.wait::st done): order it before the arrive
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_a(bar_a + B_PFULL);
        par ^= 2;
        // off the critical path: join the tile max with the partner, decide the reference for the next tile
        pending = 1.0f;
        if (more) {
          sts32(pmax_a + ((slot * 2 + ch) * 128 + r) * 4, mx);
          store_tables(slot ^ 1, nbias, ndead);
          tile_dead = named_bar_or(bar_id, GROUP_THREADS, ndead && stid < BKV);
          const float m_tile = fmaxf(mx, lds32(pmax_a + ((slot * 2 + (ch ^ 1)) * 128 + r) * 4));
          if (m_run != -INFINITY && m_tile > m_run + RESCALE_THRESHOLD) {
            pending = ex2(m_run - m_tile);
            m_run = m_tile;
          }
        } else if (have_tab) {
          store_tables(slot ^ 1, nbias, ndead);  // first tile of the next item; published by the barrier below
          tab_dead = ndead;
        }
      }

      // ---- finalise: O / l -> bf16 -------------------------------------------------------------------------------
      sts32(pmax_a + (512 + ch * 128 + r) * 4, l_run);
      tile_dead = named_bar_or(bar_id, GROUP_THREADS, have_tab && tab_dead && stid < BKV);  // next item's first tile, if stored
      const float l_tot = l_run + lds32(pmax_a + (512 + (ch ^ 1) * 128 + r) * 4);
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      ptx::mbar_wait_a(bar_a + B_OFULL, par & 1);  // the item's last PV has retired
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld_32x32(tmem_o + lane_addr + ch * 32, o);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(bar_a + B_OFREE);  // O_g is in registers: the next item's first PV may overwrite it
      uint4 pk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pk[c].x = pack_bf16(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
        pk[c].y = pack_bf16(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
        pk[c].z = pack_bf16(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
        pk[c].w = pack_bf16(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
      }
      if (q0 + quarter * 32 + 31 < N) {
        // all 32 rows of this warp are inside the clip: stage the 32 x 64-byte box (64-byte swizzle) and hand it to the TMA
        // engine -- one asynchronous store of full lines instead of four 32-way scattered STG.128 per thread, whose
        // drain stalled the start of the next item
        if (lane == 0) ptx::tma_store_wait_read();  // the previous box has left the staging buffer
        __syncwarp();
        const uint32_t obuf = smem_a + OFF_OSTG + warp * 2048 + lane * 64, sw = (lane >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obuf + ((c ^ sw) << 4)), "r"(pk[c].x), "r"(pk[c].y),
                       "r"(pk[c].z), "r"(pk[c].w)
                       : "memory");
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d_s(&map_out, smem_a + OFF_OSTG + warp * 2048, h * HD + ch * 32, b * N + q0 + quarter * 32);
          ptx::tma_store_commit();
        }
      } else if (q0 + r < N) {  // ragged tail of the clip
        __nv_bfloat16* dst = a.out + ((size_t)b * N + q0 + r) * (size_t)(a.H * HD) + h * HD + ch * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + c * 8) = pk[c];
      }
      par ^= 1;
    }
    if (lane == 0) ptx::tma_store_wait_all();  // the staging buffer must outlive the last store
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace avexk

extern "C" int avexk_attention_gated(const void* qkv, int B, int N, int H, const float* gate_w, const float* gate_b,
                                     const float* grep_a, const float* bias_vec, const uint8_t* key_pad, void* out,
                                     void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(qkv && gate_w && gate_b && grep_a && bias_vec && out, "avexk_attention_gated: null argument");
  AVEXK_CHECK_ARG(B >= 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535 && (long long)H * (2LL * N - 1) < (1LL << 31),
                  "avexk_attention_gated: bad shape B=%d N=%d H=%d", B, N, H);
  AVEXK_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "avexk_attention_gated: qkv/out must be 16-byte aligned");
  if (B == 0) return AVEXK_OK;
  static bool attr_set[64] = {};  // per device: the opt-in is a per-device function attribute
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    AVEXK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[dev_] = true;
  }
  // qkv viewed as [B][N][3*H*64]: rows past the end of a clip are out of bounds -> zero-filled by TMA
  CUtensorMap map;
  const long long C3 = 3LL * H * HD;
  int rc = make_tmap_3d_bf16(&map, qkv, C3, N, B, C3, (long long)N * C3, HD, BQ, 1, true);
  if (rc) return rc;
  CUtensorMap map_out;  // out as [B*N, H*64] bf16, 32-row x 64-byte boxes (one softmax warp's share of an O tile)
  rc = make_tmap_2d_64B(&map_out, out, (long long)B * N, (long long)H * HD, (long long)H * HD, 2, 32);
  if (rc) return rc;
  const long long items = (long long)B * H * ceil_div(N, 2 * BQ);
  AVEXK_CHECK_ARG(items < (1LL << 31), "avexk_attention_gated: too many work items");
  AttnTcArgs a{B, N, H, (int)items, gate_w, gate_b, grep_a, bias_vec, key_pad, reinterpret_cast<__nv_bfloat16*>(out)};
  const int grid = (int)(items < num_sms() ? items : num_sms());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prof_begin(st, KID_ATTN, 4.0 * B * H * (double)N * N * HD);
  attention_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(map, map_out, a);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}
