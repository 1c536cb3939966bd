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
//   TMEM (512 columns): S_A @0, S_B @128 (fp32 128x128 each), O_A @256, O_B @320 (fp32 128x64 each), P_A @384, P_B @448
//                (bf16 pairs, 64 columns each).  P has columns of its own -- the per-row gate arrives precomputed (QKV epilogue),
//                which freed the columns its logits used to take -- so S_g(t+1) is issued BEFORE O_g += P_g(t) V(t) and the
//                softmax warps wait for one MMA batch per tile instead of two.
//   Shared memory: Q (2 tiles), a 5-stage K/V ring, the bias / mask tables -- the measured limiter before P moved to TMEM
//                was the shared-memory pipe (64 % busy: bias-window loads 31 %, P stores 11 %, MMA operand reads 21 %).
// Roofline: MUFU (one ex2 per score: B*H*N^2 per layer) and FP32 issue, not the tensor pipe; see DESIGN.md section 4.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"
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
constexpr int OFF_TAB = OFF_OSTG + 16 * 2048;
constexpr int TAB_WIN = 0;                               // [2][WIN_FLOATS] float
constexpr int TAB_MASK = TAB_WIN + 2 * WIN_FLOATS * 4;  // [2][128] float
constexpr int TAB_PMAX = TAB_MASK + 2 * 128 * 4;         // exchange buffers: [2 parity][2][128] tile max, [2][128] row sums l, [128] shared estimate
constexpr int TAB_BYTES = TAB_PMAX + 1024 * 4 + 16;  // + one int: first tile with a valid key
constexpr int OFF_BAR = OFF_TAB + 2 * TAB_BYTES;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "attention_tc: shared memory budget");

struct AttnTcArgs {
  int B, N, H;
  uint32_t m_pairs, m_heads;  // reciprocals for fast_div by ceil(N / 256) and H
  int n_items;             // B * H * ceil(N / 256): read from the constant bank where needed (a register-resident copy spilled)
  const float* gate;       // [B*N, H]: gate_a * (gate_b * grep_a - 1) + 2 of every (token, head) (backbone.py:544-550)
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

// A tile is PLAIN (every key valid), RAGGED (the clip ends inside it: keys >= nv are dead -- every clip's last tile unless N is a
// multiple of 128; handled by skipping dead 16-key chunks, no mask table) or MASKED (a padding mask kills keys inside it).
enum { TILE_PLAIN = 0, TILE_MASKED = 1, TILE_RAGGED = 2 };

// One 16-key chunk of a row: y = S * scale + (gate * bias - m_ref) (+ mask), two scores per packed fp32x2 instruction; tracks the
// chunk max on y, p = 2^min(y, P_CLAMP) -> bf16 pairs, row-sum contribution.  BOUNDARY (ragged tiles only): the chunk straddles
// the end of the clip, keys >= nv get -inf.  The window values of the NEXT chunk are fetched as the current ones are consumed.
template <int KIND, bool BOUNDARY>
__device__ __forceinline__ void score_chunk(const uint32_t (&R)[16], float4 (&w)[4], uint32_t next_win, bool fetch_next, uint32_t mask_chunk,
                                            float2 g2, float2 s2, float2 nm2, int c0, int nv, float& mx, float2& sum2, uint32_t (&pk)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 lo = __ffma2_rn(make_float2(__uint_as_float(R[4 * k + 0]), __uint_as_float(R[4 * k + 1])), s2,
                           __ffma2_rn(g2, make_float2(w[k].x, w[k].y), nm2));
    float2 hi = __ffma2_rn(make_float2(__uint_as_float(R[4 * k + 2]), __uint_as_float(R[4 * k + 3])), s2,
                           __ffma2_rn(g2, make_float2(w[k].z, w[k].w), nm2));
    if (fetch_next) w[k] = lds128(next_win + k * 16);
    if (KIND == TILE_MASKED) {
      const float4 mk = lds128(mask_chunk + k * 16);
      lo = __fadd2_rn(lo, make_float2(mk.x, mk.y));
      hi = __fadd2_rn(hi, make_float2(mk.z, mk.w));
    }
    if (BOUNDARY) {
      lo.x = c0 + 4 * k + 0 < nv ? lo.x : -INFINITY;
      lo.y = c0 + 4 * k + 1 < nv ? lo.y : -INFINITY;
      hi.x = c0 + 4 * k + 2 < nv ? hi.x : -INFINITY;
      hi.y = c0 + 4 * k + 3 < nv ? hi.y : -INFINITY;
    }
    mx = fmaxf(fmaxf(mx, fmaxf(lo.x, lo.y)), fmaxf(hi.x, hi.y));
    // the reference max may lag the true max (it moves between tiles): clamp so bf16 P cannot overflow
    const float2 p0 = make_float2(ex2(fminf(lo.x, P_CLAMP)), ex2(fminf(lo.y, P_CLAMP)));
    const float2 p1 = make_float2(ex2(fminf(hi.x, P_CLAMP)), ex2(fminf(hi.y, P_CLAMP)));
    sum2 = __fadd2_rn(sum2, __fadd2_rn(p0, p1));
    pk[2 * k] = pack_bf16(p0.x, p0.y);
    pk[2 * k + 1] = pack_bf16(p1.x, p1.y);
  }
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

// Stream the 64 scores once: p = 2^(min(x - m, P_CLAMP)) -> bf16 pairs -> into the group's P columns of TMEM (chunk c of 16
// fp32 scores becomes 8 packed columns at 8c of the thread's half), where the PV MMA takes them as its A operand -- P never
// touches shared memory.  Accumulates the row sum and tracks the tile max (used to move the reference max for the NEXT tile).
// Nothing is kept in registers.
template <int KIND>
__device__ __forceinline__ void stream_tile(uint32_t taddr, uint32_t paddr, uint32_t pv_bar, uint32_t pv_parity, uint32_t sfree_bar,
                                            int lane, uint32_t win_addr, uint32_t mask_addr, float gate, float qk_scale, float m_ref,
                                            int col0, int nv, float& mx_out, float& sum_out) {
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
    const int c0 = col0 + chunk * 16;  // first key of this chunk inside the tile
    if (chunk < 3) tmem_ld_32x16(taddr + (chunk + 1) * 16, (chunk & 1) ? ra : rb);
    uint32_t pk[8];
    const uint32_t nwin = win_addr + (chunk + 1) * 64, mchunk = mask_addr + chunk * 64;
    const bool fetch = chunk < 3;
    // warp-uniform three-way split on ragged tiles: chunk past the end of the clip (P = 0, no arithmetic), chunk straddling it
    // (per-key predicate), chunk inside it (the plain code)
    if (KIND == TILE_RAGGED && c0 >= nv) {
#pragma unroll
      for (int k = 0; k < 8; ++k) pk[k] = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (fetch) w[k] = lds128(nwin + k * 16);
    } else if (KIND == TILE_RAGGED && c0 + 16 > nv) {
      if (chunk & 1) score_chunk<KIND, true>(rb, w, nwin, fetch, mchunk, g2, s2, nm2, c0, nv, mx, sum2, pk);
      else score_chunk<KIND, true>(ra, w, nwin, fetch, mchunk, g2, s2, nm2, c0, nv, mx, sum2, pk);
    } else {
      if (chunk & 1) score_chunk<KIND, false>(rb, w, nwin, fetch, mchunk, g2, s2, nm2, c0, nv, mx, sum2, pk);
      else score_chunk<KIND, false>(ra, w, nwin, fetch, mchunk, g2, s2, nm2, c0, nv, mx, sum2, pk);
    }
    if (chunk < 3) ptx::tmem_ld_wait();
    if (chunk == 2) {
      // every score of S_g(t) is in registers (chunk 3 landed with the wait above): the issuer may overwrite S_g with S_g(t+1)
      // while this warp still works on its last chunk -- a quarter of the tile's softmax time off the MMA round trip
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(sfree_bar);
    }
    if (chunk == 0) {  // P_g(t-1) must have been consumed by its PV before it is overwritten (long retired by now, as a rule)
      ptx::mbar_wait_a(pv_bar, pv_parity);
      ptx::tc_fence_after();
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(paddr + chunk * 8),
                 "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                 : "memory");
  }
  tmem_st_wait();
  mx_out = mx + m_ref;  // back to the un-shifted domain: the caller compares tile maxima with the reference
  sum_out = sum2.x + sum2.y;
}

struct Item {
  int b, h, q0;
  bool has_b;  // the pair's second tile holds at least one valid query row
};
// n / d through a host-computed reciprocal (magic = floor(2^32 / d) + 1, exact for n * d < 2^32): five instructions instead of
// the ~25 of a runtime integer division, which the softmax warps ended up repeating inside their tile loop
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t d, uint32_t magic) {
  if (d == 1) return n;  // the reciprocal of 1 does not fit 32 bits
  uint32_t q = __umulhi(n, magic);
  if (q * d > n) --q;  // (magic rounds up: the quotient can be one too large)
  return q;
}
__device__ __forceinline__ Item decode_item(int item, int npairs, int H, int N, uint32_t m_pairs, uint32_t m_heads) {
  Item it;
  const uint32_t bh = fast_div(item, npairs, m_pairs), pair = item - bh * npairs;
  it.b = fast_div(bh, H, m_heads);
  it.h = bh - it.b * H;
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
  uint64_t* o_full = bars + 24;    // [2]  PV_g(t) retired (one phase per TILE): P_g may be rewritten, O_g is consistent
  uint64_t* o_free = bars + 26;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  uint64_t* s_free = bars + 30;    // [2]  S_g(t) has been read out of TMEM by all of the group's warps (one phase per tile)
  static_assert(KV_STAGES <= 8, "barrier layout");

  int tid;  // read once through a volatile asm: otherwise the compiler re-reads %tid.x (S2R, behind the MUFU queue) per item
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5, lane = tid & 31;
  const int N = a.N;
  const int n_kv = (N + BKV - 1) / BKV;
  const int npairs = (N + 2 * BQ - 1) / (2 * BQ);
  const int n_items = a.n_items;

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
        ptx::mbar_init(&s_free[g], GROUP_WARPS);
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
        const Item it = decode_item(item, npairs, a.H, N, a.m_pairs, a.m_heads);
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
        const Item it = decode_item(item, npairs, a.H, N, a.m_pairs, a.m_heads);
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
      const uint32_t bar_a = smem_a + OFF_BAR;
      constexpr uint32_t B_QFULL = 0, B_QEMPTY = 16, B_KVFULL = 32, B_KVEMPTY = 96, B_SFULL = 160, B_PFULL = 176, B_OFULL = 192, B_OFREE = 208,
                         B_SFREE = 240;
      const uint32_t lo_q = ptx::sw128_desc_lo(smem_a + OFF_Q), lo_kv = ptx::sw128_desc_lo(smem_a + OFF_KV);
      auto issue_s = [&](int g, uint32_t gt) {  // the caller has seen kv_full of the stage
        const uint32_t st = gt % KV_STAGES;
        const uint32_t dq = lo_q + g * (TILE_BYTES >> 4), dk = lo_kv + st * (2 * TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::umma_bf16(tmem_base + g * 128, ptx::sw128_desc_from_lo(dq + 2 * k), ptx::sw128_desc_from_lo(dk + 2 * k), idesc_s,
                         k != 0 ? 1u : 0u);
        ptx::umma_commit_a(bar_a + B_SFULL + g * 8);
      };
      // Two independent state machines, one per query-tile group, over this CTA's item list (ordinal k <-> item blockIdx.x + k *
      // gridDim.x; K/V tile t of item k has the global index k * n_kv + t, which names its ring stage and barrier phase).  A group
      // only waits for ITS OWN softmax warps: it starts the next item while the other group is still finishing the current one
      // (the first version walked the items in lock step, which cost every group an MMA round trip plus the other group's
      // remaining tile time at every item boundary).  The K/V ring couples them loosely: a stage is released by whichever
      // group's PV on it comes last.
      const int n_ord = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA
      auto has_group = [&](int g, int k) -> bool {  // group 1 only exists when the pair's second tile holds a valid query row
        if (g == 0) return true;
        const uint32_t item = blockIdx.x + k * gridDim.x;
        return (int)(item - fast_div(item, npairs, a.m_pairs) * npairs) * 2 * BQ + BQ < N;
      };
      // positions are packed (item ordinal << 8 | tile): one register each, and "is ahead of" is an integer compare
      uint32_t ps[2] = {0, 0};  // next S_g
      uint32_t pp[2] = {0, 0};  // next PV_g
      const uint32_t pend = (uint32_t)n_ord << 8;
      auto advance = [&](int g, uint32_t pos) -> uint32_t {  // next tile; past the item's last tile: first tile of g's next item
        if ((int)(pos & 0xffu) + 1 < n_kv) return pos + 1;
        uint32_t k = (pos >> 8) + 1;
        while ((int)k < n_ord && !has_group(g, (int)k)) ++k;
        return k << 8;
      };
      uint32_t qpar = 0, sfpar = 0, tpar = 0, opar = 0, sany = 0;  // bit g: parities of q_full / s_free / p_full / o_free; S ever issued
      if (n_ord > 0 && !has_group(1, 0)) ps[1] = pp[1] = advance(1, (uint32_t)(n_kv - 1));
      while (pp[0] < pend || pp[1] < pend) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // ---- S_g(k, t): Q_g of the item landed (t = 0), the previous S of this group read out of TMEM, K(k, t) landed ----
          if (ps[g] < pend) {
            const uint32_t t = ps[g] & 0xffu, gt = (ps[g] >> 8) * n_kv + t;
            bool ok = !((sany >> g) & 1) || ptx::mbar_try_wait_a(bar_a + B_SFREE + g * 8, (sfpar >> g) & 1);
            if (ok && t == 0) ok = ptx::mbar_try_wait_a(bar_a + B_QFULL + g * 8, (qpar >> g) & 1);
            if (ok) ok = ptx::mbar_try_wait_a(bar_a + B_KVFULL + (gt % KV_STAGES) * 8, (gt / KV_STAGES) & 1);
            if (ok) {
              ptx::tc_fence_after();
              issue_s(g, gt);
              if ((sany >> g) & 1) sfpar ^= 1u << g;
              sany |= 1u << g;
              if (t == 0) qpar ^= 1u << g;
              if ((int)t + 1 == n_kv) ptx::umma_commit_a(bar_a + B_QEMPTY + g * 8);  // last S of the item: Q_g may be overwritten
              ps[g] = advance(g, ps[g]);
            }
          }
          // ---- O_g += P_g(k, t) V(k, t): P stored; first tile of an item: the previous item's O_g drained ----
          if (pp[g] < pend && ptx::mbar_try_wait_a(bar_a + B_PFULL + g * 8, (tpar >> g) & 1)) {
            ptx::tc_fence_after();
            const uint32_t pos = pp[g], t = pos & 0xffu;
            if (t == 0) {
              ptx::mbar_wait_a(bar_a + B_OFREE + g * 8, ((opar >> g) & 1) ^ 1);
              ptx::tc_fence_after();
              opar ^= 1u << g;
            }
            // A: P as bf16 pairs in TMEM (keys 16 ks .. 16 ks + 15 at columns 8 ks of P_g); B: V rows = keys (MN-major), 16 keys = 2048 B
            const uint32_t st = ((pos >> 8) * n_kv + t) % KV_STAGES;
            const uint32_t dv = lo_kv + st * (2 * TILE_BYTES >> 4) + (TILE_BYTES >> 4);
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk)
              ptx::umma_bf16_ts(tmem_base + 256 + g * 64, tmem_base + 384 + g * 64 + kk * 8,
                                ptx::sw128_desc_from_lo(dv + kk * (2048 >> 4)), idesc_o, (t | kk) != 0 ? 1u : 0u);
            ptx::umma_commit_a(bar_a + B_OFULL + g * 8);  // PV_g retired: P_g may be rewritten / O_g rescaled / (last tile) read
            tpar ^= 1u << g;
            pp[g] = advance(g, pos);
            // the stage is free once every group that works on the item has issued its PV on it: the one that comes last commits
            if (!has_group(g ^ 1, (int)(pos >> 8)) || pp[g ^ 1] > pos) ptx::umma_commit_a(bar_a + B_KVEMPTY + st * 8);
          }
        }
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
    const uint32_t tmem_s = tmem_base + g * 128, tmem_o = tmem_base + 256 + g * 64, tmem_p = tmem_base + 384 + g * 64;
    const uint32_t tab_a = smem_a + OFF_TAB + g * TAB_BYTES;
    const uint32_t win_a = tab_a + TAB_WIN, mask_a = tab_a + TAB_MASK, pmax_a = tab_a + TAB_PMAX;
    const bool has_pad = a.key_pad != nullptr;
    const float qk_scale = 0.125f * LOG2E;  // head_dim^-0.5 (backbone.py:403), exp2 domain
    const int shift = (r + 1) & 3;          // which shifted window copy makes (c - r + 127 + shift) a multiple of 4
    const uint32_t win_row = win_a + (win_copy_offset(shift) + ch * 64 - r + 127 + shift) * 4;
    const int bar_id = 1 + g;
    // barrier addresses are derived from bar_a (32-bit shared address); par: bit 0 = item parity, bit 1 = tile parity
    const uint32_t bar_a = smem_a + OFF_BAR + g * 8;
    constexpr uint32_t B_SFULL = 160, B_PFULL = 176, B_OFULL = 192, B_OFREE = 208, B_SFREE = 240;
    uint32_t par = 0;

    // table entries of a tile: thread stid < 255 owns entry stid of the bias window, thread stid < 128 one mask entry
    auto fetch_bias = [&](const Item& it, int t) -> float {
      const int q0 = it.q0 + g * BQ;
      const int rel = t * BKV - q0 - 127 + stid;  // j - i
      return (stid < 255 && rel > -N && rel < N) ? __ldg(a.bias_vec + (it.h * (2 * N - 1) + (N - 1) + rel)) : 0.f;
    };
    // bit 0: the key is dead (past the end of the clip, or padded) -> -inf in the mask table; bit 1: it is dead because of the
    // padding mask -- only that forces the masked path, the end of the clip is handled by the ragged fast path
    auto fetch_dead = [&](const Item& it, int t) -> int {
      const int j = t * BKV + stid;
      int dead = j >= N ? 1 : 0;
      if (stid < BKV && !dead && has_pad && a.key_pad[(size_t)it.b * N + j] != 0) dead = 3;
      return dead;
    };
    auto store_tables = [&](uint32_t slot, float bv, int dead) {
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
    bool tile_dead = false;  // the tile about to be processed holds a PADDED key (group-uniform): masked path

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      Item it = decode_item(item, npairs, a.H, N, a.m_pairs, a.m_heads);
      if (g == 1 && !it.has_b) continue;
      // pin the decoded fields: under register pressure the compiler otherwise re-derives them (two integer divisions, ~60
      // instructions) at every use inside the tile loop
      asm volatile("" : "+r"(it.b), "+r"(it.h), "+r"(it.q0));
      const int q0 = it.q0 + g * BQ, b = it.b, h = it.h;
      if (!have_tab) {
        const int dead0 = fetch_dead(it, 0);
        store_tables((par >> 1) & 1, fetch_bias(it, 0), dead0);
        tile_dead = named_bar_or(bar_id, GROUP_THREADS, (dead0 & 2) && stid < BKV);  // tables of the first tile are visible
      }
      // gate of this query row (from UNscaled q, backbone.py:544-550), precomputed by the QKV epilogue: in flight until S(0) lands
      const float gate = q0 + r < N ? __ldg(a.gate + ((size_t)b * N + q0 + r) * a.H + h) : 0.f;

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

      // m_run: reference max of the row (log2 domain).  An estimate on tile t_first (score_est); afterwards it only moves
      // (between tiles) when a tile's max exceeds it by more than RESCALE_THRESHOLD -- O and l are rescaled by
      // `pending` at the start of the next tile.  The result is exact after the final 1/l for any reference.
      float m_run = -INFINITY, l_run = 0.f, pending = 1.0f;
      int tab_dead = 0;  // this thread's `dead` flags of the next item's first tile (stored on the last tile)
      for (int t = 0; t < n_kv; ++t) {
        const bool masked = tile_dead;
        const int nv = N - t * BKV;                  // valid keys of this tile when it is the clip's last
        const bool ragged = !masked && nv < BKV;
        const bool more = t + 1 < n_kv;
        float nbias = 0.f;
        int ndead = 0;
        if (more) {  // global loads in flight during the tile
          nbias = fetch_bias(it, t + 1);
          ndead = fetch_dead(it, t + 1);
        } else if (item + (int)gridDim.x < n_items) {  // decoded here, not held in registers across the tiles
          const Item nit = decode_item(item + gridDim.x, npairs, a.H, N, a.m_pairs, a.m_heads);
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
        if (t == jc_first / BKV) {
          // Reference estimate from the 16-key chunk that holds the clip's first valid key.  Both column halves of the row read
          // the SAME chunk of S (nothing overwrites S_g(t) before both have delivered P_g(t)), so they agree without an exchange.
          const int c = jc_first % BKV;
          const bool est_masked = masked || ragged;  // the mask table also carries -inf for keys past the end of the clip
          const uint32_t ta = tmem_s + lane_addr + c, wa = wrow + (c - ch * 64) * 4, ma = mask_a + (slot * BKV + c) * 4;
          m_run = est_masked ? score_est<true>(ta, wa, ma, gate, qk_scale) : score_est<false>(ta, wa, ma, gate, qk_scale);
        }
        // O is only touched here when the reference moved (rare): PV_g(t-1) must have retired first (S_g(t) was issued ahead
        // of it, so s_full does not cover it)
        if (t > 0 && __any_sync(0xffffffffu, pending != 1.0f)) {
          ptx::mbar_wait_a(bar_a + B_OFULL, slot ^ 1);
          ptx::tc_fence_after();
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
        const uint32_t paddr = tmem_p + lane_addr + ch * 32, pvb = bar_a + B_OFULL;
        const uint32_t sfb = bar_a + B_SFREE;
        if (masked) stream_tile<TILE_MASKED>(taddr, paddr, pvb, slot ^ 1, sfb, lane, wrow, mrow, gate, qk_scale, m_eff, ch * 64, nv, mx, sum);
        else if (ragged) stream_tile<TILE_RAGGED>(taddr, paddr, pvb, slot ^ 1, sfb, lane, wrow, mrow, gate, qk_scale, m_eff, ch * 64, nv, mx, sum);
        else stream_tile<TILE_PLAIN>(taddr, paddr, pvb, slot ^ 1, sfb, lane, wrow, mrow, gate, qk_scale, m_eff, ch * 64, nv, mx, sum);
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
          tile_dead = named_bar_or(bar_id, GROUP_THREADS, (ndead & 2) && stid < BKV);
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
      tile_dead = named_bar_or(bar_id, GROUP_THREADS, have_tab && (tab_dead & 2) && stid < BKV);  // next item's first tile, if stored
      const float l_tot = l_run + lds32(pmax_a + (512 + (ch ^ 1) * 128 + r) * 4);
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      ptx::mbar_wait_a(bar_a + B_OFULL, ((par >> 1) & 1) ^ 1);  // the item's last PV has retired (one phase per tile)
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

// gate[(b, n), h] = sig_a * (sig_b * grep_a[h] - 1) + 2 with (sig_a, sig_b) = sigmoid(q . gate_w[0|1] + gate_b[0|1]) from the bf16 q of
// a qkv buffer: the stand-alone form of what the QKV GEMM epilogue computes from its fp32 accumulators inside avexk_beats_forward.
__global__ void __launch_bounds__(256)
gate_from_qkv_kernel(const __nv_bfloat16* __restrict__ qkv, long long M, int H, const float* __restrict__ gate_w,
                     const float* __restrict__ gate_b, const float* __restrict__ grep_a, float* __restrict__ gate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, head)
  if (idx >= M * H) return;
  const int h = idx % H;
  const long long row = idx / H;
  const uint4* q = reinterpret_cast<const uint4*>(qkv + row * (3LL * H * HD) + h * HD);
  float za = gate_b[0], zb = gate_b[1];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const uint4 u = __ldg(q + i);
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]));
      const int c = i * 8 + j * 2;
      za = fmaf(f.x, gate_w[c], za);
      za = fmaf(f.y, gate_w[c + 1], za);
      zb = fmaf(f.x, gate_w[HD + c], zb);
      zb = fmaf(f.y, gate_w[HD + c + 1], zb);
    }
  }
  const float ga = 1.0f / (1.0f + __expf(-za)), gb = 1.0f / (1.0f + __expf(-zb));
  gate[idx] = ga * (gb * grep_a[h] - 1.0f) + 2.0f;
}

}  // namespace

int launch_gate_from_qkv(const void* qkv, long long M, int H, const float* gate_w, const float* gate_b, const float* grep_a,
                         float* gate, cudaStream_t st) {
  if (M == 0) return AVEXK_OK;
  gate_from_qkv_kernel<<<ceil_div(M * H, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), M, H, gate_w, gate_b, grep_a, gate);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

// qkv [B*N, 3*H*64] bf16, gate [B*N, H] fp32 (per-row gate of the relative-position bias), out [B*N, H*64] bf16
int attention_launch(const void* qkv, int B, int N, int H, const float* gate, const float* bias_vec, const uint8_t* key_pad, void* out,
                     cudaStream_t st) {
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
  AVEXK_CHECK_ARG(items < (1LL << 31), "attention: too many work items");
  const uint32_t npairs = (uint32_t)ceil_div(N, 2 * BQ);
  AVEXK_CHECK_ARG(items * (long long)(npairs > (uint32_t)H ? npairs : (uint32_t)H) < (1LL << 32), "attention: too many work items for the reciprocal decode");
  AttnTcArgs a{B, N, H, (uint32_t)(0x100000000ULL / npairs) + 1u, (uint32_t)(0x100000000ULL / (uint32_t)H) + 1u, (int)items, gate, bias_vec, key_pad,
               reinterpret_cast<__nv_bfloat16*>(out)};
  const int grid = (int)(items < num_sms() ? items : num_sms());
  prof_begin(st, KID_ATTN, 4.0 * B * H * (double)N * N * HD);
  attention_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(map, map_out, a);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // building-block entry: the per-row gates get a stream-ordered temporary (avexk_beats_forward keeps them in its workspace,
  // written by the QKV GEMM epilogue)
  float* gate = nullptr;
  AVEXK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&gate), (size_t)B * N * H * sizeof(float), st));
  int rc = launch_gate_from_qkv(qkv, (long long)B * N, H, gate_w, gate_b, grep_a, gate, st);
  if (rc == AVEXK_OK) rc = attention_launch(qkv, B, N, H, gate, bias_vec, key_pad, out, st);
  cudaFreeAsync(gate, st);
  return rc;
}
