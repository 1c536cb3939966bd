// pointwise.cu -- the 1x1 convolutions of EfficientNet as a memory-bound tcgen05 kernel.
//
// out[M, N] = act((se[row / hw, :] * A[M, K]) @ W[N, K]^T * scale + shift) (+ res), A / W / res / out fp16, fp32 accumulation.
// Replaces torchvision `Conv2dNormActivation(kernel_size=1)` + BatchNorm(eval) + SiLU inside `MBConv` and, for the project
// convolution, the squeeze-excitation rescale `scale * input` that precedes it (called from efficientnet.py:163-215).
//
// These GEMMs have K and N in 16..1152 and M up to 8.2 M rows: 0.3 .. 30 FLOP per byte, i.e. HBM-bound.  The general GEMM of
// gemm_tc.cu (256-wide tiles, two accumulator stages, sixteen epilogue warps polling) spends ~6600 cycles per 256-row tile on
// them; this kernel is built around bytes in flight and instructions per output element instead:
//   * persistent CTA, 128-row tiles; one TMA producer thread keeps a ring of A k-blocks (and W k-blocks when W does not fit)
//     in flight; W stays resident in shared memory when it fits (every layer of the high-resolution stages);
//   * one MMA thread issues kind::f16 UMMAs (M = 128, N = the layer's N rounded to 16) into a ring of up to eight TMEM
//     accumulators, so the epilogue of tile t overlaps the loads and MMAs of tiles t+1 .. t+7;
//   * eight or sixteen epilogue warps (two or four per TMEM lane quarter, taking tiles in turn) each own 32 full rows of a
//     tile: BatchNorm / SiLU / residual with packed fp32x2 arithmetic, fp16 into the warp's swizzled staging boxes, one TMA
//     tensor store per [32 x 64] box: full-line writes, no barrier among the epilogue warps;
//   * squeeze-excitation: eight "scaler" warps multiply the A tile by the clip's channel scale in shared memory before the
//     MMA reads it -- the separate read-modify-write pass over the depthwise output (se_apply_kernel) disappears.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int PW_BM = 128, PW_BK = 64, PW_A_BYTES = PW_BM * PW_BK * 2;
constexpr int PW_MAX_STAGES = 8, PW_MAX_ACC = 8;
constexpr int PW_BOX_BYTES = 32 * 128;  // one store box: 32 rows x 64 fp16 columns
constexpr int PW_WARP_LOAD = 0, PW_WARP_MMA = 1, PW_WARP_SCALE0 = 2, PW_WARP_SE_EPI0 = 10, PW_WARPS = 18, PW_THREADS = PW_WARPS * 32;
constexpr int PW_EPI_SLOTS = PW_WARPS - PW_WARP_SCALE0;  // warps that may run the epilogue (sixteen without SE, eight with)

struct PwArgs {
  int M, N, K;
  int ntb, n_nt;           // N tile (multiple of 16, <= 256) and number of N tiles
  int kblocks;             // ceil(K / 64)
  int stages, stage_bytes; // ring depth; bytes per stage (A block, + W block when W is streamed)
  int w_resident, w_block_bytes;
  int acc_stride, acc_stages;
  int direct;              // rows under 64 columns: plain 16-byte stores instead of staged TMA stores
  int off_w, off_ob, off_tab, off_bar;
  uint32_t idesc;
  int m_tiles, silu, hw;
  const float *scale, *shift, *se_scale;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
  float* raw;  // optional fp32 copy of the accumulators (the pre-BatchNorm tensor a forward hook on the conv sees)
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// silu for a pair: one packed multiply, two EX2, one packed add, two RCP, one packed multiply
__device__ __forceinline__ float2 pw_silu2(float2 v) {
  const float2 q = __fmul2_rn(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  e = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
  return __fmul2_rn(v, r);
}

// wait with back-off for the two single-thread roles (producer, MMA issuer): their rings are several tiles deep, so a few hundred
// nanoseconds of wake-up latency cost nothing, while a tight poll loop takes issue slots from the epilogue warps of its sub-core
__device__ __forceinline__ void pw_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 24)) __trap();
  }
}

__global__ void __launch_bounds__(PW_THREADS, 1)
pointwise_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const PwArgs g) {
  extern __shared__ unsigned char pw_smem_raw[];
  unsigned char* smem = pw_smem_raw + ((1024u - (ptx::smem_u32(pw_smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_a = ptx::smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bar);
  uint64_t* a_full = bars;                          // [PW_MAX_STAGES]  TMA landed
  uint64_t* a_ready = bars + PW_MAX_STAGES;         // [PW_MAX_STAGES]  scaled (SE mode): what the MMA thread waits on
  uint64_t* a_empty = bars + 2 * PW_MAX_STAGES;     // [PW_MAX_STAGES]  MMAs reading the stage have completed
  uint64_t* t_full = bars + 3 * PW_MAX_STAGES;      // [PW_MAX_ACC]
  uint64_t* t_empty = t_full + PW_MAX_ACC;          // [PW_MAX_ACC]
  uint64_t* w_full = t_empty + PW_MAX_ACC;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* tab = reinterpret_cast<float*>(smem + g.off_tab);  // scale[N], shift[N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool se = g.se_scale != nullptr;
  const int n_tiles = g.m_tiles * g.n_nt;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
    ptx::prefetch_tensormap(&map_out);
    for (int s = 0; s < g.stages; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_ready[s], PW_WARP_SE_EPI0 - PW_WARP_SCALE0);  // one arrival per scaler warp
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < g.acc_stages; ++s) {
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], 4);  // one warp per TMEM lane quarter
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < g.N; i += blockDim.x) {
    tab[i] = g.scale != nullptr ? __ldg(g.scale + i) : 1.0f;
    tab[g.N + i] = g.shift != nullptr ? __ldg(g.shift + i) : 0.0f;
  }
  if (warp == PW_WARP_MMA) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == PW_WARP_LOAD) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (g.w_resident) {
        ptx::mbar_arrive_expect_tx(w_full, g.kblocks * g.w_block_bytes);
        for (int kb = 0; kb < g.kblocks; ++kb) ptx::tma_load_2d(smem + g.off_w + kb * g.w_block_bytes, &map_w, w_full, kb * PW_BK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = g.n_nt == 1 ? tile : tile / g.n_nt, nt = tile - mt * g.n_nt;  // (no division on the common single-N-tile path)
        for (int kb = 0; kb < g.kblocks; ++kb) {
          pw_wait_relaxed(&a_empty[stage], phase ^ 1);
          unsigned char* dst = smem + stage * g.stage_bytes;
          ptx::mbar_arrive_expect_tx(&a_full[stage], g.stage_bytes);
          ptx::tma_load_2d(dst, &map_a, &a_full[stage], kb * PW_BK, mt * PW_BM);
          if (!g.w_resident) ptx::tma_load_2d(dst + PW_A_BYTES, &map_w, &a_full[stage], kb * PW_BK, nt * g.ntb);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == PW_WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (g.w_resident) {
        ptx::mbar_wait(w_full, 0);
        ptx::tc_fence_after();
      }
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        pw_wait_relaxed(&t_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * g.acc_stride;
        for (int kb = 0; kb < g.kblocks; ++kb) {
          pw_wait_relaxed(se ? &a_ready[stage] : &a_full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_a + stage * g.stage_bytes;
          const uint32_t sb = g.w_resident ? smem_a + g.off_w + kb * g.w_block_bytes : sa + PW_A_BYTES;
          const uint64_t da = ptx::make_sw128_desc(sa), db = ptx::make_sw128_desc(sb);
          const int ksteps = min(PW_BK, g.K - kb * PW_BK + 15) / 16;  // the zero-filled tail of the last k-block is skipped
          for (int k = 0; k < ksteps; ++k) ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, g.idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&a_empty[stage]);
          if (kb == g.kblocks - 1) ptx::umma_commit(&t_full[acc]);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == g.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (se && warp < PW_WARP_SE_EPI0 && g.hw >= PW_BM && g.n_nt == 1) {
    // ===================== squeeze-excitation scalers: A[row, :] *= se[clip(row), :] in shared memory =====================
    // (clips of at least 128 rows: the high-resolution layers, where this path matters)
    // A thread owns ONE logical 8-column chunk j of the k-block for four rows 32 apart: it needs 8 scale values per clip and
    // stage (32 bytes) instead of 32 per row, so the values of the NEXT stage -- next k-block, or the first k-block of the
    // CTA's next tile -- fit in registers and are fetched one stage ahead: their L2 latency (the scale table was just written
    // by the SE kernel) hides behind the current stage instead of stalling it (36 % of the scaler's time in ncu).  A tile of
    // 128 rows meets at most one clip boundary (hw >= 128): rows past it use the second value set.  The 8 lanes of a
    // quarter-warp read the 8 chunks of one row: conflict-free.
    const int t = threadIdx.x - PW_WARP_SCALE0 * 32, j = t & 7, rbase = t >> 3;  // rows rbase + 32 i
    uint32_t roff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = rbase + 32 * i;
      roff[i] = row * 128 + ((j ^ (row & 7)) << 4);  // physical position of logical chunk j under the 128-byte swizzle
    }
    const int n_clips = (g.M + g.hw - 1) / g.hw;
    const int step_rows = gridDim.x * PW_BM, dq = step_rows / g.hw, dr = step_rows - dq * g.hw;
    int grow0 = blockIdx.x * PW_BM + rbase, clip = grow0 / g.hw, rem = grow0 - clip * g.hw;
    auto load_scales = [&](int clipA, int remA, int kb, float4 (&a)[2], float4 (&b)[2]) {
      const int col = kb * PW_BK + j * 8;
      if (col >= g.K || clipA >= n_clips) return;
      const float* sp = g.se_scale + (size_t)clipA * g.K + col;
      a[0] = __ldg(reinterpret_cast<const float4*>(sp));
      a[1] = __ldg(reinterpret_cast<const float4*>(sp + 4));
      if (remA + 96 >= g.hw && clipA + 1 < n_clips) {  // one of this thread's rows lies in the next clip
        b[0] = __ldg(reinterpret_cast<const float4*>(sp + g.K));
        b[1] = __ldg(reinterpret_cast<const float4*>(sp + g.K + 4));
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    float4 ca[2], cb[2];
    if ((int)blockIdx.x < n_tiles) load_scales(clip, rem, 0, ca, cb);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int ntile = tile + gridDim.x;
      int nclip = clip + dq, nrem = rem + dr;  // row state of the CTA's next tile (one grid stride further)
      if (nrem >= g.hw) { nrem -= g.hw; ++nclip; }
      for (int kb = 0; kb < g.kblocks; ++kb) {
        float4 na[2], nb[2];
        if (kb + 1 < g.kblocks) load_scales(clip, rem, kb + 1, na, nb);
        else if (ntile < n_tiles) load_scales(nclip, nrem, 0, na, nb);
        ptx::mbar_wait(&a_full[stage], phase);
        if (kb * PW_BK + j * 8 < g.K) {
          const uint32_t base = smem_a + stage * g.stage_bytes;
          uint4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (grow0 + 32 * i < g.M) v[i] = lds128u(base + roff[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (grow0 + 32 * i >= g.M) continue;
            const bool second = rem + 32 * i >= g.hw;  // this row belongs to the next clip
            const float4 s0 = second ? cb[0] : ca[0], s1 = second ? cb[1] : ca[1];
            const float2 a = unpack_h16(v[i].x), b = unpack_h16(v[i].y), c = unpack_h16(v[i].z), d = unpack_h16(v[i].w);
            uint4 o;
            o.x = pack_h16(a.x * s0.x, a.y * s0.y);
            o.y = pack_h16(b.x * s0.z, b.y * s0.w);
            o.z = pack_h16(c.x * s1.x, c.y * s1.y);
            o.w = pack_h16(d.x * s1.z, d.y * s1.w);
            sts128u(base + roff[i], o);
          }
        }
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[stage]);
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
        ca[0] = na[0]; ca[1] = na[1]; cb[0] = nb[0]; cb[1] = nb[1];
      }
      grow0 += step_rows; clip = nclip; rem = nrem;
    }
  } else if (se && warp < PW_WARP_SE_EPI0) {
    // ===================== squeeze-excitation scalers, general form (clips shorter than a tile, or several N tiles) ==========
    // Eight warps, two threads per A row (four 16-byte chunks each).  The clip's scale values are fetched BEFORE the wait on
    // the stage (they depend only on the tile), and a thread's chunks are read together, so a stage costs one shared-memory
    // round trip, not a chain of dependent global loads.  The clip index advances incrementally (no division per tile).
    const int t = threadIdx.x - PW_WARP_SCALE0 * 32, row = t >> 1, hsel = t & 1;  // row 0..127
    int stage = 0;
    uint32_t phase = 0;
    const int step_rows = gridDim.x * PW_BM, dq = step_rows / g.hw, dr = step_rows - dq * g.hw;
    int last_mt = blockIdx.x / g.n_nt;
    int grow = last_mt * PW_BM + row, clip = grow / g.hw, rem = grow - clip * g.hw;
    uint32_t joff[4];
    int jcol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = (4 * hsel + i + row) & 7;  // the 8 lanes of a quarter-warp (4 rows x 2 halves) hit 8 distinct bank groups
      joff[i] = row * 128 + j * 16;
      jcol[i] = (j ^ (row & 7)) << 3;          // logical k (inside the k-block) of that chunk under the 128-byte swizzle
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = g.n_nt == 1 ? tile : tile / g.n_nt;
      if (mt != last_mt) {  // this thread's row moved: one N tile per row block -> by exactly one grid stride, no division
        if (g.n_nt == 1) {
          clip += dq; rem += dr;
          if (rem >= g.hw) { rem -= g.hw; ++clip; }
          grow += step_rows;
        } else {
          grow = mt * PW_BM + row;
          clip = grow / g.hw;
          rem = grow - clip * g.hw;
        }
        last_mt = mt;
      }
      const float* sp = g.se_scale + (size_t)(grow < g.M ? clip : 0) * g.K;
      for (int kb = 0; kb < g.kblocks; ++kb) {
        float4 s0[4], s1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = kb * PW_BK + jcol[i];
          if (col < g.K) {
            s0[i] = __ldg(reinterpret_cast<const float4*>(sp + col));
            s1[i] = __ldg(reinterpret_cast<const float4*>(sp + col + 4));
          }
        }
        ptx::mbar_wait(&a_full[stage], phase);
        if (grow < g.M) {
          const uint32_t base = smem_a + stage * g.stage_bytes;
          uint4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (kb * PW_BK + jcol[i] < g.K) v[i] = lds128u(base + joff[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (kb * PW_BK + jcol[i] < g.K) {
              const float2 a = unpack_h16(v[i].x), b = unpack_h16(v[i].y), c = unpack_h16(v[i].z), d = unpack_h16(v[i].w);
              uint4 o;
              o.x = pack_h16(a.x * s0[i].x, a.y * s0[i].y);
              o.y = pack_h16(b.x * s0[i].z, b.y * s0[i].w);
              o.z = pack_h16(c.x * s1[i].x, c.y * s1[i].y);
              o.w = pack_h16(d.x * s1[i].z, d.y * s1[i].w);
              sts128u(base + joff[i], o);
            }
          }
        }
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[stage]);
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    // Warps 10..17 (squeeze-excitation mode) or 2..17: `ne` = up to 2 or 4 warps per TMEM lane quarter, which take the CTA's tiles
    // in turn (the SiLU epilogue is special-function-unit bound: it wants every warp the register file can hold).  A warp owns 32 FULL rows of its tile from TMEM to global memory, so there is no barrier among epilogue warps:
    // 16-column chunks (double-buffered TMEM loads) -> BatchNorm / SiLU / residual in packed fp32x2 -> fp16 into this warp's
    // staging boxes ([32 rows x 64 columns], 128-byte swizzle: conflict-free 16-byte writes) -> one TMA tensor store per
    // box, issued by lane 0 (full-line writes; the map clips columns >= N and rows >= M).  Rows narrower than 64 columns go
    // out with plain 16-byte stores instead (the warp's 32 rows are one contiguous range of global memory either way).
    // ne <= accumulator stages: a warp waits on t_full[acc] by phase PARITY, which only tells the current phase from the previous
    // one -- the previous use of the accumulator (tile it - acc_stages) must be complete when the warp, done with tile it - ne,
    // starts waiting for tile it.  (With three warps on two accumulators a warp would take tile it - 4's completion for tile it's.)
    const int first = se ? PW_WARP_SE_EPI0 : PW_WARP_SCALE0, ne_max = se ? 2 : 4, ne = g.acc_stages < ne_max ? g.acc_stages : ne_max;
    const int quarter = warp & 3, sub = (warp - first) >> 2;
    const int row = quarter * 32 + lane;
    const int nch = g.ntb >> 4;  // 16-column chunks per tile
    const uint32_t tab_a = smem_a + g.off_tab;
    const bool direct = g.direct != 0;
    const uint32_t stg = smem_a + g.off_ob + (warp - PW_WARP_SCALE0) * (2 * PW_BOX_BYTES);  // this warp's two staging boxes
    const uint32_t sw = lane & 7;
    uint32_t box_par = 0;  // staging box in use (alternates per 64-column box)
    // counters instead of divisions by run-time values: `turn` = tile ordinal modulo ne, (acc, acc_phase) = the accumulator ring
    int turn = 0, acc = -1;
    uint32_t acc_phase = 1;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (++acc == g.acc_stages) acc = 0;
      if (acc == 0) acc_phase ^= 1;
      const bool mine = turn == sub;  // (a warp with sub >= ne has no tiles)
      if (++turn == ne) turn = 0;
      if (!mine) continue;
      const int mt = g.n_nt == 1 ? tile : tile / g.n_nt, nt = tile - mt * g.n_nt;
      const int grow = mt * PW_BM + row, n0 = nt * g.ntb;
      const __nv_bfloat16* resp = g.res != nullptr && grow < g.M ? g.res + (size_t)grow * g.N + n0 : nullptr;
      __nv_bfloat16* outp = g.out + (size_t)grow * g.N + n0;
      ptx::mbar_wait(&t_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * g.acc_stride;

      auto process8 = [&](const uint32_t* v, int c8) {  // 8 columns starting at tile column c8 * 8
        const int col = n0 + c8 * 8;
        if (col >= g.N) return;  // (warp-uniform) zero-padded columns of the last N tile
        if (g.raw != nullptr && grow < g.M) {
          float4* rp = reinterpret_cast<float4*>(g.raw + (size_t)grow * g.N + col);
          rp[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          rp[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
        const float4 sc0 = lds128f(tab_a + col * 4), sc1 = lds128f(tab_a + col * 4 + 16);
        const float4 sh0 = lds128f(tab_a + (g.N + col) * 4), sh1 = lds128f(tab_a + (g.N + col) * 4 + 16);
        float2 y0 = __ffma2_rn(make_float2(__uint_as_float(v[0]), __uint_as_float(v[1])), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
        float2 y1 = __ffma2_rn(make_float2(__uint_as_float(v[2]), __uint_as_float(v[3])), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
        float2 y2 = __ffma2_rn(make_float2(__uint_as_float(v[4]), __uint_as_float(v[5])), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
        float2 y3 = __ffma2_rn(make_float2(__uint_as_float(v[6]), __uint_as_float(v[7])), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
        if (g.silu) {
          y0 = pw_silu2(y0); y1 = pw_silu2(y1); y2 = pw_silu2(y2); y3 = pw_silu2(y3);
        }
        if (resp != nullptr) {
          uint4 rv;  // streamed once: keep it out of the (small, ~28 KB beside 228 KB of shared memory) L1 that holds the SE scale rows
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "l"(resp + c8 * 8));
          y0 = __fadd2_rn(y0, unpack_h16(rv.x)); y1 = __fadd2_rn(y1, unpack_h16(rv.y));
          y2 = __fadd2_rn(y2, unpack_h16(rv.z)); y3 = __fadd2_rn(y3, unpack_h16(rv.w));
        }
        const uint4 pk = make_uint4(pack_h16(y0.x, y0.y), pack_h16(y1.x, y1.y), pack_h16(y2.x, y2.y), pack_h16(y3.x, y3.y));
        if (!direct) sts128u(stg + box_par * PW_BOX_BYTES + lane * 128 + (((c8 & 7) ^ sw) << 4), pk);
        else if (grow < g.M) *reinterpret_cast<uint4*>(outp + c8 * 8) = pk;
      };
      // one 16-column chunk; a 64-column box is four chunks: before its first chunk the box written two boxes ago must have left
      // the staging buffer, after its last chunk (or the tile's last) lane 0 hands it to the TMA engine
      auto chunk = [&](const uint32_t* v, int i) {
        if (!direct && (i & 3) == 0) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        if (g.silu && !direct && resp == nullptr && g.raw == nullptr && n0 + i * 16 + 16 <= g.N) {
          // the expand convolutions' path: all 16 columns in ONE basic block, so that the sixteen activation chains overlap
          const uint32_t ta = tab_a + (n0 + i * 16) * 4, tb = ta + g.N * 4;
          float2 y[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 sc = lds128f(ta + k * 16), sh = lds128f(tb + k * 16);
            y[2 * k] = __ffma2_rn(make_float2(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
            y[2 * k + 1] = __ffma2_rn(make_float2(__uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
          }
          // silu(y) = y * sigmoid(y) with ONE special-function op per element: t = 2^-|y log2 e| <= 1 (EX2), d = 1 + t in [1, 2],
          // 1/d on the FMA pipe -- linear seed 24/17 - 8/17 d (|err| <= 1/17) and three Newton steps (-> 1.5e-10, more exact
          // than RCP.approx) -- then sigmoid = 1/d for y >= 0, t/d for y < 0.  With EX2 + RCP the epilogue sat on the
          // special-function unit (16 lanes/clk/SM: 76 % busy in ncu); this form trades one MUFU for ~8 issue slots per pair and
          // the packed-FMA pipe had the room.  No overflow: t never exceeds 1.
          float2 t[8], r[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 q = __fmul2_rn(y[k], make_float2(1.4426950408889634f, 1.4426950408889634f));
            const float qx = -fabsf(q.x), qy = -fabsf(q.y);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[k].x) : "f"(qx));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[k].y) : "f"(qy));
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 d = __fadd2_rn(t[k], make_float2(1.0f, 1.0f)), nd = make_float2(-d.x, -d.y);
            float2 x = __ffma2_rn(d, make_float2(-8.0f / 17.0f, -8.0f / 17.0f), make_float2(24.0f / 17.0f, 24.0f / 17.0f));
#pragma unroll
            for (int it = 0; it < 3; ++it) x = __ffma2_rn(x, __ffma2_rn(nd, x, make_float2(1.0f, 1.0f)), x);
            r[k] = x;
          }
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 m = make_float2(y[k].x >= 0.f ? 1.0f : t[k].x, y[k].y >= 0.f ? 1.0f : t[k].y);
            const float2 o = __fmul2_rn(y[k], __fmul2_rn(r[k], m));
            pk[k] = pack_h16(o.x, o.y);
          }
          const uint32_t rowa = stg + box_par * PW_BOX_BYTES + lane * 128;
          sts128u(rowa + ((((2 * i) & 7) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128u(rowa + ((((2 * i + 1) & 7) ^ sw) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
        } else {
          process8(v, 2 * i);
          process8(v + 8, 2 * i + 1);
        }
        if (!direct && ((i & 3) == 3 || i == nch - 1)) {
          ptx::fence_proxy_async();  // the warp's shared-memory writes -> visible to the TMA store
          __syncwarp();
          const int c0 = n0 + (i >> 2) * 64;
          if (lane == 0 && c0 < g.N) ptx::tma_store_2d_s(&map_out, stg + box_par * PW_BOX_BYTES, c0, mt * PW_BM + quarter * 32);
          if (lane == 0) ptx::tma_store_commit();  // (possibly empty: keeps one group per box for the wait above)
          box_par ^= 1;
        }
      };
      auto release = [&]() {  // every TMEM read of this accumulator has completed: hand it back to the MMA thread
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
      };

      uint32_t va[16], vb[16];
      tmem_ld_32x16(t_addr, va);
      for (int i = 0; i < nch; i += 2) {
        ptx::tmem_ld_wait();
        if (i + 1 < nch) tmem_ld_32x16(t_addr + (i + 1) * 16, vb);
        else release();
        chunk(va, i);
        if (i + 1 < nch) {
          ptx::tmem_ld_wait();
          if (i + 2 < nch) tmem_ld_32x16(t_addr + (i + 2) * 16, va);
          else release();
          chunk(vb, i + 1);
        }
      }
    }
    if (!direct && lane == 0) ptx::tma_store_wait_all();  // the stores must have been performed before the CTA exits
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == PW_WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Whether the memory-bound kernel takes this 1x1 convolution (else the general GEMM does): it always writes an fp16 output.
bool pointwise_supported(int N, int K, const void* out, int out_16bit) {
  return out != nullptr && out_16bit && N % 8 == 0 && K % 8 == 0 && N >= 8 && K >= 8;
}

int pointwise_launch(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                     const __nv_bfloat16* res, const float* se_scale, int hw, float* raw_out, void* out, cudaStream_t st) {
  AVEXK_CHECK_ARG(pointwise_supported(N, K, out, 1), "pointwise: unsupported shape N=%d K=%d", N, K);
  AVEXK_CHECK_ARG(se_scale == nullptr || hw > 0, "pointwise: squeeze-excitation scale needs rows per clip");
  if (M == 0) return AVEXK_OK;
  PwArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.kblocks = ceil_div(K, PW_BK);
  // N tiling: one tile when N <= 256, else 128-column tiles (a multiple of the 64-column store box: a box never spills into the
  // next tile's columns)
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) { g.n_nt = 1; g.ntb = n16; }
  else { g.ntb = 128; g.n_nt = ceil_div(N, 128); }
  g.w_block_bytes = g.ntb * 128;
  const int tab_bytes = (2 * N * 4 + 15) / 16 * 16;
  const int bar_bytes = (3 * PW_MAX_STAGES + 2 * PW_MAX_ACC + 1) * 8 + 16;
  const int w_all = g.kblocks * g.w_block_bytes;
  g.direct = g.ntb < 64;
  const int ob_bytes = g.direct ? 0 : PW_EPI_SLOTS * 2 * PW_BOX_BYTES;  // two staging boxes per epilogue warp
  // shared-memory plan: staging boxes, W resident when it fits beside >= 4 stages, the rest to the load ring
  const int avail = 227 * 1024 - 1024 - tab_bytes - bar_bytes - 128 - ob_bytes;
  g.w_resident = g.n_nt == 1 && avail - w_all >= 4 * PW_A_BYTES;
  g.stage_bytes = PW_A_BYTES + (g.w_resident ? 0 : g.w_block_bytes);
  int stages = (avail - (g.w_resident ? w_all : 0)) / g.stage_bytes;
  AVEXK_CHECK_ARG(stages >= 2, "pointwise: tile does not fit shared memory (N=%d K=%d)", N, K);
  g.stages = stages > PW_MAX_STAGES ? PW_MAX_STAGES : stages;
  g.off_w = g.stages * g.stage_bytes;
  g.off_ob = g.off_w + (g.w_resident ? w_all : 0);
  g.off_tab = g.off_ob + ob_bytes;
  g.off_bar = g.off_tab + tab_bytes;
  g.acc_stride = (g.ntb + 31) / 32 * 32;
  g.acc_stages = 512 / g.acc_stride > PW_MAX_ACC ? PW_MAX_ACC : 512 / g.acc_stride;
  g.idesc = ptx::make_idesc_f16(PW_BM, g.ntb);
  g.m_tiles = ceil_div(M, PW_BM);
  g.silu = silu; g.hw = hw;
  g.scale = scale; g.shift = shift; g.se_scale = se_scale; g.res = res;
  g.out = reinterpret_cast<__nv_bfloat16*>(out);
  g.raw = raw_out;
  CUtensorMap map_a, map_w, map_out;
  int rc = make_tmap_2d_bf16(&map_a, A, M, K, K, PW_BM, PW_BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&map_w, W, N, K, K, g.ntb, PW_BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&map_out, out, M, N, N, 32, 64);  // store boxes: 32 rows x 64 columns, 128-byte swizzle
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)g.off_bar + bar_bytes;
  static bool attr_set[64] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    AVEXK_CUDA(cudaFuncSetAttribute(pointwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  const int tiles = g.m_tiles * g.n_nt;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  pointwise_kernel<<<grid, PW_THREADS, smem, st>>>(map_a, map_w, map_out, g);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

}  // namespace avexk
