// posconv.cu -- BEATs convolutional position embedding on the sm_100a tensor cores.
//
// Replaces avex/models/beats/backbone.py:52-68,172-174 + modules.py:67-94:
//   y = x + GELU(Conv1d(C, C, k=128, pad=64, groups=16)(x^T)[:, :, :N]^T)      (weight-norm resolved at load)
// as 16 implicit GEMMs, one per channel group.  For a super-tile of 256 tokens of one clip and group g
//   acc[n, co] = sum_{t<128} sum_{ci<48} x[n + t - 64, g*48 + ci] * w[g*48 + co, ci, t]
// The activation window (rows n0-64 .. n0+319 of the group-padded bf16 copy xg [B, N, 16*64]) is loaded ONCE per
// super-tile by TMA (3-D tensor map: rows outside the clip are zero-filled by hardware -- exactly the conv's zero
// padding).  Tap t is one K-block whose A operand is that same window shifted down by t rows: the UMMA shared-memory
// descriptor simply starts t*128 bytes later (the hardware swizzles on absolute address bits, so the phase still matches
// what TMA wrote).  Only the per-tap weight tile (48 x 48, 6 KB) streams through the TMA ring, and it
// is shared by the two 128-token halves of the super-tile.  UMMA shape 128 x 48 x 16, K = 48 per tap (the 16 pad
// channels are never multiplied); accumulators in TMEM (2 buffers x 2 halves x 64 columns).
// Epilogue: + bias, GELU, + x0 (residual), fp32 store.
// Roofline: shared-memory operand bandwidth (A 12 KB + W 4.5 KB per tap per half at 128 B/clk), then the tensor pipe.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace avexk {
namespace {

constexpr int BM = 128, HALVES = 2, SUPER = BM * HALVES, CG = 48, TAPS = 128, WSTAGES = 8;
constexpr int WIN_BOXES = 3, WIN_ROWS = WIN_BOXES * 128;       // rows n0-64 .. n0+319 (383 needed)
constexpr int WIN_BYTES = WIN_ROWS * 128, W_BYTES = CG * 128;  // 49152, 6144 (both multiples of 1024)
constexpr int OFF_W = 2 * WIN_BYTES, OFF_BAR = OFF_W + WSTAGES * W_BYTES;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 512;
constexpr int NTHREADS = 192;
constexpr int ACC_COLS = 64;

struct PcArgs {
  int B, N, G, tiles_per_clip;
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* win_full = bars;                  // [2]
  uint64_t* win_empty = bars + 2;             // [2]
  uint64_t* w_full = bars + 4;                // [WSTAGES]
  uint64_t* w_empty = bars + 4 + WSTAGES;     // [WSTAGES]
  uint64_t* tfull_bar = bars + 4 + 2 * WSTAGES;   // [2]
  uint64_t* tempty_bar = bars + 6 + 2 * WSTAGES;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * WSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.G * CG;
  const int num_tiles = g.B * g.tiles_per_clip * g.G;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_w);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&win_full[i], 1);
      ptx::mbar_init(&win_empty[i], 1);
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);
    }
    for (int i = 0; i < WSTAGES; ++i) {
      ptx::mbar_init(&w_full[i], 1);
      ptx::mbar_init(&w_empty[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * HALVES * ACC_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (clip b, super-tile nt, group grp); groups fastest so concurrent CTAs share the activation rows in L2
  if (warp == 0) {
    if (lane == 0) {
      auto load_window = [&](int tile, int it) {
        const int grp = tile % g.G, mt = tile / g.G;
        const int nt = mt % g.tiles_per_clip, b = mt / g.tiles_per_clip;
        const int wb = it & 1;
        ptx::mbar_wait(&win_empty[wb], ((it >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&win_full[wb], WIN_BYTES);
        for (int bx = 0; bx < WIN_BOXES; ++bx)
          ptx::tma_load_3d(smem + wb * WIN_BYTES + bx * 128 * 128, &map_x, &win_full[wb], grp * 64,
                           nt * SUPER - TAPS / 2 + bx * 128, b);
      };
      int stage = 0, it = 0;
      uint32_t phase = 0;
      if (blockIdx.x < num_tiles) load_window(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        if (tile + (int)gridDim.x < num_tiles) load_window(tile + gridDim.x, it + 1);  // one super-tile ahead
        const int grp = tile % g.G;
        for (int t = 0; t < TAPS; ++t) {
          ptx::mbar_wait(&w_empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&w_full[stage], W_BYTES);
          ptx::tma_load_2d(smem + OFF_W + stage * W_BYTES, &map_w, &w_full[stage], t * 64, grp * CG);
          if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, CG);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1, wb = it & 1;
        ptx::mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        ptx::mbar_wait(&win_full[wb], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t win_addr = ptx::smem_u32(smem + wb * WIN_BYTES);
        for (int t = 0; t < TAPS; ++t) {
          ptx::mbar_wait(&w_full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t db = ptx::make_sw128_desc(ptx::smem_u32(smem + OFF_W + stage * W_BYTES));
#pragma unroll
          for (int hf = 0; hf < HALVES; ++hf) {
            const uint64_t da = make_sw128_desc_rows(win_addr, hf * BM + t);
            const uint32_t d_tmem = tmem_base + (acc * HALVES + hf) * ACC_COLS;
#pragma unroll
            for (int k = 0; k < CG / 16; ++k) ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (t | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&w_empty[stage]);
          if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[acc]);
        ptx::umma_commit(&win_empty[wb]);
      }
    }
  } else {
    const int quarter = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int grp = tile % g.G, mt = tile / g.G;
      const int nt = mt % g.tiles_per_clip, b = mt / g.tiles_per_clip;
      ptx::mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < HALVES; ++hf) {
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * HALVES + hf) * ACC_COLS;
        const int n = nt * SUPER + hf * BM + quarter * 32 + lane;  // token inside the clip (one row per lane)
        const size_t row_off = ((size_t)b * g.N + n) * C + grp * CG;
#pragma unroll
        for (int c = 0; c < CG / 16; ++c) {
          uint32_t r[16];
          tmem_ld_32x16(t_addr + c * 16, r);
          ptx::tmem_ld_wait();
          if (n < g.N) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int col = c * 16 + i * 4;
              const float4 bs = __ldg(reinterpret_cast<const float4*>(g.bias + grp * CG + col));
              const float4 rs = __ldg(reinterpret_cast<const float4*>(g.x0 + row_off + col));
              float4 v;
              v.x = gelu_erf(__uint_as_float(r[4 * i + 0]) + bs.x) + rs.x;
              v.y = gelu_erf(__uint_as_float(r[4 * i + 1]) + bs.y) + rs.y;
              v.z = gelu_erf(__uint_as_float(r[4 * i + 2]) + bs.z) + rs.z;
              v.w = gelu_erf(__uint_as_float(r[4 * i + 3]) + bs.w) + rs.w;
              *reinterpret_cast<float4*>(g.out + row_off + col) = v;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * HALVES * ACC_COLS);
  }
}

}  // namespace

// xg [B, N, G*64] bf16 (group-padded activations), Wpc [G*48, 128*64] bf16, out = x0 + gelu(conv + bias), fp32 [B*N, G*48]
int launch_posconv(const __nv_bfloat16* xg, const __nv_bfloat16* Wpc, const float* bias, const float* x0, float* out, int B,
                   int N, int G, int cg, int taps, cudaStream_t st) {
  AVEXK_CHECK_ARG(cg == CG && taps == TAPS, "posconv kernel is specialised to 48 channels/group and 128 taps (got %d, %d)", cg, taps);
  if (B == 0 || N == 0) return AVEXK_OK;
  static bool attr_set = false;
  if (!attr_set) {
    AVEXK_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap mx, mw;
  int rc = make_tmap_3d_bf16(&mx, xg, (long long)G * 64, N, B, (long long)G * 64, (long long)N * G * 64, 64, 128, 1, true);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&mw, Wpc, (long long)G * CG, (long long)TAPS * 64, (long long)TAPS * 64, CG, 64);
  if (rc) return rc;
  PcArgs a{B, N, G, ceil_div(N, SUPER), bias, x0, out};
  const int tiles = B * a.tiles_per_clip * G;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  prof_begin(st, KID_POSCONV, 2.0 * B * N * (double)(G * CG) * CG * TAPS);
  posconv_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(mx, mw, a);
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
