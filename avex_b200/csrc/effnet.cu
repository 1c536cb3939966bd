// effnet.cu -- EfficientNet-B0/B1 feature extractor forward (torchvision topology) as NHWC fp16 kernels.
//
// Storage and tensor-core operands are FP16, not bf16 (the 16-bit buffers are typed __nv_bfloat16 in the signatures for
// historical reasons): every stored activation is post-BatchNorm / SiLU, i.e. bounded, the tensor core runs kind::f16 at the
// same rate for both formats, and fp16's three extra mantissa bits take the per-layer cosine of this 18-conv-deep, ill-conditioned
// random-init net from 0.997 (bf16) to >= 0.9999 -- inside the north_star tolerance without a separate fp32 mode.
//
// Replaces `efficientnet.Model.forward` (avex/models/efficientnet.py:163-215): the 3-channel repeat of the mel image
// (efficientnet.py:138-140), torchvision `efficientnet_b0().features` (stem conv, 16 MBConv blocks, head conv; BatchNorm
// in eval mode, SiLU, squeeze-excitation), `avgpool` and the classifier.
//   * the 3 identical input channels are never materialised: the stem kernel uses the conv weights summed over C_in and
//     applies the per-clip min-max normalisation of the mel image (audio_utils.py:167-172) on load;
//   * activations live in HBM as NHWC fp16, so every 1x1 convolution (88 % of the MACs) is a GEMM [B*H*W, C_in] x [C_out, C_in]^T:
//     the memory-bound tcgen05 kernel of pointwise.cu (folded BatchNorm, SiLU, residual add in its epilogue, the squeeze-
//     excitation rescale of the project convolution's input on its A operand, and -- when a forward hook on `block.3.0` /
//     `block.2.0` asks for it -- an fp32 copy of the raw pre-BN accumulators); the head convolution (fp32 NCHW features)
//     runs on the general GEMM of gemm_tc.cu;
//   * depthwise k x k convolutions: a TMA-fed sliding-window kernel (dwconv_tma_kernel below: input rows through a shared-
//     memory ring, weights and a rotating window of output-row accumulators in registers), BN + SiLU fused, the squeeze-
//     excitation sums leaving as 64-bit fixed-point reductions; the round-1 per-output-row kernel stays as the fallback;
//   * SE: a small MLP kernel (four clips per CTA) produces the channel scales that the project convolution applies.
// Roofline: HBM (about 29 FLOP/B overall); per kernel family the binding unit differs.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

struct avexk_effnet {
  std::vector<avexk_effnet_block_cfg> cfg;
  bool loaded = false;
  int num_classes = 0;
  std::vector<void*> allocs;
  // stem
  float *stem_w = nullptr, *stem_scale = nullptr, *stem_shift = nullptr;  // [9][32], [32], [32]
  struct Block {
    __nv_bfloat16 *expand_w = nullptr, *proj_w = nullptr;
    float *expand_scale = nullptr, *expand_shift = nullptr, *dw_w = nullptr, *dw_scale = nullptr, *dw_shift = nullptr;
    float *se1_w = nullptr, *se1_b = nullptr, *se2_w = nullptr, *se2_b = nullptr, *proj_scale = nullptr, *proj_shift = nullptr;
  };
  std::vector<Block> blocks;
  __nv_bfloat16* head_w = nullptr;
  float *head_scale = nullptr, *head_shift = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  int stem_out = 32, head_in = 320, head_out = 1280;
};

namespace avexk {
namespace {

__device__ __forceinline__ float silu(float v) { return v / (1.0f + __expf(-v)); }
// silu for a channel pair: one packed multiply, two EX2, two RCP (no range fix-ups: ex2.approx.ftz saturates to +inf / 0, and
// v * rcp(1 + inf) = -0 is the right limit)
__device__ __forceinline__ float2 silu2(float2 v) {
  const float2 q = __fmul2_rn(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  e = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
  return __fmul2_rn(v, r);
}

__device__ __forceinline__ float dec_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// BatchNorm (eval) -> per-channel scale / shift: y = x * scale + shift, eps 1e-5 (torchvision default)
__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, float eps, int C,
                               float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = w[c] / sqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = b[c] - mean[c] * s;
  }
}

// stem weights [32, 3, 3, 3] -> [9][32] summed over the three identical input channels
__global__ void stem_pack_kernel(const float* w, int Cout, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // tap * Cout + co
  if (i < 9 * Cout) {
    const int co = i % Cout, tap = i / Cout;
    out[i] = w[(co * 3 + 0) * 9 + tap] + w[(co * 3 + 1) * 9 + tap] + w[(co * 3 + 2) * 9 + tap];
  }
}

// depthwise weights [C, 1, k, k] -> [k*k][C]
__global__ void dw_pack_kernel(const float* w, int C, int kk, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * kk) {
    const int c = i % C, tap = i / C;
    out[i] = w[c * kk + tap];
  }
}

// ---------------------------------------------------------------------------------------------------------
// stem: 3x3 stride-2 pad-1 conv on the (normalised-on-load) mel image [B, 128, Tm] -> NHWC bf16 [B, Ho, Wo, 32]
// ---------------------------------------------------------------------------------------------------------
constexpr int STEM_C = 32;
__global__ void __launch_bounds__(256)
stem_kernel(const float* __restrict__ mel, const unsigned* __restrict__ minmax, int B, int H, int W, int Ho, int Wo,
            const float* __restrict__ w9, const float* __restrict__ scale, const float* __restrict__ shift,
            __nv_bfloat16* __restrict__ out, float* __restrict__ raw_nchw) {
  __shared__ __align__(16) float sw[9 * STEM_C], ss[STEM_C], sh[STEM_C];
  for (int i = threadIdx.x; i < 9 * STEM_C; i += blockDim.x) sw[i] = w9[i];
  if (threadIdx.x < STEM_C) {
    ss[threadIdx.x] = scale[threadIdx.x];
    sh[threadIdx.x] = shift[threadIdx.x];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo) return;
  const int ho = p / Wo, wo = p - ho * Wo;
  float mn = 0.f, inv = 1.f;
  if (minmax != nullptr) {
    mn = dec_ordered(minmax[2 * b]);
    inv = 1.0f / (dec_ordered(minmax[2 * b + 1]) - mn + 1e-8f);
  }
  const float* img = mel + (size_t)b * H * W;
  float v[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int hi = 2 * ho - 1 + ky, wi = 2 * wo - 1 + kx;
      const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
      v[ky * 3 + kx] = ok ? (__ldg(img + (size_t)hi * W + wi) - mn) * inv : 0.f;  // zero padding of the normalised image
    }
  // four channels per step: one 16-byte broadcast read of the tap's weights feeds two packed fp32x2 FMAs
  uint32_t packed[STEM_C / 2];
#pragma unroll
  for (int c = 0; c < STEM_C; c += 4) {
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w4 = *reinterpret_cast<const float4*>(&sw[t * STEM_C + c]);
      const float2 x2 = make_float2(v[t], v[t]);
      a0 = __ffma2_rn(x2, make_float2(w4.x, w4.y), a0);
      a1 = __ffma2_rn(x2, make_float2(w4.z, w4.w), a1);
    }
    if (raw_nchw != nullptr) {
      raw_nchw[((size_t)b * STEM_C + c) * Ho * Wo + p] = a0.x;
      raw_nchw[((size_t)b * STEM_C + c + 1) * Ho * Wo + p] = a0.y;
      raw_nchw[((size_t)b * STEM_C + c + 2) * Ho * Wo + p] = a1.x;
      raw_nchw[((size_t)b * STEM_C + c + 3) * Ho * Wo + p] = a1.y;
    }
    const float4 s4 = *reinterpret_cast<const float4*>(&ss[c]), h4 = *reinterpret_cast<const float4*>(&sh[c]);
    const float2 y0 = silu2(__ffma2_rn(a0, make_float2(s4.x, s4.y), make_float2(h4.x, h4.y)));
    const float2 y1 = silu2(__ffma2_rn(a1, make_float2(s4.z, s4.w), make_float2(h4.z, h4.w)));
    packed[c / 2] = pack_h16(y0.x, y0.y);
    packed[c / 2 + 1] = pack_h16(y1.x, y1.y);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * Ho * Wo + p) * STEM_C);
#pragma unroll
  for (int i = 0; i < STEM_C / 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

// ---------------------------------------------------------------------------------------------------------
// depthwise k x k conv + BN + SiLU, NHWC bf16, with the squeeze-excitation sums
// ---------------------------------------------------------------------------------------------------------
constexpr int DW_TW = 4;        // consecutive output pixels (along W) per thread
constexpr float SE_FIX = 16777216.0f;  // 2^24
// One thread = 8 channels x DW_TW consecutive output pixels of one row.  Per kernel row it loads the (DW_TW-1)*S + K input
// columns once (16-byte loads, channels-last) and the K weight vectors once, and reuses both across the DW_TW outputs:
// 10 instead of 25 activation loads and 12 instead of 50 weight loads per output pixel at K = 5, S = 1.
template <int K, int S>
__global__ void __launch_bounds__(256)
dwconv_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int C, int Ho, int Wo,
              const float* __restrict__ wkk, const float* __restrict__ scale, const float* __restrict__ shift,
              __nv_bfloat16* __restrict__ out, unsigned long long* __restrict__ se_sum, int quads_per_cta) {
  // squeeze-excitation sums in 40.24 fixed point: integer adds are associative, so the atomics below give the same bits
  // whatever order the pixel groups / CTAs arrive in (a float atomicAdd made results differ from run to run)
  extern __shared__ unsigned long long sse[];  // [C]
  const int b = blockIdx.y, CV = C >> 3;
  const int QG = blockDim.x / CV;  // quads in flight
  const int cv = threadIdx.x % CV, qg = threadIdx.x / CV;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sse[i] = 0ull;
  __syncthreads();
  constexpr int PAD = (K - 1) / 2, NC = (DW_TW - 1) * S + K;
  const int quads_per_row = (Wo + DW_TW - 1) / DW_TW, nquads = Ho * quads_per_row;
  const int q_end = min(nquads, (int)(blockIdx.x + 1) * quads_per_cta);
  float se[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) se[i] = 0.f;
  if (qg < QG) {
    const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + cv * 8)), sc1 = __ldg(reinterpret_cast<const float4*>(scale + cv * 8 + 4));
    const float4 sh0 = __ldg(reinterpret_cast<const float4*>(shift + cv * 8)), sh1 = __ldg(reinterpret_cast<const float4*>(shift + cv * 8 + 4));
    const __nv_bfloat16* src = in + (size_t)b * H * W * C + cv * 8;
    for (int q = blockIdx.x * quads_per_cta + qg; q < q_end; q += QG) {
      const int ho = q / quads_per_row, wo0 = (q - ho * quads_per_row) * DW_TW;
      float acc[DW_TW][8];
#pragma unroll
      for (int t = 0; t < DW_TW; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int hi = ho * S - PAD + ky;
        if (hi < 0 || hi >= H) continue;
        float4 w0[K], w1[K];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float* wp = wkk + (ky * K + kx) * C + cv * 8;
          w0[kx] = __ldg(reinterpret_cast<const float4*>(wp));
          w1[kx] = __ldg(reinterpret_cast<const float4*>(wp + 4));
        }
        const __nv_bfloat16* rowp = src + (size_t)hi * W * C;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int wi = wo0 * S - PAD + c;
          if (wi < 0 || wi >= W) continue;
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)wi * C));
          const float2 x0 = unpack_h16(raw.x), x1 = unpack_h16(raw.y), x2 = unpack_h16(raw.z), x3 = unpack_h16(raw.w);
#pragma unroll
          for (int t = 0; t < DW_TW; ++t) {
            const int kx = c - t * S;  // compile-time after unrolling
            if (kx >= 0 && kx < K) {
              acc[t][0] = fmaf(x0.x, w0[kx].x, acc[t][0]); acc[t][1] = fmaf(x0.y, w0[kx].y, acc[t][1]);
              acc[t][2] = fmaf(x1.x, w0[kx].z, acc[t][2]); acc[t][3] = fmaf(x1.y, w0[kx].w, acc[t][3]);
              acc[t][4] = fmaf(x2.x, w1[kx].x, acc[t][4]); acc[t][5] = fmaf(x2.y, w1[kx].y, acc[t][5]);
              acc[t][6] = fmaf(x3.x, w1[kx].z, acc[t][6]); acc[t][7] = fmaf(x3.y, w1[kx].w, acc[t][7]);
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < DW_TW; ++t) {
        if (wo0 + t >= Wo) continue;
        float y[8];
        y[0] = silu(fmaf(acc[t][0], sc0.x, sh0.x)); y[1] = silu(fmaf(acc[t][1], sc0.y, sh0.y));
        y[2] = silu(fmaf(acc[t][2], sc0.z, sh0.z)); y[3] = silu(fmaf(acc[t][3], sc0.w, sh0.w));
        y[4] = silu(fmaf(acc[t][4], sc1.x, sh1.x)); y[5] = silu(fmaf(acc[t][5], sc1.y, sh1.y));
        y[6] = silu(fmaf(acc[t][6], sc1.z, sh1.z)); y[7] = silu(fmaf(acc[t][7], sc1.w, sh1.w));
#pragma unroll
        for (int i = 0; i < 8; ++i) se[i] += y[i];
        *reinterpret_cast<uint4*>(out + ((size_t)b * Ho * Wo + (size_t)ho * Wo + wo0 + t) * C + cv * 8) =
            make_uint4(pack_h16(y[0], y[1]), pack_h16(y[2], y[3]), pack_h16(y[4], y[5]), pack_h16(y[6], y[7]));
      }
    }
    if (se_sum != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sse[cv * 8 + i], static_cast<unsigned long long>(__float2ll_rn(se[i] * SE_FIX)));
    }
  }
  if (se_sum != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(se_sum + (size_t)b * C + i, sse[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// depthwise conv, sliding-window form fed by TMA (the production kernel)
// ---------------------------------------------------------------------------------------------------------
// A persistent CTA works through tiles of (clip, band of output rows, G*DW_TW output columns, CSL channels).  One producer
// thread streams the tile's INPUT rows through a shared-memory ring with 4-D TMA boxes [1, 1, ncw columns, CSL channels]; the
// map's out-of-bounds zero fill is the convolution's zero padding, so the arithmetic has no edge predicates, and the ring
// keeps several rows (tens of KB per CTA) in flight without holding a register.  A consumer thread owns CH channels x DW_TW
// output columns and walks DOWN the band: the K*K weights of its channels stay in registers, every input row is read once
// (NC = (DW_TW-1)*S + K columns from shared memory, conflict-free: lanes = consecutive channels) and scattered into the
// ceil(K/S) output rows it contributes to, whose accumulators rotate through a register window (the row loop is unrolled
// over one period of that rotation, so every slot index is a compile-time constant).  Two channels ride in one packed fp32x2
// FMA.  Against the per-output-row kernel above: K (stride 1) or K/2 (stride 2) times fewer loads per output, no weight
// traffic, no address arithmetic in the row loop -- the layers become DRAM / FMA-issue bound instead of LSU bound.
__host__ __device__ constexpr int pos_mod(int a, int m) { return ((a % m) + m) % m; }
__host__ __device__ constexpr int floor_div(int a, int m) { return (a - pos_mod(a, m)) / m; }

__device__ __forceinline__ void dw_lds(uint32_t addr, float2 (&x)[1]) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  x[0] = unpack_h16(r);
}
__device__ __forceinline__ void dw_lds(uint32_t addr, float2 (&x)[2]) {
  uint32_t r0, r1;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
  x[0] = unpack_h16(r0);
  x[1] = unpack_h16(r1);
}
__device__ __forceinline__ void dw_store(__nv_bfloat16* p, const float2 (&y)[1]) { *reinterpret_cast<uint32_t*>(p) = pack_h16(y[0].x, y[0].y); }
__device__ __forceinline__ void dw_store(__nv_bfloat16* p, const float2 (&y)[2]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_h16(y[0].x, y[0].y), pack_h16(y[1].x, y[1].y));
}

constexpr int DW_MAX_RING = 8;
struct DwTmaArgs {
  int B, H, W, C, Ho, Wo;
  int csl, pv, g;            // channels per tile, channel vectors (of CH) per column group, column groups per tile
  int bh;                    // output rows per band
  int n_cs, n_ct, n_band;    // channel slices; column tiles and bands per clip
  int n_tiles, ring, slot_bytes, row_bytes, out_row_pitch;
  const float *wkk, *scale, *shift;
  __nv_bfloat16* out;
  unsigned long long* se_sum;
};
struct DwTile {
  int b, ct, band;
};
// A CTA keeps ONE channel slice (blockIdx.x % n_cs: its weights are loaded once) and strides over the (clip, band, column tile)
// positions; CTAs with consecutive blockIdx work on the other slices of the same position at the same time, so the slices of
// a pixel -- which share DRAM sectors / L2 lines -- are fetched from DRAM once.
__device__ __forceinline__ DwTile dw_decode(int sp, const DwTmaArgs& a) {
  DwTile t;
  t.ct = sp % a.n_ct;
  sp /= a.n_ct;
  t.band = sp % a.n_band;
  t.b = sp / a.n_band;
  return t;
}

template <int K, int S, int CH>
__global__ void __launch_bounds__(256, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap map_in, const DwTmaArgs a) {
  constexpr int PAD = (K - 1) / 2, V = CH / 2, TW = DW_TW;
  constexpr int NSLOT = (K + S - 1) / S;    // output rows with contributions in flight
  constexpr int PERIOD = S * NSLOT;         // input rows after which the slot rotation repeats
  constexpr int NC = (TW - 1) * S + K;      // input columns per consumer thread and row step
  extern __shared__ unsigned char dw_smem_raw[];
  unsigned char* smem = dw_smem_raw + ((128u - (ptx::smem_u32(dw_smem_raw) & 127u)) & 127u);
  const uint32_t ring_a = ptx::smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.ring * a.slot_bytes);  // full[ring], empty[ring]
  const uint32_t full_a = ptx::smem_u32(bars), empty_a = full_a + DW_MAX_RING * 8;
  const int n_cons_warps = blockDim.x / 32 - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&map_in);
    for (int i = 0; i < a.ring; ++i) {
      ptx::mbar_init(&bars[i], 1);
      ptx::mbar_init(&bars[DW_MAX_RING + i], n_cons_warps);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  if (warp == n_cons_warps) {
    // ===================== producer: one thread streams input rows of this CTA's tiles =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const int cs = blockIdx.x % a.n_cs;
      for (int sp = blockIdx.x / a.n_cs; sp < a.n_tiles; sp += gridDim.x / a.n_cs) {
        const DwTile t = dw_decode(sp, a);
        const int ho0 = t.band * a.bh, rows = min(a.bh, a.Ho - ho0);
        const int hi0 = ho0 * S - PAD, wi0 = t.ct * a.g * TW * S - PAD, nsteps = (rows - 1) * S + K;
        for (int r = 0; r < nsteps; ++r) {
          const int hi = hi0 + r;
          if ((unsigned)hi >= (unsigned)a.H) continue;  // a padding row: the consumers know, nothing is sent
          for (uint32_t spins = 0; !ptx::mbar_try_wait_a(empty_a + slot * 8, phase ^ 1);) {  // back-off: the ring is rows deep
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();  // a protocol bug fails the launch instead of hanging the GPU
          }
          ptx::mbar_arrive_expect_tx_a(full_a + slot * 8, a.row_bytes);
          ptx::tma_load_4d_a(ring_a + slot * a.slot_bytes, &map_in, full_a + slot * 8, cs * a.csl, wi0, hi, t.b);
          if (++slot == a.ring) { slot = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  const int tid = threadIdx.x;
  const bool active = tid < a.pv * a.g;
  const int cvv = active ? tid % a.pv : 0, cg = active ? tid / a.pv : 0;  // idle threads shadow thread 0 (no stores)
  const uint32_t x_off = ((cg * TW * S) * a.csl + cvv * CH) * 2;          // this thread's first column inside a ring row
  const uint32_t col_pitch = a.csl * 2;
  int slot = 0;
  uint32_t phase = 0;
  const int cs = blockIdx.x % a.n_cs, c0 = cs * a.csl + cvv * CH;
  float2 w[K * K][V], sh[V];  // BatchNorm folded: w = conv weight * scale, accumulators start from the shift
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float2 sc = __ldg(reinterpret_cast<const float2*>(a.scale + c0 + 2 * v));
    sh[v] = __ldg(reinterpret_cast<const float2*>(a.shift + c0 + 2 * v));
#pragma unroll
    for (int kk = 0; kk < K * K; ++kk)
      w[kk][v] = __fmul2_rn(__ldg(reinterpret_cast<const float2*>(a.wkk + (size_t)kk * a.C + c0 + 2 * v)), sc);
  }
  for (int sp = blockIdx.x / a.n_cs; sp < a.n_tiles; sp += gridDim.x / a.n_cs) {
    const DwTile t = dw_decode(sp, a);
    const int ho0 = t.band * a.bh, rows = min(a.bh, a.Ho - ho0), wo0 = (t.ct * a.g + cg) * TW;
    const int hi0 = ho0 * S - PAD, r_last = (rows - 1) * S + K - 1;
    const int nvalid = active ? a.Wo - wo0 : 0;  // output columns tt < nvalid are stored
    __nv_bfloat16* dst = a.out + (((size_t)t.b * a.Ho + ho0) * a.Wo + wo0) * a.C + c0;  // advances one output row per emission
    float2 se[V], acc[NSLOT][TW][V];
#pragma unroll
    for (int v = 0; v < V; ++v) se[v] = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
      for (int tt = 0; tt < TW; ++tt)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[s][tt][v] = sh[v];

    for (int rb = 0; rb <= r_last; rb += PERIOD) {
#pragma unroll
      for (int u = 0; u < PERIOD; ++u) {
        const int r = rb + u;
        const bool row_in = (unsigned)(hi0 + r) < (unsigned)a.H && r <= r_last;  // tile-uniform
        if (row_in) {
          ptx::mbar_wait_a(full_a + slot * 8, phase);
          const uint32_t xrow = ring_a + slot * a.slot_bytes + x_off;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            float2 x[V];
            dw_lds(xrow + c * col_pitch, x);
            if (c == NC - 1) {  // the row is in registers: hand the slot back to the producer
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_a(empty_a + slot * 8);
            }
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              if (pos_mod(u - ky, S) != 0) continue;  // this input row is not under tap row ky of any output row
              const int s = pos_mod(floor_div(u - ky, S), NSLOT);
#pragma unroll
              for (int tt = 0; tt < TW; ++tt) {
                const int kx = c - tt * S;
                if (kx < 0 || kx >= K) continue;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                  if (ky == 0 && kx == 0) acc[s][tt][v] = __ffma2_rn(x[v], w[0][v], sh[v]);  // first tap (re)starts the slot
                  else acc[s][tt][v] = __ffma2_rn(x[v], w[ky * K + kx][v], acc[s][tt][v]);
                }
              }
            }
          }
          if (++slot == a.ring) { slot = 0; phase ^= 1; }
        } else if (pos_mod(u, S) == 0) {  // a zero (padding) row under tap row 0: the slot still has to restart
          const int s = pos_mod(floor_div(u, S), NSLOT);
#pragma unroll
          for (int tt = 0; tt < TW; ++tt)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[s][tt][v] = sh[v];
        }
        if (pos_mod(u - (K - 1), S) == 0) {  // tap row K-1 completes an output row
          const int s = pos_mod(floor_div(u - (K - 1), S), NSLOT);
          if (r >= K - 1 && r <= r_last) {
#pragma unroll
            for (int tt = 0; tt < TW; ++tt) {
              if (tt >= nvalid) continue;
              float2 y[V];
#pragma unroll
              for (int v = 0; v < V; ++v) {
                y[v] = silu2(acc[s][tt][v]);
                se[v] = __fadd2_rn(se[v], y[v]);
              }
              dw_store(dst + tt * a.C, y);
            }
            dst += a.out_row_pitch;
          }
        }
      }
    }
    // squeeze-excitation sums: one fire-and-forget 64-bit reduction per channel and thread and tile, straight to global memory
    // (40.24 fixed point: integer adds commute, the result does not depend on arrival order).  No barrier: the tile loop of a
    // warp never waits for the other warps of the CTA.
    if (a.se_sum != nullptr && active) {
      unsigned long long* sp64 = a.se_sum + (size_t)t.b * a.C + c0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        atomicAdd(sp64 + 2 * v, static_cast<unsigned long long>(__float2ll_rn(se[v].x * SE_FIX)));
        atomicAdd(sp64 + 2 * v + 1, static_cast<unsigned long long>(__float2ll_rn(se[v].y * SE_FIX)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// squeeze-excitation MLP per clip: s = sigmoid(W2 silu(W1 avg + b1) + b2)
// ---------------------------------------------------------------------------------------------------------
constexpr int SE_CL = 4;  // clips per CTA: a weight row read from L2 serves four clips (one clip per CTA re-read 442 KB of
                          // weights per clip at C = 1152); the loops keep several independent loads in flight per thread
__global__ void __launch_bounds__(256)
se_mlp_kernel(const unsigned long long* __restrict__ se_sum, float inv_hw, int B, int C, int S, const float* __restrict__ w1,
              const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ se_scale) {
  extern __shared__ float sm[];  // avg[SE_CL][C], hid[SE_CL][S]
  float* avg = sm;
  float* hid = sm + SE_CL * C;
  const int b0 = blockIdx.x * SE_CL, nb = min(SE_CL, B - b0), tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll 4
  for (int i = tid; i < SE_CL * C; i += blockDim.x) {
    const int q = i / C, c = i - q * C;
    avg[i] = q < nb ? static_cast<float>(static_cast<double>(static_cast<long long>(se_sum[(size_t)(b0 + q) * C + c])) * (1.0 / SE_FIX)) * inv_hw
                    : 0.f;
  }
  __syncthreads();
  // hidden layer: warp w owns rows w, w + 8, ...; a lane reads four consecutive weights per 16-byte load and keeps the loads of
  // three rows in flight (the 221 KB of W1 at C = 1152 stream through the SM once; with one 4-byte load per lane and step this
  // phase alone took 17 us)
  for (int j0 = warp; j0 < S; j0 += 3 * (blockDim.x / 32)) {
    float a[3][SE_CL];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < SE_CL; ++q) a[r][q] = 0.f;
#pragma unroll 2
    for (int c = 4 * lane; c < C; c += 128) {
      float4 w[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int j = j0 + r * (blockDim.x / 32);
        w[r] = j < S ? __ldg(reinterpret_cast<const float4*>(w1 + (size_t)j * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < SE_CL; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(avg + q * C + c);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          a[r][q] = fmaf(w[r].x, v.x, a[r][q]);
          a[r][q] = fmaf(w[r].y, v.y, a[r][q]);
          a[r][q] = fmaf(w[r].z, v.z, a[r][q]);
          a[r][q] = fmaf(w[r].w, v.w, a[r][q]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = j0 + r * (blockDim.x / 32);
#pragma unroll
      for (int q = 0; q < SE_CL; ++q) {
        const float v = warp_sum(a[r][q]);
        if (lane == 0 && j < S) hid[q * S + j] = silu(v + __ldg(b1 + j));
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float a[SE_CL];
    const float bias = __ldg(b2 + c);
#pragma unroll
    for (int q = 0; q < SE_CL; ++q) a[q] = bias;
    const float* wr = w2 + (size_t)c * S;
    if ((S & 3) == 0) {  // 16-byte weight loads (S = 8, 4 .. 48 in B0: every squeeze width is a multiple of 4)
#pragma unroll 12
      for (int j = 0; j < S; j += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + j));
#pragma unroll
        for (int q = 0; q < SE_CL; ++q) {
          a[q] = fmaf(w.x, hid[q * S + j], a[q]);
          a[q] = fmaf(w.y, hid[q * S + j + 1], a[q]);
          a[q] = fmaf(w.z, hid[q * S + j + 2], a[q]);
          a[q] = fmaf(w.w, hid[q * S + j + 3], a[q]);
        }
      }
    } else {
#pragma unroll 8
      for (int j = 0; j < S; ++j) {
        const float w = __ldg(wr + j);
#pragma unroll
        for (int q = 0; q < SE_CL; ++q) a[q] = fmaf(w, hid[q * S + j], a[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < SE_CL; ++q)
      if (q < nb) se_scale[(size_t)(b0 + q) * C + c] = 1.0f / (1.0f + __expf(-a[q]));
  }
}

__global__ void __launch_bounds__(256)
se_fix_to_float_kernel(const unsigned long long* __restrict__ acc, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(static_cast<double>(static_cast<long long>(acc[i])) * (1.0 / SE_FIX));
}

// x[b, p, c] *= s[b, c]   (bf16 NHWC, 8 channels per thread)
__global__ void __launch_bounds__(256)
se_apply_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ s, long long per_clip_vec, int CV, long long total_vec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_clip_vec;
    const int cv = (int)(i % CV);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + (b * CV + cv) * 8)), s1 = __ldg(reinterpret_cast<const float4*>(s + (b * CV + cv) * 8 + 4));
    uint4 raw = reinterpret_cast<uint4*>(x)[i];
    float2 a = unpack_h16(raw.x), bq = unpack_h16(raw.y), c = unpack_h16(raw.z), d = unpack_h16(raw.w);
    raw = make_uint4(pack_h16(a.x * s0.x, a.y * s0.y), pack_h16(bq.x * s0.z, bq.y * s0.w), pack_h16(c.x * s1.x, c.y * s1.y),
                     pack_h16(d.x * s1.z, d.y * s1.w));
    reinterpret_cast<uint4*>(x)[i] = raw;
  }
}

// [B, P, C] fp32 (NHWC) -> [B, C, P] fp32 (NCHW), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ in, int P, int C, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < P && c < C) ? in[((size_t)b * P + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (p < P && c < C) out[((size_t)b * C + c) * P + p] = tile[tx][r];
  }
}

// global average pool over P then Linear: logits[b, n] = bias[n] + sum_c W[n, c] mean_p x[b, p, c]   (x fp32 NHWC)
__global__ void __launch_bounds__(256)
pool_kernel(const float* __restrict__ x, int P, int C, float* __restrict__ pooled) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int p = 0; p < P; ++p) a += x[((size_t)b * P + p) * C + c];
  pooled[(size_t)b * C + c] = a / (float)P;
}
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int C, int N,
              float* __restrict__ out) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x / 32) + warp;
  if (n >= N) return;
  float a = 0.f;
  for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w + (size_t)n * C + c), x[(size_t)b * C + c], a);
  a = warp_sum(a);
  if (lane == 0) out[(size_t)b * N + n] = a + __ldg(bias + n);
}

template <typename T>
int dev_alloc(avexk_effnet* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return AVEXK_OK;
}

int fold_bn(avexk_effnet* h, const avexk_bn_params& bn, int C, float** scale, float** shift, cudaStream_t st) {
  int rc = dev_alloc(h, scale, C);
  if (rc) return rc;
  rc = dev_alloc(h, shift, C);
  if (rc) return rc;
  AVEXK_CHECK_ARG(bn.weight && bn.bias && bn.mean && bn.var, "effnet: missing BatchNorm parameter");
  bn_fold_kernel<<<ceil_div(C, 256), 256, 0, st>>>(bn.weight, bn.bias, bn.mean, bn.var, 1e-5f, C, *scale, *shift);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int copy_f32(avexk_effnet* h, float** dst, const float* src, size_t n, cudaStream_t st) {
  AVEXK_CHECK_ARG(src != nullptr, "effnet: missing parameter tensor");
  int rc = dev_alloc(h, dst, n);
  if (rc) return rc;
  AVEXK_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return AVEXK_OK;
}

int to_f16(avexk_effnet* h, __nv_bfloat16** dst, const float* src, size_t n, cudaStream_t st) {  // 16-bit storage, fp16 values
  AVEXK_CHECK_ARG(src != nullptr, "effnet: missing conv weight");
  int rc = dev_alloc(h, dst, n);
  if (rc) return rc;
  return launch_f32_to_f16(src, *dst, (long long)n, st);
}

struct Geom {
  int H, W;
};
inline int conv_out(int n, int k, int stride) { return (n + 2 * ((k - 1) / 2) - k) / stride + 1; }

}  // namespace

// Tile geometry of the TMA kernel for one layer: channels per tile (a divisor of C, multiple of 8), column groups per tile,
// band height, ring depth.  Scored by (useful input columns / loaded columns) x (busy threads / launched consumer threads).
struct DwPlan {
  int csl = 0, pv = 0, g = 0, bh = 0, ring = 0, row_bytes = 0, slot_bytes = 0, threads = 0;
};
DwPlan plan_dwconv(int B, int C, int Ho, int Wo, int k, int stride, int ch) {
  DwPlan best;
  double best_score = -1.0;
  for (int csl = 8; csl <= C && csl <= 256; csl += 8) {
    if (C % csl || csl % ch) continue;
    const int pv = csl / ch;
    if (pv > 224) continue;
    for (int g = 1; g * pv <= 224; ++g) {
      const int twc = g * DW_TW, ncw = (twc - 1) * stride + k;
      if (ncw > 256) break;
      const int row_bytes = ncw * csl * 2;
      if (row_bytes > 24 * 1024) break;
      const int n_ct = ceil_div(Wo, twc), warps = ceil_div(pv * g, 32);
      double score = ((double)Wo * stride / ((double)n_ct * ncw)) * ((double)pv * g / (warps * 32.0));
      if (pv * g < 128) score *= 0.5 + pv * g / 256.0;  // small CTAs hide less latency
      score += 1e-6 * row_bytes / 1024.0;                 // tie-break: larger boxes
      if (score > best_score) {
        best_score = score;
        best.csl = csl; best.pv = pv; best.g = g; best.row_bytes = row_bytes;
        best.slot_bytes = (row_bytes + 127) / 128 * 128;
        best.threads = (warps + 1) * 32;
      }
    }
  }
  if (best.csl == 0) return best;
  // band height: the whole column when that still leaves >= 4 tiles per resident CTA, else split (each band re-reads k - stride rows)
  const long long want = 4LL * 2 * num_sms();
  int bh = Ho;
  auto tiles = [&](int h) { return (long long)B * (C / best.csl) * ceil_div(Wo, best.g * DW_TW) * ceil_div(Ho, h); };
  while (bh > 8 && tiles(bh) < want) bh = (bh + 1) / 2;
  best.bh = bh;
  int ring = 96 * 1024 / best.slot_bytes;
  best.ring = ring > DW_MAX_RING ? DW_MAX_RING : (ring < 2 ? 2 : ring);
  return best;
}

template <int K, int S, int CH>
int launch_dwconv_tma(const __nv_bfloat16* in, int B, int H, int W, int C, int Ho, int Wo, const DwPlan& p, const float* wkk,
                      const float* scale, const float* shift, __nv_bfloat16* out, unsigned long long* se_sum, cudaStream_t st) {
  DwTmaArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.Ho = Ho; a.Wo = Wo;
  a.csl = p.csl; a.pv = p.pv; a.g = p.g; a.bh = p.bh;
  a.n_cs = C / p.csl; a.n_ct = ceil_div(Wo, p.g * DW_TW); a.n_band = ceil_div(Ho, p.bh);
  a.n_tiles = B * a.n_ct * a.n_band;  // positions; each is worked on by n_cs CTAs
  a.ring = p.ring; a.slot_bytes = p.slot_bytes; a.row_bytes = p.row_bytes; a.out_row_pitch = Wo * C;
  a.wkk = wkk; a.scale = scale; a.shift = shift; a.out = out; a.se_sum = se_sum;
  CUtensorMap map;
  const int ncw = (p.g * DW_TW - 1) * S + K;
  int rc = make_tmap_nhwc16(&map, in, B, H, W, C, p.csl, ncw);
  if (rc) return rc;
  const size_t smem = 128 + (size_t)p.ring * p.slot_bytes + 2 * DW_MAX_RING * 8;
  static bool attr_set[64] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    AVEXK_CUDA(cudaFuncSetAttribute(dwconv_tma_kernel<K, S, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set[dev] = true;
  }
  // resident CTAs per SM: 128 registers per thread allow 512 threads; shared memory is not the limit at <= 100 KB per CTA
  int ctas_per_sm = 512 / p.threads;
  if ((size_t)ctas_per_sm * (smem + 1024) > 220 * 1024) ctas_per_sm = (int)(220 * 1024 / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int per_slice = ctas_per_sm * num_sms() / a.n_cs;  // CTAs per channel slice
  if (per_slice < 1) per_slice = 1;
  if (per_slice > a.n_tiles) per_slice = a.n_tiles;
  const int grid = per_slice * a.n_cs;
  dwconv_tma_kernel<K, S, CH><<<grid, p.threads, smem, st>>>(map, a);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

int launch_dwconv(const __nv_bfloat16* in, int B, int H, int W, int C, int k, int stride, const float* wkk, const float* scale,
                  const float* shift, __nv_bfloat16* out, unsigned long long* se_sum, cudaStream_t st) {
  AVEXK_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && (k == 3 || k == 5) && (stride == 1 || stride == 2), "dwconv: unsupported C=%d k=%d stride=%d", C, k, stride);
  if (B == 0) return AVEXK_OK;
  const int Ho = conv_out(H, k, stride), Wo = conv_out(W, k, stride);
  // AVEXK_DW_OLD=1 (measurement switch) runs the per-output-row kernel, which is also the fallback for shapes the TMA tiling
  // cannot express
  static const int use_old = [] { const char* e = getenv("AVEXK_DW_OLD"); return e ? atoi(e) : 0; }();
  if (!use_old) {
    // 3x3: four channels per thread (8-byte shared loads, half the threads per tile); 5x5: two (the 25 weight pairs + 20
    // accumulator pairs of four channels would not fit the register file at two CTAs per SM)
    const int ch = k == 3 ? 4 : 2;
    const DwPlan p = plan_dwconv(B, C, Ho, Wo, k, stride, ch);
    if (p.csl != 0) {
#define AVEXK_DW(K_, S_, CH_) launch_dwconv_tma<K_, S_, CH_>(in, B, H, W, C, Ho, Wo, p, wkk, scale, shift, out, se_sum, st)
      if (k == 3 && stride == 1) return AVEXK_DW(3, 1, 4);
      if (k == 3) return AVEXK_DW(3, 2, 4);
      if (stride == 1) return AVEXK_DW(5, 1, 2);
      return AVEXK_DW(5, 2, 2);
#undef AVEXK_DW
    }
  }
  const int nquads = Ho * ceil_div(Wo, DW_TW);
  // pixel quads per CTA: enough work per CTA to amortise its fixed cost (zeroing / flushing the squeeze-excitation sums, two
  // barriers), few enough CTAs-worth to keep every SM busy.  AVEXK_DW_QUADS overrides (measurement).
  static const int quads_env = [] { const char* e = getenv("AVEXK_DW_QUADS"); return e ? atoi(e) : 0; }();
  int qpc = quads_env > 0 ? quads_env : 256;
  const int in_flight = 256 / (C / 8) > 0 ? 256 / (C / 8) : 1;
  if (qpc < in_flight) qpc = in_flight;
  dim3 grid(ceil_div(nquads, qpc), B);
  const size_t sm = C * sizeof(unsigned long long);
  if (k == 3 && stride == 1) dwconv_kernel<3, 1><<<grid, 256, sm, st>>>(in, H, W, C, Ho, Wo, wkk, scale, shift, out, se_sum, qpc);
  else if (k == 3) dwconv_kernel<3, 2><<<grid, 256, sm, st>>>(in, H, W, C, Ho, Wo, wkk, scale, shift, out, se_sum, qpc);
  else if (stride == 1) dwconv_kernel<5, 1><<<grid, 256, sm, st>>>(in, H, W, C, Ho, Wo, wkk, scale, shift, out, se_sum, qpc);
  else dwconv_kernel<5, 2><<<grid, 256, sm, st>>>(in, H, W, C, Ho, Wo, wkk, scale, shift, out, se_sum, qpc);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

// 1x1 convolution dispatch: the memory-bound kernel (pointwise.cu) whenever the request has an fp16 output (the raw pre-BN copy
// a forward hook asks for rides along); the general GEMM otherwise (fp32 output: the head convolution).  AVEXK_PW=0 (measurement switch) forces the general GEMM.  A squeeze-excitation scale on the A
// operand is fused by the pointwise kernel; on the general path it is applied in place first.
int conv1x1_any(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                const __nv_bfloat16* res, const float* se_scale, int hw, float* raw_out, void* out, int out_16bit, cudaStream_t st) {
  static const int use_pw = [] { const char* e = getenv("AVEXK_PW"); return e ? atoi(e) : 1; }();
  if (use_pw && pointwise_supported(N, K, out, out_16bit))
    return pointwise_launch(A, W, M, N, K, scale, shift, silu, res, se_scale, hw, raw_out, out, st);
  if (se_scale != nullptr && M > 0) {
    const long long per_clip_vec = (long long)hw * (K / 8), total = (long long)M * (K / 8);
    int grid = ceil_div(total, 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    se_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(A)), se_scale, per_clip_vec, K / 8, total);
    AVEXK_LAUNCH_CHECK();
  }
  return conv1x1_launch(A, W, M, N, K, scale, shift, silu, res, raw_out, out, out_16bit, st);
}

}  // namespace avexk

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" int avexk_conv1x1_f16(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift,
                                  int silu, const void* res_bf16, float* raw_out, void* out, int out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && (out || raw_out) && M >= 0, "avexk_conv1x1_f16: null argument");
  return conv1x1_any(A, W, M, N, K, scale, shift, silu, reinterpret_cast<const __nv_bfloat16*>(res_bf16), nullptr, 0, raw_out, out,
                     out_bf16, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int avexk_conv1x1_se_f16(const void* A, const float* se_scale, int rows_per_clip, const void* W, int M, int N, int K,
                                     const float* scale, const float* shift, const void* res_f16, void* out, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(A && W && out && se_scale && M >= 0 && rows_per_clip > 0, "avexk_conv1x1_se_f16: bad argument");
  AVEXK_CHECK_ARG(pointwise_supported(N, K, out, 1), "avexk_conv1x1_se_f16: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  return pointwise_launch(A, W, M, N, K, scale, shift, 0, reinterpret_cast<const __nv_bfloat16*>(res_f16), se_scale, rows_per_clip, nullptr, out,
                          reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int avexk_dwconv_nhwc(const void* in_bf16, int B, int H, int W, int C, int k, int stride, const float* w_ckk,
                                 const float* scale, const float* shift, void* out_bf16, float* se_sum, void* workspace,
                                 void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(in_bf16 && w_ckk && scale && shift && out_bf16 && workspace, "avexk_dwconv_nhwc: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // workspace: [B*C fixed-point accumulators (8 bytes each)][C*k*k repacked weights (fp32)]
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(workspace);
  float* wkk = reinterpret_cast<float*>(acc + (size_t)B * C);
  dw_pack_kernel<<<ceil_div(C * k * k, 256), 256, 0, st>>>(w_ckk, C, k * k, wkk);
  AVEXK_LAUNCH_CHECK();
  if (se_sum) AVEXK_CUDA(cudaMemsetAsync(acc, 0, sizeof(unsigned long long) * (size_t)B * C, st));
  int rc = launch_dwconv(reinterpret_cast<const __nv_bfloat16*>(in_bf16), B, H, W, C, k, stride, wkk, scale, shift,
                         reinterpret_cast<__nv_bfloat16*>(out_bf16), se_sum ? acc : nullptr, st);
  if (rc || !se_sum || B == 0) return rc;
  se_fix_to_float_kernel<<<ceil_div((long long)B * C, 256), 256, 0, st>>>(acc, (long long)B * C, se_sum);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}

extern "C" int avexk_effnet_create(const avexk_effnet_block_cfg* blocks, int num_blocks, int stem_out, int head_out,
                                   avexk_effnet_t** out) {
  using namespace avexk;
  AVEXK_CHECK_ARG(blocks && num_blocks > 0 && out, "avexk_effnet_create: null argument");
  AVEXK_CHECK_ARG(stem_out == STEM_C, "avexk_effnet_create: stem kernel is specialised to %d output channels (got %d)", STEM_C, stem_out);
  AVEXK_CHECK_ARG(head_out % 8 == 0, "avexk_effnet_create: head width must be a multiple of 8");
  int cin = stem_out;
  for (int i = 0; i < num_blocks; ++i) {
    const avexk_effnet_block_cfg& c = blocks[i];
    AVEXK_CHECK_ARG(c.cin == cin && c.cin % 8 == 0 && c.cexp % 8 == 0 && c.cout % 8 == 0 && c.csq >= 1 && c.cexp <= 2048,
                    "avexk_effnet_create: block %d has unsupported channels (cin=%d cexp=%d cout=%d csq=%d)", i, c.cin, c.cexp, c.cout, c.csq);
    AVEXK_CHECK_ARG((c.kernel == 3 || c.kernel == 5) && (c.stride == 1 || c.stride == 2), "avexk_effnet_create: block %d kernel/stride unsupported", i);
    cin = c.cout;
  }
  auto* h = new avexk_effnet();
  h->cfg.assign(blocks, blocks + num_blocks);
  h->blocks.resize(num_blocks);
  h->stem_out = stem_out;
  h->head_in = cin;
  h->head_out = head_out;
  *out = h;
  return AVEXK_OK;
}

extern "C" void avexk_effnet_destroy(avexk_effnet_t* h) {
  if (!h) return;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

extern "C" int avexk_effnet_load_weights(avexk_effnet_t* h, const avexk_effnet_weights* w, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && w && w->blocks, "avexk_effnet_load_weights: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->loaded = false;
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  AVEXK_CHECK_ARG(w->stem_w != nullptr, "avexk_effnet_load_weights: stem weight missing");
  TRY(dev_alloc(h, &h->stem_w, 9 * h->stem_out));
  stem_pack_kernel<<<ceil_div(9 * h->stem_out, 256), 256, 0, st>>>(w->stem_w, h->stem_out, h->stem_w);
  AVEXK_LAUNCH_CHECK();
  TRY(fold_bn(h, w->stem_bn, h->stem_out, &h->stem_scale, &h->stem_shift, st));
  for (size_t i = 0; i < h->cfg.size(); ++i) {
    const avexk_effnet_block_cfg& c = h->cfg[i];
    const avexk_effnet_block_weights& s = w->blocks[i];
    avexk_effnet::Block& b = h->blocks[i];
    b = avexk_effnet::Block();
    if (c.cexp != c.cin || s.expand_w != nullptr) {
      TRY(to_f16(h, &b.expand_w, s.expand_w, (size_t)c.cexp * c.cin, st));
      TRY(fold_bn(h, s.expand_bn, c.cexp, &b.expand_scale, &b.expand_shift, st));
    }
    AVEXK_CHECK_ARG(s.dw_w != nullptr, "avexk_effnet_load_weights: block %zu depthwise weight missing", i);
    TRY(dev_alloc(h, &b.dw_w, (size_t)c.cexp * c.kernel * c.kernel));
    dw_pack_kernel<<<ceil_div(c.cexp * c.kernel * c.kernel, 256), 256, 0, st>>>(s.dw_w, c.cexp, c.kernel * c.kernel, b.dw_w);
    AVEXK_LAUNCH_CHECK();
    TRY(fold_bn(h, s.dw_bn, c.cexp, &b.dw_scale, &b.dw_shift, st));
    TRY(copy_f32(h, &b.se1_w, s.se1_w, (size_t)c.csq * c.cexp, st));
    TRY(copy_f32(h, &b.se1_b, s.se1_b, c.csq, st));
    TRY(copy_f32(h, &b.se2_w, s.se2_w, (size_t)c.cexp * c.csq, st));
    TRY(copy_f32(h, &b.se2_b, s.se2_b, c.cexp, st));
    TRY(to_f16(h, &b.proj_w, s.proj_w, (size_t)c.cout * c.cexp, st));
    TRY(fold_bn(h, s.proj_bn, c.cout, &b.proj_scale, &b.proj_shift, st));
  }
  TRY(to_f16(h, &h->head_w, w->head_w, (size_t)h->head_out * h->head_in, st));
  TRY(fold_bn(h, w->head_bn, h->head_out, &h->head_scale, &h->head_shift, st));
  h->num_classes = 0;
  if (w->cls_w != nullptr && w->num_classes > 0) {
    TRY(copy_f32(h, &h->cls_w, w->cls_w, (size_t)w->num_classes * h->head_out, st));
    TRY(copy_f32(h, &h->cls_b, w->cls_b, w->num_classes, st));
    h->num_classes = w->num_classes;
  }
#undef TRY
  AVEXK_CUDA(cudaStreamSynchronize(st));
  h->loaded = true;
  return AVEXK_OK;
}

namespace avexk {
namespace {
struct EffPlan {
  size_t act, exp, dw, raw, se, total;
  int Hf, Wf;
};
EffPlan effnet_plan(const avexk_effnet* h, int B, int H0, int W0) {
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  EffPlan p{};
  int H = conv_out(H0, 3, 2), W = conv_out(W0, 3, 2);
  size_t act = (size_t)H * W * h->stem_out, ex = 0, dw = 0, raw = 0, se = 0;
  for (const auto& c : h->cfg) {
    ex = std::max(ex, (size_t)H * W * c.cexp);
    const int Ho = conv_out(H, c.kernel, c.stride), Wo = conv_out(W, c.kernel, c.stride);
    dw = std::max(dw, (size_t)Ho * Wo * c.cexp);
    act = std::max(act, (size_t)Ho * Wo * c.cout);
    raw = std::max(raw, (size_t)Ho * Wo * c.cout);
    se = std::max(se, (size_t)c.cexp);
    H = Ho;
    W = Wo;
  }
  raw = std::max(raw, (size_t)H * W * h->head_out);
  p.Hf = H;
  p.Wf = W;
  p.act = al(act * B * 2);
  p.exp = al(ex * B * 2);
  p.dw = al(dw * B * 2);
  p.raw = al(raw * B * 4);
  p.se = al(se * B * 8);  // fixed-point SE accumulators (8 bytes); the fp32 SE scales use half of a region
  p.total = 2 * p.act + p.exp + p.dw + 2 * p.raw + 2 * p.se + al((size_t)B * h->head_out * 4) + 4096;
  return p;
}
}  // namespace
}  // namespace avexk

extern "C" int avexk_effnet_out_hw(const avexk_effnet_t* h, int H0, int W0, int* Hf, int* Wf) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && Hf && Wf && H0 > 0 && W0 > 0, "avexk_effnet_out_hw: bad argument");
  const EffPlan p = effnet_plan(h, 1, H0, W0);
  *Hf = p.Hf;
  *Wf = p.Wf;
  return AVEXK_OK;
}

extern "C" size_t avexk_effnet_workspace_bytes(const avexk_effnet_t* h, int B, int H0, int W0) {
  if (!h || B <= 0 || H0 <= 0 || W0 <= 0) return 0;
  return avexk::effnet_plan(h, B, H0, W0).total;
}

extern "C" int avexk_effnet_forward(avexk_effnet_t* h, const float* mel, const void* minmax, int B, int H0, int W0,
                                    float* features_nchw, float* logits, float* const* hook_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && h->loaded, "avexk_effnet_forward: weights not loaded");
  AVEXK_CHECK_ARG(mel && workspace && B > 0 && H0 > 0 && W0 > 0, "avexk_effnet_forward: bad argument");
  AVEXK_CHECK_ARG(features_nchw || logits || hook_out, "avexk_effnet_forward: no output requested");
  AVEXK_CHECK_ARG(!logits || h->num_classes > 0, "avexk_effnet_forward: logits requested but no classifier was loaded");
  AVEXK_CHECK_ARG(B <= 65535, "avexk_effnet_forward: B=%d exceeds grid.y", B);
  const EffPlan pl = effnet_plan(h, B, H0, W0);
  AVEXK_CHECK_ARG(workspace_bytes >= pl.total, "avexk_effnet_forward: workspace too small (%zu < %zu)", workspace_bytes, pl.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* p = reinterpret_cast<char*>(workspace);
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(p); p += pl.act;
  __nv_bfloat16* act2 = reinterpret_cast<__nv_bfloat16*>(p); p += pl.act;
  __nv_bfloat16* ex = reinterpret_cast<__nv_bfloat16*>(p); p += pl.exp;
  __nv_bfloat16* dw = reinterpret_cast<__nv_bfloat16*>(p); p += pl.dw;
  float* raw = reinterpret_cast<float*>(p); p += pl.raw;
  float* raw2 = reinterpret_cast<float*>(p); p += pl.raw;
  unsigned long long* se_sum = reinterpret_cast<unsigned long long*>(p); p += pl.se;
  float* se_scale = reinterpret_cast<float*>(p); p += pl.se;
  float* pooled = reinterpret_cast<float*>(p);
  const int nb = (int)h->cfg.size();
  int rc;
  // AVEXK_DEBUG_SYNC=1 (bring-up aid): synchronise after every launch and name the launch that failed or never returned
  static const int dbg_sync = [] { const char* e = getenv("AVEXK_DEBUG_SYNC"); return e ? atoi(e) : 0; }();
  auto dbg = [&](const char* what, int layer) -> int {
    if (!dbg_sync) return AVEXK_OK;
    fprintf(stderr, "[avexk] effnet layer %d %s ...", layer, what);
    fflush(stderr);
    const cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, " %s\n", e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    if (e != cudaSuccess) {
      set_error("effnet layer %d %s: %s", layer, what, cudaGetErrorString(e));
      return AVEXK_ECUDA;
    }
    return AVEXK_OK;
  };
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  auto to_nchw = [&](const float* src, int P, int C, float* dst) -> int {
    dim3 grid(ceil_div(P, 32), ceil_div(C, 32), B);
    nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(src, P, C, dst);
    AVEXK_LAUNCH_CHECK();
    return AVEXK_OK;
  };
  // ---- stem -------------------------------------------------------------------------------------------------------
  int H = conv_out(H0, 3, 2), W = conv_out(W0, 3, 2);
  {
    dim3 grid(ceil_div((long long)H * W, 256), B);
    stem_kernel<<<grid, 256, 0, st>>>(mel, reinterpret_cast<const unsigned*>(minmax), B, H0, W0, H, W, h->stem_w, h->stem_scale,
                                      h->stem_shift, act, hook_out ? hook_out[0] : nullptr);
    AVEXK_LAUNCH_CHECK();
  }
  // ---- MBConv blocks ----------------------------------------------------------------------------------------------
  for (int i = 0; i < nb; ++i) {
    const avexk_effnet_block_cfg& c = h->cfg[i];
    const avexk_effnet::Block& b = h->blocks[i];
    const long long M = (long long)B * H * W;
    const __nv_bfloat16* dw_in = act;
    if (b.expand_w != nullptr) {
      TRY(conv1x1_any(act, b.expand_w, (int)M, c.cexp, c.cin, b.expand_scale, b.expand_shift, 1, nullptr, nullptr, 0, nullptr, ex, 1, st));
      TRY(dbg("expand", i));
      dw_in = ex;
    }
    const int Ho = conv_out(H, c.kernel, c.stride), Wo = conv_out(W, c.kernel, c.stride);
    AVEXK_CUDA(cudaMemsetAsync(se_sum, 0, sizeof(unsigned long long) * (size_t)B * c.cexp, st));
    TRY(launch_dwconv(dw_in, B, H, W, c.cexp, c.kernel, c.stride, b.dw_w, b.dw_scale, b.dw_shift, dw, se_sum, st));
    TRY(dbg("depthwise", i));
    if (SE_CL * (c.cexp + c.csq) * sizeof(float) > 48 * 1024)  // (only wider variants than B0 get here)
      AVEXK_CUDA(cudaFuncSetAttribute(se_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    se_mlp_kernel<<<ceil_div(B, SE_CL), 256, SE_CL * (c.cexp + c.csq) * sizeof(float), st>>>(se_sum, 1.0f / (float)(Ho * Wo), B, c.cexp, c.csq,
                                                                                             b.se1_w, b.se1_b, b.se2_w, b.se2_b, se_scale);
    AVEXK_LAUNCH_CHECK();
    const long long Mo = (long long)B * Ho * Wo;
    const bool use_res = c.stride == 1 && c.cin == c.cout;
    float* hook = hook_out ? hook_out[1 + i] : nullptr;
    // project conv; the squeeze-excitation rescale of its input rides on the A operand (or runs in place first on the hook path)
    TRY(conv1x1_any(dw, b.proj_w, (int)Mo, c.cout, c.cexp, b.proj_scale, b.proj_shift, 0, use_res ? act : nullptr, se_scale, Ho * Wo,
                    hook ? raw : nullptr, act2, 1, st));
    TRY(dbg("project", i));
    if (hook) TRY(to_nchw(raw, Ho * Wo, c.cout, hook));
    std::swap(act, act2);
    H = Ho;
    W = Wo;
  }
  // ---- head -------------------------------------------------------------------------------------------------------
  {
    const long long M = (long long)B * H * W;
    float* hook = hook_out ? hook_out[nb + 1] : nullptr;
    TRY(conv1x1_launch(act, h->head_w, (int)M, h->head_out, h->head_in, h->head_scale, h->head_shift, 1, nullptr,
                       hook ? raw : nullptr, raw2, 0, st));
    if (hook) TRY(to_nchw(raw, H * W, h->head_out, hook));
    if (features_nchw) TRY(to_nchw(raw2, H * W, h->head_out, features_nchw));
    if (logits) {
      dim3 g1(ceil_div(h->head_out, 256), B);
      pool_kernel<<<g1, 256, 0, st>>>(raw2, H * W, h->head_out, pooled);
      AVEXK_LAUNCH_CHECK();
      dim3 g2(ceil_div(h->num_classes, 8), B);
      linear_kernel<<<g2, 256, 0, st>>>(pooled, h->cls_w, h->cls_b, h->head_out, h->num_classes, logits);
      AVEXK_LAUNCH_CHECK();
    }
  }
#undef TRY
  return AVEXK_OK;
}
