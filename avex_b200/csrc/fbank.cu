// fbank.cu -- fused Kaldi log-mel filterbank for sm_100a (v2: two frames per 16-lane group, packed fp32x2 arithmetic).
//
// Replaces avex/models/beats/beats.py:120-163 (`_BatchedFbank.forward`: unfold, DC removal, pre-emphasis,
// Povey window, zero-pad, rfft, |.|^2, mel matmul, clamp+log) and the affine of beats.py:323, in ONE kernel:
// the [B,F,512] frame tensor and the [B,F,257] spectrum the reference materialises never exist.  With the patch-operand
// output mode the [B,F,128] fbank itself never exists either: the kernel writes the bf16 [hi|lo|hi] operand of the
// 16x16 patch-embedding GEMM (beats.py:349-352) directly.
//
// Why it looks the way it does (the v1 kernel was FP32-issue bound at 815 warp instructions per frame):
//   * TWO frames per 16-lane group, every value a float2 (frame A, frame B): all arithmetic is FFMA2 / FADD2 / FMUL2, and
//     sm_100 takes a scalar as the broadcast operand of a packed instruction, so window / twiddle / mel constants cost no
//     registers and no extra instructions.  ~420 warp instructions per PAIR of frames.
//   * pre-emphasis moved to the staging pass: d[i] = x[i] - 0.97 x[i-1] is computed ONCE per sample (frames overlap 2.5x)
//     and stored interleaved, sD[i] = (d[i], d[i + 160]), so that one 16-byte shared load yields the (re, im) samples of
//     both frames; the per-frame DC removal becomes y[n] = (d[n] - 0.03 mu) w[n] (mu from partial sums of the staging pass).
//   * 512-pt real FFT = 256-pt complex FFT as 16 x 16 Cooley-Tukey: both radix-16 stages in registers, ONE padded
//     shared-memory transposition (16-byte accesses, conflict-free), conjugate partners for the real-FFT split fetched
//     with warp shuffles instead of a second trip through shared memory.
//   * sparse triangular mel projection (504 non-zeros instead of the reference's dense 257x128 SGEMM), lane t owns bins
//     t, t+16, ..., t+112 (balanced filter lengths), one special-function op per log.
//   * persistent CTAs (3 per SM) loop over (clip, 16-frame chunk) items; tables are loaded to shared memory once.
// Roofline: HBM; algorithmic bytes per clip = 4*T + 4*F*128.  The kernel is bound by the shared-memory /
// L1 data path (~15 KB per frame), not by HBM: see DESIGN.md section 4.
#include <float.h>
#include <math.h>

#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace avexk {
namespace {

constexpr int WIN = 400, HOP = 160, NMEL = 128, NBIN = 257;
constexpr int THREADS = 128;
constexpr int GROUPS = THREADS / 16;        // lane groups per CTA, one frame PAIR each
constexpr int FPC = 2 * GROUPS;             // 16 frames per chunk == one row of 16x16 patches
constexpr int SEG = (FPC - 1) * HOP + WIN;  // 2800 samples feed one chunk
constexpr int NPAIR = SEG - HOP;            // 2640 interleaved entries (d[i], d[i+160])
constexpr int NQ = NPAIR / 4;               // 660 staging units of four entries
constexpr int NPS = SEG / 4;                // 700 partial sums of four samples
constexpr int XROW = 17;                    // float2 per row of the transposition buffer (16 + 1 pad)
constexpr int XG = 16 * XROW;               // float2 per group (2176 B): one plane (re, then im) of the 16 x 16 transposition,
                                            // then the conjugate-partner exchange, then the power spectrum
constexpr int MAXNZ = 10;                   // longest mel filter at n_fft 512 / 128 bins
constexpr float LN2 = 0.69314718055994530942f;
constexpr int PPAD = 272;                   // power-spectrum entries per group incl. ELL overrun pad
// Mel weights, ELL by lane: lane t owns bins t + 16 i; its taps are stored contiguously (row i after row i-1) so that four
// weights arrive per 16-byte load.  STD = the filter lengths of the reference geometry (20 Hz .. 8 kHz, 128 bins, n_fft 512);
// any other table (longer filters) runs the generic profile with MAXNZ taps per bin.
__host__ __device__ constexpr int mel_std_len(int i) {  // 37 taps per lane in total
  return i < 3 ? 2 : i == 3 ? 3 : i == 4 ? 4 : i == 5 ? 6 : i == 6 ? 8 : 10;
}
constexpr int MEL_STD_STRIDE = 44;  // floats per lane: 37 taps padded; 44 t mod 32 = 12 t -> conflict-free 16-byte loads
constexpr int MEL_GEN_STRIDE = 84;  // 8 * MAXNZ taps padded; 84 t mod 32 = 20 t

constexpr int OFF_D = 0;                                // float2 [NPAIR]
constexpr int OFF_PS = OFF_D + NPAIR * 8;               // float  [NPS]
constexpr int OFF_X0 = OFF_PS + NPS * 4;                // float  [FPC]  first sample of every frame (n = 0 fix-up)
constexpr int OFF_XCH = OFF_X0 + FPC * 4;               // float2 [GROUPS * XG]
constexpr int OFF_WIN = OFF_XCH + GROUPS * XG * 8;      // float  [WIN]
constexpr int OFF_TW1 = OFF_WIN + WIN * 4;              // float2 [256]
constexpr int OFF_TW2 = OFF_TW1 + 256 * 8;              // float2 [136]
constexpr int OFF_MW = OFF_TW2 + 136 * 8;               // float  [16 * MEL_*_STRIDE]
constexpr int OFF_MS = OFF_MW + 16 * MEL_GEN_STRIDE * 4;  // int  [NMEL]
constexpr int SMEM_BYTES = OFF_MS + NMEL * 4;
static_assert(OFF_XCH % 16 == 0 && OFF_TW1 % 8 == 0 && OFF_TW2 % 8 == 0 && OFF_MW % 16 == 0, "shared-memory carve alignment");
static_assert(PPAD <= XG, "power spectrum must fit the group's transposition buffer");
static_assert(4 * (SMEM_BYTES + 1024) <= 232448, "four CTAs per SM");

enum { OUT_F32 = 0, OUT_BF16 = 1, OUT_PATCH3 = 2 };

struct Tables {
  const float* window;   // [400], pre-multiplied by 0.5 (folds the 1/2 of the real-FFT split; exact)
  const float2* tw1;     // [16][16]: tw1[q*16+t] = exp(-2 pi i t q / 256)
  const float2* tw2;     // [136]: exp(-2 pi i k / 512), k = 0..128
  const float* melw;     // [16][stride] ELL weights by lane: lane t, then bin row i (bin t + 16 i), then tap m
  const int* melstart;   // [128] first FFT bin of each filter
  int mel_std;           // 1: every filter fits the MEL_STD profile, 0: generic profile
  int win0_nonzero;      // window[0] != 0: the replicate-padded first sample needs its fix-up (never for povey / hanning)
};

// ---- packed (frame A, frame B) arithmetic --------------------------------------------------------------------------
typedef float2 P2;
__device__ __forceinline__ P2 padd(P2 a, P2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ P2 psub(P2 a, P2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ P2 pmul(P2 a, P2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ P2 pmuls(P2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ P2 pfma(P2 a, P2 b, P2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ P2 pfmas(P2 a, float s, P2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ P2 pneg(P2 a) { return make_float2(-a.x, -a.y); }

struct C2 {  // one complex value of each of the two frames
  P2 re, im;
};
__device__ __forceinline__ C2 cadd(const C2& a, const C2& b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ C2 csub(const C2& a, const C2& b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
// a * (wr + i wi), scalar twiddle shared by both frames
__device__ __forceinline__ C2 cmuls(const C2& a, float wr, float wi) {
  return {pfmas(a.re, wr, pmuls(a.im, -wi)), pfmas(a.re, wi, pmuls(a.im, wr))};
}

// forward radix-4 DFT, in place: (a0,a1,a2,a3) -> (y0,y1,y2,y3), w4 = -i
__device__ __forceinline__ void fft4(C2& a0, C2& a1, C2& a2, C2& a3) {
  const C2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = {padd(t1.re, t3.im), psub(t1.im, t3.re)};  // t1 - i t3
  a3 = {psub(t1.re, t3.im), padd(t1.im, t3.re)};  // t1 + i t3
}
// the same with a3 == 0 (the zero tail of the 400-sample frame in the first radix-4 column)
__device__ __forceinline__ void fft4_z3(C2& a0, C2& a1, C2& a2, C2& a3) {
  const C2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
  const C2 t = a1;
  a0 = cadd(t0, t);
  a2 = csub(t0, t);
  a1 = {padd(t1.re, t.im), psub(t1.im, t.re)};
  a3 = {psub(t1.re, t.im), padd(t1.im, t.re)};
}

// forward 16-point DFT in registers.  On return V[q] (natural order) lives in v[IDX16(q)].
#define IDX16(q) (4 * ((q) & 3) + ((q) >> 2))
template <bool ZERO_TAIL>  // v[13], v[14], v[15] are zero on entry (first stage: samples 416..511 of the padded frame)
__device__ __forceinline__ void fft16(C2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, R = 0.70710678118654752440f;
  fft4(v[0], v[4], v[8], v[12]);
  if (ZERO_TAIL) {
    fft4_z3(v[1], v[5], v[9], v[13]);
    fft4_z3(v[2], v[6], v[10], v[14]);
    fft4_z3(v[3], v[7], v[11], v[15]);
  } else {
    fft4(v[1], v[5], v[9], v[13]);
    fft4(v[2], v[6], v[10], v[14]);
    fft4(v[3], v[7], v[11], v[15]);
  }
  // v[j0 + 4 q0] *= w16^(j0 q0),  w16^m = (cos(pi m / 8), -sin(pi m / 8))
  v[1 + 4] = cmuls(v[1 + 4], C1, -S1);                                  // m = 1
  v[1 + 8] = {pmuls(padd(v[1 + 8].re, v[1 + 8].im), R), pmuls(psub(v[1 + 8].im, v[1 + 8].re), R)};      // m = 2: (R, -R)
  v[1 + 12] = cmuls(v[1 + 12], S1, -C1);                                // m = 3
  v[2 + 4] = {pmuls(padd(v[2 + 4].re, v[2 + 4].im), R), pmuls(psub(v[2 + 4].im, v[2 + 4].re), R)};      // m = 2
  v[2 + 8] = {v[2 + 8].im, pneg(v[2 + 8].re)};                          // m = 4: * (-i)
  v[2 + 12] = {pmuls(psub(v[2 + 12].im, v[2 + 12].re), R), pmuls(padd(v[2 + 12].re, v[2 + 12].im), -R)};  // m = 6: (-R, -R)
  v[3 + 4] = cmuls(v[3 + 4], S1, -C1);                                  // m = 3
  v[3 + 8] = {pmuls(psub(v[3 + 8].im, v[3 + 8].re), R), pmuls(padd(v[3 + 8].re, v[3 + 8].im), -R)};      // m = 6
  v[3 + 12] = cmuls(v[3 + 12], -C1, S1);                                // m = 9
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) fft4(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3]);
}

template <int MODE, bool MELSTD>
__global__ void __launch_bounds__(THREADS, 4)
fbank_kernel(const float* __restrict__ wav, long long stride, int T, int F, int Fout, int chunks, int n_items, float prescale,
             float nmean, float nscale, Tables tb, void* __restrict__ out, long long out_rows_per_clip,
             double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem[];
  float2* sD = reinterpret_cast<float2*>(smem + OFF_D);
  float* sPS = reinterpret_cast<float*>(smem + OFF_PS);
  float* sX0 = reinterpret_cast<float*>(smem + OFF_X0);
  float2* sXch = reinterpret_cast<float2*>(smem + OFF_XCH);
  float* sWin = reinterpret_cast<float*>(smem + OFF_WIN);
  float2* sTw1 = reinterpret_cast<float2*>(smem + OFF_TW1);
  float2* sTw2 = reinterpret_cast<float2*>(smem + OFF_TW2);
  float* sMW = reinterpret_cast<float*>(smem + OFF_MW);
  int* sMS = reinterpret_cast<int*>(smem + OFF_MS);
  constexpr int MSTRIDE = MELSTD ? MEL_STD_STRIDE : MEL_GEN_STRIDE;

  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 4, t = tid & 15;

  // the 2^15 scaling of beats.py:322 is folded into the window (a power of two: bit-identical to scaling the samples)
  for (int i = tid; i < WIN; i += THREADS) sWin[i] = tb.window[i] * prescale;
  for (int i = tid; i < 256; i += THREADS) sTw1[i] = tb.tw1[i];
  for (int i = tid; i < 136; i += THREADS) sTw2[i] = tb.tw2[i];
  for (int i = tid; i < 16 * MSTRIDE; i += THREADS) sMW[i] = tb.melw[i];
  for (int i = tid; i < NMEL; i += THREADS) sMS[i] = tb.melstart[i];

  const bool vec_ok = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(wav) & 15) == 0);

  // Staging loads: all 12 float4 (+ the warp-edge samples) of a chunk are issued before the first one is consumed, ahead of the
  // CTA barrier (with one dependent load per loop iteration the staging pass took half of all stall samples; holding the next
  // chunk's loads across the mel phase instead spilled).
  constexpr int NIT = (NQ + THREADS - 1) / THREADS;
  float4 A[NIT], Bv[NIT];
  float ea[NIT], eb[NIT];  // previous sample at the warp edge (lane 0 only)
  auto prefetch = [&](int item) {
    const int b = item / chunks, c = item - b * chunks;
    const long long s0 = (long long)c * FPC * HOP;
    const float* src = wav + (long long)b * stride;
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int u = k * THREADS + tid;
      const bool act = u < NQ;
      const long long sa = s0 + 4 * u, sb = sa + HOP;
      A[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      Bv[k] = A[k];
      ea[k] = eb[k] = 0.f;
      if (act) {
        if (vec_ok && sb + 3 < T) {
          A[k] = __ldg(reinterpret_cast<const float4*>(src + sa));
          Bv[k] = __ldg(reinterpret_cast<const float4*>(src + sb));
        } else {
          auto ld = [&](long long s) { return s < T ? __ldg(src + s) : 0.f; };
          A[k] = make_float4(ld(sa), ld(sa + 1), ld(sa + 2), ld(sa + 3));
          Bv[k] = make_float4(ld(sb), ld(sb + 1), ld(sb + 2), ld(sb + 3));
        }
        if (lane == 0) {
          ea[k] = sa > 0 && sa - 1 < T ? __ldg(src + sa - 1) : 0.f;
          eb[k] = sb - 1 < T ? __ldg(src + sb - 1) : 0.f;
        }
      }
    }
  };
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / chunks, c = item - b * chunks;
    const int f0 = c * FPC;
    const long long s0 = (long long)f0 * HOP;
    const float* src = wav + (long long)b * stride;
    prefetch(item);   // issued ahead of the barrier: in flight while the slowest warp of the CTA finishes the previous chunk
    __syncthreads();  // the previous chunk's readers of sD / sPS / sX0 are done (and the tables are in place)

    // ---- staging: d[i] = x[i] - 0.97 x[i-1] once per sample, stored as (d[i], d[i+160]); partial sums for the means --------
#pragma unroll
    for (int k = 0; k < NIT; ++k) {  // uniform trip count: the shuffles below need the whole warp
      const int u = k * THREADS + tid;
      const bool act = u < NQ;
      const long long sa = s0 + 4 * u;
      // previous sample: the neighbour lane's .w, except at the warp edge (and the clip start: replicate, beats.py:143)
      float pa = __shfl_up_sync(0xffffffffu, A[k].w, 1);
      float pb = __shfl_up_sync(0xffffffffu, Bv[k].w, 1);
      if (lane == 0) {
        pa = sa > 0 ? ea[k] : A[k].x;
        pb = eb[k];
      }
      if (act) {
        const P2 x0 = make_float2(A[k].x, Bv[k].x), x1 = make_float2(A[k].y, Bv[k].y), x2 = make_float2(A[k].z, Bv[k].z), x3 = make_float2(A[k].w, Bv[k].w);
        const P2 d0 = pfmas(make_float2(pa, pb), -0.97f, x0), d1 = pfmas(x0, -0.97f, x1);  // beats.py:144
        const P2 d2 = pfmas(x1, -0.97f, x2), d3 = pfmas(x2, -0.97f, x3);
        // 16-byte unit e (two entries) lives at e ^ ((e >> 3) & 1): the eight lanes of a store phase write units 2u (then 2u+1)
        // of two aligned groups of eight -- without the swizzle they would fall on the same banks pairwise; readers below fetch
        // aligned groups of eight consecutive units, for which the swizzle is a per-lane constant
        const int e0 = 2 * u, e1 = 2 * u + 1;
        reinterpret_cast<float4*>(sD)[e0 ^ ((e0 >> 3) & 1)] = make_float4(d0.x, d0.y, d1.x, d1.y);
        reinterpret_cast<float4*>(sD)[e1 ^ ((e1 >> 3) & 1)] = make_float4(d2.x, d2.y, d3.x, d3.y);
        const P2 ps = padd(padd(x0, x1), padd(x2, x3));
        sPS[u] = ps.x;
        if (u >= NQ - HOP / 4) sPS[u + HOP / 4] = ps.y;
      }
    }
    if (tb.win0_nonzero && tid < FPC) {
      const long long s = s0 + (long long)tid * HOP;
      sX0[tid] = s < T ? __ldg(src + s) : 0.f;
    }
    __syncthreads();

    // ---- one frame pair per 16 lanes ---------------------------------------------------------------------------------
    const int flA = 2 * g;  // frames f0 + flA and f0 + flA + 1
    // per-frame mean (beats.py:140): 100 partial sums per frame, 16 lanes, fixed order
    P2 dc;
    {
      P2 sm = make_float2(0.f, 0.f);
      const float* pA = sPS + 40 * flA;
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int i = t + 16 * r;
        if (i < 100) sm = padd(sm, make_float2(pA[i], pA[i + 40]));
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
        sm = padd(sm, make_float2(__shfl_xor_sync(0xffffffffu, sm.x, o), __shfl_xor_sync(0xffffffffu, sm.y, o)));
      dc = pmuls(sm, -0.03f / 400.0f);  // y[n] = (x[n] - mu) - 0.97 (x[n-1] - mu) = d[n] - 0.03 mu
    }

    C2 v[16];
    {
      // unit index 80 flA + 16 j + t: bit 3 of it is bit 3 of t, so the staging swizzle is `t ^ (t >> 3)` on the lane's offset
      const float4* dbase = reinterpret_cast<const float4*>(sD + HOP * flA) + (t ^ (t >> 3));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;  // complex point n = samples (2n, 2n+1)
        if (j < 12 || (j == 12 && t < 8)) {
          const float4 d = dbase[16 * j];  // (dA[2n], dB[2n], dA[2n+1], dB[2n+1])
          const float2 w = *reinterpret_cast<const float2*>(sWin + 2 * n);
          v[j].re = pmuls(padd(make_float2(d.x, d.y), dc), w.x);  // beats.py:144,147
          v[j].im = pmuls(padd(make_float2(d.z, d.w), dc), w.y);
        } else {
          v[j].re = make_float2(0.f, 0.f);
          v[j].im = make_float2(0.f, 0.f);
        }
      }
      if (tb.win0_nonzero && t == 0) {  // replicate pad: y[0] = 0.03 (x[0] - mu)
        const P2 x0 = make_float2(sX0[flA], sX0[flA + 1]);
        v[0].re = pmuls(pfmas(x0, 0.03f, dc), sWin[0]);
      }
    }
    // stage 1: V_t[q] = sum_j z[t+16j] w16^(jq); then twiddle w256^(tq)
    fft16<true>(v);
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      const float2 w = sTw1[q * 16 + t];
      v[IDX16(q)] = cmuls(v[IDX16(q)], w.x, w.y);
    }
    // 16 x 16 transposition through shared memory, one plane at a time (8-byte entries: the packed pair as it lies in its
    // register pair, no repacking; row stride 17 -> conflict-free both ways).  Lane t then owns q = t.
    float2* xg = sXch + g * XG;
#pragma unroll
    for (int q = 0; q < 16; ++q) xg[q * XROW + t] = v[IDX16(q)].re;
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) v[tt].re = xg[t * XROW + tt];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; ++q) xg[q * XROW + t] = v[IDX16(q)].im;
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) v[tt].im = xg[t * XROW + tt];
    __syncwarp();
    // stage 2: Z[t + 16 p] = sum_tt u[tt] w16^(tt p)
    fft16<false>(v);

    // real-FFT split on conjugate pairs (k, 256-k), k = t + 16 p, p < 8; Z is already halved through the window table.
    //   E = Z[k] + conj(Z[256-k]),  O = -i (Z[k] - conj(Z[256-k])),  X[k] = E + W^k O,  X[256-k] = conj(E - W^k O)
    // Z[256-k] lives in lane 16-t as register 15-p (lane 0 and lane 8 are their own partners; lane 0 pairs p with 16-p):
    // every lane publishes its upper eight values (two planes of 8 x 17) and reads its partner's.
    {
      float2* xre = xg;
      float2* xim = xg + 8 * XROW;
      // lane 0 is its own partner with the rows shifted by one (p pairs with 16-p): it publishes into the pad column 16 one
      // row up, so that every lane reads row 7-p at column 16-t and the 16 addresses of a load fall on 16 distinct bank pairs
      const bool l0 = t == 0;
      const int wcol = l0 ? 16 - XROW : t;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i > 0 || !l0) {
          xre[i * XROW + wcol] = v[IDX16(8 + i)].re;
          xim[i * XROW + wcol] = v[IDX16(8 + i)].im;
        }
      }
      __syncwarp();
      const int srcl = 16 - t;
      const C2 z8 = v[IDX16(8)];  // lane 0 needs its own Z[128] below
      C2 r[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {  // (p = 0 on lane 0 reads a stale slot and is replaced by A below)
        r[p].re = xre[(7 - p) * XROW + srcl];
        r[p].im = xim[(7 - p) * XROW + srcl];
      }
      __syncwarp();  // the exchange planes become the power spectrum
      float2* sP = xg;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const C2& A = v[IDX16(p)];
        C2 Bc = r[p];
        if (p == 0) {
          Bc.re = l0 ? A.re : Bc.re;
          Bc.im = l0 ? A.im : Bc.im;
        }
        const int k = t + 16 * p;
        const float2 W = sTw2[k];
        const P2 ex = padd(A.re, Bc.re), ey = psub(A.im, Bc.im);
        const P2 ox = padd(A.im, Bc.im), oy = psub(Bc.re, A.re);
        const P2 tx = pfmas(ox, W.x, pmuls(oy, -W.y)), ty = pfmas(oy, W.x, pmuls(ox, W.y));
        const P2 px = padd(ex, tx), py = padd(ey, ty), qx = psub(ex, tx), qy = psub(ey, ty);
        sP[k] = pfma(px, px, pmul(py, py));        // |X[k]|^2      beats.py:155
        sP[256 - k] = pfma(qx, qx, pmul(qy, qy));  // |X[256-k]|^2
      }
      if (l0) {  // k = 128: its own partner, W^128 = -i  ->  |X[128]|^2 = (2 re)^2 + (2 im)^2
        const P2 ex = padd(z8.re, z8.re), ox = padd(z8.im, z8.im);
        sP[128] = pfma(ex, ex, pmul(ox, ox));
      }
      if (t < PPAD - NBIN) sP[NBIN + t] = make_float2(0.f, 0.f);  // ELL overrun pad
    }
    __syncwarp();

    // ---- mel projection, log, affine, store: lane t owns bins t + 16 i -------------------------------------------------
    {
      const float2* sP = xg;
      const int fA = f0 + flA, fB = fA + 1;
      float ssum = 0.f, ssq = 0.f;
      const float4* wl = reinterpret_cast<const float4*>(sMW + t * MSTRIDE);
      float4 wq = wl[0];
      int widx = 0;  // compile-time after unrolling: position in the lane's tap list
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = t + 16 * i;
        const float2* p = sP + sMS[j];
        const int L = MELSTD ? mel_std_len(i) : MAXNZ;
        P2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < MAXNZ; ++m) {
          if (m < L) {
            const float w = (widx & 3) == 0 ? wq.x : (widx & 3) == 1 ? wq.y : (widx & 3) == 2 ? wq.z : wq.w;
            acc = pfmas(p[m], w, acc);  // beats.py:159
            ++widx;
            if ((widx & 3) == 0) wq = wl[widx >> 2];
          }
        }
        // beats.py:163, :323: log(max(mel, eps)) as one lg2.approx per value, with ln 2 and the (x - mean) * scale affine folded into
        // one packed FMA; frames past the last one are zero in the log-mel domain (eat/audio_processor.py:121-124)
        float lgA, lgB;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lgA) : "f"(fmaxf(acc.x, FLT_EPSILON)));
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lgB) : "f"(fmaxf(acc.y, FLT_EPSILON)));
        const P2 lg = make_float2(fA < F ? lgA : 0.f, fB < F ? lgB : 0.f);
        const P2 o = pfmas(lg, LN2 * nscale, make_float2(-nmean * nscale, -nmean * nscale));
        const float oA = o.x, oB = o.y;
        if (MODE == OUT_PATCH3) {
          // [hi | lo | hi] bf16 operand of the patch-embedding GEMM: row = token (b, chunk, bin row i), col = frame * 16 + bin.
          // Lane pairs swap one value so that every lane stores two adjacent columns as one 32-bit word.
          const bool odd = t & 1;
          const float got = __shfl_xor_sync(0xffffffffu, odd ? oA : oB, 1);
          const float e0 = odd ? got : oA, e1 = odd ? oB : got;  // even lane: frame A bins (t, t+1); odd lane: frame B bins (t-1, t)
          const uint32_t hi = pack_bf16(e0, e1);
          const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hi);
          const float2 hf = __bfloat1622float2(hb);
          const uint32_t lo = pack_bf16(e0 - hf.x, e1 - hf.y);
          const long long row = (long long)b * out_rows_per_clip + (long long)c * 8 + i;
          const int col = (flA + (odd ? 1 : 0)) * 16 + (t & ~1);
          uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + row * 768 + col);
          dst[0] = hi;
          dst[128] = lo;  // + 256 elements
          dst[256] = hi;  // + 512 elements
        } else {
          if (stats != nullptr) {  // per-utterance mode (nmean = 0, nscale = 1: o is the log-mel value itself)
            const float vA = fA < Fout ? oA : 0.f, vB = fB < Fout ? oB : 0.f;
            ssum += vA + vB;
            ssq = fmaf(vA, vA, fmaf(vB, vB, ssq));
          }
          const size_t ia = ((size_t)b * Fout + fA) * NMEL + j;
          if (MODE == OUT_BF16) {
            if (fA < Fout) reinterpret_cast<__nv_bfloat16*>(out)[ia] = __float2bfloat16_rn(oA);
            if (fB < Fout) reinterpret_cast<__nv_bfloat16*>(out)[ia + NMEL] = __float2bfloat16_rn(oB);
          } else {
            if (fA < Fout) reinterpret_cast<float*>(out)[ia] = oA;
            if (fB < Fout) reinterpret_cast<float*>(out)[ia + NMEL] = oB;
          }
        }
      }
      if (MODE != OUT_PATCH3 && stats != nullptr) {  // per-utterance statistics: warp shuffles, then one atomic pair per warp
        ssum = warp_sum(ssum);
        ssq = warp_sum(ssq);
        if (lane == 0) {
          atomicAdd(stats + 2 * b, (double)ssum);
          atomicAdd(stats + 2 * b + 1, (double)ssq);
        }
      }
    }
  }
}

// (x - mu) / (2 sigma), sigma unbiased; eat/audio_processor.py:132-135
__global__ void per_utt_normalise_kernel(float* __restrict__ x, const double* __restrict__ stats, long long per_clip) {
  const int b = blockIdx.y;
  const double n = (double)per_clip, S = stats[2 * b], Q = stats[2 * b + 1];
  const double mu = S / n;
  double var = (Q - S * S / n) / (n - 1.0);
  float sd = var > 0.0 ? (float)sqrt(var) : 1.0f;
  if (!(sd > 0.f)) sd = 1.0f;
  const float fmu = (float)mu, inv = 1.0f / (2.0f * sd);
  float* p = x + (long long)b * per_clip;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip; i += (long long)gridDim.x * blockDim.x)
    p[i] = (p[i] - fmu) * inv;
}

}  // namespace
}  // namespace avexk

struct avexk_fbank {
  void* dev = nullptr;  // one allocation holding all tables
  avexk::Tables tb;
};

extern "C" int avexk_fbank_num_frames(int T) { return T < avexk::WIN ? 0 : 1 + (T - avexk::WIN) / avexk::HOP; }

extern "C" int avexk_fbank_create(const float* window_host, const float* mel_fb_host, avexk_fbank_t** out) {
  using namespace avexk;
  AVEXK_CHECK_ARG(window_host && mel_fb_host && out, "avexk_fbank_create: null argument");
  std::vector<float> win(WIN), melw(16 * MEL_GEN_STRIDE, 0.f);
  std::vector<float2> tw1(256), tw2(136);
  std::vector<int> start(NMEL, 0), len(NMEL, 0);
  for (int i = 0; i < WIN; ++i) win[i] = 0.5f * window_host[i];
  for (int q = 0; q < 16; ++q)
    for (int t = 0; t < 16; ++t) {
      double a = -2.0 * M_PI * (double)(t * q) / 256.0;
      tw1[q * 16 + t] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k = 0; k < 136; ++k) {
    double a = -2.0 * M_PI * (double)k / 512.0;
    tw2[k] = make_float2((float)cos(a), (float)sin(a));
  }
  for (int j = 0; j < NMEL; ++j) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBIN; ++k)
      if (mel_fb_host[k * NMEL + j] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    if (lo < 0) { lo = 0; hi = -1; }  // empty filter (bin 3 at 20 Hz .. 8 kHz): sum = 0 -> log(eps), as the reference's matmul
    len[j] = hi - lo + 1;
    AVEXK_CHECK_ARG(len[j] <= MAXNZ, "avexk_fbank_create: mel filter %d spans %d bins (> %d)", j, len[j], MAXNZ);
    start[j] = lo;
  }
  auto* h = new avexk_fbank();
  // lane t's tap list: bin row i = 0..7 (bin t + 16 i), L_i taps each (zero padded)
  bool std_ok = true;
  for (int j = 0; j < NMEL; ++j) std_ok = std_ok && len[j] <= mel_std_len(j / 16);
  h->tb.mel_std = std_ok ? 1 : 0;
  const int mstride = std_ok ? MEL_STD_STRIDE : MEL_GEN_STRIDE;
  for (int t = 0; t < 16; ++t) {
    int pos = 0;
    for (int i = 0; i < 8; ++i) {
      const int j = 16 * i + t, L = std_ok ? mel_std_len(i) : MAXNZ;
      for (int m = 0; m < L; ++m, ++pos) melw[t * mstride + pos] = m < len[j] ? mel_fb_host[(start[j] + m) * NMEL + j] : 0.f;
    }
  }
  h->tb.win0_nonzero = window_host[0] != 0.f;
  size_t off_win = 0, off_tw = off_win + WIN * 4, off_tw2 = off_tw + 256 * 8, off_mw = off_tw2 + 136 * 8,
         off_ms = off_mw + 16 * MEL_GEN_STRIDE * 4, total = off_ms + NMEL * 4;
  cudaError_t e = cudaMalloc(&h->dev, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("avexk_fbank_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  char* d = reinterpret_cast<char*>(h->dev);
  cudaMemcpy(d + off_win, win.data(), WIN * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_tw, tw1.data(), 256 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_tw2, tw2.data(), 136 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_mw, melw.data(), 16 * MEL_GEN_STRIDE * 4, cudaMemcpyHostToDevice);
  e = cudaMemcpy(d + off_ms, start.data(), NMEL * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(h->dev);
    delete h;
    set_error("avexk_fbank_create: upload failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  h->tb.window = reinterpret_cast<const float*>(d + off_win);
  h->tb.tw1 = reinterpret_cast<const float2*>(d + off_tw);
  h->tb.tw2 = reinterpret_cast<const float2*>(d + off_tw2);
  h->tb.melw = reinterpret_cast<const float*>(d + off_mw);
  h->tb.melstart = reinterpret_cast<const int*>(d + off_ms);
  cudaFuncSetAttribute(fbank_kernel<OUT_F32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<OUT_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<OUT_PATCH3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<OUT_F32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<OUT_BF16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<OUT_PATCH3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  *out = h;
  return AVEXK_OK;
}

extern "C" void avexk_fbank_destroy(avexk_fbank_t* h) {
  if (!h) return;
  if (h->dev) cudaFree(h->dev);
  delete h;
}

namespace avexk {
namespace {
template <int MODE>
void launch(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride, int F, int Fout, int chunks, float prescale,
            float nm, float ns, void* out, long long rows_per_clip, double* stats, cudaStream_t st) {
  const long long items = (long long)B * chunks;
  const int grid = (int)(items < 4LL * num_sms() ? items : 4LL * num_sms());
  if (h->tb.mel_std)
    fbank_kernel<MODE, true><<<grid, THREADS, SMEM_BYTES, st>>>(wav, wav_stride, T, F, Fout, chunks, (int)items, prescale, nm, ns, h->tb,
                                                                 out, rows_per_clip, stats);
  else
    fbank_kernel<MODE, false><<<grid, THREADS, SMEM_BYTES, st>>>(wav, wav_stride, T, F, Fout, chunks, (int)items, prescale, nm, ns, h->tb,
                                                                  out, rows_per_clip, stats);
}
}  // namespace
}  // namespace avexk

extern "C" int avexk_fbank_forward(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride,
                                   float prescale, float norm_mean, float norm_scale, int out_frames, int per_utt,
                                   double* stats_ws, void* out, int out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && wav && out, "avexk_fbank_forward: null argument");
  AVEXK_CHECK_ARG(B >= 0 && T >= 0 && wav_stride >= T, "avexk_fbank_forward: bad shape B=%d T=%d stride=%lld", B, T, wav_stride);
  const int F = avexk_fbank_num_frames(T);
  const int Fout = out_frames > 0 ? out_frames : F;
  if (B == 0 || Fout == 0) return AVEXK_OK;
  AVEXK_CHECK_ARG(!per_utt || (stats_ws && !out_bf16), "avexk_fbank_forward: per_utt needs stats_ws and fp32 output");
  const int chunks = ceil_div(Fout, FPC);
  AVEXK_CHECK_ARG((long long)B * chunks < (1LL << 31), "avexk_fbank_forward: too many frames (B=%d, frames=%d)", B, Fout);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* stats = nullptr;
  float nm = norm_mean, ns = norm_scale;
  if (per_utt) {
    stats = stats_ws;
    AVEXK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, st));
    nm = 0.f;
    ns = 1.f;
  }
  prof_begin(st, KID_FBANK, (double)B * (4.0 * T + (out_bf16 ? 2.0 : 4.0) * Fout * NMEL));
  if (out_bf16)
    launch<OUT_BF16>(h, wav, B, T, wav_stride, F, Fout, chunks, prescale, nm, ns, out, 0, stats, st);
  else
    launch<OUT_F32>(h, wav, B, T, wav_stride, F, Fout, chunks, prescale, nm, ns, out, 0, stats, st);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  if (per_utt) {
    long long per_clip = (long long)Fout * NMEL;
    AVEXK_CHECK_ARG(B <= 65535, "avexk_fbank_forward: per_utt with B=%d exceeds grid.y", B);
    dim3 g2(ceil_div(per_clip, 256 * 8), B);
    per_utt_normalise_kernel<<<g2, 256, 0, st>>>(reinterpret_cast<float*>(out), stats, per_clip);
    AVEXK_LAUNCH_CHECK();
  }
  return AVEXK_OK;
}

extern "C" int avexk_fbank_patch_operand(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride, float prescale,
                                         float norm_mean, float norm_scale, void* out, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && wav && out, "avexk_fbank_patch_operand: null argument");
  AVEXK_CHECK_ARG(B >= 0 && T >= 0 && wav_stride >= T, "avexk_fbank_patch_operand: bad shape B=%d T=%d stride=%lld", B, T, wav_stride);
  const int F = avexk_fbank_num_frames(T);
  const int chunks = F / FPC;  // complete rows of 16x16 patches only (beats.py:349: Conv2d stride 16 drops the ragged tail)
  if (B == 0 || chunks == 0) return AVEXK_OK;
  AVEXK_CHECK_ARG((long long)B * chunks < (1LL << 31), "avexk_fbank_patch_operand: too many frames (B=%d, frames=%d)", B, F);
  AVEXK_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 3) == 0, "avexk_fbank_patch_operand: out must be 4-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // algorithmic bytes as for the fp32 fbank so that the roofline figure stays comparable; this mode writes 6 B per bin
  prof_begin(st, KID_FBANK, (double)B * (4.0 * T + 4.0 * F * NMEL));
  launch<OUT_PATCH3>(h, wav, B, T, wav_stride, F, chunks * FPC, chunks, prescale, norm_mean, norm_scale, out, (long long)chunks * 8, nullptr, st);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  return AVEXK_OK;
}
