// melspec.cu -- STFT mel spectrogram of the EfficientNet path, one fused kernel + per-clip min/max.
//
// Replaces `AudioProcessor.__call__` + `_normalize` (avex/data/audio_utils.py:106-172) for the mel_spectrogram
// representation of esp_aves2_effnetb0_all.yml (n_fft = win = 800, hop 160, hann periodic, center=True with reflect
// padding, 128 HTK mel bins over 0..8 kHz, no area norm):
//   X = stft(wav) ; P = |X|^2 [401, frames] ; mel = fb^T P ; y = ln(mel + 1e-6) ; out = (y - min y) / (max y - min y + 1e-8)
// The reference runs cuFFT on a non-power-of-two size, a dense 401x128 product on a <2 %-dense matrix, and four
// full-tensor reductions.  Here: 800-point real DFT as a two-stage Cooley-Tukey 25 x 32 per frame
//   stage 1: thread (frame, n2) -- 25-point DFT over n1 of the windowed samples x[32 n1 + n2] (real input), times W800^(n2 k1)
//   stage 2: thread (frame, k1) -- 32-point radix-2 FFT over n2 in registers -> bins k = k1 + 25 k2 (k <= 400 kept)
// then power, sparse mel (contiguous bin range per filter), log, fp32 store time-last [B, 128, frames], and the per-clip
// min/max through warp shuffles + one atomic pair per CTA.  Every sample is read once per 8-frame chunk (staged in
// shared memory with the reflect padding resolved on load).
// Roofline: FP32 issue (about 2.3 kFMA per frame in stage 1 + 0.5 k in stage 2), not HBM: 4 T + 4*128*frames bytes/clip.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace avexk {
namespace {

constexpr int NFFT = 800, HOP = 160, NBIN = 401, NMEL = 128, N1 = 25, N2 = 32;
constexpr int FPC = 8;  // frames per CTA
constexpr int SPAN = (FPC - 1) * HOP + NFFT;  // 1920 samples
constexpr int NTHREADS = 256;

struct MelTables {
  const float* window;    // [800]
  const float2* w25;      // [25] W25^j
  const float2* w800;     // [32*25] W800^(n2*k1), index n2*25 + k1
  const int* mel_start;   // [128] first FFT bin of each filter
  const int* mel_off;     // [129] offsets into mel_w
  const float* mel_w;     // packed non-zero weights
  int mel_nnz;
};

// A CTA works through CPC consecutive 8-frame chunks of one clip: the samples of chunk c+1 are fetched with cp.async into the
// other half of a double buffer while chunk c is transformed, and the mel filter tables are read from global memory once.
constexpr int CPC = 8;            // chunks per CTA
constexpr int MEL_NNZ_MAX = 1280; // packed non-zero mel weights kept in shared memory (the HTK 128 x 401 bank has ~930)

// smem layout (floats)
constexpr int SM_X = 0;                           // [2][SPAN]
constexpr int SM_Y = SM_X + 2 * SPAN;             // float2 [FPC][N2][N1]  (k1 fastest)
constexpr int SM_P = SM_Y + 2 * FPC * N2 * N1;    // [FPC][NBIN + 3]
constexpr int PSTR = NBIN + 3;
constexpr int SM_MW = SM_P + FPC * PSTR;          // [MEL_NNZ_MAX] mel weights
constexpr int SM_MS = SM_MW + MEL_NNZ_MAX;        // int [128] first bin, int [129] offsets
constexpr int SM_RED = SM_MS + NMEL + NMEL + 1;   // [16]
constexpr int SMEM_FLOATS = SM_RED + 16;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;

__device__ __forceinline__ unsigned enc_ordered(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-register 32-point DIF FFT on packed complex values (x = re, y = im): a butterfly is two packed adds and, where the twiddle is
// not trivial, one packed multiply + one packed FMA (d * w = d * c + (-d.y, d.x) * s).  Output index k2 ends up bit-reversed:
// z[bitrev5(k2)] = X[k2].
__device__ __forceinline__ void fft32_dif(float2 (&z)[32]) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & half) == 0) {
        const int j = i | half;
        const int tw = (i & (half - 1)) << s;  // exponent of W32
        const float2 a = z[i], b = z[j];
        z[i] = __fadd2_rn(a, b);
        const float2 d = __fadd2_rn(a, make_float2(-b.x, -b.y));
        // W32^tw = exp(-2 pi i tw / 32): immediates after unrolling
        constexpr float kCos[16] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
        constexpr float kSin[16] = {-0.000000000e+00f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f, -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f};
        if (tw == 0) z[j] = d;
        else if (tw == 8) z[j] = make_float2(d.y, -d.x);  // * (-i)
        else {
          const float c = kCos[tw], sn = kSin[tw];
          z[j] = __ffma2_rn(d, make_float2(c, c), __fmul2_rn(make_float2(-d.y, d.x), make_float2(sn, sn)));
        }
      }
    }
  }
}
__host__ __device__ constexpr int bitrev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

__global__ void __launch_bounds__(NTHREADS)
melspec_kernel(const float* __restrict__ wav, long long wav_stride, int T, int frames, const MelTables tb,
               float* __restrict__ out, unsigned* __restrict__ minmax) {
  extern __shared__ float sm[];
  float2* sy = reinterpret_cast<float2*>(sm + SM_Y);
  float* sp = sm + SM_P;
  float* smw = sm + SM_MW;
  int* sms = reinterpret_cast<int*>(sm + SM_MS);
  int* smo = sms + NMEL;
  float* sred = sm + SM_RED;
  const int tid = threadIdx.x, b = blockIdx.y;
  const float* x = wav + (size_t)b * wav_stride;
  const int chunk0 = blockIdx.x * CPC, n_chunks = (frames + FPC - 1) / FPC;
  const int chunk_end = min(chunk0 + CPC, n_chunks);

  // samples of one chunk -> shared memory, asynchronously; reflect padding (center=True) resolved in the source index
  auto fetch = [&](int chunk, float* dst) {
    const int base = chunk * FPC * HOP - NFFT / 2;
    for (int i = tid; i < SPAN; i += NTHREADS) {
      int j = base + i;  // index into the unpadded clip
      if (j < 0) j = -j;
      if (j >= T) j = 2 * (T - 1) - j;
      const bool ok = j >= 0 && j < T;
      const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst + i));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(x + (ok ? j : 0)), "r"(ok ? 4 : 0) : "memory");  // (0 bytes: zero fill)
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(chunk0, sm + SM_X);
  for (int i = tid; i < tb.mel_nnz; i += NTHREADS) smw[i] = tb.mel_w[i];
  for (int i = tid; i < NMEL; i += NTHREADS) sms[i] = tb.mel_start[i];
  for (int i = tid; i <= NMEL; i += NTHREADS) smo[i] = tb.mel_off[i];
  float vmin = INFINITY, vmax = -INFINITY;

  for (int chunk = chunk0; chunk < chunk_end; ++chunk) {
  const int f0 = chunk * FPC;
  const float* sx = sm + SM_X + ((chunk - chunk0) & 1) * SPAN;
  if (chunk + 1 < chunk_end) {
    fetch(chunk + 1, sm + SM_X + ((chunk + 1 - chunk0) & 1) * SPAN);  // lands while this chunk is transformed
    asm volatile("cp.async.wait_group 1;" ::: "memory");
  } else {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();  // this chunk's samples (and, the first time, the tables) are visible to every thread

  // ---- stage 1: thread (k1 half, frame pair, n2): Y[k1] = W800^(n2 k1) * sum_n1 w[32 n1 + n2] x[f*160 + 32 n1 + n2] W25^(n1 k1) ----
  // The 25-point DFT of a REAL sequence, brute force with every W25 power an immediate operand, for TWO frames at once: the
  // two frames' samples ride in one packed fp32x2 register pair and share the immediate.  Real input: only k1 = 0..12 are
  // computed (Y[25 - k1] = conj Y[k1]), and the sums run over a[n] = x[n] + x[25-n] (cosine part) and b[n] = x[n] - x[25-n]
  // (sine part), n = 1..12: 24 packed FMAs per k1 and frame pair instead of the 100 scalar FMAs of the plain double loop.
  // The two halves of the CTA split the k1 range (0..6 | 7..12).
  {
    constexpr float kC25[25] = {1.000000000e+00f, 9.685831611e-01f, 8.763066800e-01f, 7.289686274e-01f, 5.358267950e-01f, 3.090169944e-01f, 6.279051953e-02f, -1.873813146e-01f, -4.257792916e-01f, -6.374239897e-01f, -8.090169944e-01f, -9.297764859e-01f, -9.921147013e-01f, -9.921147013e-01f, -9.297764859e-01f, -8.090169944e-01f, -6.374239897e-01f, -4.257792916e-01f, -1.873813146e-01f, 6.279051953e-02f, 3.090169944e-01f, 5.358267950e-01f, 7.289686274e-01f, 8.763066800e-01f, 9.685831611e-01f};
    constexpr float kS25[25] = {-0.000000000e+00f, -2.486898872e-01f, -4.817536741e-01f, -6.845471059e-01f, -8.443279255e-01f, -9.510565163e-01f, -9.980267284e-01f, -9.822872507e-01f, -9.048270525e-01f, -7.705132428e-01f, -5.877852523e-01f, -3.681245527e-01f, -1.253332336e-01f, 1.253332336e-01f, 3.681245527e-01f, 5.877852523e-01f, 7.705132428e-01f, 9.048270525e-01f, 9.822872507e-01f, 9.980267284e-01f, 9.510565163e-01f, 8.443279255e-01f, 6.845471059e-01f, 4.817536741e-01f, 2.486898872e-01f};
    const int hh = tid >> 7, pr = (tid >> 5) & 3, n2 = tid & 31;
    const float* xa = sx + (2 * pr) * HOP + n2;
    float2 x0, av[12], bv[12];  // (frame 2 pr, frame 2 pr + 1)
    {
      const float w0 = __ldg(tb.window + n2);
      x0 = make_float2(xa[0] * w0, xa[HOP] * w0);
    }
#pragma unroll
    for (int n = 1; n <= 12; ++n) {
      const float wl = __ldg(tb.window + 32 * n + n2), wh = __ldg(tb.window + 32 * (N1 - n) + n2);
      const float2 lo = make_float2(xa[32 * n] * wl, xa[HOP + 32 * n] * wl);
      const float2 hi = make_float2(xa[32 * (N1 - n)] * wh, xa[HOP + 32 * (N1 - n)] * wh);
      av[n - 1] = __fadd2_rn(lo, hi);
      bv[n - 1] = __fadd2_rn(lo, make_float2(-hi.x, -hi.y));
    }
    float2* dst0 = sy + ((2 * pr) * N2 + n2) * N1;  // frame 2 pr; frame 2 pr + 1 is N2 * N1 further
    const float2* tw = tb.w800 + n2 * N1;
    auto emit = [&](int k1, float2 re, float2 im) {  // Y[k1] = re + i im for both frames -> twiddle -> shared memory
      const float2 t = __ldg(tw + k1);
      const float2 yr = __ffma2_rn(re, make_float2(t.x, t.x), __fmul2_rn(im, make_float2(-t.y, -t.y)));
      const float2 yi = __ffma2_rn(re, make_float2(t.y, t.y), __fmul2_rn(im, make_float2(t.x, t.x)));
      dst0[k1] = make_float2(yr.x, yi.x);
      dst0[N2 * N1 + k1] = make_float2(yr.y, yi.y);
    };
#pragma unroll
    for (int k1 = 0; k1 <= N1 / 2; ++k1) {
      if ((k1 <= 6) != (hh == 0)) continue;  // (CTA-half uniform)
      float2 re = x0, im = make_float2(0.f, 0.f);
#pragma unroll
      for (int n = 1; n <= 12; ++n) {
        re = __ffma2_rn(av[n - 1], make_float2(kC25[(n * k1) % N1], kC25[(n * k1) % N1]), re);
        if (k1 > 0) im = __ffma2_rn(bv[n - 1], make_float2(kS25[(n * k1) % N1], kS25[(n * k1) % N1]), im);
      }
      emit(k1, re, im);
      if (k1 > 0) emit(N1 - k1, re, make_float2(-im.x, -im.y));
    }
  }
  __syncthreads();

  // ---- stage 2: thread (f, k1): 32-point FFT over n2 -> power of bins k1 + 25 k2 <= 400 --------------------------
  if (tid < FPC * N1) {
    const int f = tid / N1, k1 = tid - f * N1;
    float2 z[32];
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) z[n2] = sy[(f * N2 + n2) * N1 + k1];
    fft32_dif(z);
    float* pr = sp + f * PSTR;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float2 v = z[bitrev5(k2)];
      pr[k1 + N1 * k2] = v.x * v.x + v.y * v.y;
    }
    if (k1 == 0) {
      const float2 v = z[bitrev5(16)];
      pr[400] = v.x * v.x + v.y * v.y;
    }
  }
  __syncthreads();

  // ---- mel + log + store (time-last) + per-clip min / max ---------------------------------------------------------
  for (int i = tid; i < FPC * NMEL; i += NTHREADS) {
    const int f = i & (FPC - 1), j = i >> 3;  // consecutive threads -> consecutive frames of one mel bin
    if (f0 + f < frames) {
      const int ks = sms[j], o0 = smo[j], o1 = smo[j + 1];
      const float* pr = sp + f * PSTR + ks;
      float acc = 0.f;
      for (int q = 0; q < o1 - o0; ++q) acc = fmaf(pr[q], smw[o0 + q], acc);
      const float y = logf(acc + 1e-6f);
      out[((size_t)b * NMEL + j) * frames + f0 + f] = y;
      vmin = fminf(vmin, y);
      vmax = fmaxf(vmax, y);
    }
  }
  // (the next chunk's stage 1 writes sy / sp only behind its own barriers; the sample buffer it refills was last read in this
  // chunk's stage 1, two barriers ago)
  }  // chunk loop

  if (minmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((tid & 31) == 0) {
      sred[tid >> 5] = vmin;
      sred[8 + (tid >> 5)] = vmax;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NTHREADS / 32; ++w) {
        vmin = fminf(vmin, sred[w]);
        vmax = fmaxf(vmax, sred[8 + w]);
      }
      if (vmin <= vmax) {
        atomicMin(minmax + 2 * b, enc_ordered(vmin));
        atomicMax(minmax + 2 * b + 1, enc_ordered(vmax));
      }
    }
  }
}

__global__ void melspec_init_minmax(unsigned* minmax, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    minmax[2 * i] = 0xffffffffu;
    minmax[2 * i + 1] = 0u;
  }
}

// (y - min) / (max - min + 1e-8), audio_utils.py:167-172; in place or into `out`
__global__ void melspec_normalise_kernel(const float* __restrict__ y, const unsigned* __restrict__ minmax, long long per_clip,
                                         float* __restrict__ out) {
  const int b = blockIdx.y;
  const float mn = dec_ordered(minmax[2 * b]), mx = dec_ordered(minmax[2 * b + 1]);
  const float den = mx - mn + 1e-8f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip; i += (long long)gridDim.x * blockDim.x)
    out[(size_t)b * per_clip + i] = (y[(size_t)b * per_clip + i] - mn) / den;
}

}  // namespace

float melspec_decode_ordered_host(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &v, 4);
  return f;
}

}  // namespace avexk

struct avexk_melspec {
  void* dev = nullptr;
  avexk::MelTables tb;
};

extern "C" int avexk_melspec_num_frames(int T) { return T <= avexk::NFFT / 2 ? 0 : 1 + T / avexk::HOP; }

extern "C" int avexk_melspec_create(const float* window_host, const float* mel_fb_host, avexk_melspec_t** out) {
  using namespace avexk;
  AVEXK_CHECK_ARG(window_host && mel_fb_host && out, "avexk_melspec_create: null argument");
  std::vector<float2> w25(N1), w800(N2 * N1);
  for (int j = 0; j < N1; ++j) {
    const double a = -2.0 * M_PI * (double)j / 25.0;
    w25[j] = make_float2((float)cos(a), (float)sin(a));
  }
  for (int n2 = 0; n2 < N2; ++n2)
    for (int k1 = 0; k1 < N1; ++k1) {
      const double a = -2.0 * M_PI * (double)(n2 * k1) / 800.0;
      w800[n2 * N1 + k1] = make_float2((float)cos(a), (float)sin(a));
    }
  std::vector<int> start(NMEL, 0), off(NMEL + 1, 0);
  std::vector<float> mw;
  for (int j = 0; j < NMEL; ++j) {
    int lo = -1, hi = -1;
    for (int k = 0; k < NBIN; ++k)
      if (mel_fb_host[k * NMEL + j] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    off[j] = (int)mw.size();
    if (lo >= 0) {
      start[j] = lo;
      for (int k = lo; k <= hi; ++k) mw.push_back(mel_fb_host[k * NMEL + j]);
    }
  }
  off[NMEL] = (int)mw.size();
  if (mw.empty()) mw.push_back(0.f);
  AVEXK_CHECK_ARG((int)mw.size() <= MEL_NNZ_MAX, "avexk_melspec_create: %d non-zero mel weights exceed the shared-memory table (%d)", (int)mw.size(), MEL_NNZ_MAX);
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t o_win = 0, o_w25 = al(o_win + NFFT * 4), o_w800 = al(o_w25 + N1 * 8), o_ms = al(o_w800 + N2 * N1 * 8),
               o_mo = al(o_ms + NMEL * 4), o_mw = al(o_mo + (NMEL + 1) * 4), total = al(o_mw + mw.size() * 4);
  auto* h = new avexk_melspec();
  cudaError_t e = cudaMalloc(&h->dev, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("avexk_melspec_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  char* d = reinterpret_cast<char*>(h->dev);
  cudaMemcpy(d + o_win, window_host, NFFT * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_w25, w25.data(), N1 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_w800, w800.data(), N2 * N1 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_ms, start.data(), NMEL * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_mo, off.data(), (NMEL + 1) * 4, cudaMemcpyHostToDevice);
  e = cudaMemcpy(d + o_mw, mw.data(), mw.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(h->dev);
    delete h;
    set_error("avexk_melspec_create: upload failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  h->tb.window = reinterpret_cast<const float*>(d + o_win);
  h->tb.w25 = reinterpret_cast<const float2*>(d + o_w25);
  h->tb.w800 = reinterpret_cast<const float2*>(d + o_w800);
  h->tb.mel_start = reinterpret_cast<const int*>(d + o_ms);
  h->tb.mel_off = reinterpret_cast<const int*>(d + o_mo);
  h->tb.mel_w = reinterpret_cast<const float*>(d + o_mw);
  h->tb.mel_nnz = (int)mw.size();
  cudaFuncSetAttribute(melspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  *out = h;
  return AVEXK_OK;
}

extern "C" void avexk_melspec_destroy(avexk_melspec_t* h) {
  if (!h) return;
  if (h->dev) cudaFree(h->dev);
  delete h;
}

extern "C" int avexk_melspec_forward(const avexk_melspec_t* h, const float* wav, int B, int T, long long wav_stride,
                                     int normalize, float* out, void* minmax_ws, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && wav && out, "avexk_melspec_forward: null argument");
  AVEXK_CHECK_ARG(B >= 0 && wav_stride >= T, "avexk_melspec_forward: bad shape B=%d T=%d stride=%lld", B, T, wav_stride);
  AVEXK_CHECK_ARG(T > NFFT / 2, "avexk_melspec_forward: reflect padding needs more than %d samples (T=%d)", NFFT / 2, T);
  AVEXK_CHECK_ARG(B <= 65535, "avexk_melspec_forward: B=%d exceeds grid.y", B);
  AVEXK_CHECK_ARG(!normalize || minmax_ws, "avexk_melspec_forward: normalize needs minmax_ws (B * 2 uint32)");
  if (B == 0) return AVEXK_OK;
  const int frames = avexk_melspec_num_frames(T);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned* mm = reinterpret_cast<unsigned*>(minmax_ws);
  if (mm) {
    melspec_init_minmax<<<ceil_div(B, 256), 256, 0, st>>>(mm, B);
    AVEXK_LAUNCH_CHECK();
  }
  dim3 grid(ceil_div(ceil_div(frames, FPC), CPC), B);
  prof_begin(st, KID_FBANK, (double)B * (4.0 * T + 4.0 * frames * NMEL));
  melspec_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(wav, wav_stride, T, frames, h->tb, out, mm);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  if (normalize) {
    const long long per_clip = (long long)NMEL * frames;
    dim3 g2(ceil_div(per_clip, 256 * 4), B);
    melspec_normalise_kernel<<<g2, 256, 0, st>>>(out, mm, per_clip, out);
    AVEXK_LAUNCH_CHECK();
  }
  return AVEXK_OK;
}
