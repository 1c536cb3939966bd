// fbank.cu -- fused Kaldi log-mel filterbank for sm_100a.
//
// Replaces avex/models/beats/beats.py:120-163 (`_BatchedFbank.forward`: unfold, DC removal, pre-emphasis,
// Povey window, zero-pad, rfft, |.|^2, mel matmul, clamp+log) and the affine of beats.py:323, in ONE kernel:
// the [B,F,512] frame tensor and the [B,F,257] spectrum the reference materialises never exist.
//
// Layout / mapping
//   grid  = (ceil(out_frames / 16), B); CTA = 256 threads = 16 frames of one clip.
//   load  : the 16 frames' 2800-sample span is read ONCE from HBM (float4, coalesced) into shared memory
//           (frames overlap 2.5x; adjacent CTAs re-read 240 samples, served by L2).
//   phase A (16 lanes per frame): 512-pt real FFT as a 256-pt complex FFT, 16 x 16 Cooley-Tukey with the two
//           radix-16 stages held in registers and one transposition through padded shared memory; real-FFT
//           split + power spectrum on conjugate pairs -> P[0..256] in shared memory.
//   phase B (one thread per mel bin, 8 frames each): sparse triangular mel projection (504 non-zeros instead
//           of the reference's dense 257x128 SGEMM), log, affine, coalesced 512 B row stores.
// Roofline: HBM; algorithmic bytes per clip = 4*T + 4*F*128.
#include <float.h>
#include <math.h>

#include <vector>

#include "common.cuh"

namespace avexk {
namespace {

constexpr int WIN = 400, HOP = 160, NMEL = 128, NBIN = 257;
constexpr int FPC = 16;                       // frames per CTA
constexpr int SEG = (FPC - 1) * HOP + WIN;    // 2800 samples staged per CTA
constexpr int XROW = 272;                     // float2 per frame in the exchange buffer (16 rows x 17)
constexpr int PROW = 272;                     // floats per frame in the power buffer (257 + ELL overrun pad)
constexpr int MAXNZ = 10;                     // max non-zeros of one mel filter at n_fft 512 / 128 bins
constexpr int SMEM_BYTES = SEG * 4 + FPC * XROW * 8 + FPC * PROW * 4 + WIN * 4 + 256 * 8 + 136 * 8;

struct Tables {
  const float* window;   // [400], pre-multiplied by 0.5 (folds the 1/2 of the real-FFT split; exact)
  const float2* tw;      // [16][16]: tw[q*16+t] = exp(-2 pi i t q / 256)
  const float2* tw2;     // [136]: exp(-2 pi i k / 512), k = 0..128
  const float* melw;     // [MAXNZ][128] ELL weights, zero padded
  const int* melstart;   // [128] first FFT bin of each filter
  int warp_maxlen[4];    // longest filter among each group of 32 bins
};

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {  // a * w
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

// forward radix-4 DFT, in place: (a0,a1,a2,a3) -> (y0,y1,y2,y3), w4 = -i
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
  float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
  float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
  float2 t3 = make_float2(a1.x - a3.x, a1.y - a3.y);
  a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
  a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i t3
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i t3
}

// forward 16-point DFT in registers.  On return V[q] (natural order) lives in v[IDX16(q)].
#define IDX16(q) (4 * ((q) & 3) + ((q) >> 2))
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f, R = 0.70710678118654752440f;
#pragma unroll
  for (int j0 = 0; j0 < 4; ++j0) fft4(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
  // v[j0 + 4 q0] *= w16^(j0 q0),  w16^m = (cos(pi m / 8), -sin(pi m / 8))
  v[1 + 4] = cmul(v[1 + 4], make_float2(C1, -S1));    // m = 1
  v[1 + 8] = cmul(v[1 + 8], make_float2(R, -R));      // m = 2
  v[1 + 12] = cmul(v[1 + 12], make_float2(S1, -C1));  // m = 3
  v[2 + 4] = cmul(v[2 + 4], make_float2(R, -R));      // m = 2
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);    // m = 4: * (-i)
  v[2 + 12] = cmul(v[2 + 12], make_float2(-R, -R));   // m = 6
  v[3 + 4] = cmul(v[3 + 4], make_float2(S1, -C1));    // m = 3
  v[3 + 8] = cmul(v[3 + 8], make_float2(-R, -R));     // m = 6
  v[3 + 12] = cmul(v[3 + 12], make_float2(-C1, S1));  // m = 9
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) fft4(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3]);
}

template <bool BF16>
__global__ void __launch_bounds__(256, 3)
fbank_kernel(const float* __restrict__ wav, long long stride, int T, int F, int Fout, float prescale, float nmean,
             float nscale, Tables tb, void* __restrict__ out, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* sWav = reinterpret_cast<float*>(smem);
  float2* sX = reinterpret_cast<float2*>(smem + SEG * 4);
  float* sP = reinterpret_cast<float*>(smem + SEG * 4 + FPC * XROW * 8);
  float* sWin = sP + FPC * PROW;
  float2* sTw = reinterpret_cast<float2*>(sWin + WIN);
  float2* sTw2 = sTw + 256;

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FPC;
  const long long s0 = (long long)f0 * HOP;
  const float* src = wav + (long long)b * stride;

  // ---- stage the waveform span (each sample read once), scaled by `prescale` ------------------------------
  const bool vec_ok = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(wav) & 15) == 0);
  if (vec_ok) {
    for (int i = tid; i < SEG / 4; i += 256) {
      long long s = s0 + 4 * i;
      float4 v;
      if (s + 3 < T) {
        v = __ldg(reinterpret_cast<const float4*>(src + s));
      } else {
        v.x = s + 0 < T ? src[s + 0] : 0.f;
        v.y = s + 1 < T ? src[s + 1] : 0.f;
        v.z = s + 2 < T ? src[s + 2] : 0.f;
        v.w = 0.f;
      }
      v.x *= prescale; v.y *= prescale; v.z *= prescale; v.w *= prescale;
      reinterpret_cast<float4*>(sWav)[i] = v;
    }
  } else {
    for (int i = tid; i < SEG; i += 256) {
      long long s = s0 + i;
      sWav[i] = s < T ? src[s] * prescale : 0.f;
    }
  }
  for (int i = tid; i < WIN; i += 256) sWin[i] = tb.window[i];
  sTw[tid] = tb.tw[tid];
  if (tid < 136) sTw2[tid] = tb.tw2[tid];
  __syncthreads();

  // ---- phase A: one frame per 16 lanes ---------------------------------------------------------------------
  {
    const int fl = tid >> 4, t = tid & 15;
    const float* x = sWav + fl * HOP;
    float2* sXf = sX + fl * XROW;
    float* sPf = sP + fl * PROW;
    float2 v[16];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = t + 16 * j;
      if (j < 12 || (j == 12 && t < 8)) {  // n < 200: the 400 real samples as 200 complex points
        v[j] = *reinterpret_cast<const float2*>(x + 2 * n);
        s += v[j].x + v[j].y;
      } else {
        v[j] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
    const float mu = s / 400.0f;  // beats.py:140
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = t + 16 * j;
      if (j < 12 || (j == 12 && t < 8)) {
        const float a0 = v[j].x - mu, a1 = v[j].y - mu;
        const float prev = (n == 0) ? a0 : (x[2 * n - 1] - mu);  // replicate pad, beats.py:143
        const float2 w = *reinterpret_cast<const float2*>(sWin + 2 * n);
        v[j].x = (a0 - 0.97f * prev) * w.x;  // beats.py:144,147
        v[j].y = (a1 - 0.97f * a0) * w.y;
      }
    }
    // stage 1: V_t[q] = sum_j z[t+16j] w16^(jq); then twiddle w256^(tq)
    fft16(v);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[IDX16(q)] = cmul(v[IDX16(q)], sTw[q * 16 + t]);
#pragma unroll
    for (int q = 0; q < 16; ++q) sXf[q * 17 + t] = v[IDX16(q)];
    __syncwarp();
    // stage 2 (this lane now owns q = t): Z[t + 16 p] = sum_tt u[tt] w16^(tt p)
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) v[tt] = sXf[t * 17 + tt];
    __syncwarp();
    fft16(v);
#pragma unroll
    for (int p = 0; p < 16; ++p) sXf[t + 16 * p] = v[IDX16(p)];
    __syncwarp();
    // real-FFT split on conjugate pairs (k, 256-k); Z is already halved through the window table.
    //   E = Z[k] + conj(Z[256-k]),  O = -i (Z[k] - conj(Z[256-k])),  X[k] = E + W^k O,  X[256-k] = conj(E - W^k O)
#pragma unroll
    for (int m = 0; m < 9; ++m) {
      const int k = t + 16 * m;
      if (m < 8 || t == 0) {
        const float2 A = sXf[k], Bc = sXf[(256 - k) & 255], W = sTw2[k];
        const float ex = A.x + Bc.x, ey = A.y - Bc.y;
        const float ox = A.y + Bc.y, oy = Bc.x - A.x;
        const float tx = fmaf(W.x, ox, -W.y * oy), ty = fmaf(W.x, oy, W.y * ox);
        const float px = ex + tx, py = ey + ty, qx = ex - tx, qy = ey - ty;
        sPf[k] = fmaf(px, px, py * py);        // |X[k]|^2      beats.py:155
        sPf[256 - k] = fmaf(qx, qx, qy * qy);  // |X[256-k]|^2
      }
    }
    if (t < 15) sPf[NBIN + t] = 0.f;
  }
  __syncthreads();

  // ---- phase B: one thread per mel bin, 8 frames each ------------------------------------------------------
  {
    const int j = tid & (NMEL - 1), half = tid >> 7;
    const int start = __ldg(tb.melstart + j);
    const int wq = (tid >> 5) & 3;
    const int maxlen = wq == 0 ? tb.warp_maxlen[0] : wq == 1 ? tb.warp_maxlen[1] : wq == 2 ? tb.warp_maxlen[2] : tb.warp_maxlen[3];
    float w[MAXNZ];
#pragma unroll
    for (int i = 0; i < MAXNZ; ++i) w[i] = __ldg(tb.melw + i * NMEL + j);
    float ssum = 0.f, ssq = 0.f;
#pragma unroll 1
    for (int ff = 0; ff < 8; ++ff) {
      const int fl = half * 8 + ff, f = f0 + fl;
      if (f >= Fout) break;
      float val = 0.f;  // zero padding in the log-mel domain (eat/audio_processor.py:121-124)
      if (f < F) {
        const float* p = sP + fl * PROW + start;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < MAXNZ; ++i)
          if (i < maxlen) acc = fmaf(w[i], p[i], acc);  // beats.py:159
        val = logf(fmaxf(acc, FLT_EPSILON));              // beats.py:163
      }
      ssum += val;
      ssq = fmaf(val, val, ssq);
      const float o = (val - nmean) * nscale;  // beats.py:323
      const size_t idx = ((size_t)b * Fout + f) * NMEL + j;
      if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(o);
      else reinterpret_cast<float*>(out)[idx] = o;
    }
    if (stats != nullptr) {  // per-utterance statistics: warp shuffles, then one atomic pair per warp
      ssum = warp_sum(ssum);
      ssq = warp_sum(ssq);
      if ((tid & 31) == 0) {
        atomicAdd(stats + 2 * b, (double)ssum);
        atomicAdd(stats + 2 * b + 1, (double)ssq);
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
  std::vector<float> win(WIN), melw(MAXNZ * NMEL, 0.f);
  std::vector<float2> tw(256), tw2(136);
  std::vector<int> start(NMEL, 0);
  int wmax[4] = {1, 1, 1, 1};
  for (int i = 0; i < WIN; ++i) win[i] = 0.5f * window_host[i];
  for (int q = 0; q < 16; ++q)
    for (int t = 0; t < 16; ++t) {
      double a = -2.0 * M_PI * (double)(t * q) / 256.0;
      tw[q * 16 + t] = make_float2((float)cos(a), (float)sin(a));
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
    if (lo < 0) { lo = 0; hi = 0; }
    int len = hi - lo + 1;
    AVEXK_CHECK_ARG(len <= MAXNZ, "avexk_fbank_create: mel filter %d spans %d bins (> %d)", j, len, MAXNZ);
    start[j] = lo;
    for (int i = 0; i < len; ++i) melw[i * NMEL + j] = mel_fb_host[(lo + i) * NMEL + j];
    if (len > wmax[j / 32]) wmax[j / 32] = len;
  }
  auto* h = new avexk_fbank();
  size_t off_win = 0, off_tw = off_win + WIN * 4, off_tw2 = off_tw + 256 * 8, off_mw = off_tw2 + 136 * 8,
         off_ms = off_mw + MAXNZ * NMEL * 4, total = off_ms + NMEL * 4;
  cudaError_t e = cudaMalloc(&h->dev, total);
  if (e != cudaSuccess) {
    delete h;
    set_error("avexk_fbank_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  char* d = reinterpret_cast<char*>(h->dev);
  cudaMemcpy(d + off_win, win.data(), WIN * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_tw, tw.data(), 256 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_tw2, tw2.data(), 136 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + off_mw, melw.data(), MAXNZ * NMEL * 4, cudaMemcpyHostToDevice);
  e = cudaMemcpy(d + off_ms, start.data(), NMEL * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(h->dev);
    delete h;
    set_error("avexk_fbank_create: upload failed: %s", cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  h->tb.window = reinterpret_cast<const float*>(d + off_win);
  h->tb.tw = reinterpret_cast<const float2*>(d + off_tw);
  h->tb.tw2 = reinterpret_cast<const float2*>(d + off_tw2);
  h->tb.melw = reinterpret_cast<const float*>(d + off_mw);
  h->tb.melstart = reinterpret_cast<const int*>(d + off_ms);
  for (int i = 0; i < 4; ++i) h->tb.warp_maxlen[i] = wmax[i];
  cudaFuncSetAttribute(fbank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  *out = h;
  return AVEXK_OK;
}

extern "C" void avexk_fbank_destroy(avexk_fbank_t* h) {
  if (!h) return;
  if (h->dev) cudaFree(h->dev);
  delete h;
}

extern "C" int avexk_fbank_forward(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride,
                                   float prescale, float norm_mean, float norm_scale, int out_frames, int per_utt,
                                   double* stats_ws, void* out, int out_bf16, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && wav && out, "avexk_fbank_forward: null argument");
  AVEXK_CHECK_ARG(B >= 0 && T >= 0 && wav_stride >= T, "avexk_fbank_forward: bad shape B=%d T=%d stride=%lld", B, T, wav_stride);
  const int F = avexk_fbank_num_frames(T);
  const int Fout = out_frames > 0 ? out_frames : F;
  if (B == 0 || Fout == 0) return AVEXK_OK;
  AVEXK_CHECK_ARG(B <= 65535, "avexk_fbank_forward: B=%d exceeds grid.y", B);
  AVEXK_CHECK_ARG(!per_utt || (stats_ws && !out_bf16), "avexk_fbank_forward: per_utt needs stats_ws and fp32 output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* stats = nullptr;
  float nm = norm_mean, ns = norm_scale;
  if (per_utt) {
    stats = stats_ws;
    AVEXK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, st));
    nm = 0.f;
    ns = 1.f;
  }
  dim3 grid(ceil_div(Fout, FPC), B);
  prof_begin(st, KID_FBANK, (double)B * (4.0 * T + (out_bf16 ? 2.0 : 4.0) * Fout * NMEL));
  if (out_bf16)
    fbank_kernel<true><<<grid, 256, SMEM_BYTES, st>>>(wav, wav_stride, T, F, Fout, prescale, nm, ns, h->tb, out, stats);
  else
    fbank_kernel<false><<<grid, 256, SMEM_BYTES, st>>>(wav, wav_stride, T, F, Fout, prescale, nm, ns, h->tb, out, stats);
  prof_end(st);
  AVEXK_LAUNCH_CHECK();
  if (per_utt) {
    long long per_clip = (long long)Fout * NMEL;
    dim3 g2(ceil_div(per_clip, 256 * 8), B);
    per_utt_normalise_kernel<<<g2, 256, 0, st>>>(reinterpret_cast<float*>(out), stats, per_clip);
    AVEXK_LAUNCH_CHECK();
  }
  return AVEXK_OK;
}
