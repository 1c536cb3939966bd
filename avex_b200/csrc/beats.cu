// beats.cu -- BEATs encoder forward as one C-ABI call: enqueues every kernel of the path on the caller's stream.
//
// Mirrors avex/models/beats/beats.py:325-382 (front end) and backbone.py:151-221, :350-373 (12 post-LN DeepNorm
// blocks).  Data layout in HBM (M = B*N token rows, batch-major; C = 768):
//   x    fp32 [M,C]   residual stream (never rounded to bf16: SURVEY.md section 7, bf16 tolerance analysis)
//   xb   bf16 [M,C]   copy of x feeding the next tensor-core GEMM (written by the LayerNorm kernel)
//   qkv  bf16 [M,3C]  fused q|k|v projection, token-major; attention reads 128-byte head rows straight from it
//   att  bf16 [M,C]   attention output, token-major (the reference's permute+contiguous never happens)
//   h    bf16 [M,4C]  GELU(fc1)
//   tmp  fp32 [M,C]   pre-LayerNorm sums (GEMM epilogue output)
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "kernels.cuh"

struct avexk_beats {
  avexk_beats_dims d;
  bool loaded = false;
  std::vector<void*> allocs;
  // packed weights (device)
  __nv_bfloat16 *patch_w = nullptr, *proj_w = nullptr, *posconv_w = nullptr;
  float *ln0_w, *ln0_b, *proj_b, *posconv_b, *enc_ln_w, *enc_ln_b;
  int precision = 0;  // 0: bf16 operands (default); 1: fp32 mode (3-term split GEMMs, fp32 attention and pos-conv)
  float* posconv_wf = nullptr;
  struct Layer {
    __nv_bfloat16 *qkv_w, *o_w, *fc1_w, *fc2_w;
    __nv_bfloat16 *qkv_w3 = nullptr, *o_w3 = nullptr, *fc1_w3 = nullptr, *fc2_w3 = nullptr;  // [hi|hi|lo] rows, fp32 mode
    float *qkv_b, *o_b, *fc1_b, *fc2_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b, *gate_w, *gate_b, *grep_a;
  };
  std::vector<Layer> layers;
};

namespace avexk {
namespace {

template <typename T>
int dev_alloc(avexk_beats* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return AVEXK_ECUDA;
  }
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return AVEXK_OK;
}

int copy_f32(avexk_beats* h, float** dst, const float* src, size_t n, cudaStream_t st) {
  int rc = dev_alloc(h, dst, n);
  if (rc) return rc;
  AVEXK_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return AVEXK_OK;
}

int pack_bf16_w(avexk_beats* h, __nv_bfloat16** dst, const float* src, size_t n, cudaStream_t st) {
  int rc = dev_alloc(h, dst, n);
  if (rc) return rc;
  return launch_f32_to_bf16(src, *dst, (long long)n, st);
}

struct Carver {
  char* p;
  size_t left;
  bool ok = true;
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (bytes > left) { ok = false; return nullptr; }
    T* r = reinterpret_cast<T*>(p);
    p += bytes;
    left -= bytes;
    return r;
  }
};

struct Plan {
  int B, T, F, N;
  long long M;
};

size_t workspace_bytes(const avexk_beats_dims& d, int B, int T, int precision = 0) {
  const long long F = avexk_fbank_num_frames(T), N = 8 * (F / 16), M = (long long)B * N;
  const long long C = d.embed, E = d.patch_embed, Ff = d.ffn, G = d.conv_groups;
  auto al = [](long long b) { return (size_t)((b + 255) & ~255LL); };
  size_t s = 0;
  s += al(M * 768 * 2);                 // patch operand [hi|lo|hi]
  s += al(M * E * 4);                   // patch-embed out
  s += al(M * E * 3 * 2);               // LN(512) bf16 [hi|lo|hi]
  s += al(M * C * 4);                   // x0 (hook 0) when the caller gives no buffer
  s += al(M * G * 64 * 2);              // group-padded pos-conv operand
  s += al(M * C * 4);                   // x
  s += al(M * C * 2);                   // xb
  s += al(M * 3 * C * 2);               // qkv
  s += al(M * C * 2);                   // att
  s += al(M * Ff * 2);                  // h
  s += al(M * C * 4);                   // tmp: pre-LN sums (pos-conv; GEMM epilogues when the LayerNorm is not fused)
  s += al((long long)gemm_ln_scratch_bytes((int)M));  // fused GEMM+LayerNorm: per-CTA tiles, row statistics, counters
  s += al((long long)(d.layers + 2) * B * C * 8);     // 40.24 fixed-point column sums of the fused mean-pooling (hooks + final)
  s += al(M * d.heads * 4);                           // per-(token, head) gate of the relative-position bias (QKV epilogue -> attention)
  if (precision == 1) {
    s += al(M * 3 * C * 2);   // xs   [hi|lo|hi] of x
    s += al(M * 3 * C * 4);   // qkv  fp32
    s += al(M * C * 4);       // att  fp32
    s += al(M * 3 * C * 2);   // atts [hi|lo|hi] of att
    s += al(M * Ff * 4);      // h    fp32
    s += al(M * 3 * Ff * 2);  // hs   [hi|lo|hi] of h
  }
  return s + 4096;
}

}  // namespace
}  // namespace avexk

extern "C" int avexk_beats_num_tokens(int T) { return 8 * (avexk_fbank_num_frames(T) / 16); }

extern "C" int avexk_beats_create(const avexk_beats_dims* dims, avexk_beats_t** out) {
  using namespace avexk;
  AVEXK_CHECK_ARG(dims && out, "avexk_beats_create: null argument");
  const avexk_beats_dims& d = *dims;
  AVEXK_CHECK_ARG(d.layers >= 1 && d.heads >= 1 && d.embed == d.heads * 64, "BEATs kernels need head_dim 64 (embed=%d heads=%d)", d.embed, d.heads);
  AVEXK_CHECK_ARG(d.embed % 128 == 0 && d.patch_embed % 128 == 0 && d.ffn % 64 == 0 && d.embed <= 1024 && d.patch_embed <= 1024,
                  "unsupported widths embed=%d patch_embed=%d ffn=%d", d.embed, d.patch_embed, d.ffn);
  AVEXK_CHECK_ARG(d.conv_pos == 128 && d.embed / d.conv_groups == 48, "pos-conv kernel needs 128 taps and 48 channels per group");
  auto* h = new avexk_beats();
  h->d = d;
  h->layers.resize(d.layers);
  *out = h;
  return AVEXK_OK;
}

extern "C" void avexk_beats_destroy(avexk_beats_t* h) {
  if (!h) return;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

extern "C" int avexk_beats_load_weights(avexk_beats_t* h, const avexk_beats_weights* w, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && w && w->layers, "avexk_beats_load_weights: null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->loaded = false;
  const avexk_beats_dims& d = h->d;
  const size_t C = d.embed, E = d.patch_embed, Ff = d.ffn, G = d.conv_groups, K = d.conv_pos, cg = C / G;
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  // front end in 3-term split bf16 (K tripled): [hi|hi|lo] weights against [hi|lo|hi] activations
  TRY(dev_alloc(h, &h->patch_w, E * 256 * 3));
  TRY(launch_f32_to_bf16_split3(w->patch_w, h->patch_w, (int)E, 256, st));
  TRY(copy_f32(h, &h->ln0_w, w->ln0_w, E, st));
  TRY(copy_f32(h, &h->ln0_b, w->ln0_b, E, st));
  TRY(dev_alloc(h, &h->proj_w, C * E * 3));
  TRY(launch_f32_to_bf16_split3(w->proj_w, h->proj_w, (int)C, (int)E, st));
  TRY(copy_f32(h, &h->proj_b, w->proj_b, C, st));
  TRY(copy_f32(h, &h->posconv_b, w->posconv_b, C, st));
  TRY(copy_f32(h, &h->enc_ln_w, w->enc_ln_w, C, st));
  TRY(copy_f32(h, &h->enc_ln_b, w->enc_ln_b, C, st));
  float* nrm = nullptr;
  TRY(dev_alloc(h, &nrm, K));
  TRY(dev_alloc(h, &h->posconv_w, C * K * 64));
  TRY(launch_posconv_pack(w->posconv_v, w->posconv_g, (int)C, (int)cg, (int)K, nrm, h->posconv_w, st));
  if (h->precision == 1) {
    TRY(dev_alloc(h, &h->posconv_wf, C * cg * K));
    TRY(launch_posconv_pack_f32(w->posconv_v, w->posconv_g, nrm, (int)G, (int)cg, (int)K, h->posconv_wf, st));
  }
  for (int li = 0; li < d.layers; ++li) {
    const avexk_beats_layer_weights& s = w->layers[li];
    avexk_beats::Layer& L = h->layers[li];
    TRY(dev_alloc(h, &L.qkv_w, 3 * C * C));
    TRY(launch_f32_to_bf16(s.q_w, L.qkv_w, (long long)(C * C), st));
    TRY(launch_f32_to_bf16(s.k_w, L.qkv_w + C * C, (long long)(C * C), st));
    TRY(launch_f32_to_bf16(s.v_w, L.qkv_w + 2 * C * C, (long long)(C * C), st));
    TRY(dev_alloc(h, &L.qkv_b, 3 * C));
    AVEXK_CUDA(cudaMemcpyAsync(L.qkv_b, s.q_b, C * 4, cudaMemcpyDeviceToDevice, st));
    AVEXK_CUDA(cudaMemcpyAsync(L.qkv_b + C, s.k_b, C * 4, cudaMemcpyDeviceToDevice, st));
    AVEXK_CUDA(cudaMemcpyAsync(L.qkv_b + 2 * C, s.v_b, C * 4, cudaMemcpyDeviceToDevice, st));
    TRY(pack_bf16_w(h, &L.o_w, s.o_w, C * C, st));
    TRY(copy_f32(h, &L.o_b, s.o_b, C, st));
    TRY(pack_bf16_w(h, &L.fc1_w, s.fc1_w, Ff * C, st));
    TRY(copy_f32(h, &L.fc1_b, s.fc1_b, Ff, st));
    TRY(pack_bf16_w(h, &L.fc2_w, s.fc2_w, C * Ff, st));
    TRY(copy_f32(h, &L.fc2_b, s.fc2_b, C, st));
    if (h->precision == 1) {
      TRY(dev_alloc(h, &L.qkv_w3, 3 * C * 3 * C));
      TRY(launch_f32_to_bf16_split3(s.q_w, L.qkv_w3, (int)C, (int)C, st));
      TRY(launch_f32_to_bf16_split3(s.k_w, L.qkv_w3 + C * 3 * C, (int)C, (int)C, st));
      TRY(launch_f32_to_bf16_split3(s.v_w, L.qkv_w3 + 2 * C * 3 * C, (int)C, (int)C, st));
      TRY(dev_alloc(h, &L.o_w3, C * 3 * C));
      TRY(launch_f32_to_bf16_split3(s.o_w, L.o_w3, (int)C, (int)C, st));
      TRY(dev_alloc(h, &L.fc1_w3, Ff * 3 * C));
      TRY(launch_f32_to_bf16_split3(s.fc1_w, L.fc1_w3, (int)Ff, (int)C, st));
      TRY(dev_alloc(h, &L.fc2_w3, C * 3 * Ff));
      TRY(launch_f32_to_bf16_split3(s.fc2_w, L.fc2_w3, (int)C, (int)Ff, st));
    }
    TRY(copy_f32(h, &L.ln1_w, s.ln1_w, C, st));
    TRY(copy_f32(h, &L.ln1_b, s.ln1_b, C, st));
    TRY(copy_f32(h, &L.ln2_w, s.ln2_w, C, st));
    TRY(copy_f32(h, &L.ln2_b, s.ln2_b, C, st));
    TRY(copy_f32(h, &L.grep_a, s.grep_a, d.heads, st));
    TRY(dev_alloc(h, &L.gate_w, 128));
    TRY(dev_alloc(h, &L.gate_b, 2));
    TRY(launch_gate_pack(s.grep_w, s.grep_b, L.gate_w, L.gate_b, st));
  }
#undef TRY
  AVEXK_CUDA(cudaStreamSynchronize(st));
  h->loaded = true;
  return AVEXK_OK;
}

extern "C" size_t avexk_beats_workspace_bytes(const avexk_beats_t* h, int B, int T) {
  if (!h || B <= 0 || T < 400) return 0;
  return avexk::workspace_bytes(h->d, B, T, h->precision);
}

extern "C" int avexk_beats_set_precision(avexk_beats_t* h, int fp32_mode) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && (fp32_mode == 0 || fp32_mode == 1), "avexk_beats_set_precision: bad argument");
  if (h->precision != fp32_mode) {
    h->precision = fp32_mode;
    h->loaded = false;  // the split weight copies are packed by avexk_beats_load_weights
  }
  return AVEXK_OK;
}

extern "C" int avexk_beats_forward(avexk_beats_t* h, const float* wav, int B, int T, long long wav_stride,
                                   const avexk_fbank_t* fbank, const uint8_t* key_pad, const float* bias_vec, float* out,
                                   float* const* hook_out, float* const* hook_pooled, float* pooled, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  using namespace avexk;
  AVEXK_CHECK_ARG(h && h->loaded, "avexk_beats_forward: weights not loaded");
  AVEXK_CHECK_ARG(wav && fbank && bias_vec && workspace, "avexk_beats_forward: null argument");
  AVEXK_CHECK_ARG(out || pooled || hook_out || hook_pooled, "avexk_beats_forward: no output requested");
  const avexk_beats_dims& d = h->d;
  const int N = avexk_beats_num_tokens(T);
  AVEXK_CHECK_ARG(B > 0 && N > 0, "avexk_beats_forward: clip too short (B=%d T=%d -> %d tokens)", B, T, N);
  const long long M = (long long)B * N;
  AVEXK_CHECK_ARG(M < (1LL << 31), "avexk_beats_forward: too many token rows (%lld)", M);
  AVEXK_CHECK_ARG(workspace_bytes >= avexk::workspace_bytes(d, B, T, h->precision), "avexk_beats_forward: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = d.embed, E = d.patch_embed, Ff = d.ffn, G = d.conv_groups, H = d.heads;
  const float alpha = powf(2.0f * (float)d.layers, 0.25f);  // backbone.py:306

  Carver cw{reinterpret_cast<char*>(workspace), workspace_bytes};
  __nv_bfloat16* pa = cw.take<__nv_bfloat16>((size_t)M * 768);
  float* pe = cw.take<float>((size_t)M * E);
  __nv_bfloat16* peb = cw.take<__nv_bfloat16>((size_t)M * E * 3);
  float* x0_ws = cw.take<float>((size_t)M * C);
  __nv_bfloat16* xg = cw.take<__nv_bfloat16>((size_t)M * G * 64);
  float* x = cw.take<float>((size_t)M * C);
  __nv_bfloat16* xb = cw.take<__nv_bfloat16>((size_t)M * C);
  __nv_bfloat16* qkv = cw.take<__nv_bfloat16>((size_t)M * 3 * C);
  __nv_bfloat16* att = cw.take<__nv_bfloat16>((size_t)M * C);
  __nv_bfloat16* hb = cw.take<__nv_bfloat16>((size_t)M * Ff);
  float* tmp = cw.take<float>((size_t)M * C);
  const size_t ln_ws_bytes = gemm_ln_scratch_bytes((int)M);
  char* ln_ws = cw.take<char>(ln_ws_bytes);
  long long* pool_acc = cw.take<long long>((size_t)(d.layers + 2) * B * C);  // [layers + 1 hooks | final][B][C]
  float* gate = cw.take<float>((size_t)M * d.heads);
  AVEXK_CHECK_ARG(cw.ok, "avexk_beats_forward: workspace carve failed");
  float* x0 = (hook_out && hook_out[0]) ? hook_out[0] : x0_ws;
  auto hpool = [&](int i) -> float* { return hook_pooled ? hook_pooled[i] : nullptr; };
  // mean-pooling fused into the GEMM+LayerNorm epilogue (column sums per clip, no pass over [M,C]) needs the fused epilogue and
  // at least one 32-row box per clip; otherwise the tensor is materialised and pooled by mean_pool_kernel
  bool any_pool = pooled != nullptr;
  for (int i = 0; i <= d.layers && hook_pooled; ++i) any_pool = any_pool || hook_pooled[i] != nullptr;

  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  auto gemm = [&](const void* A, int K, const __nv_bfloat16* W, int Nn, const float* bias, int gelu, float* raw, const float* res,
                  float rs, void* o, int obf) -> int {
    return gemm_bf16_launch(A, K, W, K, (int)M, Nn, K, bias, gelu, raw, res, rs, o, Nn, obf, st);
  };
  // GEMM + DeepNorm residual + LayerNorm in one launch when the width matches the fused epilogue (BEATs-base: 768)
  // AVEXK_FUSE_LN: 0 = separate LayerNorm launches, 1 = fc2 only, 2 (default) = out_proj and fc2.  Measured at config #2
  // with the TMEM-resident LayerNorm epilogue: 33.1 ms (1) vs 32.8 ms (2) per step.
  static const int fuse_level = [] { const char* e = getenv("AVEXK_FUSE_LN"); return e ? atoi(e) : 2; }();
  const bool fuse_ln = C == 768 && fuse_level > 0;
  const bool fuse_pool = fuse_ln && N >= 32 && h->precision != 1;
  if (any_pool && fuse_pool) AVEXK_CUDA(cudaMemsetAsync(pool_acc, 0, (size_t)(d.layers + 2) * B * C * sizeof(long long), st));
  int ln_first = 1;  // the scratch counters are zeroed once per forward; every launch leaves them zero
  // pool_raw / pool_y: fixed-point accumulators of the fused pooling (column sums of the raw Linear output / of the LN output)
  auto gemm_ln = [&](const void* A, int K, const __nv_bfloat16* W, const float* bias, float* raw, const float* gamma, const float* beta,
                     float* dst_f32, __nv_bfloat16* dst_bf16, long long* pool_raw, long long* pool_y) -> int {
    if (fuse_ln && (fuse_level >= 2 || K > C)) {
      const int zero = ln_first;
      ln_first = 0;
      const int r = gemm_bf16_ln_launch(A, K, W, K, (int)M, K, bias, raw, x, alpha, gamma, beta, d.ln_eps, dst_f32, dst_bf16, ln_ws, ln_ws_bytes,
                                        zero, st, pool_raw, pool_y, N);
      // AVEXK_ENOMEM: the grid of the fused epilogue cannot be co-resident on this device partition -> separate launches below
      // (only possible without fused pooling, which the caller requests only when it has checked the same condition)
      if (r != AVEXK_ENOMEM || pool_raw != nullptr || pool_y != nullptr) return r;
      if (dst_f32 == nullptr) dst_f32 = x;
    }
    int r = gemm_bf16_launch(A, K, W, K, (int)M, C, K, bias, 0, raw, x, alpha, tmp, C, 0, st);
    if (r) return r;
    return launch_layernorm(tmp, (int)M, C, gamma, beta, d.ln_eps, dst_f32, dst_bf16, st);
  };

  // ---- front end: fbank -> patch embed -> LN -> post_extract_proj (beats.py:344-359) -----------------------------
  // the fbank kernel writes the [hi|lo|hi] operand of the patch-embedding GEMM directly: no [B,F,128] tensor, no im2col pass
  TRY(avexk_fbank_patch_operand(fbank, wav, B, T, wav_stride, 32768.0f, d.fbank_mean, 1.0f / (2.0f * d.fbank_std), pa, stream));
  TRY(gemm(pa, 768, h->patch_w, E, nullptr, 0, nullptr, nullptr, 0.f, pe, 0));
  TRY(launch_layernorm(pe, (int)M, E, h->ln0_w, h->ln0_b, d.ln_eps, nullptr, peb, st, 1));
  TRY(gemm(peb, 3 * E, h->proj_w, C, h->proj_b, 0, nullptr, nullptr, 0.f, x0, 0));
  // ---- encoder prologue: mask, pos-conv + GELU + residual, LN (backbone.py:169-177) ------------------------------
  TRY(launch_group_pad(x0, key_pad, M, G, C / G, xg, st));
  if (hpool(0)) TRY(launch_mean_pool(x0, nullptr, 0, B, N, C, hpool(0), st));  // hooks pool over ALL tokens (base_model.py:419-453)
  if (h->precision != 1) {
    TRY(launch_posconv(xg, h->posconv_w, h->posconv_b, x0, tmp, B, N, G, C / G, d.conv_pos, st));
    TRY(launch_layernorm(tmp, (int)M, C, h->enc_ln_w, h->enc_ln_b, d.ln_eps, x, xb, st));
  }
  if (h->precision == 1) {
    // ---- fp32 mode: 3-term split-bf16 GEMMs (K tripled), fp32 attention / pos-conv / residual stream ----------------------
    __nv_bfloat16* xs = cw.take<__nv_bfloat16>((size_t)M * 3 * C);
    float* qkv32 = cw.take<float>((size_t)M * 3 * C);
    float* att32 = cw.take<float>((size_t)M * C);
    __nv_bfloat16* atts = cw.take<__nv_bfloat16>((size_t)M * 3 * C);
    float* h32 = cw.take<float>((size_t)M * Ff);
    __nv_bfloat16* hs = cw.take<__nv_bfloat16>((size_t)M * 3 * Ff);
    AVEXK_CHECK_ARG(cw.ok, "avexk_beats_forward: workspace carve failed (fp32 mode)");
    auto gemm3 = [&](const void* A, int K3, const __nv_bfloat16* W, int Nn, const float* bias, int gelu, float* raw, const float* res,
                     float* o) -> int {
      return gemm_bf16_launch(A, K3, W, K3, (int)M, Nn, K3, bias, gelu, raw, res, alpha, o, Nn, 0, st);
    };
    TRY(launch_posconv_fp32(x0, h->posconv_wf, h->posconv_b, tmp, B, N, G, C / G, d.conv_pos, st));
    TRY(launch_layernorm(tmp, (int)M, C, h->enc_ln_w, h->enc_ln_b, d.ln_eps, x, xs, st, 1));
    for (int li = 0; li < d.layers; ++li) {
      const avexk_beats::Layer& L = h->layers[li];
      const bool last = li == d.layers - 1;
      TRY(gemm3(xs, 3 * C, L.qkv_w3, 3 * C, L.qkv_b, 0, nullptr, nullptr, qkv32));
      TRY(launch_attention_fp32(qkv32, B, N, H, L.gate_w, L.gate_b, L.grep_a, bias_vec, key_pad, att32, st));
      TRY(launch_split3_rows(att32, M, C, atts, st));
      TRY(gemm3(atts, 3 * C, L.o_w3, C, L.o_b, 0, nullptr, x, tmp));
      TRY(launch_layernorm(tmp, (int)M, C, L.ln1_w, L.ln1_b, d.ln_eps, x, xs, st, 1));
      TRY(gemm3(xs, 3 * C, L.fc1_w3, Ff, L.fc1_b, 1, nullptr, nullptr, h32));
      TRY(launch_split3_rows(h32, M, Ff, hs, st));
      float* raw = hook_out ? hook_out[li + 1] : nullptr;
      if (!raw && hpool(li + 1)) raw = qkv32;  // free by now: scratch for the raw fc2 output that is pooled below
      TRY(gemm3(hs, 3 * Ff, L.fc2_w3, C, L.fc2_b, 0, raw, x, tmp));
      if (hpool(li + 1)) TRY(launch_mean_pool(raw, nullptr, 0, B, N, C, hpool(li + 1), st));
      float* dst = (last && out) ? out : x;
      TRY(launch_layernorm(tmp, (int)M, C, L.ln2_w, L.ln2_b, d.ln_eps, dst, last ? nullptr : xs, st, 1));
      if (last && pooled) TRY(launch_mean_pool(dst, key_pad, key_pad != nullptr, B, N, C, pooled, st));
    }
    return AVEXK_OK;
  }
  // ---- 12 post-LN DeepNorm blocks (backbone.py:350-373) --------------------------------------------------------------
  for (int li = 0; li < d.layers; ++li) {
    const avexk_beats::Layer& L = h->layers[li];
    const bool last = li == d.layers - 1;
    const GemmGate gg{L.gate_w, L.gate_b, L.grep_a, gate, H};  // the gates leave the QKV epilogue with q still in fp32
    TRY(gemm_bf16_launch(xb, C, L.qkv_w, C, (int)M, 3 * C, C, L.qkv_b, 0, nullptr, nullptr, 0.f, qkv, 3 * C, 1, st, &gg));
    TRY(attention_launch(qkv, B, N, H, gate, bias_vec, key_pad, att, st));
    TRY(gemm_ln(att, C, L.o_w, L.o_b, nullptr, L.ln1_w, L.ln1_b, x, xb, nullptr, nullptr));
    TRY(gemm(xb, C, L.fc1_w, Ff, L.fc1_b, 1, nullptr, nullptr, 0.f, hb, 1));
    float* raw = hook_out ? hook_out[li + 1] : nullptr;
    float* hp = hpool(li + 1);
    long long* acc_raw = (hp && fuse_pool) ? pool_acc + (size_t)li * B * C : nullptr;
    if (hp && !fuse_pool && !raw) raw = reinterpret_cast<float*>(qkv);  // scratch (free by now): M x 3C bf16 >= M x C fp32
    // final features: written only when the caller wants them, or when a masked mean-pool has to re-read them; the plain
    // mean-pool of the last layer is a by-product of its LayerNorm epilogue
    const bool pool_fused = last && pooled && fuse_pool && key_pad == nullptr;
    long long* acc_y = pool_fused ? pool_acc + (size_t)(d.layers + 1) * B * C : nullptr;
    float* dst = last ? (out ? out : ((pooled && !pool_fused) || !fuse_ln ? x : nullptr)) : x;
    TRY(gemm_ln(hb, Ff, L.fc2_w, L.fc2_b, raw, L.ln2_w, L.ln2_b, dst, last ? nullptr : xb, acc_raw, acc_y));
    if (acc_raw) TRY(launch_pool_finalize(acc_raw, B, C, 1.0f / N, hp, st));
    else if (hp) TRY(launch_mean_pool(raw, nullptr, 0, B, N, C, hp, st));
    if (acc_y) TRY(launch_pool_finalize(acc_y, B, C, 1.0f / N, pooled, st));
    else if (last && pooled) {
      // any_pad is decided on the host by the caller passing key_pad == NULL when nothing is padded
      TRY(launch_mean_pool(dst, key_pad, key_pad != nullptr, B, N, C, pooled, st));
    }
  }
#undef TRY
  return AVEXK_OK;
}
