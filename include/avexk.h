/*
 * avexk.h -- C ABI of libavexk.so: the B200 (sm_100a) kernels behind avex's embedding hot path.
 *
 * This is the drop-in boundary one step below the reference's Python plugin classes
 * (avex/models/beats_model.py:72 `Model`, avex/models/efficientnet.py:22 `Model`): it replaces the torch
 * ops those classes reach through avex/models/beats/{beats,backbone}.py and avex/data/audio_utils.py.
 * The reference has no FFI of its own (pure Python); the binding a maintainer adds is the ctypes stub in
 * avex_b200/_lib.py, described in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; every `const float* x` / `void* out` below is a DEVICE pointer unless the
 *    name ends in `_host`.  The caller (torch) owns all buffers and keeps them alive until the stream work is done.
 *  - every entry point is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing on the
 *    hot path (handles allocate at create / load time) and returns 0 on success, a negative AVEXK_E* code on
 *    failure; `avexk_last_error()` returns a thread-local message.
 *  - handles are thread-compatible, not thread-safe.
 */
#ifndef AVEXK_H
#define AVEXK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVEXK_OK 0
#define AVEXK_EINVAL (-1)  /* bad argument / unsupported shape */
#define AVEXK_ECUDA (-2)   /* CUDA runtime / driver error       */
#define AVEXK_ENOMEM (-3)  /* workspace too small               */

const char* avexk_last_error(void);
int avexk_version(void);
/* sha256 prefix (32 hex digits) of the sources this binary was compiled from (csrc/*.cu, csrc/*.cuh, this header);
 * avex_b200/_lib.py refuses to bind a library whose id differs from the sources next to it. */
const char* avexk_build_id(void);
/* number of kernel launches this library has enqueued since load (bench.py's `gpu_launches`). */
long long avexk_launch_count(void);

/* Optional per-kernel timing with CUDA events on the launching stream (measurement only; adds two event records
 * per launch while enabled).  kid: 0 fbank (work = algorithmic bytes), 1 gemm (FLOPs), 2 attention (FLOPs),
 * 3 layernorm (bytes), 4 posconv (FLOPs), 5 other.  avexk_profile_read synchronises on the recorded events. */
void avexk_profile_enable(int on);
int avexk_profile_read(int kid, long long* launches, double* total_ms, double* total_work);

/* ------------------------------------------------------------------------------------------------------------
 * Kaldi-style log-mel filterbank.
 * Replaces `_BatchedFbank.forward` + the affine of `BEATs.preprocess`
 *   (avex/models/beats/beats.py:120-163, :304-323) and, with window=hanning / prescale=1 / pad_to_frames=1024,
 *   `EATAudioProcessor.__call__` (avex/models/eat/audio_processor.py:72-143).
 * Geometry is fixed to the reference's: 16 kHz, 25 ms / 10 ms (400 / 160 samples), n_fft 512, 128 mel bins.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct avexk_fbank avexk_fbank_t;

/* window_host[400] and mel_fb_host[257*128] (row-major [fft_bin][mel_bin]) are HOST arrays built by the caller
 * with the reference's own fp32 formulas (beats.py:75, :82-118); the library derives the sparse mel table
 * (<= 2 non-zeros per FFT bin) and double-precision twiddles and uploads them to the current device. */
int avexk_fbank_create(const float* window_host, const float* mel_fb_host, avexk_fbank_t** out);
void avexk_fbank_destroy(avexk_fbank_t* h);

/* frames = 1 + (T - 400) / 160 (snip_edges, beats.py:136); 0 when T < 400. */
int avexk_fbank_num_frames(int T);

/* wav [B, T] fp32 with row stride `wav_stride` elements.  out [B, out_frames, 128] fp32.
 *   out[b,f,m] = (log(max(mel, FLT_EPSILON)) - norm_mean) * norm_scale        (beats.py:163, :323)
 * prescale: 32768 for BEATs (beats.py:322), 1 for EAT.  out_frames <= 0 means "num_frames(T)"; larger values
 * zero-pad in the log-mel domain before normalisation, smaller truncate (eat/audio_processor.py:121-126).
 * per_utt != 0: ignore norm_mean/scale and normalise each clip with its own mean and unbiased std,
 *   (x - mu) / (2 sigma) (eat/audio_processor.py:132-135); needs stats_ws >= B * 2 doubles (device).
 * out_bf16 != 0 stores bf16 instead of fp32. */
int avexk_fbank_forward(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride, float prescale,
                        float norm_mean, float norm_scale, int out_frames, int per_utt, double* stats_ws, void* out,
                        int out_bf16, void* stream);

/* The same front end writing the operand of the 16x16 patch-embedding GEMM directly (beats.py:349-352: Conv2d(1, 512, 16, stride 16)
 * as im2col), so that the [B, F, 128] fbank never touches HBM:
 *   out [B * N, 768] bf16, N = 8 * (frames / 16);  row = b * N + tp * 8 + fp,  col = i * 16 + j  <->  fbank[b, tp*16 + i, fp*16 + j]
 *   columns [0,256) = bf16(v), [256,512) = bf16(v - bf16(v)), [512,768) = bf16(v) again (the [hi|lo|hi] half of the 3-term split
 *   product with [hi|hi|lo] weights).  Bit-identical to splitting the fp32 output of avexk_fbank_forward. */
int avexk_fbank_patch_operand(const avexk_fbank_t* h, const float* wav, int B, int T, long long wav_stride, float prescale,
                              float norm_mean, float norm_scale, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Building blocks (each unit-testable against the oracle)
 * ---------------------------------------------------------------------------------------------------------- */

/* D = epilogue(A @ W^T): A [M,K] bf16 row-major (lda), W [N,K] bf16 row-major (nn.Linear layout), fp32
 * accumulation on the tcgen05 tensor cores (TMA-fed, accumulators in TMEM).
 *   v = acc + bias[n]                                (bias may be NULL)
 *   if gelu:      v = 0.5 v (1 + erf(v / sqrt 2))    (modules.py:191-200)
 *   if raw_out:   raw_out[m,n] = v   (fp32; the tensor a forward hook on the Linear would see)
 *   if residual:  v = v + res_scale * residual[m,n]  (fp32 residual; backbone.py:360, :372)
 *   out[m,n] = v  as fp32 (out_bf16 == 0) or bf16    (out may be NULL when only raw_out is wanted)
 * Requirements: K % 8 == 0, N % 8 == 0 (TMA zero-fills the ragged last K block / N tile), all pointers 16-byte aligned,
 * ld* % 8 == 0. */
int avexk_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias,
                    int gelu, float* raw_out, const float* residual, float res_scale, void* out, long long ldo,
                    int out_bf16, void* stream);

/* The same GEMM with the post-LN block tail fused (backbone.py:360-362, :372-373), N == 768 only:
 *   v = A @ W^T + bias;  raw_out = v (optional hook);  y = LayerNorm(v + res_scale * residual) * gamma + beta
 * y is written as fp32 (out_f32, may alias `residual`) and / or bf16 (out_bf16).  `scratch` holds the pre-LN sums of the
 * row blocks in flight (avexk_gemm_ln_scratch_bytes(M) bytes, stays L2-resident); nothing of size [M,768] fp32 other than
 * the residual read and the y write touches HBM. */
size_t avexk_gemm_ln_scratch_bytes(int M);
int avexk_gemm_bf16_ln(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias,
                       float* raw_out, const float* residual, float res_scale, const float* gamma, const float* beta, float eps,
                       float* out_f32, void* out_bf16, void* scratch, size_t scratch_bytes, void* stream);
/* The same launch with mean-pooling over the token rows of every clip fused into the epilogue (the rows are M / rows_per_clip
 * clips of rows_per_clip >= 32 consecutive rows): pooled_raw [clips,768] = mean of v (what `extract_embeddings(aggregation="mean")`
 * reduces a hooked fc2 output to, base_model.py:419-453), pooled_y [clips,768] = mean of y (beats_model.py:275); either may be
 * NULL, and y itself need not be written (out_f32 == out_bf16 == NULL).  Column sums are accumulated in 40.24 fixed point, so
 * the result does not depend on the order of the atomics.  pool_ws: 2 * clips * 768 * 8 bytes. */
int avexk_gemm_bf16_ln_pooled(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias,
                              float* raw_out, const float* residual, float res_scale, const float* gamma, const float* beta,
                              float eps, float* out_f32, void* out_bf16, void* scratch, size_t scratch_bytes, int rows_per_clip,
                              float* pooled_raw, float* pooled_y, void* pool_ws, void* stream);
/* Tuning knob: 1 (default) = CTA pairs (tcgen05 cta_group::2, 256x256 tiles), 0 = single-CTA 128x256 tiles.  Returns the
 * previous setting; any other argument only queries.  Environment override at first use: AVEXK_GEMM_PAIR=0. */
int avexk_gemm_config(int pair);

/* Row LayerNorm over the last dimension C (eps 1e-5): y = (x - mu) / sqrt(var + eps) * gamma + beta.
 * x [M,C] fp32; writes out_f32 and / or out_bf16 (either may be NULL). */
int avexk_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, float* out_f32,
                    void* out_bf16, void* stream);

/* Gated relative-position-bias attention (backbone.py:494-574), one launch for all (b, h).
 * qkv [B*N, 3*H*64] bf16 (q | k | v, each head-major within H*64), as written by the fused QKV GEMM.
 * gate_w [2,64], gate_b [2]: grep_linear rows pre-summed in groups of four (backbone.py:547-549);
 * grep_a [H]; bias_vec [H, 2N-1] with bias_vec[h, (j-i)+N-1] = table[bucket(j-i), h] (backbone.py:475-492);
 * key_pad [B,N] bytes (1 = padded key -> -inf) or NULL.  out [B*N, H*64] bf16 (token-major, ready for out_proj).
 *   out_i = softmax_j(q_i.k_j / 8 + gate_i * bias[h, j-i] + pad_j) . v_j,  gate from UNscaled q.
 * Rows of a clip whose keys are ALL padded come out as zeros (the reference's softmax is NaN there).  The gate logits are
 * formed on the tensor core with gate_w split into bf16 hi + lo parts (~16 mantissa bits). */
int avexk_attention_gated(const void* qkv, int B, int N, int H, const float* gate_w, const float* gate_b,
                          const float* grep_a, const float* bias_vec, const uint8_t* key_pad, void* out, void* stream);

/* Convolutional position embedding (backbone.py:52-68, :172-174; modules.py:67-94):
 *   out = x0 + GELU(Conv1d(C, C, k=taps, pad=taps/2, groups)(x0 along tokens)[..., :N] + bias),  weight = g * v / ||v||_(0,1)
 * x0 [B,N,C] fp32 -- rows with key_pad != 0 are zeroed IN PLACE first (backbone.py:169-170); weight_g [taps],
 * weight_v [C, C/groups, taps] fp32 (the weight_norm parametrisation); out [B,N,C] fp32.  Packs the weights on
 * every call (unit-test / building-block entry; avexk_beats_forward packs once at load).
 * workspace >= avexk_posconv_workspace_bytes(B, N, C, groups, taps). */
size_t avexk_posconv_workspace_bytes(int B, int N, int C, int groups, int taps);
int avexk_posconv(float* x0, int B, int N, int C, int groups, int taps, const float* weight_g, const float* weight_v,
                  const float* bias, const uint8_t* key_pad, float* out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BEATs encoder (avex/models/beats/beats.py:325-382 + backbone.py:151-221), whole forward in one call.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct avexk_beats avexk_beats_t;

typedef struct {
  int layers;       /* 12  */
  int embed;        /* 768 */
  int ffn;          /* 3072 */
  int heads;        /* 12  */
  int patch_embed;  /* 512 */
  int conv_pos;     /* 128 */
  int conv_groups;  /* 16  */
  float fbank_mean; /* 15.41663 */
  float fbank_std;  /* 6.55582  */
  float ln_eps;     /* 1e-5 */
} avexk_beats_dims;

/* fp32 DEVICE pointers in the reference's own layouts (state_dict names in comments). */
typedef struct {
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b; /* self_attn.{q,k,v,out}_proj.{weight [C,C],bias}  */
  const float *grep_w, *grep_b, *grep_a;                       /* self_attn.grep_linear.{weight [8,64],bias}, grep_a */
  const float *ln1_w, *ln1_b;                                  /* self_attn_layer_norm                           */
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;                  /* fc1 [Ff,C], fc2 [C,Ff]                          */
  const float *ln2_w, *ln2_b;                                  /* final_layer_norm                               */
} avexk_beats_layer_weights;

typedef struct {
  const float* patch_w;            /* backbone.patch_embedding.weight [E,1,16,16]                               */
  const float *ln0_w, *ln0_b;      /* backbone.layer_norm [E]                                                   */
  const float *proj_w, *proj_b;    /* backbone.post_extract_proj [C,E]                                          */
  const float *posconv_g;          /* encoder.pos_conv.0.parametrizations.weight.original0 [1,1,K]              */
  const float *posconv_v;          /* ...original1 [C, C/groups, K]                                             */
  const float *posconv_b;          /* encoder.pos_conv.0.bias [C]                                               */
  const float *enc_ln_w, *enc_ln_b;/* encoder.layer_norm [C]                                                    */
  const float* rel_bias_table;     /* layers.0.self_attn.relative_attention_bias.weight [buckets, H]            */
  const avexk_beats_layer_weights* layers; /* HOST array of `dims.layers` entries                              */
} avexk_beats_weights;

int avexk_beats_create(const avexk_beats_dims* dims, avexk_beats_t** out);
void avexk_beats_destroy(avexk_beats_t* h);
/* Packs bf16 copies (fused QKV [3C,C], weight-norm resolved pos-conv, gate rows pre-summed). Synchronises. */
int avexk_beats_load_weights(avexk_beats_t* h, const avexk_beats_weights* w, void* stream);

/* fp32 mode (north_star: max-abs <= 1e-3 against the fp32 reference): every nn.Linear as a 3-term split-bf16 GEMM on the
 * same tcgen05 kernel (K tripled, ~16 mantissa bits per operand), q/k/v/P, attention and the pos-conv in plain fp32 on the
 * CUDA cores, fp32 residual stream as in the default mode.  Call BEFORE avexk_beats_load_weights (which packs the split
 * weight copies); changes avexk_beats_workspace_bytes.  fp32_mode: 0 = bf16 operands (default), 1 = fp32 mode. */
int avexk_beats_set_precision(avexk_beats_t* h, int fp32_mode);

/* tokens for T samples: 8 * floor(num_frames(T) / 16). */
int avexk_beats_num_tokens(int T);
size_t avexk_beats_workspace_bytes(const avexk_beats_t* h, int B, int T);

/* wav [B,T] fp32 (row stride wav_stride).  key_pad [B,N] bytes or NULL (already reduced from the sample mask by
 * the host, beats.py:283-302).  bias_vec [H, 2N-1] fp32, host-precomputed from rel_bias_table with the
 * reference's bucket function (backbone.py:438-492).
 * out       [B,N,C] fp32 final features (may be NULL when only pooled is wanted)
 * hook_out  HOST array of layers+1 DEVICE pointers (NULL entry = not materialised):
 *             [0]   post_extract_proj output [B,N,C] (rows of padded tokens zeroed, as the reference's in-place
 *                   `x[padding_mask] = 0` makes a hook see them, backbone.py:169-170)
 *             [i+1] raw fc2 output of block i [B,N,C]   (beats_model.py:206-227)
 * hook_pooled HOST array of layers+1 DEVICE pointers [B,C] (NULL entry / NULL array = not wanted): the same tensors mean-pooled
 *           over ALL tokens (what `extract_embeddings(aggregation="mean")` makes of a hook, base_model.py:419-453) without
 *           materialising [B,N,C] -- a by-product of the fc2 epilogue (probes on layer-wise features)
 * pooled    [B,C] fp32 mean over tokens (masked mean when key_pad has padded tokens, beats_model.py:269-275); may be NULL. */
int avexk_beats_forward(avexk_beats_t* h, const float* wav, int B, int T, long long wav_stride,
                        const avexk_fbank_t* fbank, const uint8_t* key_pad, const float* bias_vec, float* out,
                        float* const* hook_out, float* const* hook_pooled, float* pooled, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * STFT mel spectrogram of the EfficientNet path.
 * Replaces `AudioProcessor.__call__` / `_normalize` (avex/data/audio_utils.py:106-172) for the configuration of
 *   api/configs/official_models/esp_aves2_effnetb0_all.yml: n_fft = win = 800, hop 160, hann (periodic), center=True
 *   (reflect pad 400), power spectrogram, 128 HTK mel bins 0..8 kHz (torchaudio MelScale, norm=None), log(x + 1e-6),
 *   per-clip min-max.  Geometry is fixed to those values.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct avexk_melspec avexk_melspec_t;

/* window_host[800] = torch.hann_window(800); mel_fb_host[401*128] row-major [fft_bin][mel_bin] =
 * torchaudio.functional.melscale_fbanks(401, 0, 8000, 128, 16000) -- HOST arrays built by the caller with the
 * reference's own ops (audio_utils.py:97-101, :165); the library derives the sparse per-filter bin ranges and the
 * double-precision DFT twiddles. */
int avexk_melspec_create(const float* window_host, const float* mel_fb_host, avexk_melspec_t** out);
void avexk_melspec_destroy(avexk_melspec_t* h);
/* frames = 1 + T / 160 (torch.stft, center=True); T must exceed 400 (reflect padding). */
int avexk_melspec_num_frames(int T);
/* wav [B,T] fp32 (row stride wav_stride) -> out [B,128,frames] fp32 (frequency-major, time last, as the reference).
 * minmax_ws: B*2 uint32 device words, receives the per-clip min / max of log(mel + 1e-6) in an order-preserving
 *   encoding (consumed by avexk_effnet_forward); may be NULL when normalize == 0.
 * normalize != 0: out = (y - min) / (max - min + 1e-8) per clip (audio_utils.py:167-172); otherwise out = y. */
int avexk_melspec_forward(const avexk_melspec_t* h, const float* wav, int B, int T, long long wav_stride, int normalize,
                          float* out, void* minmax_ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * EfficientNet building blocks (NHWC fp16 activations) and the whole feature extractor
 * (avex/models/efficientnet.py:163-215 -> torchvision efficientnet_b0/b1 .features / .avgpool / .classifier).
 * ---------------------------------------------------------------------------------------------------------- */

/* The EfficientNet path stores activations and tensor-core operands as FP16 (bounded post-BatchNorm / SiLU values; same
 * tensor-core rate as bf16, three more mantissa bits): "16" below means IEEE half.
 * 1x1 convolution == GEMM on the tcgen05 kernel: A [M,K] fp16 (NHWC rows), W [N,K] fp16 (conv weight [N,K,1,1]).
 *   acc = A @ W^T ; raw_out[m,n] = acc (fp32, the pre-BatchNorm tensor a hook on the conv sees; may be NULL)
 *   y = acc * scale[n] + shift[n] (folded BatchNorm; scale NULL = 1) ; silu != 0: y = y * sigmoid(y)
 *   res_f16 != NULL: y += res[m,n] (MBConv skip connection, fp16) ; out[m,n] = y as fp16 (out_f16 != 0) or fp32.
 * K and N must be multiples of 8. */
int avexk_conv1x1_f16(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                      const void* res_f16, float* raw_out, void* out, int out_f16, void* stream);

/* The MBConv project convolution with the squeeze-excitation rescale of its input fused on the A operand
 * (torchvision SqueezeExcitation.forward `scale * input` followed by the project Conv2dNormActivation):
 *   out[m,n] = (sum_k se_scale[m / rows_per_clip, k] * A[m,k] * W[n,k]) * scale[n] + shift[n] (+ res[m,n]), fp16 out.
 * se_scale [M / rows_per_clip, K] fp32; A is not modified. */
int avexk_conv1x1_se_f16(const void* A, const float* se_scale, int rows_per_clip, const void* W, int M, int N, int K,
                         const float* scale, const float* shift, const void* res_f16, void* out, void* stream);

/* Depthwise k x k convolution (k = 3 | 5, stride 1 | 2, padding (k-1)/2) + folded BatchNorm + SiLU.
 * in [B,H,W,C] fp16, w_ckk [C,1,k,k] fp32 (torch layout), out [B,Ho,Wo,C] fp16, Ho = (H + 2p - k) / stride + 1.
 * se_sum [B,C] fp32 (may be NULL) receives sum over output pixels of the activated output (squeeze-excitation).
 * The sums are accumulated in 40.24 fixed point, so they are bit-reproducible from run to run.
 * workspace >= 8*B*C + 4*C*k*k bytes (fixed-point accumulators, repacked weights). */
int avexk_dwconv_nhwc(const void* in_f16, int B, int H, int W, int C, int k, int stride, const float* w_ckk,
                      const float* scale, const float* shift, void* out_f16, float* se_sum, void* workspace, void* stream);

typedef struct avexk_effnet avexk_effnet_t;

typedef struct {
  int kernel, stride; /* depthwise conv */
  int cin, cexp, cout, csq; /* block input, expanded, output and squeeze channels (cexp == cin: no expand conv) */
} avexk_effnet_block_cfg;

typedef struct {
  const float *weight, *bias, *mean, *var; /* BatchNorm2d weight, bias, running_mean, running_var (eps 1e-5) */
} avexk_bn_params;

/* fp32 DEVICE pointers in torchvision's own layouts (state_dict names relative to `features.{s}.{r}.block`). */
typedef struct {
  const float* expand_w;      /* 0.0.weight [cexp,cin,1,1] (NULL when the block has no expand conv)        */
  avexk_bn_params expand_bn;  /* 0.1.*                                                                      */
  const float* dw_w;          /* depthwise conv weight [cexp,1,k,k]                                         */
  avexk_bn_params dw_bn;
  const float *se1_w, *se1_b; /* SqueezeExcitation fc1 [csq,cexp,1,1], bias                                 */
  const float *se2_w, *se2_b; /* fc2 [cexp,csq,1,1], bias                                                   */
  const float* proj_w;        /* project conv [cout,cexp,1,1]  (the `block.3.0` / `block.2.0` hook layer)   */
  avexk_bn_params proj_bn;
} avexk_effnet_block_weights;

typedef struct {
  const float* stem_w;    /* features.0.0.weight [32,3,3,3] */
  avexk_bn_params stem_bn;
  const avexk_effnet_block_weights* blocks; /* HOST array, one entry per MBConv block in forward order */
  const float* head_w;    /* features.8.0.weight [1280,320,1,1] */
  avexk_bn_params head_bn;
  const float *cls_w, *cls_b; /* classifier.1 Linear [num_classes,1280] (NULL: features only) */
  int num_classes;
} avexk_effnet_weights;

int avexk_effnet_create(const avexk_effnet_block_cfg* blocks, int num_blocks, int stem_out, int head_out, avexk_effnet_t** out);
void avexk_effnet_destroy(avexk_effnet_t* h);
/* Folds every BatchNorm into scale / shift, sums the stem weights over the 3 identical input channels
 * (efficientnet.py:138-140), repacks depthwise weights, converts 1x1 weights to bf16.  Synchronises. */
int avexk_effnet_load_weights(avexk_effnet_t* h, const avexk_effnet_weights* w, void* stream);
/* spatial size of the final feature map for an input image of H0 x W0 (128 x frames) */
int avexk_effnet_out_hw(const avexk_effnet_t* h, int H0, int W0, int* Hf, int* Wf);
size_t avexk_effnet_workspace_bytes(const avexk_effnet_t* h, int B, int H0, int W0);

/* mel [B,H0,W0] fp32: the single-channel image (log-mel, frequency-major).  minmax != NULL: the image is the
 * un-normalised output of avexk_melspec_forward(normalize=0) and the per-clip min-max normalisation is applied on load.
 * features_nchw [B,head_out,Hf,Wf] fp32 (return_features_only) and / or logits [B,num_classes]; either may be NULL.
 * hook_out: HOST array of num_blocks+2 DEVICE pointers (NULL entry = not materialised), fp32 NCHW pre-BatchNorm conv
 *   outputs: [0] stem conv (features.0.0), [1+i] project conv of block i, [num_blocks+1] head conv (features.8.0). */
int avexk_effnet_forward(avexk_effnet_t* h, const float* mel, const void* minmax, int B, int H0, int W0,
                         float* features_nchw, float* logits, float* const* hook_out, void* workspace,
                         size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVEXK_H */
