// kernels.cuh -- internal launch functions shared between translation units of libavexk.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace avexk {

// gemm_tc.cu
struct GemmGate {  // optional by-product of the fused QKV projection: per-(row, head) gate of the relative-position bias
  const float *w, *b, *grep_a;  // gate_w [2,64], gate_b [2], grep_a [heads]
  float* out;                   // [M, heads]
  int heads;
};
int gemm_bf16_launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const float* bias, int gelu,
                     float* raw_out, const float* residual, float res_scale, void* out, long long ldo, int out_bf16,
                     cudaStream_t st, const GemmGate* gate = nullptr);
size_t gemm_ln_scratch_bytes(int M);
int gemm_bf16_ln_launch(const void* A, long long lda, const void* W, long long ldw, int M, int K, const float* bias, float* raw_out,
                        const float* residual, float res_scale, const float* gamma, const float* beta, float eps,
                        float* ln_out_f32, __nv_bfloat16* ln_out_bf16, void* scratch, size_t scratch_bytes, int zero_counters,
                        cudaStream_t st, long long* pool_raw = nullptr, long long* pool_y = nullptr, int pool_rows = 0);
// pointwise.cu: the memory-bound form of the 1x1 convolution (fp16 out, optional squeeze-excitation scale on the A operand)
bool pointwise_supported(int N, int K, const void* out, int out_16bit);
int pointwise_launch(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                     const __nv_bfloat16* res, const float* se_scale, int hw, float* raw_out, void* out, cudaStream_t st);
int conv1x1_launch(const void* A, const void* W, int M, int N, int K, const float* scale, const float* shift, int silu,
                   const __nv_bfloat16* res, float* raw_out, void* out, int out_bf16, cudaStream_t st);
// attention_tc.cu
int launch_gate_from_qkv(const void* qkv, long long M, int H, const float* gate_w, const float* gate_b, const float* grep_a,
                         float* gate, cudaStream_t st);
int attention_launch(const void* qkv, int B, int N, int H, const float* gate, const float* bias_vec, const uint8_t* key_pad, void* out,
                     cudaStream_t st);
// elementwise.cu
int launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, float* out_f32,
                     void* out_bf16, cudaStream_t st, int split3 = 0);
int launch_f32_to_bf16_split3(const float* src, __nv_bfloat16* dst, int N, int K, cudaStream_t st);
int launch_group_pad(float* x0, const uint8_t* key_pad, long long M, int G, int cg, __nv_bfloat16* xg, cudaStream_t st);
// 40.24 fixed-point column sums (fused pooling of the GEMM+LN epilogue) -> fp32 means
int launch_pool_finalize(const long long* acc, int B, int C, float scale, float* out, cudaStream_t st);
int launch_mean_pool(const float* x, const uint8_t* key_pad, int any_pad, int B, int N, int C, float* out, cudaStream_t st);
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st);
int launch_f32_to_f16(const float* src, void* dst, long long n, cudaStream_t st);  // saturating
int launch_posconv_pack(const float* v, const float* g, int C, int cg, int K, float* nrm_ws, __nv_bfloat16* W, cudaStream_t st);
int launch_gate_pack(const float* w, const float* b, float* gw, float* gb, cudaStream_t st);
// fp32_path.cu (fp32 mode)
int launch_split3_rows(const float* src, long long M, int K, __nv_bfloat16* dst, cudaStream_t st);
int launch_attention_fp32(const float* qkv, int B, int N, int H, const float* gate_w, const float* gate_b, const float* grep_a,
                          const float* bias_vec, const uint8_t* key_pad, float* out, cudaStream_t st);
int launch_posconv_pack_f32(const float* v, const float* g, const float* nrm, int G, int cg, int K, float* W, cudaStream_t st);
int launch_posconv_fp32(const float* x0, const float* Wf, const float* bias, float* out, int B, int N, int G, int cg, int taps,
                        cudaStream_t st);
// posconv.cu
int launch_posconv(const __nv_bfloat16* xg, const __nv_bfloat16* Wpc, const float* bias, const float* x0, float* out, int B,
                   int N, int G, int cg, int taps, cudaStream_t st);

}  // namespace avexk
