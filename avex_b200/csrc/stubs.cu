// TEMPORARY: entry points not implemented yet return an error (never a fallback).
#include "common.cuh"
#define NOTIMPL(name) avexk::set_error(name ": not implemented in this build"); return AVEXK_EINVAL
extern "C" int avexk_gemm_bf16(const void*, long long, const void*, long long, int, int, int, const float*, int, float*, const float*, float, void*, long long, int, void*) { NOTIMPL("avexk_gemm_bf16"); }
extern "C" int avexk_layernorm(const float*, int, int, const float*, const float*, float, float*, void*, void*) { NOTIMPL("avexk_layernorm"); }
extern "C" int avexk_attention_gated(const void*, int, int, int, const float*, const float*, const float*, const float*, const uint8_t*, void*, void*) { NOTIMPL("avexk_attention_gated"); }
extern "C" int avexk_beats_create(const avexk_beats_dims*, avexk_beats_t**) { NOTIMPL("avexk_beats_create"); }
extern "C" void avexk_beats_destroy(avexk_beats_t*) {}
extern "C" int avexk_beats_load_weights(avexk_beats_t*, const avexk_beats_weights*, void*) { NOTIMPL("avexk_beats_load_weights"); }
extern "C" int avexk_beats_num_tokens(int T) { return 8 * (avexk_fbank_num_frames(T) / 16); }
extern "C" size_t avexk_beats_workspace_bytes(const avexk_beats_t*, int, int) { return 0; }
extern "C" int avexk_beats_forward(avexk_beats_t*, const float*, int, int, long long, const avexk_fbank_t*, const uint8_t*, const float*, float*, float* const*, float*, void*, size_t, void*) { NOTIMPL("avexk_beats_forward"); }
