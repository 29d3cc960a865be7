// extern "C" boundary of the engine (include/jen1_b200.h).  Plain pointers and sizes only; no exceptions cross.
#include <new>

#include "codec.h"
#include "engine.h"

using jen1::CodecDecoder;
using jen1::Engine;

namespace {
thread_local char g_create_error[256] = "";
inline Engine* E(void* h) { return reinterpret_cast<Engine*>(h); }
inline CodecDecoder* K(void* h) { return reinterpret_cast<CodecDecoder*>(h); }
}  // namespace

extern "C" {

int jen1_engine_create(const Jen1ModelDesc* desc, int device, int dtype, void** out_handle) {
  if (!desc || !out_handle) return 1;
  *out_handle = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_create_error, sizeof g_create_error, "no CUDA device %d (the engine has no CPU fallback)", device);
    return 2;
  }
  if (dtype != JEN1_DTYPE_F32 && dtype != JEN1_DTYPE_BF16) return 3;
  Engine* e = new (std::nothrow) Engine(*desc, device, dtype);
  if (!e) return 4;
  *out_handle = e;
  return 0;
}

void jen1_engine_destroy(void* h) { delete E(h); }

const char* jen1_last_error(void* h) { return h ? E(h)->last_error() : g_create_error; }

int jen1_engine_load_tensor(void* h, const char* name, const float* host_data, const int64_t* shape, int ndim) {
  if (!h || !name || !host_data || !shape) return 1;
  try {
    return E(h)->load_tensor(name, host_data, shape, ndim);
  } catch (...) {
    return 99;
  }
}

int jen1_engine_finalize(void* h) {
  if (!h) return 1;
  try {
    return E(h)->finalize();
  } catch (...) {
    return 99;
  }
}

size_t jen1_engine_workspace_bytes(void* h, int B, int T) {
  if (!h) return 0;
  try {
    return E(h)->workspace_bytes(B, T);
  } catch (...) {
    return 0;
  }
}

int jen1_engine_reserve(void* h, int B, int T) {
  if (!h) return 1;
  try {
    return E(h)->reserve(B, T);
  } catch (...) {
    return 99;
  }
}

int jen1_engine_set_context(void* h, const float* emb, const float* mask, int B, int S, jen1_stream_t stream) {
  if (!h || !emb) return 1;
  try {
    return E(h)->set_context(emb, mask, B, S, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int jen1_engine_set_timesteps(void* h, const int64_t* t_host, int n, jen1_stream_t stream) {
  if (!h || !t_host) return 1;
  try {
    return E(h)->set_timesteps(t_host, n, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int jen1_unet_forward(void* h, const float* x, const float* concat_cond, const int32_t* cond_rows,
                      const uint8_t* drop_mask, int B, int T, int causal, float embedding_scale, int scale_cfg,
                      float scale_phi, float* out, jen1_stream_t stream) {
  if (!h || !x || !concat_cond || !out) return 1;
  try {
    return E(h)->forward(x, concat_cond, cond_rows, drop_mask, B, T, causal, embedding_scale, scale_cfg, scale_phi,
                         out, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int jen1_sample_begin(void* h, const float* coef_host, int S, const float* concat_cond, int B, int T, int causal,
                      float embedding_scale, int scale_cfg, float scale_phi, int objective, int use_graph,
                      jen1_stream_t stream) {
  if (!h || !coef_host || !concat_cond) return 1;
  try {
    return E(h)->sample_begin(coef_host, S, concat_cond, B, T, causal, embedding_scale, scale_cfg, scale_phi,
                              objective, use_graph, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int jen1_sample_step(void* h, int step, float* x, const float* noise, const uint8_t* drop_mask,
                     jen1_stream_t stream) {
  if (!h || !x) return 1;
  try {
    return E(h)->sample_step(step, x, noise, drop_mask, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int64_t jen1_engine_launch_count(void* h) { return h ? E(h)->launch_count() : 0; }
int64_t jen1_engine_weight_bytes(void* h) { return h ? E(h)->weight_bytes() : 0; }
int64_t jen1_engine_umma_launch_count(void* h) { return h ? E(h)->umma_launch_count() : 0; }
int64_t jen1_engine_umma_attn_launch_count(void* h) { return h ? E(h)->umma_attn_launch_count() : 0; }
int64_t jen1_engine_fused_transformer_launch_count(void* h) { return h ? E(h)->fused_tr_launch_count() : 0; }

int jen1_attention_forward(void* h, const void* qkv_bf16, void* out_bf16, int B, int N, int H, int d, int causal, int impl,
                           jen1_stream_t stream) {
  if (!h || !qkv_bf16 || !out_bf16) return 1;
  try {
    return E(h)->attention(qkv_bf16, out_bf16, B, N, H, d, causal, impl, (cudaStream_t)stream);
  } catch (...) {
    return 99;
  }
}

int jen1_engine_debug_tensor(void* h, const char* name, float* host_out, int64_t capacity, int64_t* shape3) {
  if (!h || !name || !host_out || !shape3) return 1;
  try {
    return E(h)->debug_tensor(name, host_out, capacity, shape3);
  } catch (...) {
    return 99;
  }
}


// ---------------------------------------------------------------------------------------------- Encodec decoder
int jen1_codec_create(const Jen1CodecDesc* desc, int device, int precision, void** out_handle) {
  if (!desc || !out_handle) return 1;
  *out_handle = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_create_error, sizeof g_create_error, "no CUDA device %d (the codec engine has no CPU fallback)", device);
    return 2;
  }
  if (precision != JEN1_CODEC_FP32 && precision != JEN1_CODEC_TF32) return 3;
  CodecDecoder* k = new (std::nothrow) CodecDecoder(*desc, device, precision == JEN1_CODEC_FP32);
  if (!k) return 4;
  *out_handle = k;
  return 0;
}
int jen1_codec_create_encoder(const Jen1CodecDesc* desc, int device, int precision, int n_q, int codebook_size, void** out_handle) {
  if (!desc || !out_handle) return 1;
  *out_handle = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    snprintf(g_create_error, sizeof g_create_error, "no CUDA device %d (the codec engine has no CPU fallback)", device);
    return 2;
  }
  if (precision != JEN1_CODEC_FP32 && precision != JEN1_CODEC_TF32) return 3;
  CodecDecoder* k = new (std::nothrow) CodecDecoder(*desc, device, precision == JEN1_CODEC_FP32, true, n_q, codebook_size);
  if (!k) return 4;
  *out_handle = k;
  return 0;
}
int jen1_codec_encode(void* h, const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L,
                      jen1_stream_t stream) {
  if (!h) return 1;
  try {
    return K(h)->encode(audio, latent, codes, quantized, N, L, reinterpret_cast<cudaStream_t>(stream));
  } catch (...) {
    return 99;
  }
}
int jen1_codec_quantize(void* h, const float* latent, int32_t* codes, float* quantized, int N, int T, jen1_stream_t stream) {
  if (!h) return 1;
  try {
    return K(h)->quantize(latent, codes, quantized, N, T, reinterpret_cast<cudaStream_t>(stream));
  } catch (...) {
    return 99;
  }
}
void jen1_codec_destroy(void* h) { delete K(h); }
const char* jen1_codec_last_error(void* h) { return h ? K(h)->last_error() : g_create_error; }
int jen1_codec_load_tensor(void* h, const char* name, const float* host_data, const int64_t* shape, int ndim) {
  if (!h || !name || !host_data || !shape) return 1;
  try {
    return K(h)->load_tensor(name, host_data, shape, ndim);
  } catch (...) {
    return 99;
  }
}
int jen1_codec_finalize(void* h) {
  if (!h) return 1;
  try {
    return K(h)->finalize();
  } catch (...) {
    return 99;
  }
}
size_t jen1_codec_workspace_bytes(void* h, int B, int T) { return h ? K(h)->workspace_bytes(B, T) : 0; }
int jen1_codec_reserve(void* h, int B, int T) {
  if (!h) return 1;
  try {
    return K(h)->reserve(B, T);
  } catch (...) {
    return 99;
  }
}
int jen1_codec_decode(void* h, const float* latent, float* audio, int B, int T, jen1_stream_t stream) {
  if (!h) return 1;
  try {
    return K(h)->decode(latent, audio, B, T, reinterpret_cast<cudaStream_t>(stream));
  } catch (...) {
    return 99;
  }
}
int64_t jen1_codec_launch_count(void* h) { return h ? K(h)->launch_count() : 0; }
int64_t jen1_codec_weight_bytes(void* h) { return h ? K(h)->weight_bytes() : 0; }
int64_t jen1_codec_tf32_launch_count(void* h) { return h ? K(h)->tf32_launch_count() : 0; }
int64_t jen1_codec_lstm_tc_launch_count(void* h) { return h ? K(h)->lstm_tc_launch_count() : 0; }
int jen1_codec_hop(void* h) { return h ? K(h)->hop() : 0; }
int jen1_codec_lstm_cluster(void* h) { return h ? K(h)->lstm_cluster() : 0; }
}  // extern "C"
