/*
 * jen1_b200 -- C ABI of the B200-native JEN-1 denoiser engine.
 *
 * The reference (0417keito/JEN-1-pytorch) has no FFI of its own: it is pure Python.  The seams it does have
 *  are what this ABI replaces, one entry point per seam:
 *
 *   reference seam                                                   | replaced by
 *   -----------------------------------------------------------------+---------------------------------------
 *   UNetCFG1d(**ModelConfig)            utils/script_util.py:271-284 | jen1_engine_create(Jen1ModelDesc)
 *   model.load_state_dict(...)          utils/script_util.py:93-122  | jen1_engine_load_tensor + _finalize
 *   conditioning['cross_attn_cond'/'cross_attn_masks'] handed to the  | jen1_engine_set_context
 *     model every step                  jen1/diffusion/gdm/gdm.py:118-119 |   (step-invariant K/V hoisted)
 *   time -> to_time/to_mapping/MappingToScaleShift/to_time_embedding  | jen1_engine_set_timesteps
 *     every step      jen1/model/model.py:204-223,315-316; blocks.py:161-165 | (functions of t only -> table)
 *   model(x, t, embedding=..., channels_list=[...], causal=...)       | jen1_unet_forward
 *                                       jen1/model/model.py:299-376  |
 *   GaussianDiffusion.ddim_sample loop body                           | jen1_sample_begin / jen1_sample_step
 *                                       jen1/diffusion/gdm/gdm.py:202-222 |
 *   self.audio_encoder.decoder(sample_embs)        generation.py:130  | jen1_codec_create / _load_tensor / _decode
 *   get_emb: audio_encoder.encode + quantizer.decode  generation.py:145-150 | jen1_codec_create_encoder / _encode
 *
 * Conventions: every function returns 0 on success or a non-zero code; the message is available from
 * jen1_last_error(handle).  Nothing throws or aborts across the ABI.  A handle is bound to one CUDA device and
 * is NOT thread-safe (one handle per GPU).  All pointers documented "device" are device pointers owned by the
 * caller (PyTorch tensors on the host side); calls are asynchronous and ordered on the given stream.  The
 * engine owns its packed weights, caches and workspace; after jen1_engine_reserve no call allocates.
 * There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef JEN1_B200_H_
#define JEN1_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JEN1_MAX_LEVELS 16

/* Mirrors reference utils/config.py:49-74 (ModelConfig). */
typedef struct Jen1ModelDesc {
  int32_t in_channels;
  int32_t out_channels;
  int32_t channels;
  int32_t num_layers;                        /* len(multipliers) - 1 */
  int32_t multipliers[JEN1_MAX_LEVELS + 1];
  int32_t factors[JEN1_MAX_LEVELS];
  int32_t num_blocks[JEN1_MAX_LEVELS];
  int32_t attentions[JEN1_MAX_LEVELS + 1];   /* attentions[num_layers] drives the bottleneck (attentions[-1]) */
  int32_t resnet_groups;
  int32_t context_channels;                  /* input-concat channels at level 0 (129) */
  int32_t context_features_multiplier;
  int32_t context_embedding_features;        /* 1024 */
  int32_t context_embedding_max_length;      /* 128 (the time token is appended by the engine) */
  int32_t attention_heads;
  int32_t attention_multiplier;
  int32_t use_skip_scale;
} Jen1ModelDesc;

enum { JEN1_DTYPE_F32 = 0, JEN1_DTYPE_BF16 = 1 };
enum { JEN1_OBJECTIVE_NOISE = 0, JEN1_OBJECTIVE_X0 = 1, JEN1_OBJECTIVE_V = 2 };

typedef void* jen1_stream_t; /* cudaStream_t */

/* Create an engine on CUDA device `device`. `dtype` selects the storage precision of weights/activations:
 * JEN1_DTYPE_BF16 (fast path, tcgen05 kernels) or JEN1_DTYPE_F32 (strict mode, fp32 FMA kernels). */
int jen1_engine_create(const Jen1ModelDesc* desc, int device, int dtype, void** out_handle);
void jen1_engine_destroy(void* handle);
const char* jen1_last_error(void* handle);

/* Hand the engine one tensor of the reference state_dict (fp32, host memory, reference key name and shape). */
int jen1_engine_load_tensor(void* handle, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* Check completeness, fold/pack/upload the weights, build the weight-only caches. */
int jen1_engine_finalize(void* handle);

/* Bytes of workspace a (B samples, T frames) forward needs; reserve allocates it (and everything else). */
size_t jen1_engine_workspace_bytes(void* handle, int B, int T);
int jen1_engine_reserve(void* handle, int B, int T);

/* Cross-attention context: emb fp32 device [B][S][E], mask fp32 device [B][S] (1 keep / 0 pad) or NULL. */
int jen1_engine_set_context(void* handle, const float* emb, const float* mask, int B, int S, jen1_stream_t stream);
/* Conditioning rows for n timesteps (host int64): row i holds everything derived from t[i]. */
int jen1_engine_set_timesteps(void* handle, const int64_t* t_host, int n, jen1_stream_t stream);

/* One UNetCFG1d evaluation.  x: fp32 device [B][in_channels][T]; concat_cond: fp32 device [B][context_channels][T];
 * cond_rows: host int32 [B] conditioning-table row per sample; drop_mask: DEVICE uint8 [B] cond-dropout flags
 * (the bernoulli draw of reference utils/module.py:36-42, made by the caller so the RNG stream is the host
 * framework's) or NULL; embedding_scale == 1 disables classifier-free guidance (single pass).
 * out: fp32 device [B][out][T]. */
int jen1_unet_forward(void* handle, const float* x, const float* concat_cond, const int32_t* cond_rows,
                      const uint8_t* drop_mask, int B, int T, int causal, float embedding_scale, int scale_cfg,
                      float scale_phi, float* out, jen1_stream_t stream);

/* DDIM sampling (reference gdm.py:181-225).  coef: host fp32 [S][8] per-step scalars
 * {sqrt_recip_ac[t], sqrt_recipm1_ac[t], sqrt_ac[t], sqrt_1m_ac[t], sqrt(alpha_next), c, sigma, is_last};
 * conditioning row i of jen1_engine_set_timesteps must correspond to step i. */
int jen1_sample_begin(void* handle, const float* coef_host, int S, const float* concat_cond, int B, int T,
                      int causal, float embedding_scale, int scale_cfg, float scale_phi, int objective,
                      int use_graph, jen1_stream_t stream);
/* One step: x (fp32 device [B][C][T]) is updated in place; noise fp32 device [B][C][T] (ignored on the last
 * step, may be NULL there); drop_mask DEVICE uint8 [B] or NULL. */
int jen1_sample_step(void* handle, int step, float* x, const float* noise, const uint8_t* drop_mask,
                     jen1_stream_t stream);

/* The attention core alone -- reference jen1/model/blocks.py:355-380 (AttentionBase.forward: softmax(q k^T d^-1/2) v per head,
 * causal mask :304-319) -- on a packed DEVICE bf16 tensor qkv[B][N][3*H*d] (q | k | v, what Attention.to_q / to_kv produce);
 * out: DEVICE bf16 [B][N][H*d].  impl 0: tcgen05 kernels (single-tile up to 256 keys, key-tiled online softmax beyond),
 * 1: fp32-FMA core (the strict-mode kernel), 2: force the key-tiled kernel.  bf16 engines only. */
int jen1_attention_forward(void* handle, const void* qkv_bf16, void* out_bf16, int B, int N, int H, int d, int causal,
                           int impl, jen1_stream_t stream);

/* ---- Encodec (SEANet) decoder: latent -> audio, the step after the sampling loop.
 * Replaces `self.audio_encoder.decoder(sample_embs)` (reference generation.py:130; audio_encoder =
 * EncodecModel.encodec_model_48khz(), generation.py:34, pip encodec==0.1.1).  The description mirrors the SEANetDecoder
 * constructor arguments of that model (encodec/model.py encodec_model_48khz, encodec/modules/seanet.py). */
#define JEN1_CODEC_MAX_RATIOS 8
typedef struct Jen1CodecDesc {
  int32_t channels;              /* audio channels (2) */
  int32_t dimension;             /* latent channels (128) */
  int32_t n_filters;             /* 32 */
  int32_t n_ratios;
  int32_t ratios[JEN1_CODEC_MAX_RATIOS]; /* decoder order (8, 5, 4, 2) */
  int32_t kernel_size;           /* 7 */
  int32_t last_kernel_size;      /* 7 */
  int32_t residual_kernel_size;  /* 3 */
  int32_t compress;              /* 2 */
  int32_t lstm_layers;           /* 2 */
  float eps;                     /* GroupNorm epsilon (1e-5) */
} Jen1CodecDesc;
enum { JEN1_CODEC_FP32 = 0, JEN1_CODEC_TF32 = 1 };
/* precision: JEN1_CODEC_TF32 = convs on the TF32 tensor-core kernel (fp32 storage and accumulation), JEN1_CODEC_FP32 =
 * strict mode (fp32 FMA everywhere except the recurrent LSTM weights, which are fp16 in shared memory in both modes). */
int jen1_codec_create(const Jen1CodecDesc* desc, int device, int precision, void** out_handle);
void jen1_codec_destroy(void* handle);
const char* jen1_codec_last_error(void* handle);
/* One tensor of the decoder's state_dict (fp32 host memory; names as in the pip package: model.N.conv.conv.weight ...). */
int jen1_codec_load_tensor(void* handle, const char* name, const float* host_data, const int64_t* shape, int ndim);
int jen1_codec_finalize(void* handle);
size_t jen1_codec_workspace_bytes(void* handle, int B, int T);
int jen1_codec_reserve(void* handle, int B, int T);
/* latent: DEVICE fp32 [B][dimension][T]; audio: DEVICE fp32 [B][channels][T * hop]. */
int jen1_codec_decode(void* handle, const float* latent, float* audio, int B, int T, jen1_stream_t stream);
/* ---- Encodec encoder + residual vector quantizer: audio -> latent, reference generation.py:145-150 (`get_emb`:
 * audio_encoder.encode(...) segment by segment, then quantizer.decode(codes)).  Same description struct; tensors are named
 * as in the pip package (encoder.model.N.conv.conv.weight ..., quantizer.vq.layers.i._codebook.embed).  The handle is used
 * with jen1_codec_load_tensor / _finalize / _destroy / _last_error like a decoder handle. */
int jen1_codec_create_encoder(const Jen1CodecDesc* desc, int device, int precision, int n_q, int codebook_size, void** out_handle);
/* audio: DEVICE fp32 [N][channels][L] (one row per segment, already loudness-normalised by the caller);
 * latent: DEVICE fp32 [N][dimension][T], T = ceil(L / hop) -- the encoder output before quantisation;
 * codes: DEVICE int32 [n_q][N][T] or NULL; quantized: DEVICE fp32 [N][dimension][T] or NULL (sum of the chosen entries). */
int jen1_codec_encode(void* handle, const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L,
                      jen1_stream_t stream);
/* The residual vector quantizer alone (quantizer.encode + quantizer.decode) on a DEVICE latent [N][dimension][T]. */
int jen1_codec_quantize(void* handle, const float* latent, int32_t* codes, float* quantized, int N, int T, jen1_stream_t stream);
int64_t jen1_codec_launch_count(void* handle);
int64_t jen1_codec_tf32_launch_count(void* handle); /* how many of them were the TF32 tensor-core tap-GEMM */
int64_t jen1_codec_lstm_tc_launch_count(void* handle); /* ... and the tensor-core LSTM cluster kernel */
int64_t jen1_codec_weight_bytes(void* handle);
int jen1_codec_hop(void* handle);          /* samples per latent frame (320) */
int jen1_codec_lstm_cluster(void* handle); /* CTAs per sequence of the LSTM cluster kernel */

/* Introspection for tests / benchmarks. */
int64_t jen1_engine_launch_count(void* handle);        /* kernels launched (or replayed) so far */
int64_t jen1_engine_weight_bytes(void* handle);        /* device bytes of packed weights streamed per step */
int64_t jen1_engine_umma_launch_count(void* handle);   /* how many of those launches were the tcgen05 conv kernel */
int64_t jen1_engine_umma_attn_launch_count(void* handle); /* ... and the tcgen05 attention kernel */
int64_t jen1_engine_fused_transformer_launch_count(void* handle); /* ... and the fused Transformer1d kernel (tcgen05 linears + attention) */
int jen1_engine_debug_tensor(void* handle, const char* name, float* host_out, int64_t capacity, int64_t* shape3);

#ifdef __cplusplus
}
#endif
#endif /* JEN1_B200_H_ */
