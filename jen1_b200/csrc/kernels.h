// Launch interfaces of the hand-written kernels (all asynchronous on the given stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cuda_bf16.h>

#include "conv_params.h"

namespace jen1 {

typedef __nv_bfloat16 bf16;

// ---- fused tap-GEMM (conv / linear) ------------------------------------------------------------------
// TA: activation (and residual) storage type, TW: weight storage type, TO: output storage type.
template <typename TA, typename TW, typename TO>
cudaError_t launch_conv_generic(const ConvParams& p, cudaStream_t stream);
int conv_generic_row_tile();
// TF32 tensor-core variant for the Encodec decoder stack (conv_tf32.cu): fp32 storage, PRO_AFFINE prologue only
bool conv_tf32_supported(const ConvParams& p);
cudaError_t launch_conv_tf32(const ConvParams& p, cudaStream_t stream);
int conv_generic_col_tile();

// ---- tcgen05 / TMEM implementation for bf16 storage (conv_umma.cu): swap-AB implicit GEMM, weights streamed as
//      pre-packed 16 KB blobs by bulk async copies, split-K reduced over a thread-block cluster's distributed
//      shared memory, PDL-aware.
struct UmmaPlan {
  int ok;                 // 0: shape not supported -> generic kernel
  int NT, n_tiles;        // accumulator columns (padded flat positions) per CTA, number of N tiles
  int m_tiles;            // Cout / 128
  int splitk;             // == cluster size (1,1,splitk)
  int Lq, amin, halo;     // padded positions per batch row; tap row-offset range
  int R, PS, panel_bytes; // panel rows, rows per sub-panel, bytes of one panel (all sub-panels)
  int steps0, steps1;     // 64-channel K blocks of segment 0 / 1
  int stages, tmem_cols;
  int ring_bytes;         // weight ring (also the epilogue scratch / partial tile)
  int ch_cap;             // seg-0 input channels one CTA may own (coefficient-table stride)
  int off_rowmeta1, off_colmeta, off_rowstat, off_coef;  // shared-memory table offsets
  size_t smem;
};
UmmaPlan conv_umma_plan(const ConvParams& p, bool want_stats, int num_sms);
size_t conv_umma_packed_elems(int Cin, int Cout, int ntaps);
void conv_umma_pack(const float* w_tap_cin_cout, int Cin, int Cout, int nphase, int taps_per_phase, int wtap0,
                    int wtap_phase, int wtap_step, uint16_t* out_bf16);
cudaError_t conv_umma_init();
int conv_umma_max_cluster();
cudaError_t launch_conv_umma(const ConvParams& p, const UmmaPlan& plan, const void* w0_packed, const void* w1_packed,
                             bool out_f32, bool pdl, cudaStream_t stream, long long* timeline = nullptr);

// ---- boundary: [Bx][C][L] fp32 (reference layout) -> channels-last T [Bx][L][Cp] (channels C..Cp-1 zero-filled so
//      rows stay 16-byte aligned for the tcgen05 path) + GroupNorm statistics (FG = 1, fixed-point accumulators [Bx][1][2])
template <typename T>
cudaError_t launch_pack_ncl(const float* x, T* out, long long* stats, int Bx, int C, int Cp, int L, cudaStream_t stream);

// ---- per-row (sum, sumsq) of a [R][C] matrix -> rowpart [R][1][2]
template <typename T>
cudaError_t launch_rowstats(const T* x, float* rowpart, int R, int C, cudaStream_t stream);

// ---- learned Fourier time features (reference utils/module.py:66-72): out fp32 [n][2*half+1]
cudaError_t launch_time_features(const int64_t* t, const float* weights, float* out, int n, int half,
                                 cudaStream_t stream);

// ---- attention core: softmax(q k^T * scale) v with zero-logit key masking (reference blocks.py:355-380, 431-434)
struct AttnParams {
  const void* q;   // T, element (r, i, h, e) at q[((r*N + i) * q_ld) + q_off + h*d + e]
  int q_ld, q_off;
  int B2, Bc, N, M, H, d, C;
  float scale;
  int causal;
  int cross;
  // self-attention keys/values: rows (r*N + j) of `kv`, K at k_off, V at v_off
  const void* kv;
  int kv_ld, k_off, v_off;
  // cross-attention K/V cache (row stride kvc_ld, this layer's K at kvc_off, V at kvc_off + C)
  const void* kv_cond;   // [Bc][M-1] rows
  const void* kv_fixed;  // [M] rows
  const void* kv_time;   // [n_rows] rows (the time-token key/value per conditioning row)
  int kvc_ld, kvc_off;
  const uint8_t* drop;   // [Bc] cond-dropout flags or nullptr
  const float* mask;     // [Bc][M-1] or nullptr
  const int* cond_row;   // [B2]
  void* out;             // T [B2][N][C]
};
template <typename T>
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream);
cudaError_t attention_init();   // per-device function attributes (call after cudaSetDevice, once per engine)
cudaError_t attn_umma_init();
// tcgen05 / TMEM implementation for bf16 storage (attn_umma.cu): head dim in {16, 32, 64, 128}, up to 256 keys
bool attn_umma_supported(const AttnParams& p);
// key-tiled online-softmax tcgen05 attention for any self-attention length (attn_flash.cu): two query tiles per CTA,
// TMA-staged K / V rings, output accumulator resident in TMEM
bool attn_flash_supported(const AttnParams& p);
cudaError_t attn_flash_init();
cudaError_t launch_attention_flash(const AttnParams& p, bool pdl, cudaStream_t stream);
cudaError_t launch_attention_umma(const AttnParams& p, bool pdl, cudaStream_t stream);

// ---- fused Transformer1d (tr_umma.cu): one launch per transformer, a thread-block cluster per batch row walks the op list
//      (k=1 linears with GroupNorm / LayerNorm / raw prologues and bias / GELU / residual epilogues, tcgen05 attention)
enum { TR_GEMM = 0, TR_ATTN = 1 };
enum { TRP_RAW = 0, TRP_LN = 1, TRP_GN = 2 };
constexpr int TR_MAX_OPS = 20;
struct TrOp {
  int type;
  // TR_GEMM: dst[row][token][Cout] = epi(W * pro(src[row % src_bmod][token][K]) + bias) (+ res[row][token][Cout])
  const bf16* src;
  int src_ld, src_bmod, K, pro;
  const bf16* w;      // tcgen05 blob stream of the [Cout][K] weight (conv_umma_pack, one tap)
  const float* bias;  // [Cout] or nullptr
  int Cout, gelu;
  const bf16* res;
  int res_ld;
  bf16* dst;
  int dst_ld;
  int stats;          // add GroupNorm fine-group sums of dst to TrParams::stats_out
  // TR_ATTN: ao[row][token][C] = softmax(q k^T * scale) v per head
  int cross, M;       // keys: N (self) or context length + 1 (cross)
  const bf16* q;
  int q_ld;
  const bf16* kv;     // self: rows of the qkv buffer, K at k_off, V at v_off
  int kv_ld, k_off, v_off;
  int kvc_off;        // cross: this layer's column offset in the K/V caches (V at + C)
  bf16* ao;
};
struct TrParams {
  int B2, N, C, H, d, Bc, causal, n_ops, CS;
  int NT, panel_bytes, smem_bytes, work_bytes, kvx_bytes, prestage, cross_op, stages;  // filled by launch_tr_umma
  float scale, gn_eps;
  const long long* gn_stats;        // [Bx][32][2] fixed-point statistics of the input x (GroupNorm(32))
  const float *gn_gamma, *gn_beta;
  const bf16 *kv_cond, *kv_fixed, *kv_time;
  int kvc_ld;
  const uint8_t* drop;
  const float* mask;
  const int* cond_row;
  long long* stats_out;             // [B2][FGo][2]
  int FGo;
  long long* timeline;              // optional [TR_MAX_OPS][8] phase clocks of CTA (0, 0) (JEN1_TIMELINE debugging)
  TrOp ops[TR_MAX_OPS];
};
bool tr_umma_supported(int N, int C, int H, int M, int n_blocks);
size_t tr_umma_smem_bytes(int N, int C, int H, int M);
cudaError_t tr_umma_init();
cudaError_t launch_tr_umma(const TrParams& p, bool pdl, cudaStream_t stream);

// ---- classifier-free-guidance combine + std rescale (reference model.py:362-369) fused with the x0/eps
//      conversion, clamp and DDIM update (reference gdm.py:128-141, 212-222)
struct SamplerParams {
  const float* y;  // fp32 [B2][L][C] UNet output, rows [0,B) conditional, [B,2B) unconditional when cfg
  int B, C, L;
  int cfg;        // 1: two halves are combined; 0: y is the prediction
  float emb_scale;
  int scale_cfg;
  float phi, one_minus_phi;
  int mode;       // 0: write the combined prediction to pred_out; 1: DDIM step
  float* pred_out;      // fp32 [B][C][L]
  const float* x;       // fp32 [B][C][L] current state
  const float* noise;   // fp32 [B][C][L] (unused on the last step)
  float* x_out;         // next state (may alias x)
  const float* coef;    // [S][8]: sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac, sqrt_1m_ac, sqrt_alpha_next, c, sigma, last
  const int* step;      // device scalar: row of `coef`
  int objective;        // 0 noise, 1 x0, 2 v
};
cudaError_t launch_sampler(const SamplerParams& p, cudaStream_t stream);

}  // namespace jen1
