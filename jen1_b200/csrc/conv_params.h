// Parameter block of the fused "tap-GEMM" operator that implements every conv / linear on the hot path.
//
// One launch computes, for batch row b, output phase z, logical row m and output channel n:
//
//   out[b, o(m,z), n] = epi( bias[n] + sum_{seg} sum_{tap j} sum_{c} W_seg[wtap(z,j)][c][n] * A_seg(b, m*in_stride + shift(j), c) )
//                       (+ residual[b, o, n])
//   o(m,z)  = m*out_stride + out_off0 + z*out_off_phase            (rows outside [0, Lout) are skipped)
//   A_seg   = prologue( concat_channels(src0, src1) ) with ZERO outside [0, L) -- i.e. the zero padding is applied
//             AFTER the normalisation / FiLM / activation, exactly as reference blocks.py:137-145 -> :44-51 does.
//
// This covers (reference jen1/model/blocks.py):
//   * Conv1d wrapper k in {1,3}, causal or centred (:34-53)      in_stride=1, shift(j) = j - pad_left
//   * Downsample1d k=2f+1, stride f (:55-66)                     in_stride=f, shift(j) = j - pad_left
//   * ConvTranspose1d k=2f, stride f (:88-95) as f output phases out_stride=f, phase z uses taps {z, z+f} at shifts {0,-1}
//   * nn.Linear on [B, N, C] tokens (k=1)
// with the GroupNorm-apply / FiLM / SiLU (ConvBlock1d :137-145) or LayerNorm-normalise (Attention :427) prologue,
// channel concat of two sources without materialising torch.cat (UpsampleBlock1d.add_skip :732-734, model.py:240),
// a second K-segment for ResnetBlock1d.to_out (:229-231), bias / GELU / residual epilogue, and per-tile
// GroupNorm / per-row LayerNorm partial statistics of the OUTPUT for the next consumer.
#pragma once
#include <stdint.h>

namespace jen1 {

struct ConvSrc {
  const void* ptr;     // T [Bsrc][L][C] channels-last
  const long long* stats;  // [Bsrc][FG][2] fixed-point (sum, sumsq) accumulators of this tensor (common.cuh), or nullptr
  int C;               // channels (0 = source absent)
  int FG;              // fine groups in `stats`
  int bmod;            // batch row used = b % bmod
  float scale;         // raw values are multiplied by this (skip scale 2^-1/2, reference blocks.py:734)
};

struct ConvSeg {
  ConvSrc s[2];
  const void* w;  // T [n_wtaps][Cin][Cout]
  const void* wT; // fp32 [n_wtaps][Cout][Cin]: the same weights, Cin contiguous (TF32 kernel only)
  int Cin;        // s[0].C + s[1].C
  int L;          // rows per batch in the sources
  int ntaps;
  int in_stride;
  int shift0, shift_step;             // shift(j)  = shift0 + j*shift_step
  int wtap0, wtap_phase, wtap_step;   // wtap(z,j) = wtap0 + z*wtap_phase + j*wtap_step
  // ---- window into a longer stored tensor (generic kernel only; the Encodec decoder reads the trimmed middle of an
  //      untrimmed transposed-conv output): logical row i lives at stored row row0 + i of Lstore rows per batch.
  //      Lstore == 0 means Lstore = L, row0 = 0.  GroupNorm statistics always cover the Lstore stored rows.
  int row0, Lstore;
};

enum { PRO_AFFINE = 0, PRO_ROWNORM = 1 };
enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_ELU = 3 };
enum { PAD_ZERO = 0, PAD_REFLECT = 1 };

struct ConvParams {
  ConvSeg seg[2];
  int nseg;
  // ---- prologue of seg[0] (seg[1] is always raw * scale)
  int mode;            // PRO_AFFINE: y = act(a[b][c]*x + s[b][c]);  PRO_ROWNORM: y = (x - mu[row]) * rstd[row]
  int G;               // GroupNorm groups over the concatenated channels (0 = no norm)
  int gn_real_c;       // > 0: channels that really exist (the rest is zero padding): element count of a group is
                       //      gn_real_c / G * L instead of Cin / G * L (only used with G = 1)
  float eps;
  const float* gamma;  // [Cin]
  const float* beta;   // [Cin]
  const float* film;   // FiLM table: scale = film[row*film_stride + c], shift = film[row*film_stride + Cin + c]
  int film_stride;
  const int* cond_row; // [B] table row per batch row (device)
  int act;             // ACT_NONE / ACT_SILU applied after the affine
  // ---- generic kernel only (Encodec decoder, csrc/codec.cu)
  int pad_mode;        // PAD_REFLECT: rows outside [0, L) mirror (of the signal zero-extended to Lext rows when L is
  int Lext;            //              shorter than the padding, as encodec's pad1d does); Lext == 0 means L
  int sum2;            // seg[0] = prologue(s[0]) + prologue(s[1]) (same C; GroupNorm(1) each with its own affine and
  const float* gamma2; //              statistics; a source without statistics is taken as is) instead of a channel concat
  const float* beta2;
  const float* rowpart;  // PRO_ROWNORM: [Bsrc][L][rp_nct][2] per-row partial (sum, sumsq) of seg[0].s[0]
  int rp_nct;
  float ln_eps;
  // ---- output
  int B, Lm, nphase, out_stride, out_off0, out_off_phase, Lout, Cout;
  const float* bias;   // [Cout] or nullptr
  int epi_act;         // ACT_NONE / ACT_GELU
  const void* res;     // TA [Bres][Lout][Cout] added after the activation, or nullptr
  int res_bmod;
  void* out;           // TO [B][Lout][Cout]          (exactly one of out / out_ncl is set)
  float* out_ncl;      // fp32 [B][Cout][Lout]
  long long* stats_out;  // [B][FGo][2] fixed-point accumulators (zeroed by the host before the launch chain) or nullptr
  int FGo;
  int stat_slots;      // > 1 (generic / tf32 kernels): every fine group has this many accumulator slots, [B][FGo][slots][2],
                       //   a CTA adds into slot (row tile % slots) -- spreads the atomics of very long layers; consumers
                       //   that sum ALL entries of a tensor (GroupNorm(1)) simply see FGo * slots fine groups
  float* rowpart_out;  // [B][Lout][gridDim.y][2] or nullptr
};

}  // namespace jen1
