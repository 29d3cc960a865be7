// TF32 tensor-core variant of the fused tap-GEMM for the Encodec decoder stack (csrc/codec.cu).
//
// Same operator and parameter block as conv_generic.cu (conv_params.h), restricted to what the decoder needs: fp32
// channels-last storage, one K segment, PRO_AFFINE prologue (GroupNorm(1) apply, optional sum of two normalised
// sources, ELU, zero or reflect padding, windowed source), bias epilogue, fixed-point GroupNorm statistics of the output.
// The decoder's layers are long and thin (up to 1.45 M rows of 16-256 channels): too narrow for the 128-lane tcgen05
// tile of conv_umma.cu without repacking, and FMA-bound on the fp32 kernel.  Here the CTA tile is 64 rows x 64 output
// channels, the operand tiles are rounded to TF32 when they are staged in shared memory, and 8 warps (4 x 2) run
// mma.sync m16n8k8 with fp32 accumulation.  Rounding both operands to TF32 (10-bit mantissa) costs about 1e-3 relative
// error per layer; the fp32-FMA kernel stays available as the decoder's strict mode.
#include <cstdlib>

#include "common.cuh"
#include "conv_params.h"

namespace jen1 {

namespace {
constexpr int TM = 64, TK = 16, NT = 256;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
}  // namespace

#ifndef JEN1_TF32_MINB
#define JEN1_TF32_MINB 4
#endif
// NT8: n8 MMA tiles per warp (4: 64-column CTA tile, 2: 32-column tile for the layers with <= 32 output channels).
// The summation index of an MMA is free: k-slot 4j + t of a 16-channel step is mapped to channel 4t + j, so the four
// operands a thread needs for both k8 halves are four CONSECUTIVE channels -- one 16-byte shared-memory load per
// fragment row, the same 16 bytes a staging thread stores.  Tiles are [row][16 channels] (A) and [column][16 channels]
// (B, from weights packed [tap][Cout][Cin]) with a 16-float row stride: conflict-free for 16-byte accesses.
template <int NT8>
__global__ void __launch_bounds__(NT, JEN1_TF32_MINB) conv_tf32_kernel(const ConvParams p) {
  constexpr int TNW = NT8 * 16;  // CTA tile columns (2 warps across)
  extern __shared__ float dsm[];  // coefA[Cin] | coefS[Cin] | coefA2[Cin]
  __shared__ __align__(16) float As[2][TM][TK];
  __shared__ __align__(16) float Bs[2][TNW][TK];
  __shared__ float gmean[2], grstd[2];
  __shared__ float cpart[4][TNW][2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * TNW;  // N tiles fastest: the CTAs that share an input tile run together (L2 reuse)
  const ConvSeg& S = p.seg[0];
  const int Ct = S.Cin;
  float* coefA = dsm;
  float* coefS = dsm + Ct;
  float* coefA2 = dsm + 2 * Ct;
  const int lstore = S.Lstore > 0 ? S.Lstore : S.L;

  // ------------------------------------------------------------------ prologue coefficients (GroupNorm(1) per source)
  if (warp < 2) {  // warp w: source w; its fine-group accumulators are summed as integers (exact), one load per lane
    const ConvSrc& sr = S.s[warp];
    float mean = 0.f, rstd = 1.f;
    if (sr.C > 0 && sr.stats) {
      long long a = 0, q = 0;
      for (int fg = lane; fg < sr.FG; fg += 32) {
        const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(sr.stats + ((size_t)(b % sr.bmod) * sr.FG + fg) * 2));
        a += v.x;
        q += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const double n = (double)sr.C * (double)lstore;
      const double m = stat_get_d(a) / n;
      double var = stat_get_d(q) / n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    if (lane == 0) {
      gmean[warp] = mean;
      grstd[warp] = rstd;
    }
  }
  __syncthreads();
  for (int c = tid; c < Ct; c += NT) {
    float a0 = S.s[0].scale, a1 = p.sum2 ? S.s[1].scale : 0.f, sh = 0.f;
    if (S.s[0].stats) {
      const float ga = p.gamma[c] * grstd[0];
      a0 *= ga;
      sh += p.beta[c] - gmean[0] * ga;
    }
    if (p.sum2 && S.s[1].stats) {
      const float ga = p.gamma2[c] * grstd[1];
      a1 *= ga;
      sh += p.beta2[c] - gmean[1] * ga;
    }
    coefA[c] = a0;
    coefA2[c] = a1;
    coefS[c] = sh;
  }
  __syncthreads();

  // ------------------------------------------------------------------ main loop
  const int nk = (Ct + TK - 1) / TK;
  const int total = S.ntaps * nk;
  // staging: one 16-byte load per thread and operand tile (A: 64 rows x 16 channels; B: TNW columns x 16 channels)
  const int a_c4 = (tid & 3) * 4, a_r = tid >> 2;
  const int lext = p.Lext > 0 ? p.Lext : S.L;
  const float* src0 = (const float*)S.s[0].ptr + ((size_t)(b % S.s[0].bmod) * lstore + S.row0) * Ct + a_c4;
  const float* src1 =
      p.sum2 ? (const float*)S.s[1].ptr + ((size_t)(b % S.s[1].bmod) * lstore + S.row0) * Ct + a_c4 : nullptr;
  const int m0 = blockIdx.y * TM;
  const bool elu = p.act == ACT_ELU;
  const bool row_ok = m0 + a_r < p.Lm;
  const bool b_thread = tid < TNW * 4 && n0 + a_r < p.Cout;  // B staging: column a_r, channels a_c4 .. +3
  const float* wsrc = (const float*)S.wT + (size_t)(n0 + a_r) * Ct + a_c4;  // + (tap * Cout) * Ct + kc

  float4 ra, rb;
  int f_tap = 0, f_kc = 0;  // tap / first channel of the NEXT fetch (no division in the loop)
  auto fetch = [&]() {
    const int c = f_kc + a_c4;
    int irow = (m0 + a_r) * S.in_stride + S.shift0 + f_tap * S.shift_step;
    if (p.pad_mode == PAD_REFLECT) irow = irow < 0 ? -irow : (irow >= lext ? 2 * lext - 2 - irow : irow);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row_ok && irow >= 0 && irow < S.L && c < Ct) {  // Ct % 4 == 0: the four channels are in or out together
      const float4 x = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)irow * Ct + f_kc));
      const float4 ca = *reinterpret_cast<const float4*>(coefA + c);
      const float4 cs = *reinterpret_cast<const float4*>(coefS + c);
      v[0] = fmaf(ca.x, x.x, cs.x);
      v[1] = fmaf(ca.y, x.y, cs.y);
      v[2] = fmaf(ca.z, x.z, cs.z);
      v[3] = fmaf(ca.w, x.w, cs.w);
      if (src1) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(src1 + (size_t)irow * Ct + f_kc));
        const float4 c2 = *reinterpret_cast<const float4*>(coefA2 + c);
        v[0] = fmaf(c2.x, y.x, v[0]);
        v[1] = fmaf(c2.y, y.y, v[1]);
        v[2] = fmaf(c2.z, y.z, v[2]);
        v[3] = fmaf(c2.w, y.w, v[3]);
      }
      if (elu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.0f ? v[e] : __expf(v[e]) - 1.0f;
      }
    }
    ra = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_thread && c < Ct) {
      const int wt = S.wtap0 + f_tap * S.wtap_step;
      const float4 w = __ldg(reinterpret_cast<const float4*>(wsrc + (size_t)wt * p.Cout * Ct + f_kc));
      rb = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
    }
    f_kc += TK;
    if (f_kc >= Ct) {
      f_kc = 0;
      ++f_tap;
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][a_r][a_c4]) = ra;
    if (tid < TNW * 4) *reinterpret_cast<float4*>(&Bs[buf][a_r][a_c4]) = rb;
  };

  const int wm = warp & 3, wn = warp >> 2;  // warp tile: rows wm*16 .. +15, columns wn*(8*NT8) .. 
  const int g = lane >> 2, t = lane & 3;
  float acc[NT8][4];
#pragma unroll
  for (int i = 0; i < NT8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  fetch();
  stash(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) fetch();
    {
      const float4 x = *reinterpret_cast<const float4*>(&As[buf][wm * 16 + g][4 * t]);
      const float4 y = *reinterpret_cast<const float4*>(&As[buf][wm * 16 + g + 8][4 * t]);
      const uint32_t a0[4] = {__float_as_uint(x.x), __float_as_uint(y.x), __float_as_uint(x.y), __float_as_uint(y.y)};
      const uint32_t a1[4] = {__float_as_uint(x.z), __float_as_uint(y.z), __float_as_uint(x.w), __float_as_uint(y.w)};
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const float4 z = *reinterpret_cast<const float4*>(&Bs[buf][wn * (8 * NT8) + nt * 8 + g][4 * t]);
        const uint32_t b0[2] = {__float_as_uint(z.x), __float_as_uint(z.y)};
        const uint32_t b1[2] = {__float_as_uint(z.z), __float_as_uint(z.w)};
        mma_tf32(acc[nt], a0, b0);
        mma_tf32(acc[nt], a1, b1);
      }
    }
    if (it + 1 < total) stash(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  // thread owns rows {wm*16 + g, +8} x columns {wn*8*NT8 + nt*8 + 2t, +1}: acc[nt][0..1] row g, acc[nt][2..3] row g + 8
  float cS[NT8][2], cQ[NT8][2];
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt) cS[nt][0] = cS[nt][1] = cQ[nt][0] = cQ[nt][1] = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + wm * 16 + g + 8 * h;
    const int o = m * p.out_stride + p.out_off0;
    const bool rv = (m < p.Lm) && (o >= 0) && (o < p.Lout);
    float* op = (float*)p.out + ((size_t)b * p.Lout + o) * p.Cout;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const int n = n0 + wn * (8 * NT8) + nt * 8 + 2 * t;
      float x0 = 0.f, x1 = 0.f;
      if (rv && n < p.Cout) {  // Cout % 4 == 0: n and n + 1 are in or out together
        x0 = acc[nt][2 * h] + (p.bias ? p.bias[n] : 0.f);
        x1 = acc[nt][2 * h + 1] + (p.bias ? p.bias[n + 1] : 0.f);
        *reinterpret_cast<float2*>(op + n) = make_float2(x0, x1);
      }
      cS[nt][0] += x0;
      cS[nt][1] += x1;
      cQ[nt][0] += x0 * x0;
      cQ[nt][1] += x1 * x1;
    }
  }
  if (p.stats_out) {
    // column sums over the warp's 16 rows (lanes with equal t), then over the 4 row-warps, then per fine group
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s = cS[nt][e], q = cQ[nt][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g == 0) {
          cpart[wm][wn * (8 * NT8) + nt * 8 + 2 * t + e][0] = s;
          cpart[wm][wn * (8 * NT8) + nt * 8 + 2 * t + e][1] = q;
        }
      }
    __syncthreads();
    const int gs = min(p.Cout / p.FGo, TNW);  // channels per fine group inside this tile (host: gs | TNW or Cout < TNW)
    const int ngl = TNW / gs;
    if (tid < ngl) {
      const int fg = n0 / gs + tid;
      if (fg < p.FGo) {
        float a = 0.f, q = 0.f;
        for (int c = tid * gs; c < (tid + 1) * gs; ++c)
          for (int w = 0; w < 4; ++w) {
            a += cpart[w][c][0];
            q += cpart[w][c][1];
          }
        const int slots = p.stat_slots > 1 ? p.stat_slots : 1;
        long long* so = p.stats_out + (((size_t)b * p.FGo + fg) * slots + blockIdx.y % slots) * 2;
        stat_add(so, a);
        stat_add(so + 1, q);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
#ifndef JEN1_PANEL_MINB
#define JEN1_PANEL_MINB 4
#endif
// Panel variant: the CTA stages its input rows ONCE -- all Cin channels of rows [m0 + smin, m0 + 64 + smax), normalised,
// activated, padded, rounded to TF32 -- and then walks every tap and every 64-column output tile over that panel, so an
// input element is transformed once per CTA instead of once per (tap, output tile): 2x (convtr 64 -> 32) ... 64x
// (convtr 512 -> 256) less staging work, which is what bounds the streaming kernel above.  Only the weight tiles are
// streamed (double-buffered).  Panel row stride Cin + 16 floats: 16-byte fragment loads are conflict-free.
template <int NT8>
__global__ void __launch_bounds__(NT, JEN1_PANEL_MINB) conv_tf32_panel_kernel(const ConvParams p, const int smin, const int prow) {
  constexpr int TNW = NT8 * 16;
  extern __shared__ __align__(16) float psm[];  // coefA[Cin] | coefS[Cin] | coefA2[Cin] | panel[prow][Cin + 16]
  __shared__ __align__(16) float Bs[2][TNW][TK];
  __shared__ float gmean[2], grstd[2];
  __shared__ float cpart[4][TNW][2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const ConvSeg& S = p.seg[0];
  const int Ct = S.Cin, ldp = Ct + 16;
  float* coefA = psm;
  float* coefS = psm + Ct;
  float* coefA2 = psm + 2 * Ct;
  float* panel = psm + 3 * Ct;
  const int lstore = S.Lstore > 0 ? S.Lstore : S.L;
  const int m0 = blockIdx.y * TM;

  if (warp < 2) {
    const ConvSrc& sr = S.s[warp];
    float mean = 0.f, rstd = 1.f;
    if (sr.C > 0 && sr.stats) {
      long long a = 0, q = 0;
      for (int fg = lane; fg < sr.FG; fg += 32) {
        const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(sr.stats + ((size_t)(b % sr.bmod) * sr.FG + fg) * 2));
        a += v.x;
        q += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const double n = (double)sr.C * (double)lstore;
      const double m = stat_get_d(a) / n;
      double var = stat_get_d(q) / n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    if (lane == 0) {
      gmean[warp] = mean;
      grstd[warp] = rstd;
    }
  }
  __syncthreads();
  for (int c = tid; c < Ct; c += NT) {
    float a0 = S.s[0].scale, a1 = p.sum2 ? S.s[1].scale : 0.f, sh = 0.f;
    if (S.s[0].stats) {
      const float ga = p.gamma[c] * grstd[0];
      a0 *= ga;
      sh += p.beta[c] - gmean[0] * ga;
    }
    if (p.sum2 && S.s[1].stats) {
      const float ga = p.gamma2[c] * grstd[1];
      a1 *= ga;
      sh += p.beta2[c] - gmean[1] * ga;
    }
    coefA[c] = a0;
    coefA2[c] = a1;
    coefS[c] = sh;
  }
  __syncthreads();

  // ------------------------------------------------------------------ the panel
  {
    const int lext = p.Lext > 0 ? p.Lext : S.L;
    const float* src0 = (const float*)S.s[0].ptr + ((size_t)(b % S.s[0].bmod) * lstore + S.row0) * Ct;
    const float* src1 = p.sum2 ? (const float*)S.s[1].ptr + ((size_t)(b % S.s[1].bmod) * lstore + S.row0) * Ct : nullptr;
    const bool elu = p.act == ACT_ELU;
    const int c4n = Ct >> 2;  // 16-byte items per row
#pragma unroll 2
    for (int item = tid; item < prow * c4n; item += NT) {
      const int pr = item / c4n, c = (item - pr * c4n) * 4;
      int irow = m0 + pr + smin;
      if (p.pad_mode == PAD_REFLECT) irow = irow < 0 ? -irow : (irow >= lext ? 2 * lext - 2 - irow : irow);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (irow >= 0 && irow < S.L) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)irow * Ct + c));
        const float4 ca = *reinterpret_cast<const float4*>(coefA + c);
        const float4 cs = *reinterpret_cast<const float4*>(coefS + c);
        v[0] = fmaf(ca.x, x.x, cs.x);
        v[1] = fmaf(ca.y, x.y, cs.y);
        v[2] = fmaf(ca.z, x.z, cs.z);
        v[3] = fmaf(ca.w, x.w, cs.w);
        if (src1) {
          const float4 y = __ldg(reinterpret_cast<const float4*>(src1 + (size_t)irow * Ct + c));
          const float4 c2 = *reinterpret_cast<const float4*>(coefA2 + c);
          v[0] = fmaf(c2.x, y.x, v[0]);
          v[1] = fmaf(c2.y, y.y, v[1]);
          v[2] = fmaf(c2.z, y.z, v[2]);
          v[3] = fmaf(c2.w, y.w, v[3]);
        }
        if (elu) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.0f ? v[e] : __expf(v[e]) - 1.0f;
        }
      }
      *reinterpret_cast<float4*>(panel + (size_t)pr * ldp + c) = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    }
  }
  // (the first __syncthreads of the tile loop below publishes the panel)

  // ------------------------------------------------------------------ output tiles x taps x channel steps
  const int nk = Ct / TK;  // Cin % 32 == 0
  const int total = S.ntaps * nk;
  const int a_c4 = (tid & 3) * 4, a_r = tid >> 2;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (p.Cout + TNW - 1) / TNW;
  for (int nti = 0; nti < ntiles; ++nti) {
    const int n0 = nti * TNW;
    const bool b_thread = tid < TNW * 4 && n0 + a_r < p.Cout;
    const float* wsrc = (const float*)S.wT + (size_t)(n0 + a_r) * Ct + a_c4;
    float4 rb;
    int f_tap = 0, f_kc = 0;
    auto fetch = [&]() {
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_thread) {
        const int wt = S.wtap0 + f_tap * S.wtap_step;
        const float4 w = __ldg(reinterpret_cast<const float4*>(wsrc + (size_t)wt * p.Cout * Ct + f_kc));
        rb = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
      }
      f_kc += TK;
      if (f_kc >= Ct) {
        f_kc = 0;
        ++f_tap;
      }
    };
    float acc[NT8][4];
#pragma unroll
    for (int i = 0; i < NT8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    fetch();
    if (tid < TNW * 4) *reinterpret_cast<float4*>(&Bs[0][a_r][a_c4]) = rb;
    __syncthreads();
    int tap = 0, kc = 0;
    for (int it = 0; it < total; ++it) {
      const int buf = it & 1;
      if (it + 1 < total) fetch();
      {
        const float* pa = panel + (size_t)(wm * 16 + g + S.shift0 + tap * S.shift_step - smin) * ldp + kc + 4 * t;
        const float4 x = *reinterpret_cast<const float4*>(pa);
        const float4 y = *reinterpret_cast<const float4*>(pa + 8 * ldp);
        const uint32_t a0[4] = {__float_as_uint(x.x), __float_as_uint(y.x), __float_as_uint(x.y), __float_as_uint(y.y)};
        const uint32_t a1[4] = {__float_as_uint(x.z), __float_as_uint(y.z), __float_as_uint(x.w), __float_as_uint(y.w)};
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          const float4 z = *reinterpret_cast<const float4*>(&Bs[buf][wn * (8 * NT8) + nt * 8 + g][4 * t]);
          const uint32_t b0[2] = {__float_as_uint(z.x), __float_as_uint(z.y)};
          const uint32_t b1[2] = {__float_as_uint(z.z), __float_as_uint(z.w)};
          mma_tf32(acc[nt], a0, b0);
          mma_tf32(acc[nt], a1, b1);
        }
      }
      kc += TK;
      if (kc >= Ct) {
        kc = 0;
        ++tap;
      }
      if (it + 1 < total && tid < TNW * 4) *reinterpret_cast<float4*>(&Bs[buf ^ 1][a_r][a_c4]) = rb;
      __syncthreads();
    }

    // ---- epilogue of this output tile
    float cS[NT8][2], cQ[NT8][2];
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) cS[nt][0] = cS[nt][1] = cQ[nt][0] = cQ[nt][1] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + wm * 16 + g + 8 * h;
      const int o = m * p.out_stride + p.out_off0;
      const bool rv = (m < p.Lm) && (o >= 0) && (o < p.Lout);
      float* op = (float*)p.out + ((size_t)b * p.Lout + o) * p.Cout;
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const int n = n0 + wn * (8 * NT8) + nt * 8 + 2 * t;
        float x0 = 0.f, x1 = 0.f;
        if (rv && n < p.Cout) {
          x0 = acc[nt][2 * h] + (p.bias ? p.bias[n] : 0.f);
          x1 = acc[nt][2 * h + 1] + (p.bias ? p.bias[n + 1] : 0.f);
          *reinterpret_cast<float2*>(op + n) = make_float2(x0, x1);
        }
        cS[nt][0] += x0;
        cS[nt][1] += x1;
        cQ[nt][0] += x0 * x0;
        cQ[nt][1] += x1 * x1;
      }
    }
    if (p.stats_out) {
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float s = cS[nt][e], q = cQ[nt][e];
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
          }
          if (g == 0) {
            cpart[wm][wn * (8 * NT8) + nt * 8 + 2 * t + e][0] = s;
            cpart[wm][wn * (8 * NT8) + nt * 8 + 2 * t + e][1] = q;
          }
        }
      __syncthreads();
      const int gs = min(p.Cout / p.FGo, TNW);
      const int ngl = TNW / gs;
      if (tid < ngl) {
        const int fg = n0 / gs + tid;
        if (fg < p.FGo) {
          float a = 0.f, q = 0.f;
          for (int c = tid * gs; c < (tid + 1) * gs; ++c)
            for (int w = 0; w < 4; ++w) {
              a += cpart[w][c][0];
              q += cpart[w][c][1];
            }
          const int slots = p.stat_slots > 1 ? p.stat_slots : 1;
          long long* so = p.stats_out + (((size_t)b * p.FGo + fg) * slots + blockIdx.y % slots) * 2;
          stat_add(so, a);
          stat_add(so + 1, q);
        }
      }
      __syncthreads();  // cpart is rewritten by the next tile
    }
  }
}

// Supported: fp32 storage, one segment, PRO_AFFINE, nphase 1, G <= 1 (or sum2), no FiLM / residual / LayerNorm partials.
bool conv_tf32_supported(const ConvParams& p) {
  return p.nseg == 1 && p.mode == PRO_AFFINE && p.nphase == 1 && p.film == nullptr && p.res == nullptr && p.out != nullptr &&
         p.rowpart_out == nullptr && p.epi_act == ACT_NONE && (p.sum2 || p.seg[0].s[1].C == 0) && p.G <= 1 &&
         (p.act == ACT_NONE || p.act == ACT_ELU) && (p.Lm + TM - 1) / TM <= 65535 && p.seg[0].Cin % 4 == 0 && p.Cout % 4 == 0 &&
         p.seg[0].wT != nullptr && (p.stats_out == nullptr || p.Cout <= 32 || (p.Cout / p.FGo) == 64);
}

cudaError_t launch_conv_tf32(const ConvParams& p, cudaStream_t stream) {
  const ConvSeg& S = p.seg[0];
  // panel variant (input staged once per CTA): whenever an input element would otherwise be staged more than once and the
  // panel leaves room for 4 CTAs per SM (Cin <= 128; measured per layer at 4 x 30 s: -9 ... -31 %, but +5 % at Cin = 256)
  static int panel_ok = -1;
  if (panel_ok < 0) {
    const char* e = getenv("JEN1_TF32_PANEL");  // JEN1_TF32_PANEL=0: streaming kernel only (A/B knob)
    panel_ok = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int tnw = p.Cout <= 32 ? 32 : 64;
  const int ntiles = (p.Cout + tnw - 1) / tnw;
  const int step = S.shift_step;
  if (panel_ok && S.in_stride == 1 && S.ntaps <= 8 && S.ntaps * ntiles >= 2 && S.Cin % 32 == 0 && S.Cin <= 128 &&
      (S.ntaps == 1 || step == 1 || step == -1) && S.wtap_phase == 0) {
    const int last = S.shift0 + (S.ntaps - 1) * step;
    const int smin = S.shift0 < last ? S.shift0 : last, smax = S.shift0 < last ? last : S.shift0;
    const int prow = TM + (smax - smin);
    const size_t dsm = ((size_t)3 * S.Cin + (size_t)prow * (S.Cin + 16)) * sizeof(float);
    dim3 grid(1, (p.Lm + TM - 1) / TM, p.B);
    if (p.Cout <= 32) {
      cudaFuncSetAttribute(conv_tf32_panel_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      conv_tf32_panel_kernel<2><<<grid, NT, dsm, stream>>>(p, smin, prow);
    } else {
      cudaFuncSetAttribute(conv_tf32_panel_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      conv_tf32_panel_kernel<4><<<grid, NT, dsm, stream>>>(p, smin, prow);
    }
    return cudaGetLastError();
  }
  size_t dsm = (size_t)3 * S.Cin * sizeof(float);
  if (p.Cout <= 32) {
    dim3 grid((p.Cout + 31) / 32, (p.Lm + TM - 1) / TM, p.B);
    conv_tf32_kernel<2><<<grid, NT, dsm, stream>>>(p);
  } else {
    dim3 grid((p.Cout + 63) / 64, (p.Lm + TM - 1) / TM, p.B);
    conv_tf32_kernel<4><<<grid, NT, dsm, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace jen1
