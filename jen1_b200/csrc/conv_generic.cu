// Generic fp32-FMA implementation of the fused tap-GEMM operator (see conv_params.h).
//
// This is the shape-agnostic path: any channel count, any tap pattern, float or bf16 storage, fp32 math.
// It serves (a) the fp32 "strict" precision mode used for tight parity against the CPU oracle, (b) the
// conditioning-table / K-V cache builders, and (c) every shape the tcgen05 kernel (conv_umma.cu) does not take.
// CTA tile: 64 output rows (of one batch row b and one output phase z) x 64 output channels, 256 threads,
// 4x4 register micro-tile, K stepped 16 channels at a time through double-buffered shared memory.
#include "common.cuh"
#include "conv_params.h"

namespace jen1 {

namespace {
constexpr int TM = 64, TN = 64, TK = 16, NT = 256;
constexpr int LDA = TM + 4, LDB = TN + 4;
}  // namespace

template <typename TA, typename TW, typename TO>
__global__ void __launch_bounds__(NT) conv_generic_kernel(const ConvParams p) {
  extern __shared__ float dsm[];  // coefA[Cin0] | coefS[Cin0] | (sum2) coefA2[Cin0]
  __shared__ __align__(16) float As[2][TK][LDA];
  __shared__ __align__(16) float Bs[2][TK][LDB];
  __shared__ double fine[2][32][2];
  __shared__ float gmean[32], grstd[32];
  __shared__ float rmu[TM], rrs[TM];

  const int tid = threadIdx.x;
  const int b = blockIdx.z / p.nphase;
  const int z = blockIdx.z % p.nphase;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const ConvSeg& s0 = p.seg[0];
  const int Ct = s0.Cin;
  float* coefA = dsm;
  float* coefS = dsm + Ct;
  float* coefA2 = dsm + 2 * Ct;
  const int gn_rows = s0.Lstore > 0 ? s0.Lstore : s0.L;

  // ------------------------------------------------------------------ prologue coefficients
  if (p.mode == PRO_AFFINE && p.sum2) {
    // value = GN1(src0) + GN1(src1): one (mean, rstd) per source from all of its fine groups
    if (tid < 2) {
      const ConvSrc& sr = s0.s[tid];
      double a = 0.0, q = 0.0;
      float mean = 0.f, rstd = 1.f;
      if (sr.stats) {
        for (int fg = 0; fg < sr.FG; ++fg) {
          const long long* st = sr.stats + ((size_t)(b % sr.bmod) * sr.FG + fg) * 2;
          a += stat_get_d(st[0]);
          q += stat_get_d(st[1]);
        }
        const double n = (double)sr.C * (double)gn_rows;
        const double m = a / n;
        double var = q / n - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      }
      gmean[tid] = mean;
      grstd[tid] = rstd;
    }
    __syncthreads();
    for (int c = tid; c < Ct; c += NT) {
      float a0 = s0.s[0].scale, a1 = s0.s[1].scale, sh = 0.f;
      if (s0.s[0].stats) {
        const float ga = p.gamma[c] * grstd[0];
        a0 *= ga;
        sh += p.beta[c] - gmean[0] * ga;
      }
      if (s0.s[1].stats) {
        const float ga = p.gamma2[c] * grstd[1];
        a1 *= ga;
        sh += p.beta2[c] - gmean[1] * ga;
      }
      coefA[c] = a0;
      coefA2[c] = a1;
      coefS[c] = sh;
    }
  } else if (p.mode == PRO_AFFINE) {
    if (p.G > 0) {
      if (tid < 64) {  // the producers' fixed-point accumulators (order-independent => deterministic)
        const int sI = tid >> 5, fg = tid & 31;
        const ConvSrc& sr = s0.s[sI];
        double a = 0.0, q = 0.0;
        if (sr.C > 0 && fg < sr.FG) {
          const long long* st = sr.stats + ((size_t)(b % sr.bmod) * sr.FG + fg) * 2;
          a = stat_get_d(st[0]);
          q = stat_get_d(st[1]);
        }
        const double sc = (double)sr.scale;
        fine[sI][fg][0] = a * sc;
        fine[sI][fg][1] = q * sc * sc;
      }
      __syncthreads();
      if (tid < p.G) {
        const int cpg = Ct / p.G;
        const int lo = tid * cpg, hi = lo + cpg;
        double a = 0.0, q = 0.0;
        int off = 0;
        for (int s = 0; s < 2; ++s) {
          const ConvSrc& sr = s0.s[s];
          if (sr.C > 0 && p.G == 1) {  // one group: every fine group of the source, whatever their layout
            for (int f = 0; f < sr.FG; ++f) {
              a += fine[s][f][0];
              q += fine[s][f][1];
            }
          } else if (sr.C > 0) {
            const int olo = max(lo, off), ohi = min(hi, off + sr.C);
            if (ohi > olo) {
              const int gs = sr.C / sr.FG;
              for (int f = (olo - off) / gs; f < (ohi - off) / gs; ++f) {
                a += fine[s][f][0];
                q += fine[s][f][1];
              }
            }
          }
          off += sr.C;
        }
        const double n = (double)(p.gn_real_c > 0 ? p.gn_real_c / p.G : cpg) * (double)gn_rows;
        const double mean = a / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        gmean[tid] = (float)mean;
        grstd[tid] = (float)(1.0 / sqrt(var + (double)p.eps));
      }
      __syncthreads();
    }
    const int row = p.cond_row ? p.cond_row[b] : 0;
    for (int c = tid; c < Ct; c += NT) {
      const float sc = (c < s0.s[0].C) ? s0.s[0].scale : s0.s[1].scale;
      float a, s;
      if (p.G > 0) {
        const int g = c / (Ct / p.G);
        const float ga = p.gamma[c] * grstd[g];
        a = ga * sc;
        s = p.beta[c] - gmean[g] * ga;
      } else {
        a = sc;
        s = 0.0f;
      }
      if (p.film) {
        const float fs = p.film[(size_t)row * p.film_stride + c] + 1.0f;
        const float fh = p.film[(size_t)row * p.film_stride + Ct + c];
        a = a * fs;
        s = s * fs + fh;
      }
      coefA[c] = a;
      coefS[c] = s;
    }
  } else {  // PRO_ROWNORM: LayerNorm statistics of each input row from the producer's per-row partials
    if (tid < TM) {
      const int m = m0 + tid;
      float mu = 0.0f, rs = 0.0f;
      if (m < s0.L) {
        const ConvSrc& sr = s0.s[0];
        const float* rp = p.rowpart + ((size_t)(b % sr.bmod) * s0.L + m) * p.rp_nct * 2;
        float a = 0.0f, q = 0.0f;
        for (int j = 0; j < p.rp_nct; ++j) {
          a += rp[2 * j];
          q += rp[2 * j + 1];
        }
        const float inv = 1.0f / (float)sr.C;
        mu = a * inv;
        float var = q * inv - mu * mu;
        if (var < 0.0f) var = 0.0f;
        rs = 1.0f / sqrtf(var + p.ln_eps);
      }
      rmu[tid] = mu;
      rrs[tid] = rs;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ main loop
  const int nk0 = (s0.Cin + TK - 1) / TK;
  const int it0 = s0.ntaps * nk0;
  const int nk1 = (p.nseg > 1) ? (p.seg[1].Cin + TK - 1) / TK : 0;
  const int total = it0 + ((p.nseg > 1) ? p.seg[1].ntaps * nk1 : 0);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int a_c = tid & 15, a_r = tid >> 4;
  const int b_n = tid & 63, b_k = tid >> 6;
  const int tx = tid & 15, ty = tid >> 4;

  float ra[4], rb[4];
  auto fetch = [&](int it) {
    const int sg = (it >= it0) ? 1 : 0;
    const int t = sg ? it - it0 : it;
    const ConvSeg& S = p.seg[sg];
    const int nk = sg ? nk1 : nk0;
    const int tap = t / nk, kc = (t - tap * nk) * TK;
    const int shift = S.shift0 + tap * S.shift_step;
    const int wt = S.wtap0 + z * S.wtap_phase + tap * S.wtap_step;
    const int c = kc + a_c;
    const bool sum2 = sg == 0 && p.sum2;
    const bool second = !sum2 && c >= S.s[0].C;
    const ConvSrc& sr = second ? S.s[1] : S.s[0];
    const int cc = second ? c - S.s[0].C : c;
    const int lstore = S.Lstore > 0 ? S.Lstore : S.L;
    const TA* base = (const TA*)sr.ptr + ((size_t)(b % sr.bmod) * lstore + S.row0) * sr.C + cc;
    const TA* base2 = sum2 ? (const TA*)S.s[1].ptr + ((size_t)(b % S.s[1].bmod) * lstore + S.row0) * S.s[1].C + cc : nullptr;
    float ca = 0.f, cs = 0.f, ca2 = 0.f;
    if (sg == 0 && p.mode == PRO_AFFINE && c < S.Cin) {
      ca = coefA[c];
      cs = coefS[c];
      if (sum2) ca2 = coefA2[c];
    }
    const int lext = p.Lext > 0 ? p.Lext : S.L;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = a_r + 16 * i;
      const int m = m0 + r;
      int irow = m * S.in_stride + shift;
      if (sg == 0 && p.pad_mode == PAD_REFLECT) irow = irow < 0 ? -irow : (irow >= lext ? 2 * lext - 2 - irow : irow);
      float v = 0.0f;
      if (m < p.Lm && irow >= 0 && irow < S.L && c < S.Cin) {
        const float raw = ldf(base + (size_t)irow * sr.C);
        if (sg == 0) {
          if (p.mode == PRO_AFFINE) {
            v = fmaf(ca, raw, cs);
            if (sum2) v = fmaf(ca2, ldf(base2 + (size_t)irow * sr.C), v);
            if (p.act == ACT_SILU) v = silu_f(v);
            if (p.act == ACT_ELU) v = v > 0.0f ? v : __expf(v) - 1.0f;  // |abs error| < 1.2e-7 against expm1f
          } else {
            v = (raw - rmu[r]) * rrs[r];
          }
        } else {
          v = raw * sr.scale;
        }
      }
      ra[i] = v;
    }
    const int n = n0 + b_n;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = kc + b_k + 4 * i;
      rb[i] = (kk < S.Cin && n < p.Cout) ? ldf((const TW*)S.w + ((size_t)wt * S.Cin + kk) * p.Cout + n) : 0.0f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][a_c][a_r + 16 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_n] = rb[i];
  };

  fetch(0);
  stash(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    if (it + 1 < total) stash(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  const int zoff = p.out_off0 + z * p.out_off_phase;
  float colS[4] = {0, 0, 0, 0}, colQ[4] = {0, 0, 0, 0};
  float rowS[4] = {0, 0, 0, 0}, rowQ[4] = {0, 0, 0, 0};
  const bool vec_ok = (p.out != nullptr) && ((p.Cout & 3) == 0) && (n0 + tx * 4 + 3 < p.Cout);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    const int o = m * p.out_stride + zoff;
    const bool rv = (m < p.Lm) && (o >= 0) && (o < p.Lout);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float x = 0.0f;
      if (rv && n < p.Cout) {
        x = acc[i][j] + (p.bias ? p.bias[n] : 0.0f);
        if (p.epi_act == ACT_GELU) x = gelu_f(x);
        if (p.res) x += ldf((const TA*)p.res + ((size_t)(b % p.res_bmod) * p.Lout + o) * p.Cout + n);
      }
      v[j] = x;
      colS[j] += x;
      colQ[j] += x * x;
      rowS[i] += x;
      rowQ[i] += x * x;
    }
    if (rv) {
      if (p.out) {
        TO* op = (TO*)p.out + ((size_t)b * p.Lout + o) * p.Cout + n0 + tx * 4;
        if (vec_ok) {
          st4(op, v);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + tx * 4 + j < p.Cout) stf(op + j, v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + tx * 4 + j;
          if (n < p.Cout) p.out_ncl[((size_t)b * p.Cout + n) * p.Lout + o] = v[j];
        }
      }
    }
  }

  if (p.stats_out) {  // GroupNorm partials of this tile at fine-group granularity -> fixed-point accumulators
    float* part = &As[0][0][0];  // [16][64][2] floats = 8 KB (main loop is finished with As)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      part[(ty * 64 + tx * 4 + j) * 2] = colS[j];
      part[(ty * 64 + tx * 4 + j) * 2 + 1] = colQ[j];
    }
    __syncthreads();
    float* csum = &Bs[0][0][0];  // [64][2]
    if (tid < 64) {
      float a = 0.f, q = 0.f;
      for (int y = 0; y < 16; ++y) {
        a += part[(y * 64 + tid) * 2];
        q += part[(y * 64 + tid) * 2 + 1];
      }
      csum[tid * 2] = a;
      csum[tid * 2 + 1] = q;
    }
    __syncthreads();
    const int gs = p.Cout / p.FGo;  // host guarantees gs | 64 and gs <= 64
    const int ngl = TN / gs;
    if (tid < ngl) {
      const int fg = n0 / gs + tid;
      if (fg < p.FGo) {
        float a = 0.f, q = 0.f;
        for (int c = tid * gs; c < (tid + 1) * gs; ++c) {
          a += csum[c * 2];
          q += csum[c * 2 + 1];
        }
        const int slots = p.stat_slots > 1 ? p.stat_slots : 1;
        long long* so = p.stats_out + (((size_t)b * p.FGo + fg) * slots + blockIdx.x % slots) * 2;
        stat_add(so, a);
        stat_add(so + 1, q);
      }
    }
    __syncthreads();
  }
  if (p.rowpart_out) {  // per-row LayerNorm partials of the output over this CTA's channel tile
    float* rp = &As[0][0][0];  // [64][16][2]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rp[((ty * 4 + i) * 16 + tx) * 2] = rowS[i];
      rp[((ty * 4 + i) * 16 + tx) * 2 + 1] = rowQ[i];
    }
    __syncthreads();
    if (tid < TM) {
      const int m = m0 + tid;
      const int o = m * p.out_stride + zoff;
      if (m < p.Lm && o >= 0 && o < p.Lout) {
        float a = 0.f, q = 0.f;
        for (int x = 0; x < 16; ++x) {
          a += rp[(tid * 16 + x) * 2];
          q += rp[(tid * 16 + x) * 2 + 1];
        }
        float* ro = p.rowpart_out + (((size_t)b * p.Lout + o) * gridDim.y + blockIdx.y) * 2;
        ro[0] = a;
        ro[1] = q;
      }
    }
  }
}

template <typename TA, typename TW, typename TO>
cudaError_t launch_conv_generic(const ConvParams& p, cudaStream_t stream) {
  dim3 grid((p.Lm + TM - 1) / TM, (p.Cout + TN - 1) / TN, p.B * p.nphase);
  size_t dsm = (size_t)(p.sum2 ? 3 : 2) * p.seg[0].Cin * sizeof(float);
  conv_generic_kernel<TA, TW, TO><<<grid, NT, dsm, stream>>>(p);
  return cudaGetLastError();
}

template cudaError_t launch_conv_generic<float, float, float>(const ConvParams&, cudaStream_t);
template cudaError_t launch_conv_generic<bf16, bf16, bf16>(const ConvParams&, cudaStream_t);
template cudaError_t launch_conv_generic<float, bf16, bf16>(const ConvParams&, cudaStream_t);
template cudaError_t launch_conv_generic<bf16, bf16, float>(const ConvParams&, cudaStream_t);

int conv_generic_row_tile() { return TM; }
int conv_generic_col_tile() { return TN; }

}  // namespace jen1
