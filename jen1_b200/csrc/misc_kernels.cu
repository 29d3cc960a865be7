// Boundary, statistics, attention-core and sampler kernels of the JEN-1 denoiser hot path (sm_100a).
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

// =====================================================================================================
// pack: [Bx][C][L] fp32 (reference tensor layout) -> channels-last T [Bx][L][C]; one GroupNorm partial
// (sum, sumsq over the tile, all channels: FG = 1) per 32-row tile.
// =====================================================================================================
namespace {
constexpr int PK_ROWS = 32;
}
int pack_rows_per_entry() { return PK_ROWS; }

template <typename T>
__global__ void __launch_bounds__(256) pack_ncl_kernel(const float* __restrict__ x, T* __restrict__ out,
                                                      float* __restrict__ stats, int C, int L) {
  extern __shared__ float tile[];  // [PK_ROWS][C + 1]
  __shared__ float red[8][2];
  const int b = blockIdx.y, l0 = blockIdx.x * PK_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = C + 1;
  float s = 0.f, q = 0.f;
  for (int c = warp; c < C; c += 8) {
    const int l = l0 + lane;
    const float v = (l < L) ? x[((size_t)b * C + c) * L + l] : 0.f;
    tile[lane * ld + c] = v;
    s += v;
    q += v * v;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    red[warp][0] = s;
    red[warp][1] = q;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < PK_ROWS * C; idx += 256) {
    const int r = idx / C, c = idx - r * C;
    if (l0 + r < L) stf(out + ((size_t)b * L + l0 + r) * C + c, tile[r * ld + c]);
  }
  if (stats && threadIdx.x == 0) {
    float a = 0.f, qq = 0.f;
    for (int w = 0; w < 8; ++w) {
      a += red[w][0];
      qq += red[w][1];
    }
    float* so = stats + ((size_t)b * gridDim.x + blockIdx.x) * 2;
    so[0] = a;
    so[1] = qq;
  }
}

template <typename T>
cudaError_t launch_pack_ncl(const float* x, T* out, float* stats, int Bx, int C, int L, cudaStream_t stream) {
  dim3 grid((L + PK_ROWS - 1) / PK_ROWS, Bx);
  size_t smem = (size_t)PK_ROWS * (C + 1) * sizeof(float);
  pack_ncl_kernel<T><<<grid, 256, smem, stream>>>(x, out, stats, C, L);
  return cudaGetLastError();
}
template cudaError_t launch_pack_ncl<float>(const float*, float*, float*, int, int, int, cudaStream_t);
template cudaError_t launch_pack_ncl<bf16>(const float*, bf16*, float*, int, int, int, cudaStream_t);

// =====================================================================================================
// per-row (sum, sumsq): one warp per row
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) rowstats_kernel(const T* __restrict__ x, float* __restrict__ rowpart, int R,
                                                      int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float s = 0.f, q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = ldf(x + (size_t)row * C + c);
    s += v;
    q += v * v;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    rowpart[(size_t)row * 2] = s;
    rowpart[(size_t)row * 2 + 1] = q;
  }
}
template <typename T>
cudaError_t launch_rowstats(const T* x, float* rowpart, int R, int C, cudaStream_t stream) {
  rowstats_kernel<T><<<(R + 7) / 8, 256, 0, stream>>>(x, rowpart, R, C);
  return cudaGetLastError();
}
template cudaError_t launch_rowstats<float>(const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_rowstats<bf16>(const bf16*, float*, int, int, cudaStream_t);

// =====================================================================================================
// learned Fourier features of the raw integer timestep (reference utils/module.py:66-72):
//   [t, sin(((t*w)*2)*pi), cos(((t*w)*2)*pi)] with every product rounded to fp32 in that order.
// =====================================================================================================
__global__ void time_features_kernel(const int64_t* __restrict__ t, const float* __restrict__ w,
                                     float* __restrict__ out, int n, int half) {
  const int i = blockIdx.x;
  const float tf = (float)t[i];
  const int F = 2 * half + 1;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float f = __fmul_rn(__fmul_rn(__fmul_rn(tf, w[j]), 2.0f), 3.14159265358979323846f);
    out[(size_t)i * F + 1 + j] = sinf(f);
    out[(size_t)i * F + 1 + half + j] = cosf(f);
  }
  if (threadIdx.x == 0) out[(size_t)i * F] = tf;
}
cudaError_t launch_time_features(const int64_t* t, const float* weights, float* out, int n, int half,
                                 cudaStream_t stream) {
  time_features_kernel<<<n, 64, 0, stream>>>(t, weights, out, n, half);
  return cudaGetLastError();
}

// =====================================================================================================
// attention core.  One CTA = 16 query rows of one (batch row, head); 4 warps, each warp walks 4 query rows.
// Keys are visited in tiles of 32 (one key per lane for the logits, one value column set per lane for PV)
// with an fp32 online softmax.  Padded keys are multiplied by the context mask (logit 0, value 0) and stay in
// the softmax, as the reference does; causal masking uses -FLT_MAX like reference add_mask (blocks.py:304-312).
// =====================================================================================================
namespace {
constexpr int AT_QB = 16, AT_KT = 32, AT_MAXD = 128;
}

template <typename T>
__global__ void __launch_bounds__(128) attention_kernel(const AttnParams p) {
  extern __shared__ float sm[];
  const int d = p.d, ldk = d + 1;
  float* Qs = sm;                  // [AT_QB][d]
  float* Ks = Qs + AT_QB * d;      // [AT_KT][d+1]
  float* Vs = Ks + AT_KT * ldk;    // [AT_KT][d+1]
  const int r = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * AT_QB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bc = r % p.Bc;
  const bool fixed = p.cross && ((r >= p.Bc) || (p.drop && p.drop[bc]));
  const int S = p.M - 1;

  for (int idx = tid; idx < AT_QB * d; idx += 128) {
    const int qi = idx / d, e = idx - qi * d;
    const int i = i0 + qi;
    Qs[idx] = (i < p.N) ? ldf((const T*)p.q + ((size_t)r * p.N + i) * p.q_ld + p.q_off + h * d + e) : 0.f;
  }

  float m_run[4], l_run[4], acc[4][AT_MAXD / 32];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int t = 0; t < AT_MAXD / 32; ++t) acc[a][t] = 0.f;
  }

  for (int j0 = 0; j0 < p.M; j0 += AT_KT) {
    __syncthreads();
    for (int idx = tid; idx < AT_KT * d; idx += 128) {
      const int jj = idx / d, e = idx - jj * d;
      const int j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < p.M) {
        const T* rowp;
        float mk = 1.f;
        int ko, vo;
        if (!p.cross) {
          rowp = (const T*)p.kv + ((size_t)r * p.N + j) * p.kv_ld;
          ko = p.k_off;
          vo = p.v_off;
        } else {
          if (j < S) {
            rowp = fixed ? (const T*)p.kv_fixed + (size_t)j * p.kvc_ld
                         : (const T*)p.kv_cond + ((size_t)bc * S + j) * p.kvc_ld;
            if (p.mask) mk = p.mask[(size_t)bc * S + j];
          } else {
            rowp = fixed ? (const T*)p.kv_fixed + (size_t)S * p.kvc_ld
                         : (const T*)p.kv_time + (size_t)p.cond_row[r] * p.kvc_ld;
          }
          ko = p.kvc_off;
          vo = p.kvc_off + p.C;
        }
        kv = ldf(rowp + ko + h * d + e) * mk;
        vv = ldf(rowp + vo + h * d + e) * mk;
      }
      Ks[jj * ldk + e] = kv;
      Vs[jj * ldk + e] = vv;
    }
    __syncthreads();
    const int j = j0 + lane;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int qi = warp * 4 + a;
      const int i = i0 + qi;
      if (i >= p.N) continue;  // warp-uniform
      float s = 0.f;
      const float* qrow = Qs + qi * d;
      const float* krow = Ks + lane * ldk;
      for (int e = 0; e < d; ++e) s = fmaf(qrow[e], krow[e], s);
      s *= p.scale;
      if (p.causal && j > i + (p.M - p.N)) s = -FLT_MAX;
      if (j >= p.M) s = -INFINITY;
      const float mt = warp_max(s);
      const float m_new = fmaxf(m_run[a], mt);
      const float corr = expf(m_run[a] - m_new);
      const float pj = expf(s - m_new);
      l_run[a] = l_run[a] * corr + warp_sum(pj);
      m_run[a] = m_new;
#pragma unroll
      for (int t = 0; t < AT_MAXD / 32; ++t) acc[a][t] *= corr;
      for (int jj = 0; jj < AT_KT; ++jj) {
        const float pb = __shfl_sync(0xffffffffu, pj, jj);
#pragma unroll
        for (int t = 0; t < AT_MAXD / 32; ++t) {
          const int e = lane + 32 * t;
          if (e < d) acc[a][t] = fmaf(pb, Vs[jj * ldk + e], acc[a][t]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + warp * 4 + a;
    if (i >= p.N) continue;
    const float inv = 1.0f / l_run[a];
#pragma unroll
    for (int t = 0; t < AT_MAXD / 32; ++t) {
      const int e = lane + 32 * t;
      if (e < d) stf((T*)p.out + ((size_t)r * p.N + i) * p.C + h * d + e, acc[a][t] * inv);
    }
  }
}

template <typename T>
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.d > AT_MAXD) return cudaErrorInvalidValue;
  dim3 grid((p.N + AT_QB - 1) / AT_QB, p.H, p.B2);
  size_t smem = (size_t)(AT_QB * p.d + 2 * AT_KT * (p.d + 1)) * sizeof(float);
  attention_kernel<T><<<grid, 128, smem, stream>>>(p);
  return cudaGetLastError();
}
template cudaError_t launch_attention<float>(const AttnParams&, cudaStream_t);
template cudaError_t launch_attention<bf16>(const AttnParams&, cudaStream_t);

// =====================================================================================================
// sampler epilogue: one thread per (sample, frame); channels walked three times (means, deviations, update).
// The arithmetic order of the reference expressions is kept (explicit _rn intrinsics, no FMA contraction):
//   out_cfg = out_masked + (out - out_masked) * scale                               model.py:362
//   pred    = phi * (out_cfg * (std(out) / std(out_cfg))) + (1 - phi) * out_cfg      model.py:366-369
//   x0      = clamp(sqrt_recip_ac * x - sqrt_recipm1_ac * pred)                      gdm.py:89-93, 131
//   x'      = x0 * sqrt(alpha_next) + c * pred + sigma * noise                       gdm.py:220-222
// =====================================================================================================
__global__ void __launch_bounds__(128) sampler_kernel(const SamplerParams p) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (l >= p.L) return;
  const int C = p.C;
  const float* yc = p.y + ((size_t)b * p.L + l) * C;
  const float* yu = p.y + ((size_t)(b + p.B) * p.L + l) * C;
  float ratio = 1.0f;
  if (p.cfg && p.scale_cfg) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < C; ++c) {
      const float o = ldf(yc + c), u = ldf(yu + c);
      const float g = __fadd_rn(u, __fmul_rn(__fsub_rn(o, u), p.emb_scale));
      s1 += o;
      s2 += g;
    }
    const float m1 = s1 / (float)C, m2 = s2 / (float)C;
    float v1 = 0.f, v2 = 0.f;
    for (int c = 0; c < C; ++c) {
      const float o = ldf(yc + c), u = ldf(yu + c);
      const float g = __fadd_rn(u, __fmul_rn(__fsub_rn(o, u), p.emb_scale));
      v1 += (o - m1) * (o - m1);
      v2 += (g - m2) * (g - m2);
    }
    const float sd1 = sqrtf(v1 / (float)(C - 1)), sd2 = sqrtf(v2 / (float)(C - 1));
    ratio = __fdiv_rn(sd1, sd2);
  }
  float k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0, k6 = 0;
  bool last = false;
  if (p.mode == 1) {
    const float* cf = p.coef + (size_t)(*p.step) * 8;
    k0 = cf[0]; k1 = cf[1]; k2 = cf[2]; k3 = cf[3]; k4 = cf[4]; k5 = cf[5]; k6 = cf[6];
    last = cf[7] != 0.f;
  }
  for (int c = 0; c < C; ++c) {
    float pred;
    if (p.cfg) {
      const float o = ldf(yc + c), u = ldf(yu + c);
      const float g = __fadd_rn(u, __fmul_rn(__fsub_rn(o, u), p.emb_scale));
      pred = p.scale_cfg ? __fadd_rn(__fmul_rn(p.phi, __fmul_rn(g, ratio)), __fmul_rn(p.one_minus_phi, g)) : g;
    } else {
      pred = ldf(yc + c);
    }
    const size_t xi = ((size_t)b * C + c) * p.L + l;
    if (p.mode == 0) {
      p.pred_out[xi] = pred;
      continue;
    }
    const float xv = p.x[xi];
    float x0, eps;
    if (p.objective == 0) {
      eps = pred;
      x0 = __fsub_rn(__fmul_rn(k0, xv), __fmul_rn(k1, eps));
      x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    } else {
      if (p.objective == 1) {
        x0 = pred;
      } else {
        x0 = __fsub_rn(__fmul_rn(k2, xv), __fmul_rn(k3, pred));
      }
      x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      eps = __fdiv_rn(__fsub_rn(__fmul_rn(k0, xv), x0), k1);
    }
    float xn;
    if (last) {
      xn = x0;
    } else {
      xn = __fadd_rn(__fadd_rn(__fmul_rn(x0, k4), __fmul_rn(k5, eps)), __fmul_rn(k6, p.noise[xi]));
    }
    p.x_out[xi] = xn;
  }
}

cudaError_t launch_sampler(const SamplerParams& p, cudaStream_t stream) {
  dim3 grid((p.L + 127) / 128, p.B);
  sampler_kernel<<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace jen1
