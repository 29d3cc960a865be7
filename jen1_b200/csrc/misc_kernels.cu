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
                                                      long long* __restrict__ stats, int C, int Cp, int L) {
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
  for (int idx = threadIdx.x; idx < PK_ROWS * Cp; idx += 256) {
    const int r = idx / Cp, c = idx - r * Cp;
    if (l0 + r < L) stf(out + ((size_t)b * L + l0 + r) * Cp + c, c < C ? tile[r * ld + c] : 0.0f);
  }
  if (stats && threadIdx.x == 0) {
    float a = 0.f, qq = 0.f;
    for (int w = 0; w < 8; ++w) {
      a += red[w][0];
      qq += red[w][1];
    }
    stat_add(stats + (size_t)b * 2, a);
    stat_add(stats + (size_t)b * 2 + 1, qq);
  }
}

template <typename T>
cudaError_t launch_pack_ncl(const float* x, T* out, long long* stats, int Bx, int C, int Cp, int L, cudaStream_t stream) {
  dim3 grid((L + PK_ROWS - 1) / PK_ROWS, Bx);
  size_t smem = (size_t)PK_ROWS * (C + 1) * sizeof(float);
  pack_ncl_kernel<T><<<grid, 256, smem, stream>>>(x, out, stats, C, Cp, L);
  return cudaGetLastError();
}
template cudaError_t launch_pack_ncl<float>(const float*, float*, long long*, int, int, int, int, cudaStream_t);
template cudaError_t launch_pack_ncl<bf16>(const float*, bf16*, long long*, int, int, int, int, cudaStream_t);

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
// Keys are visited in tiles of 64: K/V rows are fetched with 16-byte loads (8 channels), four independent
// loads in flight per thread, into fp32 shared tiles; logits use one key per lane, PV one column set per lane,
// with an fp32 online softmax.  Padded keys are multiplied by the context mask (logit 0, value 0) and stay in
// the softmax, as the reference does; causal masking uses -FLT_MAX like reference add_mask (blocks.py:304-312).
// =====================================================================================================
namespace {
constexpr int AT_QB = 16, AT_KT = 64, AT_MAXD = 128;

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __low2float(h[e]);
    v[2 * e + 1] = __high2float(h[e]);
  }
}
}  // namespace

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
  const int cpr = d >> 3;  // 16-byte chunks per head row

  for (int idx = tid; idx < AT_QB * cpr; idx += 128) {
    const int qi = idx / cpr, part = idx - qi * cpr;
    const int i = i0 + qi;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < p.N) ld8((const T*)p.q + ((size_t)r * p.N + i) * p.q_ld + p.q_off + h * d + part * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) Qs[qi * d + part * 8 + e] = v[e];
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
    const int nchunk = AT_KT * cpr;
    for (int base = 0; base < nchunk; base += 128 * 2) {
      float kv[2][8], vv[2][8];
      float mk[2];
      int jj[2], part[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = base + u * 128 + tid;
        jj[u] = idx / cpr;
        part[u] = idx - jj[u] * cpr;
        const int j = j0 + jj[u];
        mk[u] = 1.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) kv[u][e] = vv[u][e] = 0.f;
        if (idx < nchunk && j < p.M) {
          const T* rowp;
          int ko, vo;
          if (!p.cross) {
            rowp = (const T*)p.kv + ((size_t)r * p.N + j) * p.kv_ld;
            ko = p.k_off;
            vo = p.v_off;
          } else {
            if (j < S) {
              rowp = fixed ? (const T*)p.kv_fixed + (size_t)j * p.kvc_ld
                           : (const T*)p.kv_cond + ((size_t)bc * S + j) * p.kvc_ld;
              if (p.mask) mk[u] = p.mask[(size_t)bc * S + j];
            } else {
              rowp = fixed ? (const T*)p.kv_fixed + (size_t)S * p.kvc_ld
                           : (const T*)p.kv_time + (size_t)p.cond_row[r] * p.kvc_ld;
            }
            ko = p.kvc_off;
            vo = p.kvc_off + p.C;
          }
          ld8(rowp + ko + h * d + part[u] * 8, kv[u]);
          ld8(rowp + vo + h * d + part[u] * 8, vv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (base + u * 128 + tid < nchunk) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            Ks[jj[u] * ldk + part[u] * 8 + e] = kv[u][e] * mk[u];
            Vs[jj[u] * ldk + part[u] * 8 + e] = vv[u][e] * mk[u];
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < AT_KT / 32; ++half) {
      const int jb = j0 + half * 32;
      if (jb >= p.M) break;  // CTA-uniform
      const int j = jb + lane;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int qi = warp * 4 + a;
        const int i = i0 + qi;
        if (i >= p.N) continue;  // warp-uniform
        float s = 0.f;
        const float* qrow = Qs + qi * d;
        const float* krow = Ks + (half * 32 + lane) * ldk;
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
        const int kmax = min(32, p.M - jb);
        for (int k2 = 0; k2 < kmax; ++k2) {
          const float pb = __shfl_sync(0xffffffffu, pj, k2);
          const float* vrow = Vs + (half * 32 + k2) * ldk;
#pragma unroll
          for (int t = 0; t < AT_MAXD / 32; ++t) {
            const int e = lane + 32 * t;
            if (e < d) acc[a][t] = fmaf(pb, vrow[e], acc[a][t]);
          }
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

// Function attributes are per device: called once per engine from Engine::finalize() after cudaSetDevice.
cudaError_t attention_init() {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

template <typename T>
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  if (p.d > AT_MAXD || (p.d & 7) || (p.q_ld & 7) || (p.q_off & 7) || (p.C & 7)) return cudaErrorInvalidValue;
  if (!p.cross && ((p.kv_ld & 7) || (p.k_off & 7) || (p.v_off & 7))) return cudaErrorInvalidValue;
  if (p.cross && ((p.kvc_ld & 7) || (p.kvc_off & 7))) return cudaErrorInvalidValue;
  dim3 grid((p.N + AT_QB - 1) / AT_QB, p.H, p.B2);
  size_t smem = (size_t)(AT_QB * p.d + 2 * AT_KT * (p.d + 1)) * sizeof(float);
  attention_kernel<T><<<grid, 128, smem, stream>>>(p);
  return cudaGetLastError();
}
template cudaError_t launch_attention<float>(const AttnParams&, cudaStream_t);
template cudaError_t launch_attention<bf16>(const AttnParams&, cudaStream_t);

// =====================================================================================================
// sampler epilogue.  One CTA = 32 frames of one sample: the channels-last UNet output rows are staged through
// shared memory (coalesced on both sides: [frame][channel] in, [channel][frame] out), one warp computes the
// per-frame channel statistics, then every thread updates its (channel, frame) elements.
// The arithmetic order of the reference expressions is kept (explicit _rn intrinsics, no FMA contraction):
//   out_cfg = out_masked + (out - out_masked) * scale                               model.py:362
//   pred    = phi * (out_cfg * (std(out) / std(out_cfg))) + (1 - phi) * out_cfg      model.py:366-369
//   x0      = clamp(sqrt_recip_ac * x - sqrt_recipm1_ac * pred)                      gdm.py:89-93, 131
//   x'      = x0 * sqrt(alpha_next) + c * pred + sigma * noise                       gdm.py:220-222
// =====================================================================================================
namespace {
constexpr int SP_TL = 32;
}
__global__ void __launch_bounds__(256) sampler_kernel(const SamplerParams p) {
  extern __shared__ float ssm[];  // yc [SP_TL][C+1] | yu [SP_TL][C+1] | ratio [SP_TL]
  const int C = p.C, ld = C + 1;
  float* yc = ssm;
  float* yu = yc + SP_TL * ld;
  float* ratio = yu + SP_TL * ld;
  const int b = blockIdx.y, l0 = blockIdx.x * SP_TL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = min(SP_TL, p.L - l0);
  for (int idx = tid; idx < nl * C; idx += 256) {
    const int fr = idx / C, c = idx - fr * C;
    const float o = p.y[((size_t)b * p.L + l0 + fr) * C + c];
    yc[fr * ld + c] = o;
    if (p.cfg) {
      const float u = p.y[((size_t)(b + p.B) * p.L + l0 + fr) * C + c];
      yu[fr * ld + c] = __fadd_rn(u, __fmul_rn(__fsub_rn(o, u), p.emb_scale));  // out_cfg
    }
  }
  __syncthreads();
  if (p.cfg && p.scale_cfg) {
    for (int fr = warp; fr < nl; fr += 8) {
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < C; c += 32) {
        s1 += yc[fr * ld + c];
        s2 += yu[fr * ld + c];
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      const float m1 = s1 / (float)C, m2 = s2 / (float)C;
      float v1 = 0.f, v2 = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float a = yc[fr * ld + c] - m1, g = yu[fr * ld + c] - m2;
        v1 += a * a;
        v2 += g * g;
      }
      v1 = warp_sum(v1);
      v2 = warp_sum(v2);
      if (lane == 0) ratio[fr] = __fdiv_rn(sqrtf(v1 / (float)(C - 1)), sqrtf(v2 / (float)(C - 1)));
    }
    __syncthreads();
  }
  float k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0, k6 = 0;
  bool last = false;
  if (p.mode == 1) {
    const float* cf = p.coef + (size_t)(*p.step) * 8;
    k0 = cf[0]; k1 = cf[1]; k2 = cf[2]; k3 = cf[3]; k4 = cf[4]; k5 = cf[5]; k6 = cf[6];
    last = cf[7] != 0.f;
  }
  const int fr = lane;
  if (fr < nl) {
    const float rt = (p.cfg && p.scale_cfg) ? ratio[fr] : 1.0f;
    for (int c = warp; c < C; c += 8) {
      float pred;
      if (p.cfg) {
        const float g = yu[fr * ld + c];
        pred = p.scale_cfg ? __fadd_rn(__fmul_rn(p.phi, __fmul_rn(g, rt)), __fmul_rn(p.one_minus_phi, g)) : g;
      } else {
        pred = yc[fr * ld + c];
      }
      const size_t xi = ((size_t)b * C + c) * p.L + l0 + fr;
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
}

cudaError_t launch_sampler(const SamplerParams& p, cudaStream_t stream) {
  dim3 grid((p.L + SP_TL - 1) / SP_TL, p.B);
  const size_t smem = (size_t)(2 * SP_TL * (p.C + 1) + SP_TL) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  sampler_kernel<<<grid, 256, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace jen1
