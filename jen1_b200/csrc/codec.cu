// Encodec (SEANet) codec engine: decoder, latent [B][128][T] -> audio [B][2][T*hop], and
// encoder + residual vector quantizer, audio segments [N][2][L] -> latent [N][128][ceil(L/hop)] -> codes / quantised latent
// (reference generation.py:145-150).  The encoder is the mirror walk over the same helpers (walk_encoder).
//
// Reference: generation.py:130 `self.audio_encoder.decoder(sample_embs)` with the 48 kHz model of pip encodec==0.1.1
// (generation.py:34; the package is not vendored -- the algorithm restated here is encodec/modules/seanet.py
// SEANetDecoder, conv.py SConv1d / SConvTranspose1d / pad1d / unpad1d, lstm.py SLSTM, norm.py; see
// jen1_b200/codec_config.py for the layer list and oracle/codec_oracle.py for the CPU restatement).
//
// Data layout: channels-last fp32 [B][L][C].  Every conv output is kept RAW (before its GroupNorm(1, C)) together with
// fixed-point (sum, sumsq) accumulators written by the producing kernel's epilogue (common.cuh); the consumer applies
// GroupNorm + ELU + reflect padding while it loads its operand tile, so no normalisation / activation / padding pass
// ever touches HBM:
//   * SConv1d            = tap-GEMM, reflect index mapping in the prologue (conv_generic.cu PAD_REFLECT)
//   * SConvTranspose1d   = stride-r phases x 2 taps; the UNTRIMMED output is stored (GroupNorm statistics cover it, as
//                          encodec normalises before unpad1d) and consumers read the trimmed window (ConvSeg.row0/Lstore)
//   * SEANetResnetBlock  = three tap-GEMMs; "shortcut + block" is never materialised: the next layer's prologue adds the
//                          two normalised tensors (ConvParams.sum2)
//   * SLSTM              = W_ih x_t for all t as one k=1 tap-GEMM, then a persistent thread-block CLUSTER (16 CTAs) for
//                          up to 8 sequences: W_hh lives in registers as fp16 mma fragments, sliced over the CTAs; h_t is
//                          broadcast with bulk copies over distributed shared memory that complete on mbarriers
//                          (lstm_tc_kernel; lstm_cluster_kernel is the first version, kept as A/B partner and fallback);
//                          the skip connection is folded into the next layer's prologue (sum2, plain second source).
// Convs run on the TF32 tensor-core tap-GEMM (conv_tf32.cu; mma.sync m16n8k8, fp32 accumulation) or, in strict mode, on
// the fp32-FMA tap-GEMM (conv_generic.cu); the last 32 -> 2 conv has its own kernel.
#include "codec.h"

#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

namespace {

// ------------------------------------------------------------------------------------------------ LSTM recurrence
struct LstmParams {
  const float* gx;    // [B][T][4H] = W_ih x_t + b_ih + b_hh (gate order i, f, g, o)
  const uint4* whh;   // [CS][H/8][R] 8 fp16 each: columns 8*k8 .. 8*k8+7 of local row lr = g*U + u  (R = 4U)
  float* hout;        // [B][T][H]
  int T, H, CS, U;
};

__device__ __forceinline__ uint32_t lstm_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void lstm_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void lstm_st_remote(uint32_t local_addr, uint32_t rank, float4 v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// One cluster of CS CTAs per sequence; CTA `r` owns hidden units [r*U, (r+1)*U).  256 threads: thread (row = tid % R,
// half = tid / R) accumulates half of the H columns of gate row `row`; the U "unit" threads then combine the gates.
__global__ void __launch_bounds__(256) lstm_cluster_kernel(const LstmParams P) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int H = P.H, U = P.U, R = 4 * U, CS = P.CS;
  uint4* W = reinterpret_cast<uint4*>(lsm);                                    // [H/8][R]
  float* hbuf = reinterpret_cast<float*>(lsm + (size_t)(H / 8) * R * 16);      // [2][H]
  float* part = hbuf + 2 * H;                                                  // [2][R]
  float* hnew = part + 2 * R;                                                  // [U]
  const int tid = threadIdx.x;
  const int r = (int)lstm_cluster_rank();
  const int b = blockIdx.x / CS;

  const uint4* wsrc = P.whh + (size_t)r * (H / 8) * R;
  for (int i = tid; i < (H / 8) * R; i += 256) W[i] = wsrc[i];
  for (int i = tid; i < 2 * H; i += 256) hbuf[i] = 0.0f;
  lstm_cluster_sync();  // every CTA's buffers are initialised before anybody writes into them remotely

  const int row = tid % R, half = tid / R;
  const bool mv = tid < 2 * R;
  const int k_lo = half * (H / 16), k_hi = k_lo + H / 16;
  const float* gxb = P.gx + (size_t)b * P.T * 4 * H + r * U + tid;  // unit threads: tid < U
  float c = 0.0f;
  float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
  if (tid < U && P.T > 0) {
    gi = __ldcg(gxb);
    gf = __ldcg(gxb + H);
    gg = __ldcg(gxb + 2 * H);
    go = __ldcg(gxb + 3 * H);
  }
  for (int t = 0; t < P.T; ++t) {
    const float* hc = hbuf + (t & 1) * H;
    float ni = 0.f, nf = 0.f, ng = 0.f, no = 0.f;
    if (tid < U && t + 1 < P.T) {  // next step's input gates: in flight during this step's matvec
      const float* g = gxb + (size_t)(t + 1) * 4 * H;
      ni = __ldcg(g);
      nf = __ldcg(g + H);
      ng = __ldcg(g + 2 * H);
      no = __ldcg(g + 3 * H);
    }
    if (mv) {
      float a0 = 0.f, a1 = 0.f;
      for (int k8 = k_lo; k8 < k_hi; ++k8) {
        const uint4 w = W[(size_t)k8 * R + row];
        const float4 h0 = *reinterpret_cast<const float4*>(hc + k8 * 8);
        const float4 h1 = *reinterpret_cast<const float4*>(hc + k8 * 8 + 4);
        const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
        const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
        const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&w.z));
        const float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&w.w));
        a0 = fmaf(w0.x, h0.x, a0);
        a1 = fmaf(w0.y, h0.y, a1);
        a0 = fmaf(w1.x, h0.z, a0);
        a1 = fmaf(w1.y, h0.w, a1);
        a0 = fmaf(w2.x, h1.x, a0);
        a1 = fmaf(w2.y, h1.y, a1);
        a0 = fmaf(w3.x, h1.z, a0);
        a1 = fmaf(w3.y, h1.w, a1);
      }
      part[half * R + row] = a0 + a1;
    }
    __syncthreads();
    if (tid < U) {
      const float pi = (part[tid] + part[R + tid]) + gi;
      const float pf = (part[U + tid] + part[R + U + tid]) + gf;
      const float pg = (part[2 * U + tid] + part[R + 2 * U + tid]) + gg;
      const float po = (part[3 * U + tid] + part[R + 3 * U + tid]) + go;
      c = sigmoid_f(pf) * c + sigmoid_f(pi) * tanhf(pg);
      const float h = sigmoid_f(po) * tanhf(c);
      hnew[tid] = h;
      P.hout[((size_t)b * P.T + t) * H + r * U + tid] = h;
      gi = ni;
      gf = nf;
      gg = ng;
      go = no;
    }
    __syncthreads();
    // broadcast this CTA's U new hidden values into every CTA's next-step buffer (16 bytes per remote store)
    if (tid < CS * (U / 4)) {
      const int dst = tid / (U / 4), q = tid - dst * (U / 4);
      const float4 v = *reinterpret_cast<const float4*>(hnew + q * 4);
      float* nxt = hbuf + ((t + 1) & 1) * H + r * U + q * 4;
      lstm_st_remote((uint32_t)__cvta_generic_to_shared(nxt), (uint32_t)dst, v);
    }
    lstm_cluster_sync();  // all reads of hbuf[t&1] are done and all of hbuf[(t+1)&1] has landed
  }
}


// ------------------------------------------------------------------------------------------------ LSTM, tensor cores
// Same cluster organisation, but the recurrent matvec runs on the tensor cores for up to 8 sequences at once:
//   gates[R x 8] = W_hh slice [R x H] (fp16, REGISTER-resident mma A fragments, loaded once) x h_{t-1} [H x 8] (fp16 in
//   shared memory, one column per batch row), fp32 accumulation, mma.sync m16n8k16.
// One cluster therefore serves 8 batch rows (the smem kernel above needs a cluster per row), no weight byte moves after
// the prologue, and the per-step critical path is: H/16 MMAs per warp (4 independent accumulators) -> gate math on
// U x 8 threads -> 8-byte DSMEM broadcast of the new fp16 hidden values -> one cluster barrier.
struct LstmTcParams {
  const float* gx;   // [B][gxT][4H]: input projections of steps gx_t0 .. gx_t0 + gxT - 1
  const float* whh;  // [4H][H] fp32 (PyTorch weight_hh layout)
  float* hout;       // [B][T][H]
  float* cstate;     // [B][H] cell state carried between the chunk launches of one sequence
  int T, t0, t1;     // this launch runs steps [t0, t1); t0 > 0 resumes from hout[t0 - 1] and cstate
  int gxT, gx_t0;
  int CS, U, B;
};

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * fast_sigmoid(2.0f * x) - 1.0f; }

template <int KS, int U>  // KS = H / 16, U = hidden units per CTA (cluster size = H / U)
__global__ void __launch_bounds__(256, 1) lstm_tc_kernel(const LstmTcParams P) {
  extern __shared__ __align__(16) uint8_t lsm[];
  constexpr int H = KS * 16, R = 4 * U, CS = H / U, ldg = R + 4;
  // hidden state of step t: [source CTA][batch column][U (+ pad)] fp16 -- a source CTA's contribution is one contiguous
  // chunk (one bulk copy per destination), the pad keeps the mma B-fragment loads bank-conflict free
  constexpr int LDU = (U % 8 == 0) ? U + 8 : ((U + 7) / 8) * 8;
  constexpr int CHUNK = 8 * LDU;  // halves per source CTA
  constexpr uint32_t CHUNK_BYTES = CHUNK * 2;
  __half* hs = reinterpret_cast<__half*>(lsm);                              // [2][CS][8][LDU]
  __half* hnew = hs + 2 * CS * CHUNK;                                       // [2][8][LDU]
  float* gsm = reinterpret_cast<float*>(hnew + 2 * CHUNK);                  // [8][ldg]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(gsm + 8 * ldg);              // [2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g_ = lane >> 2, t4 = lane & 3;
  const int r = (int)lstm_cluster_rank();
  const int b0 = (blockIdx.x / CS) * 8;
  const int Bn = min(8, P.B - b0);
  constexpr int nw = R / 16;

  uint32_t a[KS][4];
  const int lr0 = warp * 16 + g_, lr1 = lr0 + 8;
  if (warp < nw) {
    const float* w0 = P.whh + ((size_t)(lr0 / U) * H + r * U + (lr0 % U)) * H;
    const float* w1 = P.whh + ((size_t)(lr1 / U) * H + r * U + (lr1 % U)) * H;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c = ks * 16 + 2 * t4;
      a[ks][0] = pack_h2(__ldg(w0 + c), __ldg(w0 + c + 1));
      a[ks][1] = pack_h2(__ldg(w1 + c), __ldg(w1 + c + 1));
      a[ks][2] = pack_h2(__ldg(w0 + c + 8), __ldg(w0 + c + 9));
      a[ks][3] = pack_h2(__ldg(w1 + c + 8), __ldg(w1 + c + 9));
    }
  }
  for (int i = tid; i < (2 * CS * CHUNK + 2 * CHUNK) / 2; i += 256) reinterpret_cast<uint32_t*>(hs)[i] = 0u;
  if (P.t0 > 0) {  // resume: h_{t0-1} of every unit (all CTAs' chunks) into buffer 0
    __syncthreads();
    for (int i = tid; i < Bn * H; i += 256) {
      const int n = i / H, k = i - n * H;
      hs[(size_t)(k / U) * CHUNK + n * LDU + k % U] = __float2half_rn(__ldcg(P.hout + ((size_t)(b0 + n) * P.T + P.t0 - 1) * H + k));
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  lstm_cluster_sync();

  // gate role: thread (n, u) finishes unit r*U + u of batch row b0 + n
  const int gn = tid / U, gu = tid - gn * U;
  const bool gate = gn < Bn;
  const float* gxp = P.gx + ((size_t)(b0 + (gate ? gn : 0)) * P.gxT + (P.t0 - P.gx_t0)) * 4 * H + r * U + gu;
  float* hop = P.hout + ((size_t)(b0 + (gate ? gn : 0)) * P.T + P.t0) * H + r * U + gu;
  float* csp = P.cstate + (size_t)(b0 + (gate ? gn : 0)) * H + r * U + gu;
  float c = 0.0f, gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
  const int nsteps = P.t1 - P.t0;
  if (gate && P.t0 > 0) c = __ldcg(csp);
  if (gate && nsteps > 0) {
    gi = __ldcg(gxp);
    gf = __ldcg(gxp + H);
    gg = __ldcg(gxp + 2 * H);
    go = __ldcg(gxp + 3 * H);
  }
  for (int t = 0; t < nsteps; ++t) {  // t: step inside this launch
    const int pc = t & 1, pn = pc ^ 1;
    const __half* cur = hs + (size_t)pc * CS * CHUNK;
    const uint32_t bar_n = (uint32_t)__cvta_generic_to_shared(&mbar[pn]);
    if (tid == 0)  // this step's CS incoming chunks complete the phase of the NEXT buffer's barrier
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_n), "r"(CS * CHUNK_BYTES) : "memory");
    float ni = 0.f, nf = 0.f, ng = 0.f, no = 0.f;
    if (gate && t + 1 < nsteps) {
      const float* g = gxp + (size_t)(t + 1) * 4 * H;
      ni = __ldcg(g);
      nf = __ldcg(g + H);
      ng = __ldcg(g + 2 * H);
      no = __ldcg(g + 3 * H);
    }
    if (warp < nw) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const __half* hb = cur + (size_t)g_ * LDU + 2 * t4;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bq0, bq1;
        if constexpr (U >= 16) {  // k = ks*16 + 2*t4 (+8): source CTA and offset inside its chunk are compile-time
          constexpr int dummy = 0;
          (void)dummy;
          const int o0 = ((ks * 16) / U) * CHUNK + (ks * 16) % U;
          bq0 = *reinterpret_cast<const uint32_t*>(hb + o0);
          bq1 = *reinterpret_cast<const uint32_t*>(hb + o0 + 8);
        } else {
          const int k0 = ks * 16 + 2 * t4, k1 = k0 + 8;
          bq0 = *reinterpret_cast<const uint32_t*>(cur + (size_t)(k0 / U) * CHUNK + g_ * LDU + k0 % U);
          bq1 = *reinterpret_cast<const uint32_t*>(cur + (size_t)(k1 / U) * CHUNK + g_ * LDU + k1 % U);
        }
        mma_f16(acc[ks & 3], a[ks], bq0, bq1);
      }
      float* g0 = gsm + (size_t)(2 * t4) * ldg;
      g0[lr0] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
      g0[ldg + lr0] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
      g0[lr1] = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
      g0[ldg + lr1] = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    }
    __syncthreads();
    __half* hn = hnew + (size_t)pc * CHUNK;
    if (gate) {
      const float* gs = gsm + (size_t)gn * ldg + gu;
      const float pi = gs[0] + gi, pf = gs[U] + gf, pg = gs[2 * U] + gg, po = gs[3 * U] + go;
      c = fast_sigmoid(pf) * c + fast_sigmoid(pi) * fast_tanh(pg);
      const float h = fast_sigmoid(po) * fast_tanh(c);
      hop[(size_t)t * H] = h;
      hn[gn * LDU + gu] = __float2half_rn(h);
      gi = ni;
      gf = nf;
      gg = ng;
      go = no;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> readable by the bulk copies
    __syncthreads();
    if (tid < CS) {  // one bulk copy per destination CTA: this CTA's chunk of the next step's hidden state
      const uint32_t dst_local = (uint32_t)__cvta_generic_to_shared(hs + ((size_t)pn * CS + r) * CHUNK);
      uint32_t dst, rbar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(dst_local), "r"((uint32_t)tid));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar_n), "r"((uint32_t)tid));
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "r"((uint32_t)__cvta_generic_to_shared(hn)), "r"(CHUNK_BYTES), "r"(rbar)
                   : "memory");
    }
    {  // wait for all CS chunks of h_t (phase parity of this barrier's use number t / 2)
      const uint32_t par = (uint32_t)((t >> 1) & 1);
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar_n), "r"(par)
            : "memory");
      }
    }
  }
  if (gate) *csp = c;
  lstm_cluster_sync();  // nobody leaves while a peer may still be copying into its shared memory
}

typedef void (*LstmTcFn)(const LstmTcParams);
// instantiated shapes: the 48 kHz model (H = 512, 16 CTAs x 32 units) and the tiny test configuration (H = 16, 4 x 4)
LstmTcFn lstm_tc_fn(int H, int U) {
  if (H == 512 && U == 32) return lstm_tc_kernel<32, 32>;
  if (H == 256 && U == 16) return lstm_tc_kernel<16, 16>;
  if (H == 16 && U == 4) return lstm_tc_kernel<1, 4>;
  return nullptr;
}
size_t lstm_tc_smem(int H, int U) {
  const int CS = H / U, LDU = (U % 8 == 0) ? U + 8 : ((U + 7) / 8) * 8, CHUNK = 8 * LDU;
  return (size_t)(2 * CS * CHUNK + 2 * CHUNK) * 2 + (size_t)8 * (4 * U + 4) * 4 + 16 + 16;
}

// ------------------------------------------------------------------------------------------------ residual VQ
// ResidualVectorQuantizer.encode + decode (encodec/quantization/core_vq.py; reference generation.py:146-149): per stage the
// nearest codebook entry of the residual (euclidean: argmax of 2 r.e - |e|^2, lowest index on ties), residual -= entry,
// quantized += entry.  fp32 throughout (an argmax does not forgive rounded operands).  One CTA = 64 frames of one row:
// residual and running sum live in shared memory [dim][frame]; the codebook streams through in 64-entry chunks.
struct RvqParams {
  const float* emb;        // [N][D][T]
  const float* codebooks;  // [nq][K][D]
  const float* enorm;      // [nq][K]
  float* quant;            // [N][D][T] or nullptr
  int32_t* codes;          // [nq][N][T] or nullptr
  int N, T, D, K, nq;
};
constexpr int RQ_F = 64, RQ_C = 64;  // frames per CTA, codebook entries per chunk

// 256 threads = 16 frame groups (4 frames) x 16 entry groups (4 entries): a 4 x 4 register tile per thread, one 16-byte
// load of residuals + 4 scalar loads of entries per 16 FMAs.
__global__ void __launch_bounds__(256) rvq_kernel(const RvqParams P) {
  extern __shared__ __align__(16) float rsm[];
  const int D = P.D;
  constexpr int lde = RQ_C + 1;
  float* R = rsm;                    // [D][64] residual
  float* Q = R + D * RQ_F;           // [D][64] quantized
  float* E = Q + D * RQ_F;           // [D][65] codebook chunk, transposed
  float* bs = E + D * lde;           // [16][64] best score per entry group
  int* bi = reinterpret_cast<int*>(bs + 16 * RQ_F);  // [16][64] its index
  int* pick = bi + 16 * RQ_F;        // [64]
  const int tid = threadIdx.x, tf = tid & 15, tc = tid >> 4;
  const int n = blockIdx.y, t0 = blockIdx.x * RQ_F;
  for (int i = tid; i < D * RQ_F; i += 256) {
    const int d = i >> 6, ff = i & 63;
    R[i] = (t0 + ff < P.T) ? P.emb[((size_t)n * D + d) * P.T + t0 + ff] : 0.f;
    Q[i] = 0.f;
  }
  __syncthreads();
  for (int q = 0; q < P.nq; ++q) {
    const float* cb = P.codebooks + (size_t)q * P.K * D;
    float best[4] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    int besti[4] = {0, 0, 0, 0};
    for (int c0 = 0; c0 < P.K; c0 += RQ_C) {
      for (int i = tid; i < RQ_C * D; i += 256) {  // coalesced along the entry's dimensions, stored transposed
        const int c = i / D, d = i - c * D;
        E[d * lde + c] = __ldg(cb + (size_t)(c0 + c) * D + d);
      }
      __syncthreads();
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
#pragma unroll 4
      for (int d = 0; d < D; ++d) {
        const float4 r = *reinterpret_cast<const float4*>(R + d * RQ_F + 4 * tf);
        const float* e = E + d * lde + 4 * tc;
        const float e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
        acc[0][0] = fmaf(r.x, e0, acc[0][0]); acc[0][1] = fmaf(r.x, e1, acc[0][1]); acc[0][2] = fmaf(r.x, e2, acc[0][2]); acc[0][3] = fmaf(r.x, e3, acc[0][3]);
        acc[1][0] = fmaf(r.y, e0, acc[1][0]); acc[1][1] = fmaf(r.y, e1, acc[1][1]); acc[1][2] = fmaf(r.y, e2, acc[1][2]); acc[1][3] = fmaf(r.y, e3, acc[1][3]);
        acc[2][0] = fmaf(r.z, e0, acc[2][0]); acc[2][1] = fmaf(r.z, e1, acc[2][1]); acc[2][2] = fmaf(r.z, e2, acc[2][2]); acc[2][3] = fmaf(r.z, e3, acc[2][3]);
        acc[3][0] = fmaf(r.w, e0, acc[3][0]); acc[3][1] = fmaf(r.w, e1, acc[3][1]); acc[3][2] = fmaf(r.w, e2, acc[3][2]); acc[3][3] = fmaf(r.w, e3, acc[3][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = c0 + 4 * tc + j;
        const float en = __ldg(P.enorm + (size_t)q * P.K + idx);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float sc = 2.0f * acc[a][j] - en;
          if (sc > best[a]) {  // ascending index inside a thread: strict > keeps the first maximum
            best[a] = sc;
            besti[a] = idx;
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      bs[tc * RQ_F + 4 * tf + a] = best[a];
      bi[tc * RQ_F + 4 * tf + a] = besti[a];
    }
    __syncthreads();
    if (tid < RQ_F) {
      float b = bs[tid];
      int ix = bi[tid];
      for (int g = 1; g < 16; ++g) {
        const float s2 = bs[g * RQ_F + tid];
        const int i2 = bi[g * RQ_F + tid];
        if (s2 > b || (s2 == b && i2 < ix)) {
          b = s2;
          ix = i2;
        }
      }
      pick[tid] = ix;
      if (P.codes && t0 + tid < P.T) P.codes[((size_t)q * P.N + n) * P.T + t0 + tid] = ix;
    }
    __syncthreads();
    {
      const int f = tid & 63;
      const float* e = cb + (size_t)pick[f] * D;
      for (int d = tid >> 6; d < D; d += 4) {
        const float v = __ldg(e + d);
        R[d * RQ_F + f] -= v;
        Q[d * RQ_F + f] += v;
      }
    }
    __syncthreads();
  }
  if (P.quant)
    for (int i = tid; i < D * RQ_F; i += 256) {
      const int d = i >> 6, ff = i & 63;
      if (t0 + ff < P.T) P.quant[((size_t)n * D + d) * P.T + t0 + ff] = Q[i];
    }
}

// ------------------------------------------------------------------------------------------------ final GroupNorm
// audio[b][c][l] = gamma[c] * (raw[b][l][c] - mean_b) * rstd_b + beta[c]   (GroupNorm(1, C) of the last conv, NCL fp32 out)
__global__ void final_norm_kernel(const float* __restrict__ raw, const long long* __restrict__ stats, int FG,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int C, int L,
                                  float* __restrict__ out) {
  const int b = blockIdx.y;
  double a = 0.0, q = 0.0;
  for (int f = 0; f < FG; ++f) {
    a += stat_get_d(stats[((size_t)b * FG + f) * 2]);
    q += stat_get_d(stats[((size_t)b * FG + f) * 2 + 1]);
  }
  const double n = (double)C * (double)L;
  const double mean = a / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x)
    for (int c = 0; c < C; ++c)
      out[((size_t)b * C + c) * L + l] = gamma[c] * ((raw[((size_t)b * L + l) * C + c] - mu) * rstd) + beta[c];
}

// ------------------------------------------------------------------------------------------------ narrow-output conv
// The last SConv1d (32 -> 2 channels, k7) would waste 62 of the tap-GEMM's 64 tile columns.  One thread per output row:
// the CTA stages rows [m0 - pad, m0 + 256 + k - 1 - pad) of value = ELU(GN(a) + GN(b)) (reflect mapping) in shared
// memory once (row stride Cin + 1: conflict-free), the weights sit next to them, each thread runs the k*Cin*Cout MACs.
struct NarrowParams {
  const float* a;
  const float* b;          // second operand of the sum (or nullptr)
  const long long* sa;     // statistics of a / b (nullptr: take as is)
  const long long* sb;
  int FGa, FGb;
  const float *ga, *ba, *gb, *bb;  // GroupNorm affines of a / b
  const float* w;          // [k][Cin][Cout]
  const float* bias;
  float* out;              // raw [B][L][Cout]
  long long* stats_out;    // [B][1][2]
  int L, Lstore, row0, Cin, Cout, k, pad_left, slots;
  float eps;
};
constexpr int NR_ROWS = 256;

__global__ void __launch_bounds__(NR_ROWS) narrow_conv_kernel(const NarrowParams P) {
  extern __shared__ float nsm[];
  const int Cin = P.Cin, ld = Cin + 1, rows = NR_ROWS + P.k - 1;
  float* tile = nsm;                         // [rows][ld]
  float* wsm = tile + (size_t)rows * ld;     // [k][Cin][Cout]
  float* coef = wsm + P.k * Cin * P.Cout;    // [3][Cin]: a0, a1, shift
  __shared__ float mr[2][2];
  __shared__ float red[NR_ROWS / 32][2];
  const int tid = threadIdx.x, b = blockIdx.y, m0 = blockIdx.x * NR_ROWS;
  if (tid < 2) {
    const long long* st = tid ? P.sb : P.sa;
    const int FG = tid ? P.FGb : P.FGa;
    float mean = 0.f, rstd = 1.f;
    if (st) {
      double a = 0.0, q = 0.0;
      for (int f = 0; f < FG; ++f) {
        a += stat_get_d(st[((size_t)b * FG + f) * 2]);
        q += stat_get_d(st[((size_t)b * FG + f) * 2 + 1]);
      }
      const double n = (double)Cin * (double)P.Lstore;
      const double m = a / n;
      double var = q / n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)P.eps));
    }
    mr[tid][0] = mean;
    mr[tid][1] = rstd;
  }
  for (int i = tid; i < P.k * Cin * P.Cout; i += NR_ROWS) wsm[i] = P.w[i];
  __syncthreads();
  for (int c = tid; c < Cin; c += NR_ROWS) {
    float a0 = 1.f, a1 = P.b ? 1.f : 0.f, sh = 0.f;
    if (P.sa) {
      a0 = P.ga[c] * mr[0][1];
      sh += P.ba[c] - mr[0][0] * a0;
    }
    if (P.b && P.sb) {
      a1 = P.gb[c] * mr[1][1];
      sh += P.bb[c] - mr[1][0] * a1;
    }
    coef[c] = a0;
    coef[Cin + c] = a1;
    coef[2 * Cin + c] = sh;
  }
  __syncthreads();
  const size_t base = ((size_t)b * P.Lstore + P.row0) * Cin;
  for (int i = tid; i < rows * Cin; i += NR_ROWS) {
    const int r = i / Cin, c = i - r * Cin;
    int ir = m0 + r - P.pad_left;
    ir = ir < 0 ? -ir : (ir >= P.L ? 2 * P.L - 2 - ir : ir);  // reflect (L > pad here: the audio-rate layers)
    float v = 0.f;
    if (ir >= 0 && ir < P.L) {
      v = fmaf(coef[c], P.a[base + (size_t)ir * Cin + c], coef[2 * Cin + c]);
      if (P.b) v = fmaf(coef[Cin + c], P.b[base + (size_t)ir * Cin + c], v);
      v = v > 0.0f ? v : __expf(v) - 1.0f;
    }
    tile[r * ld + c] = v;
  }
  __syncthreads();
  const int m = m0 + tid;
  float s = 0.f, q = 0.f;
  if (m < P.L) {
    float acc[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n] = (P.bias && n < P.Cout) ? P.bias[n] : 0.f;
    for (int j = 0; j < P.k; ++j) {
      const float* trow = tile + (tid + j) * ld;
      const float* wj = wsm + j * Cin * P.Cout;
      for (int c = 0; c < Cin; ++c) {
        const float x = trow[c];
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < P.Cout) acc[n] = fmaf(x, wj[c * P.Cout + n], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < P.Cout) {
        P.out[((size_t)b * P.L + m) * P.Cout + n] = acc[n];
        s += acc[n];
        q += acc[n] * acc[n];
      }
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if ((tid & 31) == 0) {
    red[tid >> 5][0] = s;
    red[tid >> 5][1] = q;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, qq = 0.f;
    for (int w = 0; w < NR_ROWS / 32; ++w) {
      a += red[w][0];
      qq += red[w][1];
    }
    long long* so = P.stats_out + ((size_t)b * P.slots + blockIdx.x % P.slots) * 2;
    stat_add(so, a);
    stat_add(so + 1, qq);
  }
}

int pick_cluster(int H) {
  // largest power-of-two cluster (<= 16) whose per-CTA W_hh slice (4U x H fp16) fits shared memory, with U % 4 == 0
  for (int cs = 16; cs >= 1; cs >>= 1) {
    if (H % cs) continue;
    const int U = H / cs;
    if (U % 4 || 8 * U > 256) continue;
    if ((size_t)4 * U * H * 2 > 160 * 1024) continue;
    return cs;
  }
  return 0;
}

}  // namespace

// ================================================================================================== CodecDecoder
CodecDecoder::CodecDecoder(const Jen1CodecDesc& d, int device, int strict, bool encoder, int n_q, int codebook_size)
    : d_(d), device_(device), encoder_(encoder), n_q_(n_q), K_(codebook_size), strict_(strict != 0) {
  const char* e = getenv("JEN1_LSTM");  // JEN1_LSTM=smem: the shared-memory-weights kernel (A/B partner of the tensor-core one)
  lstm_smem_kernel_ = e && strcmp(e, "smem") == 0;
  lstm_no_overlap_ = e && strcmp(e, "serial") == 0;  // JEN1_LSTM=serial: tensor-core kernel, layers one after the other
}

CodecDecoder::~CodecDecoder() {
  cudaSetDevice(device_);
  for (cudaEvent_t e : ev_) cudaEventDestroy(e);
  if (side_) cudaStreamDestroy(side_);
  for (void* p : owned_) cudaFree(p);
  if (arena_) cudaFree(arena_);
}

int CodecDecoder::fail(const std::string& m) {
  err_ = m;
  return 1;
}

bool CodecDecoder::ck(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  err_ = std::string(what) + ": " + cudaGetErrorString(e);
  (void)cudaGetLastError();
  return false;
}

int CodecDecoder::load_tensor(const char* name, const float* data, const int64_t* shape, int ndim) {
  if (finalized_) return fail("load_tensor after finalize");
  if (!name || !data || ndim < 1 || ndim > 3) return fail("load_tensor: bad arguments");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(data, data + n);
  host_[name] = std::move(t);
  return 0;
}

const CodecDecoder::HostTensor* CodecDecoder::get(const std::string& name, std::initializer_list<int64_t> shape) {
  auto it = host_.find(name);
  if (it == host_.end()) {
    fail("missing tensor " + name);
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    fail("tensor " + name + " has the wrong shape");
    return nullptr;
  }
  return &it->second;
}

float* CodecDecoder::upload(const std::vector<float>& v) {
  float* p = nullptr;
  if (!ck(cudaMalloc((void**)&p, v.size() * sizeof(float)), "cudaMalloc(weights)")) return nullptr;
  owned_.push_back(p);
  if (!ck(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice), "upload")) return nullptr;
  weight_bytes_ += (int64_t)v.size() * 4;
  return p;
}

bool CodecDecoder::make_conv(const std::string& prefix, const char* conv, const char* norm, int cin, int cout, int k,
                             bool transposed, ConvW* out) {
  const HostTensor* w = transposed ? get(prefix + conv + ".weight", {cin, cout, k}) : get(prefix + conv + ".weight", {cout, cin, k});
  const HostTensor* bs = get(prefix + conv + ".bias", {cout});
  const HostTensor* g = get(prefix + norm + ".weight", {cout});
  const HostTensor* be = get(prefix + norm + ".bias", {cout});
  if (!w || !bs || !g || !be) return false;
  std::vector<float> packed((size_t)k * cin * cout);
  std::vector<float> bias = bs->data;
  if (transposed) {  // ConvTranspose1d [Cin][Cout][k = 2r] -> [tap j][Cin][(z, n)] = w[c][n][z + j*r]  (see convtr())
    const int r = k / 2;
    bias.resize((size_t)r * cout);
    for (int z = 0; z < r; ++z)
      for (int n = 0; n < cout; ++n) bias[(size_t)z * cout + n] = bs->data[n];
    for (int j = 0; j < 2; ++j)
      for (int c = 0; c < cin; ++c)
        for (int z = 0; z < r; ++z)
          for (int n = 0; n < cout; ++n)
            packed[((size_t)j * cin + c) * r * cout + (size_t)z * cout + n] = w->data[((size_t)c * cout + n) * k + z + j * r];
  } else {  // Conv1d [Cout][Cin][k] -> tap-major [k][Cin][Cout]
    for (int j = 0; j < k; ++j)
      for (int c = 0; c < cin; ++c)
        for (int n = 0; n < cout; ++n) packed[((size_t)j * cin + c) * cout + n] = w->data[((size_t)n * cin + c) * k + j];
  }
  out->w = upload(packed);
  {  // the same taps with Cin contiguous: [tap][columns][Cin]
    const int taps = transposed ? 2 : k, cols = transposed ? (k / 2) * cout : cout;
    std::vector<float> tp(packed.size());
    for (int j = 0; j < taps; ++j)
      for (int c = 0; c < cin; ++c)
        for (int n = 0; n < cols; ++n) tp[((size_t)j * cols + n) * cin + c] = packed[((size_t)j * cin + c) * cols + n];
    out->wT = upload(tp);
  }
  out->bias = upload(bias);
  out->gamma = upload(g->data);
  out->beta = upload(be->data);
  out->cin = cin;
  out->cout = cout;
  out->k = k;
  return out->w && out->wT && out->bias && out->gamma && out->beta;
}

int CodecDecoder::finalize() {
  if (finalized_) return 0;
  if (encoder_) {  // SEANetEncoder: model.0 conv, then (resblock, ELU, down conv) per ratio, LSTM, ELU, conv
    lstm_prefix_ = "encoder.model." + std::to_string(3 * d_.n_ratios + 1) + ".lstm.";
  }
  if (cudaSetDevice(device_) != cudaSuccess) return fail("no CUDA device " + std::to_string(device_) + " (this library has no CPU path)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device_ >= ndev) return fail("no CUDA device (this library has no CPU path)");
  if (d_.n_ratios < 1 || d_.n_ratios > 8 || d_.lstm_layers < 0 || d_.lstm_layers > 4) return fail("bad decoder description");
  H_ = d_.n_filters << d_.n_ratios;
  hop_ = 1;
  for (int i = 0; i < d_.n_ratios; ++i) hop_ *= d_.ratios[i];
  if (H_ % 16) return fail("hidden size must be a multiple of 16");
  if (encoder_) return finalize_encoder();
  if (!make_conv("model.0", ".conv.conv", ".conv.norm", d_.dimension, H_, d_.kernel_size, false, &first_)) return 1;
  if (load_lstm()) return 1;
  // ---- upsampling stages
  int idx = 2, c = H_;
  for (int i = 0; i < d_.n_ratios; ++i) {
    const int r = d_.ratios[i];
    Stage S;
    S.ratio = r;
    const std::string pt = "model." + std::to_string(idx + 1), pr = "model." + std::to_string(idx + 2);
    const int hid = (c / 2) / d_.compress;
    if (!make_conv(pt, ".convtr.convtr", ".convtr.norm", c, c / 2, 2 * r, true, &S.up)) return 1;
    if (!make_conv(pr + ".block.1", ".conv.conv", ".conv.norm", c / 2, hid, d_.residual_kernel_size, false, &S.res1)) return 1;
    if (!make_conv(pr + ".block.3", ".conv.conv", ".conv.norm", hid, c / 2, 1, false, &S.res2)) return 1;
    if (!make_conv(pr + ".shortcut", ".conv.conv", ".conv.norm", c / 2, c / 2, 1, false, &S.shortcut)) return 1;
    stages_.push_back(S);
    idx += 3;
    c /= 2;
  }
  if (!make_conv("model." + std::to_string(idx + 1), ".conv.conv", ".conv.norm", c, d_.channels, d_.last_kernel_size, false, &last_))
    return 1;
  return finish_finalize();
}

int CodecDecoder::load_lstm() {
  // ---- LSTM
  CS_ = pick_cluster(H_);
  if (d_.lstm_layers > 0 && CS_ == 0) return fail("LSTM hidden size does not fit the cluster kernel");
  U_ = CS_ ? H_ / CS_ : 0;
  for (int l = 0; l < d_.lstm_layers; ++l) {
    const std::string p = lstm_prefix_;
    const std::string s = "_l" + std::to_string(l);
    const HostTensor* wih = get(p + "weight_ih" + s, {4 * H_, H_});
    const HostTensor* whh = get(p + "weight_hh" + s, {4 * H_, H_});
    const HostTensor* bih = get(p + "bias_ih" + s, {4 * H_});
    const HostTensor* bhh = get(p + "bias_hh" + s, {4 * H_});
    if (!wih || !whh || !bih || !bhh) return 1;
    LstmW L;
    std::vector<float> wt((size_t)H_ * 4 * H_), bsum((size_t)4 * H_);
    for (int n = 0; n < 4 * H_; ++n) {
      bsum[n] = bih->data[n] + bhh->data[n];
      for (int c = 0; c < H_; ++c) wt[(size_t)c * 4 * H_ + n] = wih->data[(size_t)n * H_ + c];
    }
    L.wih = upload(wt);
    L.wih_T = upload(wih->data);
    L.bias = upload(bsum);
    L.whh_f32 = upload(whh->data);
    const int R = 4 * U_;
    std::vector<__half> pk((size_t)CS_ * (H_ / 8) * R * 8);
    for (int r = 0; r < CS_; ++r)
      for (int k8 = 0; k8 < H_ / 8; ++k8)
        for (int lr = 0; lr < R; ++lr) {
          const int g = lr / U_, u = lr % U_;
          const int grow = g * H_ + r * U_ + u;
          for (int e = 0; e < 8; ++e)
            pk[(((size_t)r * (H_ / 8) + k8) * R + lr) * 8 + e] = __float2half_rn(whh->data[(size_t)grow * H_ + k8 * 8 + e]);
        }
    void* dp = nullptr;
    if (!ck(cudaMalloc(&dp, pk.size() * sizeof(__half)), "cudaMalloc(W_hh)")) return 1;
    owned_.push_back(dp);
    if (!ck(cudaMemcpy(dp, pk.data(), pk.size() * sizeof(__half), cudaMemcpyHostToDevice), "upload W_hh")) return 1;
    weight_bytes_ += (int64_t)pk.size() * 2;
    L.whh = reinterpret_cast<const uint4*>(dp);
    if (!L.wih || !L.bias) return 1;
    lstm_.push_back(L);
  }
  return 0;
}

int CodecDecoder::finish_finalize() {
  if (CS_ > 0) {
    const size_t smem = lstm_smem();
    if (!ck(cudaFuncSetAttribute(lstm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lstm smem attribute"))
      return 1;
    if (CS_ > 8 &&
        !ck(cudaFuncSetAttribute(lstm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "lstm cluster attribute"))
      return 1;
  }
  if (cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking) != cudaSuccess) side_ = nullptr;
  for (int i = 0; i <= kLstmChunks && side_; ++i) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate");
    ev_.push_back(e);
  }
  host_.clear();
  finalized_ = true;
  return 0;
}

// SEANetEncoder + residual vector quantizer.  The audio input is padded from `channels` to 4 channels (zero weights) so
// that the first conv takes the tensor-core path too.
int CodecDecoder::finalize_encoder() {
  if (n_q_ < 1 || n_q_ > 32 || K_ < 64 || K_ % 64) return fail("bad quantizer description");
  cin0_ = (d_.channels + 3) / 4 * 4;
  {  // first conv with zero-padded input channels
    const std::string pfx = "encoder.model.0";
    const HostTensor* w = get(pfx + ".conv.conv.weight", {d_.n_filters, d_.channels, d_.kernel_size});
    if (!w) return 1;
    HostTensor padded;
    padded.shape = {d_.n_filters, cin0_, d_.kernel_size};
    padded.data.assign((size_t)d_.n_filters * cin0_ * d_.kernel_size, 0.f);
    for (int n = 0; n < d_.n_filters; ++n)
      for (int c = 0; c < d_.channels; ++c)
        for (int j = 0; j < d_.kernel_size; ++j)
          padded.data[((size_t)n * cin0_ + c) * d_.kernel_size + j] = w->data[((size_t)n * d_.channels + c) * d_.kernel_size + j];
    host_[pfx + ".conv.conv.weight"] = padded;
    if (!make_conv(pfx, ".conv.conv", ".conv.norm", cin0_, d_.n_filters, d_.kernel_size, false, &first_)) return 1;
  }
  int idx = 0, c = d_.n_filters;
  for (int i = d_.n_ratios - 1; i >= 0; --i) {  // reversed decoder ratios
    const int r = d_.ratios[i];
    Stage S;
    S.ratio = r;
    const std::string pr = "encoder.model." + std::to_string(idx + 1), pd = "encoder.model." + std::to_string(idx + 3);
    const int hid = c / d_.compress;
    if (!make_conv(pr + ".block.1", ".conv.conv", ".conv.norm", c, hid, d_.residual_kernel_size, false, &S.res1)) return 1;
    if (!make_conv(pr + ".block.3", ".conv.conv", ".conv.norm", hid, c, 1, false, &S.res2)) return 1;
    if (!make_conv(pr + ".shortcut", ".conv.conv", ".conv.norm", c, c, 1, false, &S.shortcut)) return 1;
    if (!make_conv(pd, ".conv.conv", ".conv.norm", c, 2 * c, 2 * r, false, &S.up)) return 1;
    stages_.push_back(S);
    idx += 3;
    c *= 2;
  }
  if (c != H_) return fail("encoder channel bookkeeping");
  if (load_lstm()) return 1;
  if (!make_conv("encoder.model." + std::to_string(idx + 3), ".conv.conv", ".conv.norm", H_, d_.dimension, d_.last_kernel_size, false,
                 &last_))
    return 1;
  std::vector<float> cb((size_t)n_q_ * K_ * d_.dimension), en((size_t)n_q_ * K_);
  for (int q = 0; q < n_q_; ++q) {
    const HostTensor* e = get("quantizer.vq.layers." + std::to_string(q) + "._codebook.embed", {K_, d_.dimension});
    if (!e) return 1;
    for (int j = 0; j < K_; ++j) {
      double n2 = 0.0;
      for (int dd = 0; dd < d_.dimension; ++dd) {
        const float v = e->data[(size_t)j * d_.dimension + dd];
        cb[((size_t)q * K_ + j) * d_.dimension + dd] = v;
        n2 += (double)v * v;
      }
      en[(size_t)q * K_ + j] = (float)n2;
    }
  }
  codebooks_ = upload(cb);
  enorm_ = upload(en);
  if (!codebooks_ || !enorm_) return 1;
  return finish_finalize();
}

size_t CodecDecoder::lstm_smem() const {
  const int R = 4 * U_;
  return (size_t)(H_ / 8) * R * 16 + (size_t)2 * H_ * 4 + (size_t)2 * R * 4 + (size_t)U_ * 4 + 16;
}

// ---- workspace: a bump arena sized by a dry run of the same walk (nothing is reused: 2 GB per 30 s sample)
float* CodecDecoder::falloc(size_t n) {
  const size_t bytes = (n * sizeof(float) + 255) & ~(size_t)255;
  float* p = dry_ ? nullptr : reinterpret_cast<float*>(arena_ + off_);
  off_ += bytes;
  return p;
}
long long* CodecDecoder::salloc(int B, int FG) {
  const size_t bytes = ((size_t)B * FG * 2 * sizeof(long long) + 255) & ~(size_t)255;
  long long* p = dry_ ? nullptr : reinterpret_cast<long long*>(arena_ + soff_);
  soff_ += bytes;
  return p;
}

size_t CodecDecoder::workspace_bytes(int B, int T) {
  dry_ = true;
  off_ = 0;
  soff_ = 0;
  walk(nullptr, nullptr, B, T, nullptr);
  dry_ = false;
  stats_bytes_need_ = soff_;
  return off_ + soff_ + 4096;
}

int CodecDecoder::reserve(int B, int T) {
  if (!finalized_) return fail("reserve before finalize");
  if (B < 1 || T < 1) return fail("reserve: bad shape");
  cudaSetDevice(device_);
  const size_t need = workspace_bytes(B, T);
  if (need > arena_bytes_) {
    if (arena_) cudaFree(arena_);
    arena_ = nullptr;
    arena_bytes_ = 0;
    if (!ck(cudaMalloc((void**)&arena_, need), "cudaMalloc(codec workspace)")) return 1;
    arena_bytes_ = need;
  }
  return 0;
}

// TF32 tensor-core tap-GEMM by default; fp32 FMA in strict mode and for shapes the TF32 kernel does not take
cudaError_t CodecDecoder::launch_conv(const ConvParams& p, cudaStream_t st) {
  if (!strict_ && conv_tf32_supported(p)) {
    ++tf32_launches_;
    return launch_conv_tf32(p, st);
  }
  return launch_conv_generic<float, float, float>(p, st);
}

CodecDecoder::Act CodecDecoder::conv(const Act& in, const Act* in2, int act, const ConvW& W, int pad_left, bool reflect,
                                     bool want_stats, cudaStream_t st, int stride) {
  // stride > 1 (encoder down convs): ceil(L / stride) frames; the padding on the right grows by what completes the last
  // frame (encodec get_extra_padding_for_conv1d) -- with reflect index mapping that needs no special case
  Act o;
  o.C = W.cout;
  o.L = (in.L + stride - 1) / stride;
  o.Lstore = o.L;
  o.row0 = 0;
  const int fgo = W.cout >= 64 ? W.cout / 64 : 1;
  const int slots = 32 / fgo > 1 ? 32 / fgo : 1;  // spread the statistics atomics of the long layers (<= 32 entries)
  o.FG = fgo * slots;
  o.gamma = W.gamma;
  o.beta = W.beta;
  o.ptr = falloc((size_t)B_ * o.Lstore * o.C);
  o.stats = want_stats ? salloc(B_, o.FG) : nullptr;
  if (dry_) return o;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  fill_src(p, in, in2, act);
  ConvSeg& S = p.seg[0];
  S.w = W.w;
  S.wT = W.wT;
  S.ntaps = W.k;
  S.in_stride = stride;
  S.shift0 = -pad_left;
  S.shift_step = 1;
  S.wtap0 = 0;
  S.wtap_phase = 0;
  S.wtap_step = 1;
  p.pad_mode = reflect ? PAD_REFLECT : PAD_ZERO;
  const int pad_right = (o.L - 1) * stride + W.k - pad_left - in.L;  // includes the extra padding of strided convs
  const int max_pad = pad_left > pad_right ? pad_left : pad_right;
  p.Lext = (reflect && in.L <= max_pad) ? max_pad + 1 : 0;  // encodec pad1d: zero-extend a too-short signal first
  p.B = B_;
  p.Lm = o.L;
  p.nphase = 1;
  p.out_stride = 1;
  p.Lout = o.Lstore;
  p.Cout = W.cout;
  p.bias = W.bias;
  p.out = o.ptr;
  p.stats_out = o.stats;
  p.FGo = fgo;
  p.stat_slots = slots;
  if (!ck(launch_conv(p, st), "codec conv launch")) ok_ = false;
  ++launches_;
  return o;
}

CodecDecoder::Act CodecDecoder::convtr(const Act& in, const Act* in2, int act, const ConvW& W, int r, cudaStream_t st) {
  // out[m*r + z][n] = W[z] x[m] + W[z + r] x[m - 1], m in [0, L]: (L + 1) * r untrimmed rows; trim r - r/2 left, r/2 right.
  // Row m of the [L + 1][r * Cout] matrix of a k = 2 conv with r * Cout "virtual" output channels (z, n) IS rows
  // m*r .. m*r + r - 1 of the output: one tap-GEMM with full-width N tiles, the input transformed once for all phases.
  const int cv = r * W.cout;
  Act o;
  o.C = W.cout;
  o.Lstore = (in.L + 1) * r;
  o.row0 = r - r / 2;
  o.L = in.L * r;
  const int fgo = cv >= 64 ? cv / 64 : 1;
  const int slots = 32 / fgo > 1 ? 32 / fgo : 1;
  o.FG = fgo * slots;
  o.gamma = W.gamma;
  o.beta = W.beta;
  o.ptr = falloc((size_t)B_ * o.Lstore * o.C);
  o.stats = salloc(B_, o.FG);
  if (dry_) return o;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  fill_src(p, in, in2, act);
  ConvSeg& S = p.seg[0];
  S.w = W.w;
  S.wT = W.wT;  // packed [2][Cin][r * Cout] (make_conv, transposed)
  S.ntaps = 2;
  S.in_stride = 1;
  S.shift0 = 0;
  S.shift_step = -1;
  S.wtap0 = 0;
  S.wtap_phase = 0;
  S.wtap_step = 1;
  p.pad_mode = PAD_ZERO;
  p.B = B_;
  p.Lm = in.L + 1;
  p.nphase = 1;
  p.out_stride = 1;
  p.Lout = in.L + 1;
  p.Cout = cv;
  p.bias = W.bias;
  p.out = o.ptr;
  p.stats_out = o.stats;
  p.FGo = fgo;
  p.stat_slots = slots;
  if (!ck(launch_conv(p, st), "codec convtr launch")) ok_ = false;
  ++launches_;
  return o;
}

// the narrow last conv (ELU prologue, reflect padding); falls back to the tap-GEMM for shapes the narrow kernel does not take
CodecDecoder::Act CodecDecoder::last_conv(const Act& in, const Act* in2, const ConvW& W, int pad_left, cudaStream_t st) {
  const int max_pad = pad_left > (W.k - 1 - pad_left) ? pad_left : (W.k - 1 - pad_left);
  const size_t smem = ((size_t)(NR_ROWS + W.k - 1) * (in.C + 1) + (size_t)W.k * in.C * W.cout + 3 * (size_t)in.C) * sizeof(float);
  if (W.cout > 4 || in.L <= max_pad || smem > 46 * 1024) return conv(in, in2, ACT_ELU, W, pad_left, true, true, st);
  Act o;
  o.C = W.cout;
  o.L = o.Lstore = in.L;
  o.FG = 32;
  o.gamma = W.gamma;
  o.beta = W.beta;
  o.ptr = falloc((size_t)B_ * o.L * o.C);
  o.stats = salloc(B_, o.FG);
  if (dry_) return o;
  NarrowParams P;
  memset(&P, 0, sizeof(P));
  P.a = in.ptr;
  P.sa = in.stats;
  P.FGa = in.FG;
  P.ga = in.gamma;
  P.ba = in.beta;
  if (in2) {
    P.b = in2->ptr;
    P.sb = in2->stats;
    P.FGb = in2->FG;
    P.gb = in2->gamma;
    P.bb = in2->beta;
  }
  P.w = W.w;
  P.bias = W.bias;
  P.out = o.ptr;
  P.stats_out = o.stats;
  P.L = in.L;
  P.Lstore = in.Lstore;
  P.row0 = in.row0;
  P.Cin = in.C;
  P.Cout = W.cout;
  P.k = W.k;
  P.pad_left = pad_left;
  P.slots = o.FG;
  P.eps = d_.eps;
  dim3 grid((unsigned)((in.L + NR_ROWS - 1) / NR_ROWS), (unsigned)B_);
  narrow_conv_kernel<<<grid, NR_ROWS, smem, st>>>(P);
  if (!ck(cudaGetLastError(), "narrow conv launch")) ok_ = false;
  ++launches_;
  return o;
}

// prologue of seg 0: value(in) [+ value(in2)], then `act`.  value(x) = GroupNorm(1, C)(x.raw) when x carries statistics
void CodecDecoder::fill_src(ConvParams& p, const Act& in, const Act* in2, int act) {
  ConvSeg& S = p.seg[0];
  p.nseg = 1;
  p.mode = PRO_AFFINE;
  p.act = act;
  p.eps = d_.eps;
  S.Cin = in.C;
  S.L = in.L;
  S.Lstore = in.Lstore;
  S.row0 = in.row0;
  S.s[0].ptr = in.ptr;
  S.s[0].stats = in.stats;
  S.s[0].C = in.C;
  S.s[0].FG = in.FG;
  S.s[0].bmod = B_;
  S.s[0].scale = 1.0f;
  p.gamma = in.gamma;
  p.beta = in.beta;
  if (in2) {
    p.sum2 = 1;
    S.s[1].ptr = in2->ptr;
    S.s[1].stats = in2->stats;
    S.s[1].C = in2->C;
    S.s[1].FG = in2->FG;
    S.s[1].bmod = B_;
    S.s[1].scale = 1.0f;
    p.gamma2 = in2->gamma;
    p.beta2 = in2->beta;
  } else {
    p.G = in.stats ? 1 : 0;
  }
}

// One [t0, t1) launch of the tensor-core LSTM cluster kernel for layer `l`
bool CodecDecoder::launch_lstm_tc(int l, const float* gx, int gxT, int gx_t0, float* hout, float* cstate, int t0, int t1, int B,
                                  int T, cudaStream_t st) {
  LstmTcFn tc = lstm_tc_fn(H_, U_);
  LstmTcParams Q;
  Q.gx = gx;
  Q.whh = lstm_[l].whh_f32;
  Q.hout = hout;
  Q.cstate = cstate;
  Q.T = T;
  Q.t0 = t0;
  Q.t1 = t1;
  Q.gxT = gxT;
  Q.gx_t0 = gx_t0;
  Q.CS = CS_;
  Q.U = U_;
  Q.B = B;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(((B + 7) / 8) * CS_));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = lstm_tc_smem(H_, U_);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS_;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CS_ > 8) cudaFuncSetAttribute(tc, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const bool good = ck(cudaLaunchKernelEx(&cfg, tc, Q), "lstm (tensor core) launch");
  ++launches_;
  ++lstm_tc_launches_;
  return good;
}

// The LSTM layers over y0 (normalised by the consumer).  Returns the last layer's hidden sequence [B][T][H].
// Two layers on the tensor-core kernel and a long sequence: the sequence is cut into chunks and layer 2 runs one chunk
// behind layer 1 on a second stream (its input projection per chunk in between), so the two recurrences overlap.
CodecDecoder::Act CodecDecoder::lstm_stack(const Act& y0, int B, int T, cudaStream_t st) {
  const bool tc = !lstm_smem_kernel_ && lstm_tc_fn(H_, U_) != nullptr;
  const int nchunk = (tc && lstm_.size() == 2 && T >= 1024 && side_ && !lstm_no_overlap_) ? kLstmChunks : 1;
  auto proj = [&](int l, const Act& in, cudaStream_t s) {
    ConvW W;
    W.w = lstm_[l].wih;
    W.wT = lstm_[l].wih_T;
    W.bias = lstm_[l].bias;
    W.gamma = W.beta = nullptr;
    W.cin = H_;
    W.cout = 4 * H_;
    W.k = 1;
    return conv(in, nullptr, ACT_NONE, W, 0, false, false, s);
  };
  auto seq = [&]() {
    Act h;
    h.C = H_;
    h.L = h.Lstore = T;
    h.ptr = falloc((size_t)B * T * H_);
    return h;
  };
  if (nchunk > 1) {
    Act gx1 = proj(0, y0, st);
    Act h1 = seq(), h2 = seq();
    float* c1 = falloc((size_t)B * H_);
    float* c2 = falloc((size_t)B * H_);
    for (int c = 0; c < nchunk; ++c) {
      const int t0 = (int)((long long)T * c / nchunk), t1 = (int)((long long)T * (c + 1) / nchunk);
      if (!dry_) {
        if (!launch_lstm_tc(0, gx1.ptr, T, 0, h1.ptr, c1, t0, t1, B, T, st)) ok_ = false;
        cudaEventRecord(ev_[c], st);
        cudaStreamWaitEvent(side_, ev_[c], 0);
      }
      Act win = h1;  // rows [t0, t1) of layer 1's output
      win.row0 = t0;
      win.L = t1 - t0;
      Act gx2 = proj(1, win, side_);
      if (!dry_ && !launch_lstm_tc(1, gx2.ptr, t1 - t0, t0, h2.ptr, c2, t0, t1, B, T, side_)) ok_ = false;
    }
    if (!dry_) {
      cudaEventRecord(ev_[nchunk], side_);
      cudaStreamWaitEvent(st, ev_[nchunk], 0);
    }
    return h2;
  }
  Act cur = y0;
  for (size_t l = 0; l < lstm_.size(); ++l) {
    Act gx = proj((int)l, cur, st);
    Act h = seq();
    float* cs = falloc((size_t)B * H_);
    if (!dry_) {
      if (tc) {
        if (!launch_lstm_tc((int)l, gx.ptr, T, 0, h.ptr, cs, 0, T, B, T, st)) ok_ = false;
      } else {
        LstmParams P;
        P.gx = gx.ptr;
        P.whh = lstm_[l].whh;
        P.hout = h.ptr;
        P.T = T;
        P.H = H_;
        P.CS = CS_;
        P.U = U_;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(B * CS_));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = lstm_smem();
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CS_;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // (function attributes are per process: another decoder instance may have set a smaller limit)
        cudaFuncSetAttribute(lstm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem());
        if (!ck(cudaLaunchKernelEx(&cfg, lstm_cluster_kernel, P), "lstm launch")) ok_ = false;
        ++launches_;
      }
    }
    cur = h;
  }
  return cur;
}

void CodecDecoder::walk(const float* latent, float* audio, int B, int T, cudaStream_t st) {
  B_ = B;
  // ---- latent [B][D][T] -> channels-last
  Act x;
  x.C = d_.dimension;
  x.L = x.Lstore = T;
  x.ptr = falloc((size_t)B * T * x.C);
  if (!dry_) {
    if (!ck(launch_pack_ncl<float>(latent, x.ptr, nullptr, B, x.C, x.C, T, st), "codec pack")) ok_ = false;
    ++launches_;
  }
  const int k0 = d_.kernel_size;
  Act y0 = conv(x, nullptr, ACT_NONE, first_, (k0 - 1) - (k0 - 1) / 2, true, true, st);
  // ---- SLSTM: value = lstm(GN(y0)) + GN(y0); the sum is taken by the next layer's prologue
  Act hseq;
  const bool have_h = !lstm_.empty();
  if (have_h) hseq = lstm_stack(y0, B, T, st);
  // ---- upsampling stages
  Act a = y0;               // first operand of the running value
  Act a2 = hseq;            // optional second operand (value = GN(a) + value(a2))
  bool two = have_h;
  for (const Stage& S : stages_) {
    Act up = convtr(a, two ? &a2 : nullptr, ACT_ELU, S.up, S.ratio, st);
    const int kr = S.res1.k;
    Act r1 = conv(up, nullptr, ACT_ELU, S.res1, (kr - 1) - (kr - 1) / 2, true, true, st);
    Act r2 = conv(r1, nullptr, ACT_ELU, S.res2, 0, true, true, st);
    Act sc = conv(up, nullptr, ACT_NONE, S.shortcut, 0, true, true, st);
    a = sc;
    a2 = r2;
    two = true;
  }
  const int kl = d_.last_kernel_size;
  Act fin = last_conv(a, two ? &a2 : nullptr, last_, (kl - 1) - (kl - 1) / 2, st);
  if (!dry_) {
    dim3 grid((unsigned)std::min<long long>(((long long)fin.L + 255) / 256, 4096), (unsigned)B);
    final_norm_kernel<<<grid, 256, 0, st>>>(fin.ptr, fin.stats, fin.FG, fin.gamma, fin.beta, d_.eps, fin.C, fin.L, audio);
    if (!ck(cudaGetLastError(), "final norm launch")) ok_ = false;
    ++launches_;
  }
}

// ---- encoder walk: audio [N][channels][L] -> raw last conv (+ statistics) -> normalised latent (NCL) -> residual VQ
void CodecDecoder::walk_encoder(const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L, cudaStream_t st) {
  B_ = N;
  Act x;
  x.C = cin0_;
  x.L = x.Lstore = L;
  x.ptr = falloc((size_t)N * L * x.C);
  if (!dry_) {
    if (!ck(launch_pack_ncl<float>(audio, x.ptr, nullptr, N, d_.channels, cin0_, L, st), "codec pack")) ok_ = false;
    ++launches_;
  }
  const int k0 = d_.kernel_size;
  Act v = conv(x, nullptr, ACT_NONE, first_, (k0 - 1) - (k0 - 1) / 2, true, true, st);
  for (const Stage& S : stages_) {
    const int kr = S.res1.k, r = S.ratio;
    Act r1 = conv(v, nullptr, ACT_ELU, S.res1, (kr - 1) - (kr - 1) / 2, true, true, st);
    Act r2 = conv(r1, nullptr, ACT_ELU, S.res2, 0, true, true, st);
    Act sc = conv(v, nullptr, ACT_NONE, S.shortcut, 0, true, true, st);
    // ELU(shortcut + block) -> strided conv k = 2r: padding total r, right r/2 (+ extra), left r - r/2
    v = conv(sc, &r2, ACT_ELU, S.up, r - r / 2, true, true, st, r);
  }
  Act a = v, a2;
  bool two = false;
  if (!lstm_.empty()) {
    a2 = lstm_stack(v, N, v.L, st);
    two = true;
  }
  const int kl = d_.last_kernel_size;
  Act fin = conv(a, two ? &a2 : nullptr, ACT_ELU, last_, (kl - 1) - (kl - 1) / 2, true, true, st);
  if (!dry_) {
    dim3 grid((unsigned)std::min<long long>(((long long)fin.L + 255) / 256, 4096), (unsigned)N);
    final_norm_kernel<<<grid, 256, 0, st>>>(fin.ptr, fin.stats, fin.FG, fin.gamma, fin.beta, d_.eps, fin.C, fin.L, latent);
    if (!ck(cudaGetLastError(), "final norm launch")) ok_ = false;
    ++launches_;
    if ((codes || quantized) && quantize(latent, codes, quantized, N, fin.L, st)) ok_ = false;
  }
}

int CodecDecoder::quantize(const float* latent, int32_t* codes, float* quantized, int N, int T, cudaStream_t st) {
  if (!finalized_ || !encoder_) return fail("quantize needs a finalized encoder engine");
  if (!latent || N < 1 || T < 1) return fail("quantize: bad arguments");
  cudaSetDevice(device_);
  RvqParams P;
  P.emb = latent;
  P.codebooks = codebooks_;
  P.enorm = enorm_;
  P.quant = quantized;
  P.codes = codes;
  P.N = N;
  P.T = T;
  P.D = d_.dimension;
  P.K = K_;
  P.nq = n_q_;
  const size_t smem = ((size_t)2 * P.D * RQ_F + (size_t)P.D * (RQ_C + 1) + 16 * RQ_F) * sizeof(float) + (16 * RQ_F + RQ_F) * sizeof(int);
  cudaFuncSetAttribute(rvq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 g2((unsigned)((T + RQ_F - 1) / RQ_F), (unsigned)N);
  rvq_kernel<<<g2, 256, smem, st>>>(P);
  ++launches_;
  return ck(cudaGetLastError(), "rvq launch") ? 0 : 1;
}

size_t CodecDecoder::encode_workspace_bytes(int N, int L) {
  dry_ = true;
  off_ = 0;
  soff_ = 0;
  walk_encoder(nullptr, nullptr, nullptr, nullptr, N, L, nullptr);
  dry_ = false;
  stats_bytes_need_ = soff_;
  return off_ + soff_ + 4096;
}

int CodecDecoder::encode(const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L, cudaStream_t st) {
  if (!finalized_ || !encoder_) return fail("encode needs a finalized encoder engine");
  if (!audio || !latent || N < 1 || L < 1) return fail("encode: bad arguments");
  cudaSetDevice(device_);
  const size_t need = encode_workspace_bytes(N, L);
  if (need > arena_bytes_) {
    if (arena_) cudaFree(arena_);
    arena_ = nullptr;
    arena_bytes_ = 0;
    if (!ck(cudaMalloc((void**)&arena_, need), "cudaMalloc(codec workspace)")) return 1;
    arena_bytes_ = need;
  }
  ok_ = true;
  off_ = 0;
  const size_t act_bytes = arena_bytes_ - stats_bytes_need_ - 2048;
  soff_ = act_bytes & ~(size_t)255;
  if (!ck(cudaMemsetAsync(arena_ + soff_, 0, stats_bytes_need_, st), "codec statistics memset")) return 1;
  walk_encoder(audio, latent, codes, quantized, N, L, st);
  return ok_ ? 0 : 1;
}

int CodecDecoder::decode(const float* latent, float* audio, int B, int T, cudaStream_t st) {
  if (encoder_) return fail("decode called on an encoder engine");
  if (!finalized_) return fail("decode before finalize");
  if (!latent || !audio || B < 1 || T < 1) return fail("decode: bad arguments");
  cudaSetDevice(device_);
  if (reserve(B, T)) return 1;
  ok_ = true;
  off_ = 0;
  // statistics zone at the tail of the arena: one memset per decode
  const size_t act_bytes = arena_bytes_ - stats_bytes_need_ - 2048;
  soff_ = act_bytes & ~(size_t)255;
  if (!ck(cudaMemsetAsync(arena_ + soff_, 0, stats_bytes_need_, st), "codec statistics memset")) return 1;
  walk(latent, audio, B, T, st);
  return ok_ ? 0 : 1;
}

}  // namespace jen1
