// tcgen05 / TMEM attention core for bf16 storage on sm_100a (reference jen1/model/blocks.py:355-380 with the
// key/value masking of :431-434): softmax(q k^T * scale) v for one (batch row, head, 128-query tile) per CTA.
//
//   S[query, key] = Q K^T     tcgen05.mma, M = 128 queries (TMEM lanes), N = keys padded to 16 (<= 256 TMEM columns),
//                             K = head dim; Q and K staged K-major in 128-byte-swizzled shared-memory tiles
//   P = softmax(S * scale)    warp-specialised: the four softmax warps own one query row per thread (TMEM lane ==
//                             thread), two passes over the TMEM row (max, then exp / sum), fp32; P is written bf16
//                             into its own K-major swizzled tile
//   O = P V                   tcgen05.mma, M = 128 queries, N = head dim, K = keys; V is consumed as an MN-major
//                             operand straight from its [key][channel] layout (no transpose), O re-uses S's columns
//   out = O / rowsum(P)       tcgen05.ld -> registers -> 16-byte bf16 stores
//
// Padded context keys are multiplied by the context mask (logit 0, value 0) and STAY in the softmax, exactly as the
// reference does; causal masking uses -FLT_MAX like reference add_mask (blocks.py:304-312); keys beyond the real
// key count are excluded (-inf).  Cross-attention keys/values come from the hoisted caches: per sample either the
// prompt rows + the per-step time-token row, or (cond-dropout / unconditional CFG half) the learned null embedding.
//
// Warp roles (256 threads): all 8 warps stage Q / K / V (16-byte loads, 8 in flight per thread); warps 0-3 then
// run softmax + epilogue; warp 4 allocates TMEM and its lane 0 issues the MMAs.  PDL-aware: the kernel signals its
// dependents right after TMEM allocation and waits for its producer before the first activation load.
#include <float.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

namespace {

constexpr int kAtThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1).  For a K-major operand
// the leading-dimension field is unused; for an MN-major operand it is the stride between 64-element MN blocks and
// the stride field is the distance between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct AttnGeom {
  int KP;         // keys padded to a multiple of 16 (MMA N of S, MMA K extent of PV)
  int KR;         // key rows of the K / V tiles (KP rounded up to 8)
  int DB;         // 64-channel blocks of the head dim
  int cpr_shift;  // log2(head dim / 8): 16-byte chunks per head row
  int tmem_cols;
  uint32_t off_k, off_v, off_p;  // byte offsets of the tiles (Q at 0)
  uint32_t smem;
};

__global__ void __launch_bounds__(kAtThreads, 1) attn_umma_kernel(const __grid_constant__ AttnParams p,
                                                                   const __grid_constant__ AttnGeom g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * 128, h = blockIdx.y, r = blockIdx.z;
  const int d = p.d;
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + g.off_k;
  uint8_t* Vs = smem + g.off_v;
  uint8_t* Ps = smem + g.off_p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.smem - 64);  // in_full, s_full, p_full, o_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (tid == 128) {  // the MMA-issuing thread owns the barriers (it is the first to wait on them)
    mbar_init(&bars[0], kAtThreads);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 128);
    mbar_init(&bars[3], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- per-CTA key/value source (values written before this step's kernel chain: safe ahead of the wait)
  const int bc = r % p.Bc;
  const bool fixed = p.cross && ((r >= p.Bc) || (p.drop && p.drop[bc]));
  const int S = p.M - 1;
  const int trow_time = p.cross ? p.cond_row[r] : 0;
  const int cpr = 1 << g.cpr_shift;

  // ---- stage Q, K, V with cp.async (16-byte chunks, rolled loops, everything in flight at once, no registers -- this
  //      code runs once per CTA, so its instruction footprint matters more than its instruction count).  Cross-attention
  //      keys / values come from per-step constant caches, so they are requested BEFORE the PDL wait (while the producer
  //      of Q is still running); only what the previous kernel wrote is loaded after it.  Query rows beyond N are never
  //      stored (an MMA row only feeds its own output row); key rows beyond M are zero-filled (P is 0 there and
  //      0 * garbage could be NaN).
  const int sh = g.cpr_shift;
  auto cp16 = [&](uint8_t* dst, const bf16* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
  };
  auto stage_kv = [&]() {
    for (int i = tid; i < (g.KR << (sh + 1)); i += kAtThreads) {
      const int v = i >= (g.KR << sh) ? 1 : 0;  // 0: K tile, 1: V tile
      const int j = i - (v ? (g.KR << sh) : 0);
      const int rw = j >> sh, part = j & (cpr - 1);
      uint8_t* dst = (v ? Vs : Ks) + ((uint32_t)(part >> 3) * (uint32_t)g.KR + (uint32_t)rw) * 128u + (uint32_t)(((part & 7) ^ (rw & 7)) * 16);
      if (rw >= p.M) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      const bf16* sp;
      if (!p.cross) {
        sp = (const bf16*)p.kv + ((size_t)r * p.N + rw) * p.kv_ld + (v ? p.v_off : p.k_off);
      } else {
        const bf16* rowp;
        if (rw < S)
          rowp = fixed ? (const bf16*)p.kv_fixed + (size_t)rw * p.kvc_ld : (const bf16*)p.kv_cond + ((size_t)bc * S + rw) * p.kvc_ld;
        else
          rowp = fixed ? (const bf16*)p.kv_fixed + (size_t)S * p.kvc_ld : (const bf16*)p.kv_time + (size_t)trow_time * p.kvc_ld;
        sp = rowp + p.kvc_off + (v ? p.C : 0);
      }
      cp16(dst, sp + h * d + part * 8);
    }
  };
  // chunks beyond the head dim inside the last 64-channel block must read as zero (d = 16 / 32)
  if (cpr < 8) {
    const int rows_all = 128 + 2 * g.KR;
    const int zc = 8 - cpr;
    for (int it = tid; it < rows_all * zc; it += kAtThreads) {
      const int row_all = it / zc, ch = cpr + (it - row_all * zc);
      uint8_t* tile = row_all < 128 ? Qs : (row_all < 128 + g.KR ? Ks : Vs);
      const int row = row_all < 128 ? row_all : (row_all < 128 + g.KR ? row_all - 128 : row_all - 128 - g.KR);
      *reinterpret_cast<uint4*>(tile + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) * 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (p.cross) stage_kv();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!p.cross) stage_kv();
  {
    const int nqr = min(128, p.N - i0);
    const bf16* qsrc = (const bf16*)p.q + ((size_t)r * p.N + i0) * p.q_ld + p.q_off + h * d;
    for (int i = tid; i < (nqr << sh); i += kAtThreads) {
      const int rw = i >> sh, part = i & (cpr - 1);
      cp16(Qs + ((uint32_t)(part >> 3) * 128u + (uint32_t)rw) * 128u + (uint32_t)(((part & 7) ^ (rw & 7)) * 16),
           qsrc + (size_t)rw * p.q_ld + part * 8);
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
  if (p.cross && p.mask) {
    // masked context keys: K and V rows times the mask (logit 0, value 0 -- reference blocks.py:431-434); every thread
    // revisits exactly the chunks it requested itself
    for (int i = tid; i < (g.KR << (sh + 1)); i += kAtThreads) {
      const int v = i >= (g.KR << sh) ? 1 : 0;
      const int j = i - (v ? (g.KR << sh) : 0);
      const int rw = j >> sh, part = j & (cpr - 1);
      if (rw >= S) continue;
      const float mk = __ldg(p.mask + (size_t)bc * S + rw);
      if (mk == 1.0f) continue;
      uint4* cell = reinterpret_cast<uint4*>((v ? Vs : Ks) + ((uint32_t)(part >> 3) * (uint32_t)g.KR + (uint32_t)rw) * 128u +
                                             (uint32_t)(((part & 7) ^ (rw & 7)) * 16));
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(cell);
      uint4 o;
      o.x = pack2(__low2float(hh[0]) * mk, __high2float(hh[0]) * mk);
      o.y = pack2(__low2float(hh[1]) * mk, __high2float(hh[1]) * mk);
      o.z = pack2(__low2float(hh[2]) * mk, __high2float(hh[2]) * mk);
      o.w = pack2(__low2float(hh[3]) * mk, __high2float(hh[3]) * mk);
      *cell = o;
    }
  }
  fence_async_smem();
  mbar_arrive(&bars[0]);

  if (warp == 4) {
    // ======================================================================== MMA issuer
    // (warp-uniform control flow, one elected lane issues, descriptors built once: under a divergent `lane == 0` branch
    //  every tcgen05 instruction is wrapped in an elect-and-branch loop -- ~85 clocks per MMA, see attn_flash.cu)
    // instruction descriptor: fp32 accumulate, bf16 A/B, M = 128; bit 16 = B operand is MN-major
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.KP >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(d >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t qd = make_desc_sw128(smem_u32(Qs), 16u, 1024u), kd = make_desc_sw128(smem_u32(Ks), 16u, 1024u);
    const uint64_t pd = make_desc_sw128(smem_u32(Ps), 16u, 1024u), vd = make_desc_sw128(smem_u32(Vs), (uint32_t)g.KR * 128u, 1024u);
    const uint32_t kblk16 = (uint32_t)g.KR * 8u;  // one 64-channel block of K rows in descriptor address units (16 bytes)
    const int nk = d >> 4;
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)  // 64-channel block kk >> 2, 32 bytes per K step inside the swizzle atom
        if (kk < nk)
          umma_bf16(tmem_base, qd + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2), kd + (uint64_t)((uint32_t)(kk >> 2) * kblk16 + (uint32_t)(kk & 3) * 2u),
                    idesc_s, kk > 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    __syncwarp();
    mbar_wait(&bars[2], 0);
    tc_fence_after();
    if (elect_one()) {
      const int n16 = g.KP >> 4;
#pragma unroll 4
      for (int k16 = 0; k16 < n16; ++k16)
        umma_bf16(tmem_base, pd + (uint64_t)((k16 >> 2) * 1024 + (k16 & 3) * 2), vd + (uint64_t)(k16 * 128), idesc_o, k16 > 0 ? 1u : 0u);
      umma_commit(&bars[3]);
    }
    __syncwarp();
  } else if (warp < 4) {
    // ======================================================================== softmax + epilogue (thread == query row)
    const int i = i0 + tid;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sc = p.scale * 1.4426950408889634f;  // softmax in base 2
    const int jmax = p.causal ? i + (p.M - p.N) : p.M - 1;  // last key this query may see
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = 0; c0 < g.KP; c0 += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int key = c0 + j;
        float s = v[j] * sc;
        if (key > jmax) s = -FLT_MAX;
        if (key < p.M) mx = fmaxf(mx, s);
      }
    }
    float sum = 0.f;
    for (int c0 = 0; c0 < g.KP; c0 += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c0, v);
      float e[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int key = c0 + j;
        float s = v[j] * sc;
        if (key > jmax) s = -FLT_MAX;
        float pj = (key < p.M) ? exp2f(s - mx) : 0.0f;
        pj = bf16_round(pj);
        sum += pj;
        e[j] = pj;
      }
      uint8_t* prow = Ps + (size_t)(c0 >> 6) * 128 * 128 + (size_t)tid * 128;
      const int ch = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(prow + (((ch) ^ (tid & 7)) * 16)) =
          make_uint4(pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
      *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ (tid & 7)) * 16)) =
          make_uint4(pack2(e[8], e[9]), pack2(e[10], e[11]), pack2(e[12], e[13]), pack2(e[14], e[15]));
    }
    tc_fence_before();
    fence_async_smem();
    mbar_arrive(&bars[2]);
    const float inv = 1.0f / sum;
    mbar_wait(&bars[3], 0);
    tc_fence_after();
    // tcgen05.ld is warp-collective: every lane loads, only rows < N store
    bf16* orow = (bf16*)p.out + ((size_t)r * p.N + (i < p.N ? i : 0)) * p.C + h * d;
    for (int c0 = 0; c0 < d; c0 += 16) {
      float v[16];
      tmem_ld16(trow + (uint32_t)c0, v);
      if (i < p.N) {
        uint4 o0 = make_uint4(pack2(v[0] * inv, v[1] * inv), pack2(v[2] * inv, v[3] * inv), pack2(v[4] * inv, v[5] * inv),
                              pack2(v[6] * inv, v[7] * inv));
        uint4 o1 = make_uint4(pack2(v[8] * inv, v[9] * inv), pack2(v[10] * inv, v[11] * inv),
                              pack2(v[12] * inv, v[13] * inv), pack2(v[14] * inv, v[15] * inv));
        *reinterpret_cast<uint4*>(orow + c0) = o0;
        *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

}  // namespace

bool attn_umma_supported(const AttnParams& p) {
  const int d = p.d;
  if (!(d == 16 || d == 32 || d == 64 || d == 128)) return false;
  if (p.M < 1 || p.M > 256 || p.N < 1) return false;
  if ((p.q_ld & 7) || (p.q_off & 7) || (p.C & 7)) return false;
  if (!p.cross && ((p.kv_ld & 7) || (p.k_off & 7) || (p.v_off & 7))) return false;
  if (p.cross && ((p.kvc_ld & 7) || (p.kvc_off & 7))) return false;
  return true;
}

// Function attributes are per device: called once per engine from Engine::finalize() after cudaSetDevice (a process
// may hold one engine per GPU).
cudaError_t attn_umma_init() {
  return cudaFuncSetAttribute(attn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_attention_umma(const AttnParams& p, bool pdl, cudaStream_t stream) {
  if (!attn_umma_supported(p)) return cudaErrorInvalidValue;
  AttnGeom g;
  memset(&g, 0, sizeof(g));
  g.KP = (p.M + 15) / 16 * 16;
  g.KR = g.KP;  // multiple of 16, hence of 8
  g.DB = (p.d + 63) / 64;
  g.cpr_shift = p.d == 16 ? 1 : (p.d == 32 ? 2 : (p.d == 64 ? 3 : 4));
  int tc = 32;
  while (tc < g.KP || tc < p.d) tc <<= 1;
  g.tmem_cols = tc;
  const uint32_t q_bytes = (uint32_t)g.DB * 128u * 128u;
  const uint32_t k_bytes = (uint32_t)g.DB * (uint32_t)g.KR * 128u;
  const uint32_t p_bytes = (uint32_t)((g.KP + 63) / 64) * 128u * 128u;
  g.off_k = q_bytes;
  const uint32_t qk = (q_bytes + k_bytes + 1023u) / 1024u * 1024u;
  const uint32_t pq = (p_bytes + 1023u) / 1024u * 1024u;
  // P has its own tile (aliasing it over the dead Q / K tiles saves 48 KB the kernel does not need -- one CTA per SM --
  // and makes compute-sanitizer racecheck, which cannot see mbarrier / tcgen05.commit ordering, report the reuse)
  g.off_v = qk;
  g.off_p = g.off_v + (k_bytes + 1023u) / 1024u * 1024u;
  g.smem = g.off_p + pq + 64u;
  if (g.smem + 1024 > 227 * 1024) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((p.N + 127) / 128, p.H, p.B2);
  cfg.blockDim = dim3(kAtThreads);
  cfg.dynamicSmemBytes = g.smem + 1024;  // slack for the 1024-byte alignment of the dynamic window
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn_umma_kernel, p, g);
}

}  // namespace jen1
