// tcgen05 / TMEM implementation of the fused tap-GEMM operator (see conv_params.h) for bf16 storage on sm_100a.
//
// Formulation ("swap-AB"): for one output phase z the accumulator is D[cout][q] (cout on the 128 TMEM lanes,
// padded flat position q on the TMEM columns):
//
//     D[cout, q] = sum_{seg} sum_{tap j} sum_{cin}  W_seg[wtap(z,j)][cout, cin] * P_seg[q + off(j), cin]
//
//   * A operand = weights.  Pre-packed at load time into 16 KB blobs (128 cout x 64 cin, the canonical K-major
//     128-byte-swizzle layout) in exactly the order a CTA consumes them, so one thread streams them with
//     1-D bulk async copies (cp.async.bulk -> UBLKCP) through an mbarrier ring -- and starts doing so BEFORE the
//     programmatic-dependent-launch wait, i.e. while the previous layer is still running.
//   * B operand = the activation panel.  Producer warps read raw channels-last bf16 rows once, apply the
//     GroupNorm-apply / FiLM / SiLU (or LayerNorm, or skip-scale) prologue in fp32 registers, re-zero the conv
//     padding rows AFTER the activation (reference blocks.py:137-145 -> :44-51) and store bf16 into a
//     "row panel": for each 8-channel chunk a column of 16-byte rows.  In that layout (SBO = 128 B, LBO = panel
//     stride) a conv tap is nothing but a 16-byte-granular shift of the descriptor start address, so the k taps
//     reuse one panel; strided down-convs keep one sub-panel per residue (row mod stride).
//   * Batch rows are folded into the position axis with a per-row halo (q = b*Lq + m, Lq = Lm + halo), so the deep
//     UNet levels (L = 1..24) still fill an MMA N tile, and their weight streaming is spread over the whole GPU
//     by split-K: partial tiles go to an L2-resident workspace and the last-arriving CTA of a tile (atomic ticket)
//     reduces them in fixed order and runs the epilogue (bias / GELU / residual / GroupNorm + LayerNorm partial
//     statistics for the next consumer).
//
// Warp roles (192 threads): warps 0-3 build panels, then run the epilogue (TMEM lane quarter = warp index);
// warp 4 lane 0 streams weights; warp 5 allocates TMEM and its lane 0 issues tcgen05.mma.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

namespace {

constexpr int kThreads = 192;
constexpr int kProducers = 128;
constexpr int kABytes = 128 * 64 * 2;  // one weight blob
constexpr int kMaxSlots = 16;          // distinct batch rows one N tile may touch

struct UmmaArgs {
  ConvParams p;
  UmmaPlan pl;
  const bf16* w0;  // packed blobs of seg 0: [m_tile][phase][cin block][tap]
  const bf16* w1;  // packed blobs of seg 1: [m_tile][cin block]
  float* ws;       // split-K partial tiles
  int* counters;   // split-K tickets (zero between launches)
  int out_f32;
  long long* timeline;  // optional per-launch phase clocks of CTA (0,0,0) (JEN1_TIMELINE debugging), else nullptr
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Pure polling (test_wait): try_wait's hardware suspend was measured to wake ~1.5 us late when the phase is
// completed by tcgen05.commit / bulk-copy transactions, which is as long as a whole layer of the deep UNet levels.
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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

// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = stride between 8-row groups, LBO = stride
// between the two 8-element K chunks of one K=16 instruction.  (cute::UMMA::SmemDescriptor, version 1.)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// K-major, 128-byte swizzle: rows of 128 B (64 bf16), the 16-byte chunk c of row r lives at chunk c ^ (r & 7);
// SBO = 1024 B between 8-row groups; K advances inside the swizzle atom by adding bytes to the start address.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TapGeom {
  int rho, off;
};
__device__ __forceinline__ TapGeom tap_geom(int shift0, int shift_step, int j, int f, int amin) {
  const int d = shift0 + j * shift_step;
  int rho = d % f;
  if (rho < 0) rho += f;
  const int a = (d - rho) / f;
  TapGeom g;
  g.rho = rho;
  g.off = a - amin;
  return g;
}

__device__ __forceinline__ float ldf_cg(const bf16* p) {
  const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 2) conv_umma_kernel(const __grid_constant__ UmmaArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams& p = A.p;
  const UmmaPlan& pl = A.pl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* tl = (A.timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? A.timeline : nullptr;
#define TL_MARK(i) do { if (tl) tl[i] = clock64(); } while (0)
#define TL_GLOBAL(i) do { if (tl) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); tl[i] = (long long)g_; } } while (0)
  if (tid == 0) TL_MARK(0);

  // ---- work item
  const int nt = blockIdx.x, mt = blockIdx.y;
  const int z = blockIdx.z / pl.splitk, sk = blockIdx.z % pl.splitk;
  const int NT = pl.NT, Lq = pl.Lq;
  const int q0 = nt * NT;
  const int nsteps = pl.steps0 + pl.steps1;
  const int st0 = (int)(((long long)sk * nsteps) / pl.splitk);
  const int st1 = (int)(((long long)(sk + 1) * nsteps) / pl.splitk);
  const int my_steps = st1 - st0;
  const int ntaps0 = p.seg[0].ntaps;

  // ---- shared memory carve-up
  uint8_t* a_ring = smem;                                           // stages * 16 KB
  uint8_t* panels = a_ring + (size_t)pl.stages * kABytes;          // 2 * panel_bytes
  const uint32_t panel_bytes = (uint32_t)pl.panel_bytes;
  uint8_t* misc = panels + 2 * (size_t)panel_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);            // [8]
  uint64_t* a_empty = a_full + 8;                                   // [8]
  uint64_t* p_full = a_empty + 8;                                   // [2]
  uint64_t* p_empty = p_full + 2;                                   // [2]
  uint64_t* acc_full = p_empty + 2;                                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);  // [1]
  int* ticket_slot = reinterpret_cast<int*>(tmem_slot + 1);         // [1]
  float* gmean = reinterpret_cast<float*>(misc + 256);              // [kMaxSlots][32]
  float* grstd = gmean + kMaxSlots * 32;                            // [kMaxSlots][32]
  double* fine = reinterpret_cast<double*>(grstd + kMaxSlots * 32); // [kMaxSlots][2 sources][32 fine groups][2]
  int* scrow = reinterpret_cast<int*>(fine + kMaxSlots * 2 * 32 * 2);  // [kMaxSlots] conditioning-table row per batch row
  // epilogue scratch aliases the weight ring (all MMAs have completed by then)
  float* sred = reinterpret_cast<float*>(a_ring);                   // [kMaxSlots][128][2]
  float* rowred = sred + kMaxSlots * 128 * 2;                       // [4][NT][2]

  if (tid == kProducers) {  // warp 4 lane 0
    for (int i = 0; i < pl.stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(&p_full[0], kProducers);
    mbar_init(&p_full[1], kProducers);
    mbar_init(&p_empty[0], 1);
    mbar_init(&p_empty[1], 1);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)pl.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM is held: dependents may now be scheduled next to us without a TMEM-allocation deadlock.
  pdl_launch_dependents();
  if (tid == 0) TL_MARK(1);

  if (warp == 4) {
    // ======================================================================== weight streamer
    if (lane == 0) {
      int i = 0;
      for (int t = st0; t < st1; ++t) {
        const bool s1 = t >= pl.steps0;
        const int ntp = s1 ? 1 : ntaps0;
        const bf16* src = s1 ? A.w1 + ((size_t)mt * pl.steps1 + (t - pl.steps0)) * (kABytes / 2)
                             : A.w0 + (((size_t)(mt * p.nphase + z) * pl.steps0 + t) * ntaps0) * (kABytes / 2);
        for (int j = 0; j < ntp; ++j, ++i) {
          const int s = i % pl.stages, k = i / pl.stages;
          if (k > 0) mbar_wait(&a_empty[s], (uint32_t)((k - 1) & 1));
          mbar_expect_tx(&a_full[s], kABytes);
          bulk_g2s(a_ring + (size_t)s * kABytes, src + (size_t)j * (kABytes / 2), kABytes, &a_full[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ======================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t lbo_b = (uint32_t)pl.PS * 16u;
      int i = 0;
      uint32_t acc = 0;
      for (int t = st0; t < st1; ++t) {
        const int n = t - st0, pb = n & 1;
        const bool s1 = t >= pl.steps0;
        const ConvSeg& S = p.seg[s1 ? 1 : 0];
        const int ntp = s1 ? 1 : ntaps0;
        const int f = s1 ? 1 : S.in_stride;
        const int amin = s1 ? 0 : pl.amin;
        mbar_wait(&p_full[pb], (uint32_t)((n >> 1) & 1));
        tc_fence_after();
        const uint32_t pbase = smem_u32(panels + (size_t)pb * panel_bytes);
        for (int j = 0; j < ntp; ++j, ++i) {
          const int s = i % pl.stages, k = i / pl.stages;
          TapGeom g;
          if (s1) {
            g.rho = 0;
            g.off = 0;
          } else {
            g = tap_geom(S.shift0, S.shift_step, j, f, amin);
          }
          mbar_wait(&a_full[s], (uint32_t)(k & 1));
          tc_fence_after();
          if (i == 0) TL_MARK(9);
          const uint32_t abase = smem_u32(a_ring + (size_t)s * kABytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_desc_sw128(abase + (uint32_t)kk * 32u);
            uint64_t bd;
            if (pl.bsw == 0) {
              bd = make_desc(pbase + ((uint32_t)(g.rho * 8 + kk * 2) * (uint32_t)pl.PS + (uint32_t)g.off) * 16u, lbo_b, 128u);
            } else {  // swizzled panel: a tap is a whole-row (128 B) shift of the start address
              const uint32_t sa = pbase + ((uint32_t)g.rho * (uint32_t)pl.PS + (uint32_t)g.off) * 128u + (uint32_t)kk * 32u;
              bd = make_desc_sw128(sa, pl.bsw == 2 ? (sa >> 7) & 7u : 0u);
            }
            umma_bf16(tmem_base, ad, bd, idesc, acc);
            acc = 1;
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&p_empty[pb]);
      }
      umma_commit(acc_full);
      TL_MARK(10);
      if (tl) {
        mbar_wait(acc_full, 0);
        TL_MARK(13);
      }
    }
  } else {
    // ======================================================================== panel producers, then epilogue
    // Every load below is batched: addresses and validity are computed first, then all loads of a batch are issued
    // back to back, then consumed -- a dependent L2/HBM round trip costs ~0.4-1 us and the layer chain is long.
    const ConvSeg& S0 = p.seg[0];
    const int Ct = S0.Cin;
    const int b_first = q0 / Lq;
    int b_last = (q0 + NT + pl.halo - 1) / Lq;
    if (b_last > p.B - 1) b_last = p.B - 1;
    const int nbl = b_last - b_first + 1;
    const int kc = tid & 7, rr = tid >> 3;
    const bool affine = p.mode == PRO_AFFINE;
    const bool has_gn = affine && p.G > 0;
    const bool has_film = affine && p.film != nullptr;
    const int cpg = has_gn ? Ct / p.G : 1;
    const int cl = tid;             // channel within the 128-wide M tile (== TMEM lane)
    const int nch = mt * 128 + cl;  // output channel
    float gam[8], bet[8];
    auto load_gamma_beta = [&](int t) {  // GroupNorm affine of this thread's 8 channels in K step t (weights)
      const int c0 = t * 64 + kc * 8;
      if (has_gn && t < pl.steps0 && c0 < Ct) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + 4));
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.beta + c0));
        const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + 4));
        gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
        bet[0] = e0.x; bet[1] = e0.y; bet[2] = e0.z; bet[3] = e0.w; bet[4] = e1.x; bet[5] = e1.y; bet[6] = e1.z; bet[7] = e1.w;
      }
    };
    // constants that do not depend on earlier kernels are fetched before the PDL wait
    const float bias = p.bias ? __ldg(p.bias + nch) : 0.0f;
    load_gamma_beta(st0);

    pdl_wait();  // everything below reads what the previous kernels wrote
    if (tid == 0) { TL_MARK(2); TL_GLOBAL(11); }

    if (tid < nbl) scrow[tid] = p.cond_row ? __ldcg(p.cond_row + b_first + tid) : 0;
    // ---- GroupNorm statistics of the input for the batch rows this tile touches.  Deterministic two-level reduce:
    //      128 threads = 64 (source, fine group) items x 2 interleaved halves of the producer's per-tile partials,
    //      combined by one shuffle; then one thread per group folds its fine groups.
    if (has_gn) {
      // pass 1: fine-group sums of every (batch row, source, fine group) item; an item is split over `parts`
      // lanes (interleaved entries) that are combined with shuffles in a fixed order
      const int nitem = nbl * 64;
      int parts = 1;
      while (parts < 8 && nitem * parts * 2 <= kProducers) parts *= 2;
      for (int base = 0; base < nitem * parts; base += kProducers) {
        const int idx = base + tid;
        const int item = idx / parts, part = idx - item * parts;
        const int bl = item >> 6, fs = (item >> 5) & 1, ffg = item & 31;
        double a = 0.0, q = 0.0;
        if (item < nitem) {
          const ConvSrc& fsr = S0.s[fs];
          if (fsr.C > 0 && ffg < fsr.FG) {
            const int b = b_first + bl;
            const float2* st = reinterpret_cast<const float2*>(fsr.stats) + (size_t)(b % fsr.bmod) * fsr.n_ent * fsr.FG + ffg;
            for (int e0 = part; e0 < fsr.n_ent; e0 += parts * 16) {  // 16 independent L2 loads in flight
              float2 buf[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const int e = e0 + u * parts;
                buf[u] = e < fsr.n_ent ? __ldcg(st + (size_t)e * fsr.FG) : make_float2(0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                a += (double)buf[u].x;
                q += (double)buf[u].y;
              }
            }
          }
        }
        if (tid == 0 && base == 0) TL_MARK(14);
        for (int o = 1; o < parts; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (item < nitem && part == 0) {
          const double sc = (double)S0.s[fs].scale;
          fine[item * 2] = a * sc;
          fine[item * 2 + 1] = q * sc * sc;
        }
      }
      bar_sync_producers();
      if (tid == 0) TL_MARK(15);
      // pass 2: one thread per (batch row, group)
      for (int idx = tid; idx < nbl * p.G; idx += kProducers) {
        const int bl = idx / p.G, g = idx - bl * p.G;
        const int lo = g * cpg, hi = lo + cpg;
        double ga = 0.0, gq = 0.0;
        int off = 0;
        for (int s = 0; s < 2; ++s) {
          const ConvSrc& sr = S0.s[s];
          if (sr.C > 0) {
            const int olo = max(lo, off), ohi = min(hi, off + sr.C);
            if (ohi > olo) {
              const int gs = sr.C / sr.FG;
              for (int fg = (olo - off) / gs; fg < (ohi - off) / gs; ++fg) {
                ga += fine[((bl * 2 + s) * 32 + fg) * 2];
                gq += fine[((bl * 2 + s) * 32 + fg) * 2 + 1];
              }
            }
          }
          off += sr.C;
        }
        const double n = (double)(p.gn_real_c > 0 ? p.gn_real_c / p.G : cpg) * (double)S0.L;
        const double mean = ga / n;
        double var = gq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        gmean[bl * 32 + g] = (float)mean;
        grstd[bl * 32 + g] = rsqrtf((float)var + p.eps);
      }
      bar_sync_producers();
    } else if (has_film) {
      bar_sync_producers();  // scrow
    }

    if (tid == 0) TL_MARK(3);
    // ---- panels
    float ca[8], cs[8];
    int coef_b = -1;
    for (int t = st0; t < st1; ++t) {
      const int n = t - st0, pb = n & 1;
      const bool s1 = t >= pl.steps0;
      const ConvSeg& S = p.seg[s1 ? 1 : 0];
      const int f = s1 ? 1 : S.in_stride;
      const int amin = s1 ? 0 : pl.amin;
      const int R = s1 ? NT : pl.R;
      const int cb = s1 ? t - pl.steps0 : t;
      const int c0 = cb * 64 + kc * 8;  // channel in the concatenated input
      const bool second = c0 >= S.s[0].C;
      const ConvSrc& sr = second ? S.s[1] : S.s[0];
      const int cc = second ? c0 - S.s[0].C : c0;
      const bool chan_ok = cc < sr.C;
      if (n > 0) load_gamma_beta(t);
      coef_b = -1;
      // affine coefficients of (batch row b, this thread's 8 channels): a = gamma*rstd*scale*(film_s+1), ...
      // split in two so the FiLM loads fly together with the first batch of activation loads
      float fsv[8], fhv[8];
      auto film_issue = [&](int b) {
        if (has_film) {
          const float* fp = p.film + (size_t)scrow[b - b_first] * p.film_stride + c0;
          const float4 a0 = __ldcg(reinterpret_cast<const float4*>(fp)), a1 = __ldcg(reinterpret_cast<const float4*>(fp + 4));
          const float4 h0 = __ldcg(reinterpret_cast<const float4*>(fp + Ct)), h1 = __ldcg(reinterpret_cast<const float4*>(fp + Ct + 4));
          fsv[0] = a0.x; fsv[1] = a0.y; fsv[2] = a0.z; fsv[3] = a0.w; fsv[4] = a1.x; fsv[5] = a1.y; fsv[6] = a1.z; fsv[7] = a1.w;
          fhv[0] = h0.x; fhv[1] = h0.y; fhv[2] = h0.z; fhv[3] = h0.w; fhv[4] = h1.x; fhv[5] = h1.y; fhv[6] = h1.z; fhv[7] = h1.w;
        }
      };
      auto coef_finish = [&](int b) {
        const int bl = b - b_first;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = sr.scale, sft = 0.0f;
          if (has_gn) {
            const int g = (c0 + e) / cpg;
            const float ga = gam[e] * grstd[bl * 32 + g];
            a = ga * sr.scale;
            sft = bet[e] - gmean[bl * 32 + g] * ga;
          }
          if (has_film) {
            const float fs1 = fsv[e] + 1.0f;
            a = a * fs1;
            sft = sft * fs1 + fhv[e];
          }
          ca[e] = a;
          cs[e] = sft;
        }
        coef_b = b;
      };
      const bool need_coef = !s1 && affine && chan_ok && (has_gn || has_film);
      if (need_coef) film_issue(b_first);  // common case: one batch row per tile
      if (n >= 2) mbar_wait(&p_empty[pb], (uint32_t)(((n >> 1) - 1) & 1));
      uint8_t* pan = panels + (size_t)pb * panel_bytes;
      const int rows_all = f * R;  // (residue, row) pairs flattened: idx = rho * R + r
      for (int i0 = rr; i0 < rows_all; i0 += 128) {
        uint4 raw[8];
        int bb[8];
        bool ok[8];
        float mu[8], rs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + 16 * u;
          const int rho = idx / R, r = idx - rho * R;
          const int q = q0 + r;
          const int b = q / Lq;
          const int ml = q - b * Lq;
          const int irow = (ml + amin) * f + rho;
          bb[u] = b;
          ok[u] = (idx < rows_all) && chan_ok && b < p.B && irow >= 0 && irow < S.L;
          raw[u] = make_uint4(0u, 0u, 0u, 0u);
          mu[u] = 0.f;
          rs[u] = 1.f;
          if (ok[u]) {
            const size_t rowi = (size_t)(b % sr.bmod) * S.L + irow;
            raw[u] = __ldcg(reinterpret_cast<const uint4*>((const bf16*)sr.ptr + rowi * sr.C + cc));
            if (!s1 && !affine) {  // LayerNorm statistics of the row from the producer's per-tile partials
              const float2* rp = reinterpret_cast<const float2*>(p.rowpart) + rowi * p.rp_nct;
              float a = 0.f, qq = 0.f;
#pragma unroll 8
              for (int jj = 0; jj < p.rp_nct; ++jj) {
                const float2 v2 = __ldcg(rp + jj);
                a += v2.x;
                qq += v2.y;
              }
              const float inv = 1.0f / (float)sr.C;
              const float m1 = a * inv;
              float var = qq * inv - m1 * m1;
              if (var < 0.0f) var = 0.0f;
              mu[u] = m1;
              rs[u] = 1.0f / sqrtf(var + p.ln_eps);
            }
          }
        }
        if (need_coef && i0 == rr) coef_finish(b_first);
        if (tl && tid == 0 && n == 0 && i0 == rr) {
          if (raw[0].x == 0x12345678u) tl[31] = 1;  // consume the load
          TL_MARK(16);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + 16 * u;
          if (idx >= rows_all) break;
          const int rho = idx / R, r = idx - rho * R;
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (ok[u]) {
            float v[8];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] = __low2float(h[e]);
              v[2 * e + 1] = __high2float(h[e]);
            }
            if (s1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] *= sr.scale;
            } else if (affine) {
              if (has_gn || has_film) {
                if (coef_b != bb[u]) {
                  film_issue(bb[u]);
                  coef_finish(bb[u]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(ca[e], v[e], cs[e]);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= sr.scale;
              }
              if (p.act == ACT_SILU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = (v[e] - mu[u]) * rs[u];
            }
            o.x = pack2(v[0], v[1]);
            o.y = pack2(v[2], v[3]);
            o.z = pack2(v[4], v[5]);
            o.w = pack2(v[6], v[7]);
          }
          if (pl.bsw == 0) {
            *reinterpret_cast<uint4*>(pan + ((size_t)(rho * 8 + kc) * pl.PS + r) * 16) = o;
          } else {
            *reinterpret_cast<uint4*>(pan + ((size_t)rho * pl.PS + r) * 128 + (size_t)((kc ^ (r & 7)) * 16)) = o;
          }
        }
      }
      if (tid == 0 && n == 0) TL_MARK(17);
      fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&p_full[pb]);
      if (tid == 0 && n == 0) TL_MARK(4);
    }
    if (tid == 0) TL_MARK(5);

    // ---- epilogue.  Column metadata and residual values of a 16-column chunk are fetched as one batch; for the
    //      first chunk that happens while the tensor core is still working.
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int tile_id = (z * pl.m_tiles + mt) * pl.n_tiles + nt;
    const int zoff = p.out_off0 + z * p.out_off_phase;
    const bool want_stats = p.stats_out != nullptr;
    const bool want_rows = p.rowpart_out != nullptr;
    int eb = q0 / Lq, eml = q0 - eb * Lq;  // running (batch row, padded position) of the next column
    int oi[16];
    float rv[16];
    uint32_t vmask = 0, bmask = 0;
    auto chunk_meta = [&]() {
      vmask = 0;
      bmask = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int o = eml * p.out_stride + zoff;
        const bool valid = (eb < p.B) && (eml < p.Lm) && (o >= 0) && (o < p.Lout);
        oi[j] = valid ? (eb * p.Lout + o) * p.Cout + nch : 0;
        int ro = valid ? ((eb % p.res_bmod) * p.Lout + o) * p.Cout + nch : 0;
        if (valid) vmask |= 1u << j;
        ++eml;
        if (eml == Lq) {
          bmask |= 1u << j;
          eml = 0;
          ++eb;
        }
        rv[j] = (p.res && valid) ? ldf_cg((const bf16*)p.res + ro) : 0.0f;
      }
    };
    chunk_meta();

    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tid == 0) TL_MARK(6);
    bool final_cta = true;
    const float* wsbase = nullptr;
    if (pl.splitk > 1) {
      float* wp = A.ws + ((size_t)tile_id * pl.splitk + sk) * NT * 128;
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) __stcg(wp + (size_t)(c0 + j) * 128 + cl, v[j]);
      }
      __threadfence();
      bar_sync_producers();
      if (tid == 0) {
        const int tk = atomicAdd(A.counters + tile_id, 1);
        *ticket_slot = tk;
        if (tk == pl.splitk - 1) A.counters[tile_id] = 0;  // ready for the next launch
      }
      bar_sync_producers();
      final_cta = (*ticket_slot == pl.splitk - 1);
      if (tid == 0) TL_MARK(7);
      if (final_cta) {
        __threadfence();
        wsbase = A.ws + (size_t)tile_id * pl.splitk * NT * 128;
      }
    }
    if (final_cta) {
      int sb = q0 / Lq;  // batch row of the statistics run in progress
      float colS = 0.f, colQ = 0.f;
      for (int c0 = 0; c0 < NT; c0 += 16) {
        if (c0 > 0) chunk_meta();
        float v[16];
        if (wsbase == nullptr) {
          tmem_ld16(trow + (uint32_t)c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
          for (int s0 = 0; s0 < pl.splitk; s0 += 4) {  // 64 independent L2 loads in flight, summed in split order
            float tv[4][16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float* wp = wsbase + (size_t)(s0 + u) * NT * 128 + (size_t)c0 * 128 + cl;
              const bool on = s0 + u < pl.splitk;
#pragma unroll
              for (int j = 0; j < 16; ++j) tv[u][j] = (on && ((vmask >> j) & 1u)) ? __ldcg(wp + (size_t)j * 128) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += tv[u][j];
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool valid = (vmask >> j) & 1u;
          float x = 0.0f;
          if (valid) {
            x = v[j] + bias;
            if (p.epi_act == ACT_GELU) x = gelu_f(x);
            x += rv[j];
            if (A.out_f32) {
              ((float*)p.out)[oi[j]] = x;
            } else {
              ((bf16*)p.out)[oi[j]] = __float2bfloat16_rn(x);
            }
          }
          colS += x;
          colQ += x * x;
          if (want_rows) {
            const float rs = warp_sum(x), rq = warp_sum(x * x);
            if (lane == 0) {
              rowred[((size_t)warp * NT + c0 + j) * 2] = rs;
              rowred[((size_t)warp * NT + c0 + j) * 2 + 1] = rq;
            }
          }
          // flush the per-batch-row column sums at a row boundary / at the end of the tile
          const bool last_col = (c0 + j == NT - 1);
          const bool bnd = (bmask >> j) & 1u;
          if (bnd || last_col) {
            if (want_stats && sb <= b_last && sb - b_first < kMaxSlots) {
              sred[((size_t)(sb - b_first) * 128 + cl) * 2] = colS;
              sred[((size_t)(sb - b_first) * 128 + cl) * 2 + 1] = colQ;
            }
            colS = 0.f;
            colQ = 0.f;
            if (bnd) ++sb;
          }
        }
      }
      if (want_stats || want_rows) bar_sync_producers();
      if (want_stats) {
        // per (batch row, fine group) partial of this tile -> entry e = nt - t_first(b); the last tile of a batch
        // row also zeroes the unused trailing entries so consumers can sum a fixed n_ent.
        const int gs = p.Cout / p.FGo;
        const int ngl = 128 / gs;
        const int n_ent = pl.E_max * p.nphase;
        const int nb_out = min(p.B - 1, (q0 + NT - 1) / Lq) - b_first + 1;
        for (int idx = tid; idx < nb_out * ngl; idx += kProducers) {
          const int bl = idx / ngl, gl = idx - bl * ngl;
          const int bb = b_first + bl;
          float a = 0.f, q = 0.f;
          for (int c = gl * gs; c < (gl + 1) * gs; ++c) {
            a += sred[((size_t)bl * 128 + c) * 2];
            q += sred[((size_t)bl * 128 + c) * 2 + 1];
          }
          const int t_first = (bb * Lq) / NT;
          int t_last = ((bb + 1) * Lq - 1) / NT;
          if (t_last > pl.n_tiles - 1) t_last = pl.n_tiles - 1;
          const int e = nt - t_first;
          const int fg = (mt * 128) / gs + gl;
          float* so = p.stats_out + (((size_t)bb * n_ent + e * p.nphase + z) * p.FGo + fg) * 2;
          so[0] = a;
          so[1] = q;
          if (nt == t_last) {
            for (int e2 = e + 1; e2 < pl.E_max; ++e2) {
              float* s2 = p.stats_out + (((size_t)bb * n_ent + e2 * p.nphase + z) * p.FGo + fg) * 2;
              s2[0] = 0.f;
              s2[1] = 0.f;
            }
          }
        }
      }
      if (want_rows) {
        for (int col = tid; col < NT; col += kProducers) {
          const int q = q0 + col;
          const int bb = q / Lq, m2 = q - bb * Lq;
          const int o = m2 * p.out_stride + zoff;
          if (bb < p.B && m2 < p.Lm && o >= 0 && o < p.Lout) {
            float a = 0.f, qq = 0.f;
            for (int w = 0; w < 4; ++w) {
              a += rowred[((size_t)w * NT + col) * 2];
              qq += rowred[((size_t)w * NT + col) * 2 + 1];
            }
            float* ro = p.rowpart_out + (((size_t)bb * p.Lout + o) * pl.m_tiles + mt) * 2;
            ro[0] = a;
            ro[1] = qq;
          }
        }
      }
    }
  }

  if (tid == 0) { TL_MARK(8); TL_GLOBAL(12); }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)pl.tmem_cols)
                 : "memory");
  }
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

// ---------------------------------------------------------------------------------------------- planning
UmmaPlan conv_umma_plan(const ConvParams& p, bool want_stats, size_t ws_capacity_bytes, int num_sms) {
  UmmaPlan pl;
  memset(&pl, 0, sizeof(pl));
  const ConvSeg& S0 = p.seg[0];
  if (p.out == nullptr && p.out_ncl != nullptr) return pl;  // [B][C][L] output is a boundary format: generic path
  if (p.Cout % 128 != 0 || p.B < 1) return pl;
  for (int sg = 0; sg < p.nseg; ++sg)
    for (int k = 0; k < 2; ++k) {
      const ConvSrc& sr = p.seg[sg].s[k];
      if (sr.C > 0 && (sr.C % 8 != 0)) return pl;
    }
  if (S0.s[0].C <= 0) return pl;
  if (p.G > 32) return pl;
  if (want_stats && (p.FGo <= 0 || p.Cout % p.FGo != 0 || 128 % (p.Cout / p.FGo) != 0)) return pl;
  const int f = S0.in_stride;
  if (f < 1 || f > 8 || S0.ntaps < 1 || S0.ntaps > 32) return pl;
  int amin = 1 << 30, amax = -(1 << 30);
  for (int j = 0; j < S0.ntaps; ++j) {
    const int a = floordiv(S0.shift0 + j * S0.shift_step, f);
    amin = a < amin ? a : amin;
    amax = a > amax ? a : amax;
  }
  pl.amin = amin;
  pl.halo = amax - amin;
  pl.Lq = p.Lm + pl.halo;
  if (p.nseg > 1) {
    const ConvSeg& S1 = p.seg[1];
    if (S1.ntaps != 1 || S1.in_stride != 1 || S1.shift0 != 0 || p.out_stride != 1 || p.out_off0 != 0 || p.nphase != 1 ||
        S1.L != p.Lout || p.Lm != p.Lout)
      return pl;
  }
  pl.steps0 = (S0.Cin + 63) / 64;
  pl.steps1 = p.nseg > 1 ? (p.seg[1].Cin + 63) / 64 : 0;
  pl.m_tiles = p.Cout / 128;
  const long long nq = (long long)p.B * pl.Lq;
  // N tile: aim at >= one CTA per SM, bounded by the panel size (strided convs keep `f` sub-panels)
  int nt_cap = 256 / f;
  if (nt_cap < 16) nt_cap = 16;
  long long want = (nq * pl.m_tiles * p.nphase + num_sms - 1) / num_sms;
  int NT = round_up((int)(want < 16 ? 16 : want), 16);
  if (NT > nt_cap) NT = nt_cap;
  if (NT > 128 && NT < 256) NT = round_up(NT, 32);
  if ((long long)NT > round_up((int)nq, 16)) NT = round_up((int)nq, 16);
  if (NT < 16) NT = 16;
  // distinct batch rows per tile must fit the epilogue scratch
  auto slots = [&](int nt) {
    const int s = (nt + pl.halo + pl.Lq - 1) / pl.Lq + 1;
    return s < p.B ? s : p.B;
  };
  while (NT > 16 && slots(NT) > kMaxSlots) NT -= 16;
  if (slots(NT) > kMaxSlots) return pl;
  pl.NT = NT;
  pl.n_tiles = (int)((nq + NT - 1) / NT);
  pl.R = NT + pl.halo;
  static int bsw_env = -1;
  if (bsw_env < 0) {
    const char* e = getenv("JEN1_BSW");
    bsw_env = e ? atoi(e) : 1;
  }
  pl.bsw = bsw_env;
  if (pl.bsw == 0) {
    pl.PS = pl.R | 1;  // odd panel stride (in 16-byte units): conflict-free producer stores
    pl.panel_bytes = round_up(f * 8 * pl.PS * 16, 1024);
  } else {
    pl.PS = round_up(pl.R, 8);  // rows per sub-panel (128-byte rows, 128-byte swizzle, 1024-byte aligned)
    pl.panel_bytes = f * pl.PS * 128;
  }
  int tc = 32;
  while (tc < NT) tc <<= 1;
  pl.tmem_cols = tc;
  pl.E_max = (pl.Lq - 1) / NT + 2;
  // split-K: spread weight streaming over the GPU when the output tiles alone do not fill it
  const int nsteps = pl.steps0 + pl.steps1;
  const long long tiles = (long long)pl.n_tiles * pl.m_tiles * p.nphase;
  int sk = (int)(num_sms / tiles);
  if (sk < 1) sk = 1;
  if (sk > nsteps) sk = nsteps;
  if (sk > 32) sk = 32;
  while (sk > 1 && (size_t)tiles * sk * NT * 128 * sizeof(float) > ws_capacity_bytes) --sk;
  if (tiles > 65536) return pl;
  pl.splitk = sk;
  pl.ws_bytes = sk > 1 ? (size_t)tiles * sk * NT * 128 * sizeof(float) : 0;
  // shared memory: two panels + as many 16 KB weight stages as fit in ~half an SM (two CTAs co-reside under PDL)
  const int misc = 256 + 2 * kMaxSlots * 32 * 4 + kMaxSlots * 2 * 32 * 2 * 8 + kMaxSlots * 4;
  const int budget = 110 * 1024;
  int stages = (budget - 2 * pl.panel_bytes - misc) / kABytes;
  const int scratch = (kMaxSlots * 128 * 2 + 4 * NT * 2) * 4;  // epilogue scratch aliases the ring
  const int min_stages = (scratch + kABytes - 1) / kABytes;
  if (stages < 2) stages = 2;
  if (stages < min_stages) stages = min_stages;
  if (stages > 6) stages = 6;
  pl.stages = stages;
  pl.smem = (size_t)stages * kABytes + 2 * (size_t)pl.panel_bytes + misc + 1024;
  if (pl.smem > 227 * 1024) return pl;
  pl.ok = 1;
  return pl;
}

size_t conv_umma_packed_elems(int Cin, int Cout, int ntaps) {
  return (size_t)(Cout / 128) * ((Cin + 63) / 64) * ntaps * (kABytes / 2);
}

// Host-side packing of a conv weight W[tap][cin][cout] (fp32, the engine's logical layout) into the blob stream
// consumed by the kernel: blob (mt, phase z, cin block cb, tap j) = rows cout (128) x k cin (64) in the 128-byte
// swizzled K-major layout: element (r, k) at r * 64 + ((k / 8) ^ (r % 8)) * 8 + (k % 8).  `wtap(z, j) = wtap0 + z * wtap_phase + j * wtap_step` selects the source tap.
void conv_umma_pack(const float* w, int Cin, int Cout, int nphase, int taps_per_phase, int wtap0, int wtap_phase,
                    int wtap_step, uint16_t* out_bf16) {
  const int ncb = (Cin + 63) / 64, nmt = Cout / 128;
  size_t blob = 0;
  for (int mt = 0; mt < nmt; ++mt)
    for (int z = 0; z < nphase; ++z)
      for (int cb = 0; cb < ncb; ++cb)
        for (int j = 0; j < taps_per_phase; ++j, ++blob) {
          const int tap = wtap0 + z * wtap_phase + j * wtap_step;
          uint16_t* dst = out_bf16 + blob * (kABytes / 2);
          for (int k = 0; k < 64; ++k) {
            const int cin = cb * 64 + k;
            for (int r = 0; r < 128; ++r) {
              float v = 0.0f;
              if (cin < Cin) v = w[((size_t)tap * Cin + cin) * Cout + mt * 128 + r];
              __nv_bfloat16 h = __float2bfloat16_rn(v);
              dst[(size_t)r * 64 + (size_t)(((k / 8) ^ (r & 7)) * 8) + (k % 8)] = *reinterpret_cast<uint16_t*>(&h);
            }
          }
        }
}

cudaError_t conv_umma_init() {
  return cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_conv_umma(const ConvParams& p, const UmmaPlan& pl, const void* w0, const void* w1, float* ws,
                             int* counters, bool out_f32, bool pdl, cudaStream_t stream, long long* timeline) {
  UmmaArgs a;
  a.p = p;
  a.pl = pl;
  a.w0 = (const bf16*)w0;
  a.w1 = (const bf16*)w1;
  a.ws = ws;
  a.counters = counters;
  a.out_f32 = out_f32 ? 1 : 0;
  a.timeline = timeline;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pl.n_tiles, pl.m_tiles, p.nphase * pl.splitk);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel, a);
}

}  // namespace jen1
