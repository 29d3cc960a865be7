// Fused Transformer1d for bf16 storage on sm_100a (reference jen1/model/blocks.py:497-537 with TransformerBlock :483-489,
// Attention :415-437 / AttentionBase :355-380 and FeedForward :295-301): ONE launch runs the whole block chain
//
//     t = conv1x1(GroupNorm32(x));  per block: t += SelfAttn(LN(t));  t += CrossAttn(LN(t), ctx);  t += FF(t);  out = conv1x1(t)
//
// instead of the 10 dependent launches of the unfused path.  Every op of the chain only mixes data INSIDE one batch row
// (tokens of one sample), so a thread-block cluster owns one row and walks the op list with cluster-scope barriers
// where the unfused path has kernel boundaries:
//
//   * cluster (CS,1,1) per batch row, CS = 8 or 16 CTAs; op outputs go through global memory (L2) exactly as in the unfused
//     path, made visible by fence + mbarrier arrive.release.cluster / try_wait.acquire.cluster (one mbarrier per CTA,
//     CS remote arrivals per op) -- only the warps that consume the data wait, so
//   * the weight streamer (one thread, cp.async.bulk of the same 16 KB tcgen05 blobs conv_umma.cu uses) runs AHEAD across
//     op boundaries through a ring that is never aliased: the weight stream of the whole transformer is one
//     uninterrupted sequence per CTA, started before griddepcontrol.wait;
//   * a linear op is split over the cluster by output tile (128 channels = the 128 TMEM lanes): D[cout][token] =
//     W[cout][K] * P[token][K]^T with the row's WHOLE activation panel P (<= 80 tokens x K <= 1024 channels, <= 48 KB) resident
//     in shared memory, so there is no split-K exchange at all; GroupNorm-apply / LayerNorm are computed from the panel
//     itself (the LayerNorm statistics need no side channel: a CTA sees complete rows);
//   * attention runs one head per CTA with tcgen05 QK^T / PV exactly as attn_umma.cu does (zero-logit key masking,
//     causal mask, cond-dropout / CFG K/V source select), staged from the q/k/v the cluster just wrote.
//
// Warp roles (192 threads): warps 0-3 build panels / stage attention tiles, run softmax and every epilogue; warp 4 lane 0
// streams weights; warp 5 allocates TMEM and its lane 0 issues every tcgen05.mma.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

namespace {

constexpr int kThreads = 192;
constexpr int kProd = 128;
constexpr int kABytes = 128 * 64 * 2;
constexpr int kMinStages = 4, kMaxStages = 8;  // weight ring depth (runtime: what fits next to the work region)
constexpr int kTmemCols = 256;
constexpr int kSfgBytes = 4 * 32 * 2 * 4;  // per-warp GroupNorm fine-group sums of one M tile

// ---------------------------------------------------------------------------------------------- PTX helpers
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
// acquire at cluster scope: pairs with the remote arrive.release.cluster of the op barrier
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// relaxed: the publishing thread issues ONE fence.acq_rel.gpu before the CS arrives (fence + relaxed atomic = release
// pattern); a .release on each of the 16 remote arrives was measured at ~0.2 us apiece, 3 us per op
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
__device__ __forceinline__ void bar_prod() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
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
// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor version 1); see attn_umma.cu
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __low2float(h[e]);
    v[2 * e + 1] = __high2float(h[e]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

struct AttnTiles {  // byte offsets inside the work region (see attn_umma.cu for the layout rules)
  int KP, DB, cpr_shift;
  uint32_t off_k, off_v, off_p;
};
__device__ __forceinline__ AttnTiles attn_tiles(int M, int d) {
  AttnTiles g;
  g.KP = (M + 15) / 16 * 16;
  g.DB = (d + 63) / 64;
  g.cpr_shift = d == 16 ? 1 : (d == 32 ? 2 : (d == 64 ? 3 : 4));
  const uint32_t q_bytes = (uint32_t)g.DB * 128u * 128u;
  const uint32_t k_bytes = (uint32_t)g.DB * (uint32_t)g.KP * 128u;
  const uint32_t p_bytes = (uint32_t)((g.KP + 63) / 64) * 128u * 128u;
  g.off_k = q_bytes;
  const uint32_t qk = (q_bytes + k_bytes + 1023u) / 1024u * 1024u;
  const uint32_t pq = (p_bytes + 1023u) / 1024u * 1024u;
  g.off_p = 0;
  g.off_v = qk > pq ? qk : pq;
  return g;
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 1) tr_umma_kernel(const __grid_constant__ TrParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CS = P.CS;
  const int crank = (int)cluster_ctarank();
  const int row = blockIdx.y;  // batch row of this cluster
  const int N = P.N, NT = P.NT;

  // ---- shared memory carve-up: [weight ring][work region: panel | staging | stat cells  /  attention tiles][barriers]
  uint8_t* ring = smem;
  const int kStages = P.stages;
  uint8_t* work = smem + (size_t)kStages * kABytes;
  uint8_t* panel = work;                                           // [K/64 blocks][NT rows][128 B]
  float* stage = reinterpret_cast<float*>(work + P.panel_bytes);   // [NT][128] fp32
  float* sfg = stage + (size_t)NT * 128;                           // [4 warps][32 fine groups][2]
  float* rowred = sfg + 4 * 32 * 2;                                // [NT][4][2] LayerNorm row partials
  uint8_t* resbuf = reinterpret_cast<uint8_t*>(rowred + (size_t)NT * 4 * 2);  // [NT][128] bf16 residual rows of one M tile
  const bool prestage = P.prestage != 0;                           // cross-attention K / V tiles staged at kernel start
  uint8_t* kvx = work + P.work_bytes;                              // [K tile][V tile] of head `crank` (prestage only)
  const uint32_t kvx_half = (uint32_t)P.kvx_bytes >> 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.smem_bytes - 256);
  uint64_t* a_full = bars;            // [kMaxStages]
  uint64_t* a_empty = bars + 8;       // [kMaxStages]
  uint64_t* panel_full = bars + 16;   // producers -> MMA (GEMM panel or attention tiles staged), 128 arrivals
  uint64_t* acc_full = bars + 17;     // MMA -> epilogue (tcgen05.commit)
  uint64_t* acc_empty = bars + 18;    // epilogue -> MMA (accumulator drained), 128 arrivals
  uint64_t* p_full = bars + 19;       // softmax -> MMA (P tile written), 128 arrivals
  uint64_t* opbar = bars + 20;        // cluster-wide op barrier: CS remote arrivals per op
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  if (tid == kProd) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(panel_full, kProd);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kProd);
    mbar_init(p_full, kProd);
    mbar_init(opbar, (uint32_t)CS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  // every CTA's barriers must exist before anybody arrives on them remotely
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int n_ops = P.n_ops;

  // Warps 4 and 5 run their warp-uniform control flow with all lanes and one ELECTED lane issues (see conv_umma.cu: a
  // divergent `lane == 0` issuer costs ~85 clocks per tcgen05.mma -- the "2.3 us accumulate phase" of the first version)
  if (warp == 4) {
    // ======================================================================== weight streamer (runs ahead of the ops)
    int s = 0, k = 0;
    for (int oi = 0; oi < n_ops; ++oi) {
      const TrOp& op = P.ops[oi];
      if (op.type != TR_GEMM) continue;
      const int kb_n = op.K >> 6, mtiles = op.Cout >> 7;
      for (int mt = crank; mt < mtiles; mt += CS) {
        const bf16* src = op.w + (size_t)mt * kb_n * (kABytes / 2);
        for (int kb = 0; kb < kb_n; ++kb) {
          if (k > 0) mbar_wait(&a_empty[s], (uint32_t)((k - 1) & 1));
          if (elect_one()) {
            mbar_expect_tx(&a_full[s], kABytes);
            bulk_g2s(ring + (size_t)s * kABytes, src + (size_t)kb * (kABytes / 2), kABytes, &a_full[s]);
          }
          __syncwarp();
          if (++s == kStages) {
            s = 0;
            ++k;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ======================================================================== MMA issuer
    int s = 0, k = 0;
    uint32_t n_panel = 0, n_acc_use = 0, n_p = 0;  // completed phases of panel_full / uses of the accumulator / p_full
    auto commit = [&](uint64_t* bar) {
      if (elect_one()) umma_commit(bar);
      __syncwarp();
    };
    const uint64_t ring_d = make_desc_sw128(smem_u32(ring), 16u, 1024u), panel_d = make_desc_sw128(smem_u32(panel), 16u, 1024u);
    for (int oi = 0; oi < n_ops; ++oi) {
      const TrOp& op = P.ops[oi];
      if (op.type == TR_GEMM) {
        const int kb_n = op.K >> 6, mtiles = op.Cout >> 7;
        if (crank >= mtiles) continue;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
        mbar_wait(panel_full, n_panel & 1);
        ++n_panel;
        tc_fence_after();
        if (P.timeline && crank == 0 && row == 0 && lane == 0) P.timeline[oi * 8 + 6] = clock64();
        for (int mt = crank; mt < mtiles; mt += CS) {
          if (n_acc_use > 0) {  // the previous accumulator contents have been read out
            mbar_wait(acc_empty, (n_acc_use - 1) & 1);
            tc_fence_after();
          }
          for (int kb = 0; kb < kb_n; ++kb) {
            mbar_wait(&a_full[s], (uint32_t)(k & 1));
            tc_fence_after();
            const uint64_t ad = ring_d + (uint64_t)((uint32_t)s * (kABytes >> 4));
            const uint64_t bd = panel_d + (uint64_t)((uint32_t)kb * (uint32_t)NT * 8u);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
              umma_commit(&a_empty[s]);
            }
            __syncwarp();
            if (++s == kStages) {
              s = 0;
              ++k;
            }
          }
          commit(acc_full);
          ++n_acc_use;
          if (P.timeline && crank == 0 && row == 0 && mt == crank && lane == 0) P.timeline[oi * 8 + 7] = clock64();
        }
      } else {
        // attention: heads crank, crank + CS, ...
        const AttnTiles g = attn_tiles(op.M, P.d);
        const int d = P.d;
        const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.KP >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(d >> 3) << 17) | ((128u >> 4) << 24);
        const int nk = d >> 4, n16 = g.KP >> 4;
        for (int h = crank; h < P.H; h += CS) {
          mbar_wait(panel_full, n_panel & 1);
          ++n_panel;
          tc_fence_after();
          if (n_acc_use > 0) {
            mbar_wait(acc_empty, (n_acc_use - 1) & 1);
            tc_fence_after();
          }
          const bool ext = prestage && op.cross && h == crank;
          const uint32_t Qs = smem_u32(work), Ps = Qs + g.off_p;
          const uint32_t Ks = ext ? smem_u32(kvx) : Qs + g.off_k, Vs = ext ? smem_u32(kvx) + kvx_half : Qs + g.off_v;
          const uint64_t qd = make_desc_sw128(Qs, 16u, 1024u), kd = make_desc_sw128(Ks, 16u, 1024u);
          const uint64_t pd = make_desc_sw128(Ps, 16u, 1024u), vd = make_desc_sw128(Vs, (uint32_t)g.KP * 128u, 1024u);
          const uint32_t kblk16 = (uint32_t)g.KP * 8u;  // one 64-channel block of K rows in descriptor address units
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < nk)
                umma_bf16(tmem_base, qd + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2),
                          kd + (uint64_t)((uint32_t)(kk >> 2) * kblk16 + (uint32_t)(kk & 3) * 2u), idesc_s, kk > 0 ? 1u : 0u);
            umma_commit(acc_full);  // S ready
          }
          __syncwarp();
          ++n_acc_use;
          mbar_wait(p_full, n_p & 1);  // P written (and S fully read)
          ++n_p;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll 4
            for (int k16 = 0; k16 < n16; ++k16)
              umma_bf16(tmem_base, pd + (uint64_t)((k16 >> 2) * 1024 + (k16 & 3) * 2), vd + (uint64_t)(k16 * 128), idesc_o, k16 > 0 ? 1u : 0u);
            umma_commit(acc_full);  // O ready (second use of the accumulator by this head)
          }
          __syncwarp();
          ++n_acc_use;
        }
      }
    }
  } else {
    // ======================================================================== producers / softmax / epilogues
    uint32_t n_acc = 0;   // acc_full phases consumed
    const int q4 = lane * 4;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t opbar_remote[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) opbar_remote[j] = map_cluster(smem_u32(opbar), (uint32_t)(j < CS ? j : 0));

    // Attention tile staging (16-byte chunks; K-major 128-byte-swizzled rows as in attn_umma.cu) with cp.async: rolled
    // loops, every chunk in flight at once, no registers.  kinds: bit 0 = Q rows, bit 1 = K and V rows.  The caller waits
    // (cp.async.wait_all) and then calls finish_attn for the masked-key scaling.
    auto stage_attn = [&](const TrOp& op, int h, int kinds, uint8_t* qb, uint8_t* kb, uint8_t* vb) {
      const AttnTiles g = attn_tiles(op.M, P.d);
      const int d = P.d, M = op.M, S = M - 1;
      const int sh = g.cpr_shift, cpr = 1 << sh;
      const int bc = row % P.Bc;
      const bool fixed = op.cross && ((row >= P.Bc) || (P.drop && P.drop[bc]));
      const int trow_time = op.cross ? P.cond_row[row] : 0;
      if (kinds & 1) {
        const bf16* qsrc = op.q + (size_t)row * N * op.q_ld + h * d;
        for (int i = tid; i < (N << sh); i += kProd) {
          const int rw = i >> sh, part = i & (cpr - 1);
          uint8_t* dst = qb + ((uint32_t)(part >> 3) * 128u + (uint32_t)rw) * 128u + (uint32_t)(((part & 7) ^ (rw & 7)) * 16);
          cp_async16(dst, qsrc + (size_t)rw * op.q_ld + part * 8);
        }
      }
      if (kinds & 2) {
        for (int i = tid; i < (g.KP << (sh + 1)); i += kProd) {
          const int v = i >= (g.KP << sh) ? 1 : 0;  // 0: K tile, 1: V tile
          const int j = i - (v ? (g.KP << sh) : 0);
          const int rw = j >> sh, part = j & (cpr - 1);
          uint8_t* dst = (v ? vb : kb) + ((uint32_t)(part >> 3) * (uint32_t)g.KP + (uint32_t)rw) * 128u + (uint32_t)(((part & 7) ^ (rw & 7)) * 16);
          if (rw >= M) {  // rows beyond the key count must be exact zeros (P is 0 there, 0 * garbage could be NaN)
            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          const bf16* sp;
          if (!op.cross) {
            sp = op.kv + ((size_t)row * N + rw) * op.kv_ld + (v ? op.v_off : op.k_off);
          } else {
            const bf16* rowp;
            if (rw < S)
              rowp = fixed ? P.kv_fixed + (size_t)rw * P.kvc_ld : P.kv_cond + ((size_t)bc * S + rw) * P.kvc_ld;
            else
              rowp = fixed ? P.kv_fixed + (size_t)S * P.kvc_ld : P.kv_time + (size_t)trow_time * P.kvc_ld;
            sp = rowp + op.kvc_off + (v ? P.C : 0);
          }
          cp_async16(dst, sp + h * d + part * 8);
        }
      }
      if (cpr < 8) {  // chunks beyond the head dim inside the 64-channel block must read as zero (d = 16 / 32)
        const int zc = 8 - cpr;
        const int r_lo = (kinds & 1) ? 0 : 128, r_hi = (kinds & 2) ? 128 + 2 * g.KP : 128;
        for (int it = r_lo * zc + tid; it < r_hi * zc; it += kProd) {
          const int row_all = it / zc, chz = cpr + (it - row_all * zc);
          uint8_t* tile = row_all < 128 ? qb : (row_all < 128 + g.KP ? kb : vb);
          const int rw = row_all < 128 ? row_all : (row_all < 128 + g.KP ? row_all - 128 : row_all - 128 - g.KP);
          *reinterpret_cast<uint4*>(tile + (uint32_t)rw * 128u + (uint32_t)((chz ^ (rw & 7)) * 16)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // masked context keys: K and V rows are multiplied by the mask (logit 0, value 0 -- reference blocks.py:431-434);
    // every thread revisits exactly the chunks it requested itself (visible to it after cp.async.wait_all)
    auto finish_attn = [&](const TrOp& op, uint8_t* kb, uint8_t* vb) {
      if (!op.cross || P.mask == nullptr) return;
      const AttnTiles g = attn_tiles(op.M, P.d);
      const int sh = g.cpr_shift, cpr = 1 << sh, S = op.M - 1, bc = row % P.Bc;
      for (int i = tid; i < (g.KP << (sh + 1)); i += kProd) {
        const int v = i >= (g.KP << sh) ? 1 : 0;
        const int j = i - (v ? (g.KP << sh) : 0);
        const int rw = j >> sh, part = j & (cpr - 1);
        if (rw >= S) continue;
        const float mk = __ldg(P.mask + (size_t)bc * S + rw);
        if (mk == 1.0f) continue;
        uint4* cell = reinterpret_cast<uint4*>((v ? vb : kb) + ((uint32_t)(part >> 3) * (uint32_t)g.KP + (uint32_t)rw) * 128u +
                                               (uint32_t)(((part & 7) ^ (rw & 7)) * 16));
        float f[8];
        unpack8(*cell, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] *= mk;
        *cell = pack8(f);
      }
    };
    // The cross-attention K / V tiles of this CTA's head are per-step constants (prompt / null-embedding caches + the
    // time-token row written before the step's kernel chain): staged now, while the previous kernel is still running.
    if (prestage && crank < P.H) {
      stage_attn(P.ops[P.cross_op], crank, 2, work, kvx, kvx + kvx_half);
      asm volatile("cp.async.wait_all;" ::: "memory");
      finish_attn(P.ops[P.cross_op], kvx, kvx + kvx_half);
      fence_async_smem();
    }

    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads what earlier kernels wrote
    long long* tl = (P.timeline && crank == 0 && row == 0 && tid == 0) ? P.timeline : nullptr;
#define TR_MARK(j) do { if (tl) tl[oi * 8 + (j)] = clock64(); } while (0)

    for (int oi = 0; oi < n_ops; ++oi) {
      const TrOp& op = P.ops[oi];
      TR_MARK(0);
      if (oi > 0) mbar_wait_cluster(opbar, (uint32_t)((oi - 1) & 1));  // the previous op is complete cluster-wide
      TR_MARK(1);

      if (op.type == TR_GEMM) {
        const int K = op.K, mtiles = op.Cout >> 7;
        if (crank < mtiles) {
          // ------------------------------------------------------------------ panel: the row's whole input, transformed
          // raw rows -> panel with cp.async (one L2 round trip for everything, rolled loop); GroupNorm-apply / LayerNorm
          // then run in place over shared memory.  A thread always meets the same 16-byte channel chunk `ch`.
          const int lg = K == 256 ? 5 : (K == 512 ? 6 : 7);  // log2(16-byte chunks per token row)
          const int cpr = 1 << lg, total = N << lg;
          const int ch = tid & (cpr - 1);
          const bf16* src = op.src + (size_t)(row % op.src_bmod) * N * op.src_ld;
          uint8_t* pcol = panel + (uint32_t)(ch >> 3) * (uint32_t)NT * 128u;  // this thread's 64-channel block
          for (int i = tid; i < total; i += kProd) {
            const int r = i >> lg;
            cp_async16(pcol + (uint32_t)r * 128u + (uint32_t)(((ch & 7) ^ (r & 7)) * 16), src + (size_t)r * op.src_ld + ch * 8);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          float ga[8], gb[8];  // GN: per-channel (a, s) of this thread's chunk: y = a * x + s
          if (op.pro == TRP_GN) {
            const int gsz = K >> 5;  // channels per group (8, 16 or 32)
            const int g = (ch * 8) / gsz;
            const longlong2 acc = __ldcg(reinterpret_cast<const longlong2*>(P.gn_stats + ((size_t)(row % op.src_bmod) * 32 + g) * 2));
            const float inv_n = 1.0f / ((float)gsz * (float)N);
            const float mean = stat_get(acc.x) * inv_n;
            float var = fmaf(-mean, mean, stat_get(acc.y) * inv_n);
            if (var < 0.f) var = 0.f;
            const float rstd = rsqrtf(var + P.gn_eps);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float gm = __ldg(P.gn_gamma + ch * 8 + e) * rstd;
              ga[e] = gm;
              gb[e] = fmaf(-mean, gm, __ldg(P.gn_beta + ch * 8 + e));
            }
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
          if (op.pro != TRP_RAW) {
            // (every thread only revisits the chunks it copied itself: no barrier needed before this pass)
            const int nw = cpr >> 5;  // warps per token row (1, 2 or 4)
            if (op.pro == TRP_LN) {
              const int wir = warp & (nw - 1);
#pragma unroll 1
              for (int i = tid; i < total; i += kProd) {  // warp-uniform trip count (total is a multiple of 32)
                const int r = i >> lg;
                float v[8];
                unpack8(*reinterpret_cast<const uint4*>(pcol + (uint32_t)r * 128u + (uint32_t)(((ch & 7) ^ (r & 7)) * 16)), v);
                float a = 0.f, q = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  a += v[e];
                  q = fmaf(v[e], v[e], q);
                }
                a = warp_sum(a);
                q = warp_sum(q);
                if (lane == 0) *reinterpret_cast<float2*>(rowred + ((size_t)r * 4 + wir) * 2) = make_float2(a, q);
              }
              bar_prod();
            }
            const float inv_k = 1.0f / (float)K;
#pragma unroll 1
            for (int i = tid; i < total; i += kProd) {
              const int r = i >> lg;
              uint4* cell = reinterpret_cast<uint4*>(pcol + (uint32_t)r * 128u + (uint32_t)(((ch & 7) ^ (r & 7)) * 16));
              float v[8];
              unpack8(*cell, v);
              if (op.pro == TRP_LN) {
                float a = 0.f, q = 0.f;
                for (int w = 0; w < nw; ++w) {
                  const float2 pr = *reinterpret_cast<const float2*>(rowred + ((size_t)r * 4 + w) * 2);
                  a += pr.x;
                  q += pr.y;
                }
                const float mean = a * inv_k;
                float var = q * inv_k - mean * mean;
                if (var < 0.f) var = 0.f;
                const float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(ga[e], v[e], gb[e]);
              }
              *cell = pack8(v);
            }
          }
          fence_async_smem();
          mbar_arrive(panel_full);
          TR_MARK(2);

          // ------------------------------------------------------------------ epilogue of every M tile of this CTA
          const unsigned short* resp = reinterpret_cast<const unsigned short*>(op.res);
          const size_t rbase = (size_t)row * N;
          for (int mt = crank; mt < mtiles; mt += CS) {
            const float4 bias4 = op.bias ? __ldg(reinterpret_cast<const float4*>(op.bias + mt * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (resp) {  // residual rows of this tile -> shared memory (fire-and-forget; not aliased with anything live)
              for (int i = tid; i < N * 16; i += kProd)
                cp_async16(resbuf + (size_t)i * 16, resp + (rbase + (i >> 4)) * op.res_ld + mt * 128 + (i & 15) * 8);
              asm volatile("cp.async.commit_group;" ::: "memory");
            }
            mbar_wait(acc_full, n_acc & 1);
            ++n_acc;
            tc_fence_after();
            if (mt == crank) TR_MARK(3);
            for (int c0 = 0; c0 < NT; c0 += 16) {
              float v[16];
              tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
              for (int j = 0; j < 16; ++j) stage[(size_t)(c0 + j) * 128 + tid] = v[j];
            }
            tc_fence_before();
            mbar_arrive(acc_empty);  // the accumulator may be overwritten by the next tile / op
            asm volatile("cp.async.wait_all;" ::: "memory");
            bar_prod();
            float aS[4] = {0.f, 0.f, 0.f, 0.f}, aQ[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int c = warp; c < N; c += 4) {
              const float4 acc = *reinterpret_cast<const float4*>(stage + (size_t)c * 128 + q4);
              float x[4] = {acc.x + bias4.x, acc.y + bias4.y, acc.z + bias4.z, acc.w + bias4.w};
              if (op.gelu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = gelu_f(x[e]);
              }
              if (resp) {
                const uint2 rv = *reinterpret_cast<const uint2*>(resbuf + (size_t)c * 256 + lane * 8);
                x[0] += __uint_as_float(rv.x << 16);
                x[1] += __uint_as_float(rv.x & 0xffff0000u);
                x[2] += __uint_as_float(rv.y << 16);
                x[3] += __uint_as_float(rv.y & 0xffff0000u);
              }
              *reinterpret_cast<uint2*>(op.dst + (rbase + c) * op.dst_ld + mt * 128 + q4) = make_uint2(pack2(x[0], x[1]), pack2(x[2], x[3]));
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                aS[e] += x[e];
                aQ[e] = fmaf(x[e], x[e], aQ[e]);
              }
            }
            if (op.stats) {  // GroupNorm fine-group sums of the output for the next consumer (fixed-point accumulators)
              const int gs = op.Cout / P.FGo, lpg = gs >> 2, ngl = 128 / gs;
              float t1 = (aS[0] + aS[1]) + (aS[2] + aS[3]);
              float t2 = (aQ[0] + aQ[1]) + (aQ[2] + aQ[3]);
              for (int o = 1; o < lpg; o <<= 1) {
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o);
              }
              if ((lane & (lpg - 1)) == 0) *reinterpret_cast<float2*>(sfg + ((size_t)warp * 32 + lane / lpg) * 2) = make_float2(t1, t2);
              bar_prod();
              if (tid < ngl) {
                float a = 0.f, q = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  const float2 v2 = *reinterpret_cast<const float2*>(sfg + ((size_t)w * 32 + tid) * 2);
                  a += v2.x;
                  q += v2.y;
                }
                long long* so = P.stats_out + ((size_t)row * P.FGo + (mt * 128) / gs + tid) * 2;
                stat_add(so, a);
                stat_add(so + 1, q);
              }
            }
            bar_prod();  // staging tile / residual tile / stat cells are free again
          }
        }
      } else {
        // -------------------------------------------------------------------- attention (one head per CTA at a time)
        const AttnTiles g = attn_tiles(op.M, P.d);
        const int d = P.d, M = op.M;
        for (int h = crank; h < P.H; h += CS) {
          // stage Q (and K, V unless this head's cross-attention K/V tiles were staged at kernel start)
          const bool ext = prestage && op.cross && h == crank;
          uint8_t* kb = ext ? kvx : work + g.off_k;
          uint8_t* vb = ext ? kvx + kvx_half : work + g.off_v;
          stage_attn(op, h, ext ? 1 : 3, work, kb, vb);
          asm volatile("cp.async.wait_all;" ::: "memory");
          if (!ext) finish_attn(op, kb, vb);
          fence_async_smem();
          mbar_arrive(panel_full);
          if (h == crank) TR_MARK(2);

          // softmax: thread == query row (TMEM lane)
          const int i = tid;
          const float sc = P.scale * 1.4426950408889634f;
          const int jmax = (P.causal && !op.cross) ? i + (M - N) : M - 1;
          uint8_t* Ps = work + g.off_p;
          mbar_wait(acc_full, n_acc & 1);
          ++n_acc;
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
              if (key < M) mx = fmaxf(mx, s);
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
              float pj = (key < M) ? exp2f(s - mx) : 0.0f;
              pj = bf16_round(pj);
              sum += pj;
              e[j] = pj;
            }
            uint8_t* prow = Ps + (size_t)(c0 >> 6) * 128 * 128 + (size_t)tid * 128;
            const int chn = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(prow + (((chn) ^ (tid & 7)) * 16)) =
                make_uint4(pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
            *reinterpret_cast<uint4*>(prow + (((chn + 1) ^ (tid & 7)) * 16)) =
                make_uint4(pack2(e[8], e[9]), pack2(e[10], e[11]), pack2(e[12], e[13]), pack2(e[14], e[15]));
          }
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(acc_empty);  // S has been read: the accumulator columns may be reused by P V
          mbar_arrive(p_full);
          if (h == crank) TR_MARK(3);
          const float inv = 1.0f / sum;
          mbar_wait(acc_full, n_acc & 1);
          ++n_acc;
          tc_fence_after();
          bf16* orow = op.ao + ((size_t)row * N + (i < N ? i : 0)) * P.C + h * d;
          for (int c0 = 0; c0 < d; c0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (i < N) {
              float o0[8], o1[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                o0[e] = v[e] * inv;
                o1[e] = v[8 + e] * inv;
              }
              *reinterpret_cast<uint4*>(orow + c0) = pack8(o0);
              *reinterpret_cast<uint4*>(orow + c0 + 8) = pack8(o1);
            }
          }
          tc_fence_before();
          mbar_arrive(acc_empty);  // O has been read
          bar_prod();              // the tiles may be restaged (next head / next op's panel)
        }
      }

      // ---------------------------------------------------------------------- op complete: publish to the cluster
      TR_MARK(4);
      bar_prod();  // all 128 threads have issued their global stores / atomics of this op
      if (tid == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        TR_MARK(5);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < CS) mbar_arrive_remote(opbar_remote[j]);
      }
    }
    // nobody may leave while remote CTAs can still arrive on its op barrier: wait for the last op's phase
    mbar_wait_cluster(opbar, (uint32_t)((n_ops - 1) & 1));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
bool tr_umma_supported(int N, int C, int H, int M, int n_blocks) {
  if (!(C == 256 || C == 512 || C == 1024)) return false;  // K / 8 chunks per row must divide (or equal) 128 threads
  if (C % H != 0) return false;
  const int d = C / H;
  if (!(d == 16 || d == 32 || d == 64 || d == 128)) return false;
  if (N < 1 || N > 128 || M < 1 || M > 256) return false;   // one 128-query tile, keys within one TMEM accumulator
  if (n_blocks < 1 || 2 + 8 * n_blocks > TR_MAX_OPS) return false;
  return true;
}

static size_t tr_work_bytes(int N, int C, int H, int M) {
  const int NT = (N + 15) / 16 * 16;
  const int d = C / H;
  const size_t panel = (size_t)(C / 64) * NT * 128;
  const size_t gemm = panel + (size_t)NT * 512 + kSfgBytes + (size_t)NT * 4 * 8 + (size_t)NT * 256 + 1024;
  const int Mx = M > N ? M : N;
  const int KP = (Mx + 15) / 16 * 16, DB = (d + 63) / 64;
  const size_t q_bytes = (size_t)DB * 128 * 128, k_bytes = (size_t)DB * KP * 128, p_bytes = (size_t)((KP + 63) / 64) * 128 * 128;
  size_t qk = (q_bytes + k_bytes + 1023) / 1024 * 1024, pq = (p_bytes + 1023) / 1024 * 1024;
  const size_t attn = (qk > pq ? qk : pq) + (k_bytes + 1023) / 1024 * 1024;
  return ((gemm > attn ? gemm : attn) + 1023) / 1024 * 1024;
}
static size_t tr_kvx_bytes(int C, int H, int M) {  // K + V tiles of one head for the cross-attention keys
  const int d = C / H, KP = (M + 15) / 16 * 16, DB = (d + 63) / 64;
  return 2 * (((size_t)DB * KP * 128 + 1023) / 1024 * 1024);
}
size_t tr_umma_smem_bytes(int N, int C, int H, int M) { return (size_t)kMinStages * kABytes + tr_work_bytes(N, C, H, M) + 256; }

cudaError_t tr_umma_init() {
  cudaError_t e = cudaFuncSetAttribute(tr_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  (void)cudaFuncSetAttribute(tr_umma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  (void)cudaGetLastError();
  return cudaSuccess;
}

cudaError_t launch_tr_umma(const TrParams& p_in, bool pdl, cudaStream_t stream) {
  TrParams p = p_in;
  p.NT = (p.N + 15) / 16 * 16;
  p.panel_bytes = (p.C / 64) * p.NT * 128;
  int Mx = 1;
  for (int i = 0; i < p.n_ops; ++i)
    if (p.ops[i].type == TR_ATTN && p.ops[i].M > Mx) Mx = p.ops[i].M;
  p.work_bytes = (int)tr_work_bytes(p.N, p.C, p.H, Mx);
  const size_t limit = (size_t)227 * 1024;
  if ((size_t)kMinStages * kABytes + p.work_bytes + 256 > limit) return cudaErrorInvalidValue;
  // Shared memory budget, in order of what the op chain gains most from: (1) a weight ring deep enough to hold one whole
  // M tile of the next op (C / 64 blobs: the stream then never stalls a tile on L2 latency), up to 8 stages; (2) the
  // cross-attention K / V tiles staged at kernel start (only with exactly one cross-attention op); (3) more ring stages.
  int stages = kMinStages;
  const int want = p.C / 64 < kMaxStages ? (p.C / 64 > kMinStages ? p.C / 64 : kMinStages) : kMaxStages;
  while (stages < want && (size_t)(stages + 1) * kABytes + p.work_bytes + 256 <= limit) ++stages;
  p.prestage = 0;
  p.kvx_bytes = 0;
  p.cross_op = -1;
  int ncross = 0;
  for (int i = 0; i < p.n_ops; ++i)
    if (p.ops[i].type == TR_ATTN && p.ops[i].cross) {
      ++ncross;
      p.cross_op = i;
    }
  if (ncross == 1) {
    const size_t kvx = tr_kvx_bytes(p.C, p.H, p.ops[p.cross_op].M);
    if ((size_t)stages * kABytes + p.work_bytes + 256 + kvx <= limit) {
      p.prestage = 1;
      p.kvx_bytes = (int)kvx;
    }
  }
  while (stages < kMaxStages && (size_t)(stages + 1) * kABytes + p.work_bytes + 256 + p.kvx_bytes <= limit) ++stages;
  p.stages = stages;
  p.smem_bytes = (int)((size_t)stages * kABytes + p.work_bytes + 256 + p.kvx_bytes);
  // Cluster size: 16 CTAs per row when all rows' clusters can be resident at once, else 8 (a second wave of clusters
  // doubles the launch's duration; co-residency of 16-CTA clusters is limited by the SMs per GPC).
  if (p.CS == 16) {
    static int active16 = -1;
    static int active16_smem = -1;
    if (active16 < 0 || active16_smem != p.smem_bytes) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(16, 64, 1);
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = p.smem_bytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, tr_umma_kernel, &q) != cudaSuccess) n = 0;
      (void)cudaGetLastError();
      active16 = n;
      active16_smem = p.smem_bytes;
      if (getenv("JEN1_TRACE")) fprintf(stderr, "[jen1] fused transformer: %d clusters of 16 CTAs can be resident (smem %d B)\n", n, p.smem_bytes);
    }
    if (active16 < p.B2) p.CS = 8;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.CS, p.B2, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = p.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)p.CS;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, tr_umma_kernel, p);
}

}  // namespace jen1
