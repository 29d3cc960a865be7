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
//     padding rows AFTER the activation (reference blocks.py:137-145 -> :44-51) and store bf16 into a K-major,
//     128-byte-swizzled "row panel" (one 128-byte row = 64 channels of one position).  A conv tap is a whole-row
//     shift of the descriptor start address, so the k taps reuse one panel; strided down-convs keep one
//     sub-panel per residue (row mod stride).
//   * Batch rows are folded into the position axis with a per-row halo (q = b*Lq + m, Lq = Lm + halo), so the deep
//     UNet levels (L = 1..24) still fill an MMA N tile, and their weight streaming is spread over the whole GPU
//     by split-K ACROSS A THREAD-BLOCK CLUSTER: the splitk CTAs of one output tile form a cluster (1,1,splitk),
//     park their fp32 partial tiles in their own shared memory, and after one cluster barrier every CTA reduces
//     and finishes a column slice of the tile by reading all partials over distributed shared memory in fixed
//     order (deterministic) -- no global workspace, no atomics, and the epilogue itself is spread over the cluster.
//   * Everything that is pure index arithmetic (which input row / batch row / panel row a panel slot maps to, which
//     output row a column maps to, tap geometry) and every weight-only operand (bias, GroupNorm gamma/beta) is
//     tabulated in shared memory BEFORE griddepcontrol.wait, i.e. off the critical path of the layer chain; after
//     the wait the producers only do table look-ups, 16-byte loads and FMAs.  Loads of panel unit k+1 are in flight
//     while unit k is transformed.
//
// Warp roles (192 threads): warps 0-3 build panels, then run the epilogue (TMEM lane quarter = warp index);
// warp 4 streams weights; warp 5 allocates TMEM and issues tcgen05.mma (one elected lane each, warp-uniform control flow).
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
constexpr int kMaxCluster = 16;
constexpr int kSredBytes = kMaxSlots * 128 * 2 * 4;
// fixed part of the "misc" shared-memory block (see the carve-up in the kernel)
constexpr int kMiscBar = 256;
constexpr int kMiscStat = 2 * kMaxSlots * 32 * 4;            // gmean, grstd
constexpr int kMiscFine = kMaxSlots * 2 * 32 * 2 * 4;        // fine-group sums
constexpr int kMiscFixed = kMiscBar + kMiscStat + kMiscFine + 32 * 8 + 64 * 8;  // + tap table + group ranges

int g_max_cluster = 8;

struct UmmaArgs {
  ConvParams p;
  UmmaPlan pl;
  const bf16* w0;  // packed blobs of seg 0: [m_tile][phase][cin block][tap]
  const bf16* w1;  // packed blobs of seg 1: [m_tile][cin block]
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
// JEN1_ISSUER_WAIT (A/B, profiles/r02_ab_issue_and_planner.txt): how the weight-stream / MMA warps wait.  0 = pure polling
// (test_wait), 1 = suspending try_wait for the waits completed by thread arrivals (p_full) only, 2 = for all their waits.
// The issuer warps share schedulers 0 and 1 with producer warps 0 and 1 of this CTA and of the co-resident one (whose issuer
// warps already wait while this kernel runs): polling there takes issue slots from the warps on the critical path.
// Measured: 2 is 1.1 % / 1.7 % faster per step than 0 (config 3 / 2), 1 is neutral.  (The late wake-up of try_wait noted at
// mbar_wait was observed with the former single-lane issuers; the producer / epilogue warps keep polling.)
#ifndef JEN1_ISSUER_WAIT
#define JEN1_ISSUER_WAIT 2
#endif
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void issuer_wait(uint64_t* bar, uint32_t parity, int kind) {  // kind 0: thread arrivals, 1: async completions
#if JEN1_ISSUER_WAIT == 2
  mbar_wait_suspend(bar, parity);
#elif JEN1_ISSUER_WAIT == 1
  if (kind == 0) mbar_wait_suspend(bar, parity); else mbar_wait(bar, parity);
#else
  mbar_wait(bar, parity);
#endif
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
// one lane of a converged warp (elect.sync): the issue pattern the compiler treats as uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// Cluster barrier for shared-memory hand-offs.  `barrier.cluster.arrive.release` compiles to MEMBAR.ALL.GPU (waits for
// every outstanding global access of the thread: ~1 us in the layer chain); the data exchanged here lives in shared
// memory only, so a CTA-scope fence (the stores are performed at this SM's shared memory, the one point every remote
// ld.shared::cluster of them goes through) followed by the relaxed barrier is sufficient.
// JEN1_CLUSTER_CANONICAL (A/B, profiles/r02_ab_cluster_fence.txt): 1 = the PTX-model form, barrier.cluster.arrive.release
// + wait.acquire, for the data hand-off barrier; 2 = also for the closing barrier.  0 = CTA-scope fence + relaxed arrive.
#ifndef JEN1_CLUSTER_CANONICAL
#define JEN1_CLUSTER_CANONICAL 0
#endif
__device__ __forceinline__ void cluster_sync_all() {
#if JEN1_CLUSTER_CANONICAL >= 1
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
  asm volatile("fence.acq_rel.cta;\n\tbarrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void cluster_arrive_relaxed() {
#if JEN1_CLUSTER_CANONICAL >= 2
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
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

// K-major, 128-byte swizzle: rows of 128 B (64 bf16), the 16-byte chunk c of row r lives at chunk c ^ (r & 7);
// SBO = 1024 B between 8-row groups; K advances inside the swizzle atom by adding bytes to the start address.
// (cute::UMMA::SmemDescriptor, version 1.)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 -- one issue slot for two lanes of arithmetic; the producer warps are
// bound by their own instruction latency, one warp per scheduler)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// SiLU whose result is rounded to bf16 right away: x*sigmoid(x) = h + h*tanh(h), h = x/2, with the single-instruction
// hardware tanh (MUFU.TANH, rel. error ~2^-11 -- below bf16's 2^-9 rounding step)
__device__ __forceinline__ float tanh_fast(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return t;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

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
  const int SK = pl.splitk;
  const int z = (SK > 1) ? (int)blockIdx.z / SK : (int)blockIdx.z;
  const int sk = (SK > 1) ? (int)cluster_ctarank() : 0;  // cluster = (1,1,SK): rank == blockIdx.z % SK
  const int NT = pl.NT, Lq = pl.Lq;
  const int q0 = nt * NT;
  const int nsteps = pl.steps0 + pl.steps1;
  const int st0 = (int)(((long long)sk * nsteps) / SK);
  const int st1 = (int)(((long long)(sk + 1) * nsteps) / SK);
  const int my_steps = st1 - st0;
  const int ntaps0 = p.seg[0].ntaps;
  const int f0 = p.seg[0].in_stride;
  const int rows0 = f0 * pl.R;  // panel slots of a seg-0 K step: idx = rho * R + r
  // this CTA's slice of the seg-0 input channels
  const int ch_base = st0 * 64;
  const int my_ch = (min(st1, pl.steps0) - st0) * 64;  // <= 0: no seg-0 step in this split

  // ---- shared memory carve-up
  uint8_t* a_ring = smem;                                  // pl.ring_bytes (>= stages * 16 KB)
  uint8_t* panels = a_ring + (size_t)pl.ring_bytes;       // 2 * panel_bytes
  const uint32_t panel_bytes = (uint32_t)pl.panel_bytes;
  uint8_t* misc = panels + 2 * (size_t)panel_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);            // [8]
  uint64_t* a_empty = a_full + 8;                                   // [8]
  uint64_t* p_full = a_empty + 8;                                   // [2]
  uint64_t* p_empty = p_full + 2;                                   // [2]
  uint64_t* acc_full = p_empty + 2;                                 // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);  // [1]
  float* gmean = reinterpret_cast<float*>(misc + kMiscBar);         // [kMaxSlots][32]
  float* grstd = gmean + kMaxSlots * 32;                            // [kMaxSlots][32]
  float* fine = grstd + kMaxSlots * 32;                             // [kMaxSlots][2 sources][32 fine groups][2]
  int2* tapg = reinterpret_cast<int2*>(fine + kMaxSlots * 2 * 32 * 2);  // [32] (panel row offset, unused) per tap
  int2* grange = tapg + 32;                                         // [32 groups][2 sources] fine-group range
  uint8_t* tabs = misc + kMiscFixed;
  // [rows0]: (element offset of the input row in source 0 | -1, ... in source 1, b | panel row << 8, row index in source 0)
  int4* rowmeta = reinterpret_cast<int4*>(tabs);
  int4* rowmeta1 = reinterpret_cast<int4*>(tabs + pl.off_rowmeta1);        // [NT] same for the seg-1 K steps
  int4* colmeta = reinterpret_cast<int4*>(tabs + pl.off_colmeta);          // [NT]: (out offset | -1, residual offset, slot, batch row)
  float2* rowstat = reinterpret_cast<float2*>(tabs + pl.off_rowstat);      // [rows0] LayerNorm (mean, rstd) per slot
  // [slots][ch_cap / 2] (a_even, a_odd, s_even, s_odd): y = a*x + s per channel pair (operands of one FFMA2); with SiLU the
  // halved coefficients, i.e. h = x/2 of silu(x) = h + h*tanh(h)
  float* coef = reinterpret_cast<float*>(tabs + pl.off_coef);
  const int ch_cap = pl.ch_cap;
  // epilogue scratch aliases the weight ring (all MMAs have completed by then)
  // [0, kSredBytes): per-warp GroupNorm fine-group sums; then the fp32 staging tile [columns][128] (the whole partial
  // tile with split-K, one chunk of columns otherwise)
  float* part = reinterpret_cast<float*>(a_ring + kSredBytes);

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

  // weight-only / pre-chain values of the epilogue threads (thread == output channel of the M tile)
  const float4 bias4 = (warp < 4 && p.bias) ? __ldg(reinterpret_cast<const float4*>(p.bias + mt * 128) + lane)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);

  // Warps 4 and 5 run their (warp-uniform) control flow with all lanes and ONE ELECTED lane issues the bulk copies / MMAs:
  // under a divergent `lane == 0` branch the compiler wraps every UBLKCP / UTCHMMA / UTCBAR in an elect-and-branch loop and
  // rebuilds its operands per instruction (measured in attn_flash.cu: ~85 clocks per MMA, slower than the tensor pipe).
  if (warp == 4) {
    // ======================================================================== weight streamer
    int s = 0, k = 0;
    for (int t = st0; t < st1; ++t) {
      const bool s1 = t >= pl.steps0;
      const int ntp = s1 ? 1 : ntaps0;
      const bf16* src = s1 ? A.w1 + ((size_t)mt * pl.steps1 + (t - pl.steps0)) * (kABytes / 2)
                           : A.w0 + (((size_t)(mt * p.nphase + z) * pl.steps0 + t) * ntaps0) * (kABytes / 2);
      for (int j = 0; j < ntp; ++j) {
        if (k > 0) issuer_wait(&a_empty[s], (uint32_t)((k - 1) & 1), 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[s], kABytes);
          bulk_g2s(a_ring + (size_t)s * kABytes, src + (size_t)j * (kABytes / 2), kABytes, &a_full[s]);
        }
        __syncwarp();
        if (++s == pl.stages) {
          s = 0;
          ++k;
        }
      }
    }
  } else if (warp == 5) {
    // ======================================================================== MMA issuer
    // tap geometry: panel row offset of tap j (sub-panel of its residue + whole-row shift); one tap per lane
    if (lane < ntaps0) {
      const int d = p.seg[0].shift0 + lane * p.seg[0].shift_step;
      int rho = d % f0;
      if (rho < 0) rho += f0;
      const int a = (d - rho) / f0;
      tapg[lane] = make_int2(rho * pl.PS + (a - pl.amin), 0);
    }
    __syncwarp();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad0 = make_desc_sw128(smem_u32(a_ring)), bd0 = make_desc_sw128(smem_u32(panels));
    int s = 0, k = 0;
    uint32_t acc = 0;
    for (int t = st0; t < st1; ++t) {
      const int n = t - st0, pb = n & 1;
      const bool s1 = t >= pl.steps0;
      const int ntp = s1 ? 1 : ntaps0;
      issuer_wait(&p_full[pb], (uint32_t)((n >> 1) & 1), 0);
      tc_fence_after();
      const uint32_t pbase16 = (uint32_t)pb * (panel_bytes >> 4);  // descriptor address units (16 bytes)
      for (int j = 0; j < ntp; ++j) {
        const uint32_t prow = s1 ? 0u : (uint32_t)tapg[j].x;
        issuer_wait(&a_full[s], (uint32_t)(k & 1), 1);
        tc_fence_after();
        if (t == st0 && j == 0 && lane == 0) TL_MARK(9);
        // K advances by 32 bytes inside the swizzle atom: +2 in the (address >> 4) field of the descriptor
        const uint64_t ad = ad0 + (uint64_t)((uint32_t)s * (kABytes >> 4));
        const uint64_t bd = bd0 + (uint64_t)(pbase16 + prow * 8u);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (acc | (uint32_t)kk) ? 1u : 0u);
          umma_commit(&a_empty[s]);
        }
        __syncwarp();
        acc = 1;
        if (++s == pl.stages) {
          s = 0;
          ++k;
        }
      }
      if (elect_one()) umma_commit(&p_empty[pb]);
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
    if (lane == 0) TL_MARK(10);
  } else {
    // ======================================================================== panel producers, then epilogue
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
    const bool need_coef = (has_gn || has_film) && my_ch > 0;
    const int cpg = has_gn ? Ct / p.G : 1;
    const int zoff = p.out_off0 + z * p.out_off_phase;
    const float hs = (affine && p.act == ACT_SILU) ? 0.5f : 1.0f;  // SiLU works on x/2: folded into the coefficients

    // ---- tables that do not depend on earlier kernels (built while the previous layer is still running)
    // (the conditioning rows are written before the step's kernel chain starts)
    // (q / Lq by a short subtract loop from the tile's first batch row: a tile touches <= kMaxSlots rows, and integer
    //  division is ~40 instructions of pure latency on this path)
    const int ml_first = q0 - b_first * Lq;
    auto split_q = [&](int r, int& b, int& ml) {  // q0 + r = b * Lq + ml
      b = b_first;
      ml = ml_first + r;
      while (ml >= Lq) {
        ml -= Lq;
        ++b;
      }
    };
    for (int idx = tid; idx < rows0; idx += kProducers) {
      int rho = 0, r = idx;
      while (r >= pl.R) {  // idx = rho * R + r, rho < in_stride
        r -= pl.R;
        ++rho;
      }
      int b, ml;
      split_q(r, b, ml);
      const int irow = (ml + pl.amin) * f0 + rho;
      const bool ok = b < p.B && irow >= 0 && irow < S0.L;
      int b0 = b, b1 = b;
      if (b0 >= S0.s[0].bmod) b0 -= S0.s[0].bmod;
      if (b1 >= S0.s[1].bmod) b1 -= S0.s[1].bmod;
      rowmeta[idx] = make_int4(ok ? (b0 * S0.L + irow) * S0.s[0].C : -1, ok ? (b1 * S0.L + irow) * S0.s[1].C : 0,
                               (b & 255) | ((rho * pl.PS + r) << 8), ok ? b0 * S0.L + irow : 0);
    }
    if (p.nseg > 1) {
      for (int r = tid; r < NT; r += kProducers) {
        int b, ml;
        split_q(r, b, ml);
        const bool ok = b < p.B && ml < p.seg[1].L;
        const ConvSeg& S1 = p.seg[1];
        int b0 = b, b1 = b;
        if (b0 >= S1.s[0].bmod) b0 -= S1.s[0].bmod;
        if (b1 >= S1.s[1].bmod) b1 -= S1.s[1].bmod;
        rowmeta1[r] = make_int4(ok ? (b0 * S1.L + ml) * S1.s[0].C : -1, ok ? (b1 * S1.L + ml) * S1.s[1].C : 0,
                                (b & 255) | (r << 8), 0);
      }
    }
    for (int c = tid; c < NT; c += kProducers) {
      int eb, eml;
      split_q(c, eb, eml);
      const int o = eml * p.out_stride + zoff;
      const bool valid = (eb < p.B) && (eml < p.Lm) && (o >= 0) && (o < p.Lout);
      int rb = eb;
      if (rb >= p.res_bmod) rb -= p.res_bmod;
      colmeta[c] = make_int4(valid ? (eb * p.Lout + o) * p.Cout : -1, valid ? (rb * p.Lout + o) * p.Cout : 0,
                             eb - b_first, valid ? eb * p.Lout + o : 0);
    }
    // weight-only halves of the affine coefficients: P = gamma*(1+film_scale), Q = beta*(1+film_scale) + film_shift
    // (the FiLM table and the conditioning rows are written before the step's kernel chain starts)
    if (need_coef) {
      for (int bl = 0; bl < nbl; ++bl) {
        const int crow = p.cond_row ? __ldg(p.cond_row + b_first + bl) : 0;
        const float* fp = has_film ? p.film + (size_t)crow * p.film_stride : nullptr;
        for (int c = tid; c < my_ch; c += kProducers) {
          const int ch = ch_base + c;
          float P = 0.f, Q = 0.f;
          if (ch < Ct) {
            P = has_gn ? __ldg(p.gamma + ch) : 1.0f;
            Q = has_gn ? __ldg(p.beta + ch) : 0.0f;
            if (has_film) {
              const float f1 = __ldg(fp + ch) + 1.0f;
              P *= f1;
              Q = fmaf(Q, f1, __ldg(fp + Ct + ch));
            }
            if (!has_gn) {  // final: no statistics to wait for
              P *= (ch >= S0.s[0].C ? S0.s[1] : S0.s[0]).scale * hs;
              Q *= hs;
            }
          }
          float* cd = coef + ((size_t)bl * ch_cap + (c & ~1)) * 2 + (c & 1);
          cd[0] = P;
          cd[2] = Q;
        }
      }
    }
    if (has_gn && tid < p.G) {  // fine-group range of GroupNorm group `tid` in each source
      const int lo = tid * cpg, hi = lo + cpg;
      int off = 0;
      for (int sI = 0; sI < 2; ++sI) {
        const ConvSrc& sr = S0.s[sI];
        int2 rg = make_int2(0, 0);
        if (sr.C > 0) {
          const int olo = max(lo, off), ohi = min(hi, off + sr.C);
          if (ohi > olo) {
            const int gsz = sr.C / sr.FG;
            rg = make_int2((olo - off) / gsz, (ohi - off) / gsz);
          }
        }
        grange[tid * 2 + sI] = rg;
        off += sr.C;
      }
    }
    const float inv_n = has_gn ? 1.0f / ((float)(p.gn_real_c > 0 ? p.gn_real_c / p.G : cpg) * (float)S0.L) : 0.f;
    const int cpg_shift = (has_gn && (cpg & (cpg - 1)) == 0) ? __ffs(cpg) - 1 : -1;
    bar_sync_producers();

    // ---- panel unit helpers.  A unit = up to 128 panel slots (rows) of one K step; a thread owns the 16-byte
    //      channel chunk `kc` of slots rr + 16u.
    auto step_rows = [&](int n) { return (st0 + n >= pl.steps0) ? NT : rows0; };
    auto issue = [&](int n, int ib, uint4 (&raw)[8], uint32_t& okm) {
      const int t = st0 + n;
      const bool s1 = t >= pl.steps0;
      const ConvSeg& S = p.seg[s1 ? 1 : 0];
      const int c0 = (s1 ? t - pl.steps0 : t) * 64 + kc * 8;
      const bool second = c0 >= S.s[0].C;
      const ConvSrc& sr = second ? S.s[1] : S.s[0];
      const int cc = second ? c0 - S.s[0].C : c0;
      const bool chan_ok = cc < sr.C;
      const int4* meta = s1 ? rowmeta1 : rowmeta;
      const int rows = s1 ? NT : rows0;
      const bf16* base = (const bf16*)sr.ptr + cc;
      okm = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = ib * 128 + rr + 16 * u;
        if (idx >= rows) break;
        raw[u] = make_uint4(0u, 0u, 0u, 0u);
        if (chan_ok) {
          const int4 m = meta[idx];
          if (m.x >= 0) {
            raw[u] = __ldcg(reinterpret_cast<const uint4*>(base + (size_t)(uint32_t)(second ? m.y : m.x)));
            okm |= 1u << u;
          }
        }
      }
    };
    auto consume = [&](int n, int ib, const uint4 (&raw)[8], uint32_t okm) {
      const int t = st0 + n, pb = n & 1;
      const bool s1 = t >= pl.steps0;
      const ConvSeg& S = p.seg[s1 ? 1 : 0];
      const int c0 = (s1 ? t - pl.steps0 : t) * 64 + kc * 8;
      const bool second = c0 >= S.s[0].C;
      const float sscale = (second ? S.s[1] : S.s[0]).scale;
      const int4* meta = s1 ? rowmeta1 : rowmeta;
      const int rows = s1 ? NT : rows0;
      if (ib == 0 && n >= 2) mbar_wait(&p_empty[pb], (uint32_t)(((n >> 1) - 1) & 1));
      uint8_t* pan = panels + (size_t)pb * panel_bytes;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = ib * 128 + rr + 16 * u;
        if (idx >= rows) break;
        const int mz = meta[idx].z;
        const int prow = mz >> 8;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if ((okm >> u) & 1u) {
          const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
          uint64_t x2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) x2[e] = pk2(__uint_as_float(w4[e] << 16), __uint_as_float(w4[e] & 0xffff0000u));
          if (s1) {
            const uint64_t sc2 = pk2(sscale, sscale);
#pragma unroll
            for (int e = 0; e < 4; ++e) x2[e] = fmul2(x2[e], sc2);
          } else if (affine) {
            if (need_coef) {
              const float4* cf = reinterpret_cast<const float4*>(coef + ((size_t)((mz & 255) - b_first) * ch_cap + (c0 - ch_base)) * 2);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 kq = cf[e];
                x2[e] = ffma2(pk2(kq.x, kq.y), x2[e], pk2(kq.z, kq.w));
              }
            } else {
              const uint64_t sc2 = pk2(sscale * hs, sscale * hs);
#pragma unroll
              for (int e = 0; e < 4; ++e) x2[e] = fmul2(x2[e], sc2);
            }
            if (p.act == ACT_SILU) {  // x2 holds h = x/2: silu(x) = h + h*tanh(h) (hardware tanh)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float h0, h1;
                upk2(x2[e], h0, h1);
                x2[e] = ffma2(x2[e], pk2(tanh_fast(h0), tanh_fast(h1)), x2[e]);
              }
            }
          } else {
            const float2 ms = rowstat[idx];
            const uint64_t nm2 = pk2(-ms.x, -ms.x), rs2 = pk2(ms.y, ms.y);
#pragma unroll
            for (int e = 0; e < 4; ++e) x2[e] = fmul2(fadd2(x2[e], nm2), rs2);
          }
          float y0, y1;
          upk2(x2[0], y0, y1);
          o.x = pack2(y0, y1);
          upk2(x2[1], y0, y1);
          o.y = pack2(y0, y1);
          upk2(x2[2], y0, y1);
          o.z = pack2(y0, y1);
          upk2(x2[3], y0, y1);
          o.w = pack2(y0, y1);
        }
        *reinterpret_cast<uint4*>(pan + (size_t)prow * 128 + (size_t)((kc ^ (prow & 7)) * 16)) = o;
      }
      if ((ib + 1) * 128 >= rows) {  // last unit of this K step
        fence_async_smem();           // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(&p_full[pb]);
      }
    };
    auto advance = [&](int& n, int& ib) {
      if ((ib + 1) * 128 >= step_rows(n)) {
        ib = 0;
        ++n;
      } else {
        ++ib;
      }
    };

    pdl_wait();  // everything below reads what the previous kernels wrote
    if (tid == 0) { TL_MARK(2); TL_GLOBAL(11); }

    // ---- panel pipeline.  One call site each for the loads (`issue`) and the transform (`consume`) keeps the kernel's
    //      instruction footprint small; the loads of unit k+1 are in flight while unit k is transformed, and the
    //      loads of unit 0 are in flight during the statistics reduction of the first pass.
    uint4 rA[8], rB[8];
    uint32_t mA = 0, mB = 0;
    int cn = 0, cib = 0;  // unit being consumed
    int in_ = 0, iib = 0; // next unit to load
    bool primed = false;
    for (;;) {
      const bool can_issue = in_ < my_steps;
      if (can_issue) issue(in_, iib, rB, mB);
      if (primed) {
        consume(cn, cib, rA, mA);
      } else {
        if (!affine) {
          // LayerNorm statistics of every panel row from the producer's per-tile partials (once per CTA)
          const ConvSrc& sr = S0.s[0];
          const float inv = 1.0f / (float)sr.C;
          for (int idx = tid; idx < rows0; idx += kProducers) {
            const int4 m = rowmeta[idx];
            float2 ms = make_float2(0.f, 1.f);
            if (m.x >= 0) {
              const float2* rp = reinterpret_cast<const float2*>(p.rowpart) + (size_t)m.w * p.rp_nct;
              float a = 0.f, qq = 0.f;
              for (int j0 = 0; j0 < p.rp_nct; j0 += 8) {
                float2 v2[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) v2[jj] = (j0 + jj < p.rp_nct) ? __ldcg(rp + j0 + jj) : make_float2(0.f, 0.f);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  a += v2[jj].x;
                  qq += v2[jj].y;
                }
              }
              const float m1 = a * inv;
              float var = qq * inv - m1 * m1;
              if (var < 0.0f) var = 0.0f;
              ms = make_float2(m1, 1.0f / sqrtf(var + p.ln_eps));
            }
            rowstat[idx] = ms;
          }
        }

        // ---- GroupNorm statistics of the input for the batch rows this tile touches.  Deterministic two-level reduce:
        //      fine-group sums of every (batch row, source, fine group) item from the producer's partial entries (an item
        //      is split over `parts` lanes combined by shuffles in a fixed order), then one thread per group.
        if (has_gn) {
          // one item per (batch row, source, fine group): a single 16-byte load of the producers' fixed-point
          // accumulators (common.cuh) -- no per-tile entries to reduce
          const int two_src = S0.s[1].C > 0 ? 1 : 0;
          const int nitem = nbl * (32 << two_src);
          for (int idx = tid; idx < nitem; idx += kProducers) {
            const int bl = idx >> (5 + two_src), fs = two_src ? (idx >> 5) & 1 : 0, fg = idx & 31;
            const ConvSrc& fsr = S0.s[fs];
            float2 v = make_float2(0.f, 0.f);
            if (fsr.C > 0 && fg < fsr.FG) {
              int b = b_first + bl;
              if (b >= fsr.bmod) b -= fsr.bmod;
              const longlong2 acc = __ldcg(reinterpret_cast<const longlong2*>(fsr.stats + ((size_t)b * fsr.FG + fg) * 2));
              const float sc = fsr.scale;
              v = make_float2(stat_get(acc.x) * sc, stat_get(acc.y) * sc * sc);
            }
            *reinterpret_cast<float2*>(fine + ((bl * 2 + fs) * 32 + fg) * 2) = v;
          }
          bar_sync_producers();
          // pass 2: one thread per (batch row, group)
          for (int idx = tid; idx < nbl * p.G; idx += kProducers) {
            const int bl = idx / p.G, g = idx - bl * p.G;
            float ga = 0.f, gq = 0.f;
#pragma unroll
            for (int sI = 0; sI < 2; ++sI) {
              const int2 rg = grange[g * 2 + sI];
              for (int fg = rg.x; fg < rg.y; ++fg) {
                ga += fine[((bl * 2 + sI) * 32 + fg) * 2];
                gq += fine[((bl * 2 + sI) * 32 + fg) * 2 + 1];
              }
            }
            // fp32 is ample here: this kernel only serves bf16 storage (the strict fp32 mode runs the generic kernel)
            const float mean = ga * inv_n;
            float var = fmaf(-mean, mean, gq * inv_n);
            if (var < 0.0f) var = 0.0f;
            gmean[bl * 32 + g] = mean;
            grstd[bl * 32 + g] = rsqrtf(var + p.eps);
          }
        }
        bar_sync_producers();  // gmean / grstd, rowstat
        if (tid == 0) TL_MARK(3);

        // ---- per-(batch row, channel) affine coefficients of this CTA's channel slice, finished in place:
        //      y = a*x + s,  a = scale*rstd*P,  s = Q - mean*rstd*P   (P, Q tabulated before the wait)
        if (need_coef && has_gn) {
          for (int bl = 0; bl < nbl; ++bl) {
            for (int c = tid; c < my_ch; c += kProducers) {
              const int ch = ch_base + c;
              const int g = cpg_shift >= 0 ? ch >> cpg_shift : ch / cpg;
              float* cd = coef + ((size_t)bl * ch_cap + (c & ~1)) * 2 + (c & 1);
              const float rp = (ch < Ct) ? grstd[bl * 32 + g] * cd[0] : 0.f;
              const float scale = (ch >= S0.s[0].C ? S0.s[1] : S0.s[0]).scale;
              const float sh = (ch < Ct) ? fmaf(-gmean[bl * 32 + g], rp, cd[2]) : 0.f;
              cd[0] = rp * scale * hs;
              cd[2] = sh * hs;
            }
          }
          bar_sync_producers();
        }
        if (tid == 0) TL_MARK(4);
      }
      if (!can_issue) break;
#pragma unroll
      for (int u = 0; u < 8; ++u) rA[u] = rB[u];
      mA = mB;
      cn = in_;
      cib = iib;
      advance(in_, iib);
      primed = true;
    }
    if (tid == 0) TL_MARK(5);
  }

  // ============================================================================================ epilogue
  // One COMPACT code path for every case (the instruction cache is 32 KB per SM and this code runs once per CTA: rolled
  // loops, no per-case copies).  The fp32 accumulator tile leaves TMEM (lane == output channel) through a shared-memory
  // staging tile [column][128 channels]; it is then finished "row-wise": a warp owns one output COLUMN per round and each
  // lane 4 consecutive channels, so
  //   * split-K partials of all cluster ranks are read with 16-byte DSMEM loads (rank order: deterministic),
  //   * bias / GELU / residual / bf16 conversion work on 4 channels at a time and the store is one 256-byte
  //     contiguous row segment per warp instead of 2-byte scattered stores,
  //   * the residual rows arrive through cp.async in a shared-memory tile while the accumulator is being staged,
  //   * GroupNorm fine-group partials are combined with a fixed shuffle pattern inside the warp and LayerNorm row
  //     partials are a plain warp reduction.
  // Without split-K the tile is staged in chunks of CH columns (what the ring can hold, residual double-buffered); with
  // split-K every CTA stages its whole partial tile, one cluster barrier makes all of them visible, and each CTA finishes
  // columns [cb, ce).
  const int cols_per = (SK > 1) ? (NT + SK - 1) / SK : NT;
  const int cb = (SK > 1) ? min(NT, sk * cols_per) : 0;
  const int ce = min(NT, cb + cols_per);
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const bool want_stats = p.stats_out != nullptr;
  const bool want_rows = p.rowpart_out != nullptr;
  const int b_first_e = q0 / Lq;
  const int nb_out = min(p.B - 1, (q0 + NT - 1) / Lq) - b_first_e + 1;
  const int gs = want_stats ? p.Cout / p.FGo : 128;  // channels per fine group of the output (4, 8, 16 or 32)
  const int ngl = 128 / gs;                           // fine groups inside this M tile
  const int lpg = gs >> 2;                            // lanes per fine group
  float* sfg = reinterpret_cast<float*>(a_ring);      // [4 warps][kMaxSlots][32 fine groups][2]  (kSredBytes)
  // staged columns per chunk: split-K stages the whole tile; otherwise 1 KB per column (fp32 row + two bf16 residual rows)
  const int CH = (SK > 1) ? NT : min(NT, ((pl.ring_bytes - kSredBytes) >> 10) & ~15);
  uint8_t* resbuf = reinterpret_cast<uint8_t*>(part + (size_t)CH * 128);  // [2 (SK == 1)][columns][128] bf16
  const int wq = warp & 3, q4 = lane * 4;
  const unsigned short* resp = reinterpret_cast<const unsigned short*>(p.res);
  // residual rows of columns [c0, c0 + n) -> resbuf (16 bytes per cp.async, fire-and-forget)
  auto res_fetch = [&](int c0, int n, uint8_t* dst) {
    if (resp == nullptr) return;
    for (int i = tid; i < n * 16; i += kProducers) {
      const int4 cm = colmeta[c0 + (i >> 4)];
      if (cm.x >= 0) {
        const unsigned short* src = resp + (size_t)(uint32_t)cm.y + mt * 128 + (i & 15) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + (size_t)i * 16)), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (warp < 4) {
    mbar_wait(acc_full, 0);  // every MMA has completed: the ring (now scratch) is free, the accumulator is final
    tc_fence_after();
    if (tid == 0) TL_MARK(6);
    // (the residual tile lives in the ring: it can only be requested now; its latency overlaps the TMEM staging and the
    //  cluster exchange)
    res_fetch(cb, (SK > 1) ? ce - cb : min(CH, NT), resbuf);
    if (want_stats) {
      float4* zf = reinterpret_cast<float4*>(sfg);
      for (int i = tid; i < kSredBytes / 16; i += kProducers) zf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (SK > 1) {
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) part[(size_t)(c0 + j) * 128 + tid] = v[j];
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
  }
  if (SK > 1) {
    if (tid == 0) TL_MARK(16);
    cluster_sync_all();  // every CTA's partial tile is visible cluster-wide (and this CTA's residual tile CTA-wide)
    if (tid == 0) TL_MARK(7);
  }

  if (warp < 4) {
    const uint32_t mine = smem_u32(part);
    int sb = -1;  // slot (batch row of the tile) of the statistics run in progress
    float aS[4] = {0.f, 0.f, 0.f, 0.f}, aQ[4] = {0.f, 0.f, 0.f, 0.f};
    auto flush_stats = [&]() {  // warp-uniform: the fine-group sums of this warp's run -> its own (warp, slot) cell
      if (want_stats && sb >= 0 && sb < nb_out && sb < kMaxSlots) {
        float s1 = (aS[0] + aS[1]) + (aS[2] + aS[3]);
        float s2 = (aQ[0] + aQ[1]) + (aQ[2] + aQ[3]);
        for (int o = 1; o < lpg; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if ((lane & (lpg - 1)) == 0) {
          float2* d = reinterpret_cast<float2*>(sfg) + ((wq * kMaxSlots + sb) * 32 + lane / lpg);
          *d = make_float2(d->x + s1, d->y + s2);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        aS[e] = 0.f;
        aQ[e] = 0.f;
      }
    };
    int rbsel = 0;
#pragma unroll 1
    for (int cs = 0; cs < ce - cb; cs += CH) {  // one chunk with split-K
      const uint8_t* rcur = resbuf + (size_t)rbsel * CH * 256;
      if (SK == 1) {
        if (cs > 0) bar_sync_producers();  // the previous chunk has been consumed (staging tile + other residual buffer)
        const int cn = min(CH, NT - cs);
        for (int c0 = 0; c0 < cn; c0 += 16) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(cs + c0), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) part[(size_t)(c0 + j) * 128 + tid] = v[j];
        }
        asm volatile("cp.async.wait_all;" ::: "memory");  // this chunk's residual rows have landed
        bar_sync_producers();
        if (cs + CH < NT) res_fetch(cs + CH, min(CH, NT - cs - CH), resbuf + (size_t)(rbsel ^ 1) * CH * 256);
        rbsel ^= 1;
      }
      const int c_end = min(ce, cb + cs + CH);
#pragma unroll 1
      for (int c = cb + cs + wq; c < c_end; c += 4) {  // warp-uniform
        const int4 cm = colmeta[c];
        float4 acc;
        if (SK == 1) {
          acc = *reinterpret_cast<const float4*>(part + (size_t)(c - cs) * 128 + q4);
        } else {
          const uint32_t off = mine + (uint32_t)((c * 128 + q4) * 4);
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int s0 = 0; s0 < SK; s0 += 4) {  // four ranks in flight, summed in rank order: deterministic
            float4 t0 = ld_cluster_f32x4(map_cluster(off, (uint32_t)s0)), t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1, t3 = t1;
            if (s0 + 1 < SK) t1 = ld_cluster_f32x4(map_cluster(off, (uint32_t)(s0 + 1)));
            if (s0 + 2 < SK) t2 = ld_cluster_f32x4(map_cluster(off, (uint32_t)(s0 + 2)));
            if (s0 + 3 < SK) t3 = ld_cluster_f32x4(map_cluster(off, (uint32_t)(s0 + 3)));
            acc.x = (((acc.x + t0.x) + t1.x) + t2.x) + t3.x;
            acc.y = (((acc.y + t0.y) + t1.y) + t2.y) + t3.y;
            acc.z = (((acc.z + t0.z) + t1.z) + t2.z) + t3.z;
            acc.w = (((acc.w + t0.w) + t1.w) + t2.w) + t3.w;
          }
        }
        if (cm.z != sb) {
          flush_stats();
          sb = cm.z;
        }
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        if (cm.x >= 0) {
          x[0] = acc.x + bias4.x;
          x[1] = acc.y + bias4.y;
          x[2] = acc.z + bias4.z;
          x[3] = acc.w + bias4.w;
          if (p.epi_act == ACT_GELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = gelu_f(x[e]);
          }
          if (resp) {
            const uint2 rr = *reinterpret_cast<const uint2*>(rcur + (size_t)(c - cb - cs) * 256 + lane * 8);
            x[0] += __uint_as_float(rr.x << 16);
            x[1] += __uint_as_float(rr.x & 0xffff0000u);
            x[2] += __uint_as_float(rr.y << 16);
            x[3] += __uint_as_float(rr.y & 0xffff0000u);
          }
          const size_t oo = (size_t)(uint32_t)cm.x + mt * 128 + q4;
          if (A.out_f32) {
            *reinterpret_cast<float4*>((float*)p.out + oo) = make_float4(x[0], x[1], x[2], x[3]);
          } else {
            *reinterpret_cast<uint2*>((bf16*)p.out + oo) = make_uint2(pack2(x[0], x[1]), pack2(x[2], x[3]));
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          aS[e] += x[e];
          aQ[e] = fmaf(x[e], x[e], aQ[e]);
        }
        if (want_rows) {  // LayerNorm partial of this row over the 128 channels of the M tile
          const float rs = warp_sum((x[0] + x[1]) + (x[2] + x[3]));
          const float rq2 = warp_sum(fmaf(x[0], x[0], x[1] * x[1]) + fmaf(x[2], x[2], x[3] * x[3]));
          if (lane == 0 && cm.x >= 0)
            *reinterpret_cast<float2*>(p.rowpart_out + ((size_t)(uint32_t)cm.w * pl.m_tiles + mt) * 2) = make_float2(rs, rq2);
        }
      }
    }
    flush_stats();
    if (tid == 0) TL_MARK(13);
  }
  // this CTA no longer reads remote partial tiles: arrive now, wait only before leaving (nobody may exit while its
  // partial tile can still be read)
  if (SK > 1) cluster_arrive_relaxed();
  if (warp < 4 && want_stats) {
    bar_sync_producers();
    // per (batch row, fine group) partial of this CTA's columns -> fixed-point accumulators (fire-and-forget atomics)
    for (int idx = tid; idx < nb_out * ngl; idx += kProducers) {
      const int bl = idx / ngl, gl = idx - bl * ngl;
      float a = 0.f, q = 0.f;
      if (bl < kMaxSlots) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float2 v = reinterpret_cast<const float2*>(sfg)[(w * kMaxSlots + bl) * 32 + gl];
          a += v.x;
          q += v.y;
        }
      }
      long long* so = p.stats_out + ((size_t)(b_first_e + bl) * p.FGo + (mt * 128) / gs + gl) * 2;
      stat_add(so, a);
      stat_add(so + 1, q);
    }
  }
  if (tid == 0) TL_MARK(14);
  if (SK > 1) cluster_wait();

  if (tid == 0) { TL_MARK(8); TL_GLOBAL(12); }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)pl.tmem_cols)
                 : "memory");
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- planning
UmmaPlan conv_umma_plan(const ConvParams& p, bool want_stats, int num_sms) {
  UmmaPlan pl;
  memset(&pl, 0, sizeof(pl));
  const ConvSeg& S0 = p.seg[0];
  if (p.out == nullptr && p.out_ncl != nullptr) return pl;  // [B][C][L] output is a boundary format: generic path
  if (p.Cout % 128 != 0 || p.B < 1 || p.B > 256) return pl;
  if ((long long)p.B * p.Lout * p.Cout >= (1ll << 31)) return pl;  // 32-bit element offsets in the column table
  for (int sg = 0; sg < p.nseg; ++sg)
    for (int k = 0; k < 2; ++k) {
      const ConvSrc& sr = p.seg[sg].s[k];
      if (sr.C > 0 && (sr.C % 8 != 0)) return pl;
      if (sr.C > 0 && (sr.bmod < 1 || p.B > 2 * sr.bmod)) return pl;  // `b % bmod` is one conditional subtract
      if (sr.C > 0 && (long long)sr.bmod * p.seg[sg].L * sr.C >= (1ll << 31)) return pl;  // 32-bit offsets in the row table
    }
  if (p.res && (p.res_bmod < 1 || p.B > 2 * p.res_bmod)) return pl;
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
  // ---- tile / split-K selection.  The layer chain is latency bound, so the plan minimises the critical path of ONE
  //      CTA instead of maximising the CTA count: K steps per CTA (each costs a load -> transform -> MMA round),
  //      weight bytes a CTA must stream beyond what the ring prefetches under PDL, panel rows and epilogue columns.
  //      Split-K (<= cluster size) is preferred over narrow N tiles: it divides both the K loop and the weight stream,
  //      whereas every extra N tile re-streams the whole weight slice from L2.
  int nt_cap = 256 / f;
  if (nt_cap < 16) nt_cap = 16;
  if ((long long)nt_cap > round_up((int)(nq < 16 ? 16 : nq), 16)) nt_cap = round_up((int)(nq < 16 ? 16 : nq), 16);
  auto slots = [&](int nt) {  // distinct batch rows per tile must fit the per-slot tables
    const int sl = (nt + pl.halo + pl.Lq - 1) / pl.Lq + 1;
    return sl < p.B ? sl : p.B;
  };
  const int nsteps = pl.steps0 + pl.steps1;
  const bool need_coef = p.mode == PRO_AFFINE && (p.G > 0 || p.film != nullptr);
  const int ntaps = S0.ntaps;
  auto configure = [&](int NT, UmmaPlan& c) -> bool {
    c = pl;
    if (slots(NT) > kMaxSlots) return false;
    c.NT = NT;
    c.n_tiles = (int)((nq + NT - 1) / NT);
    c.R = NT + c.halo;
    c.PS = round_up(c.R, 8);  // rows per sub-panel (128-byte rows, 128-byte swizzle, 1024-byte aligned)
    c.panel_bytes = f * c.PS * 128;
    if (c.panel_bytes < NT * 128) c.panel_bytes = round_up(NT * 128, 1024);
    int tc = 32;
    while (tc < NT) tc <<= 1;
    c.tmem_cols = tc;
    const long long tiles = (long long)c.n_tiles * c.m_tiles * p.nphase;
    if (tiles > 65535) return false;
    // split-K over a cluster: spread weight streaming over the GPU when the output tiles alone do not fill it
    int sk = (int)(num_sms / tiles);
    if (sk < 1) sk = 1;
    if (sk > nsteps) sk = nsteps;
    if (sk > g_max_cluster) sk = g_max_cluster;
    // the per-slot coefficient table of a CTA's channel slice must stay small: split K further if needed
    const int nslot = slots(NT);
    auto ch_cap_of = [&](int sp) {
      int ms = (nsteps + sp - 1) / sp;
      if (ms > c.steps0) ms = c.steps0;
      return ms * 64;
    };
    const int coef_budget = 24 * 1024;
    while (need_coef && nslot * ch_cap_of(sk) * 8 > coef_budget && sk < nsteps && sk < g_max_cluster) ++sk;
    if (need_coef && nslot * ch_cap_of(sk) * 8 > coef_budget) return false;
    c.splitk = sk;
    c.ch_cap = ch_cap_of(sk);
    // tables
    const int rows0 = f * c.R;
    int off = rows0 * 16;
    c.off_rowmeta1 = off;
    off += p.nseg > 1 ? NT * 16 : 0;
    c.off_colmeta = off;
    off += NT * 16;
    c.off_rowstat = off;
    off += p.mode == PRO_ROWNORM ? round_up(rows0 * 8, 16) : 0;
    c.off_coef = off;
    off += need_coef ? nslot * c.ch_cap * 8 : 0;
    const int misc = kMiscFixed + off;
    // shared memory: two panels + as many 16 KB weight stages as fit in ~half an SM (two CTAs co-reside under PDL);
    // the epilogue scratch (+ the fp32 partial tile of the cluster reduction) aliases the ring
    const int budget = 110 * 1024;
    int stages = (budget - 2 * c.panel_bytes - misc) / kABytes;
    // statistics cells + staging tile + residual tile (see the epilogue)
    const int scratch = kSredBytes + (sk > 1 ? NT * 512 + ((NT + sk - 1) / sk) * 256 : 16 * 1024);
    if (stages < 2) stages = 2;
    if (stages > 6) stages = 6;
    c.stages = stages;
    c.ring_bytes = stages * kABytes;
    if (c.ring_bytes < scratch) c.ring_bytes = round_up(scratch, 1024);
    c.smem = (size_t)c.ring_bytes + 2 * (size_t)c.panel_bytes + misc + 1024;
    return c.smem <= (size_t)(sk > 1 ? 113 : 227) * 1024;
  };
  auto cost_us = [&](const UmmaPlan& c) {
    const long long tiles = (long long)c.n_tiles * c.m_tiles * p.nphase;
    // Two CTAs fit an SM (<= 113 KB each).  The panel / epilogue phases are bound by per-warp instruction latency
    // (one warp per scheduler), so two RUNNING CTAs per SM nearly double the throughput of those phases; the price is
    // that the next layer's CTAs can no longer sit next to this layer's (exposed prologue, no weight prefetch).
    static int oversub = -1;
    if (oversub < 0) {
      const char* e = getenv("JEN1_OVERSUB");
      oversub = e ? atoi(e) : 1;
    }
    static double o_mul = -1.0, o_add = -1.0;
    if (o_mul < 0.0) {
      const char* e1 = getenv("JEN1_OVERSUB_MUL");
      const char* e2 = getenv("JEN1_OVERSUB_ADD");
      o_mul = e1 ? atof(e1) : 0.85;
      o_add = e2 ? atof(e2) : 0.0;
    }
    const long long ctas = tiles * c.splitk;
    const double conc = (oversub && c.smem <= (size_t)113 * 1024) ? 2.0 : 1.0;
    const double waves = (double)((ctas + (long long)(conc * num_sms) - 1) / (long long)(conc * num_sms));
    const double over_mul = (conc > 1.0 && ctas > num_sms) ? o_mul : 1.0, over_add = (conc > 1.0 && ctas > num_sms) ? o_add : 0.0;
    const int steps = (nsteps + c.splitk - 1) / c.splitk;
    const double rows = (double)(c.R * f);
    const double wkb = (double)steps * ntaps * 16.0;
    const double prefetch = c.stages * 16.0;
    const double cols = c.splitk > 1 ? (double)((c.NT + c.splitk - 1) / c.splitk) : (double)c.NT;
    static double kstep_us = -1.0, xch_us = -1.0;  // planner experiments: JEN1_KSTEP_US (fixed cost of a K step), JEN1_XCH_US (per exchanged column)
    if (kstep_us < 0.0) {
      const char* e1 = getenv("JEN1_KSTEP_US");
      const char* e2 = getenv("JEN1_XCH_US");
      kstep_us = e1 ? atof(e1) : 0.45;  // (0.6 before the MMA issue became warp-uniform; A/B in one GPU call: -0.75 % per step)
      xch_us = e2 ? atof(e2) : 0.16;
    }
    double t = steps * (kstep_us + 0.15 * rows / 16.0);      // K loop: fixed round trip + panel rows
    t += (wkb > prefetch ? (wkb - prefetch) / 60.0 : 0.0);  // weight bytes beyond the PDL prefetch at ~60 GB/s per SM
    t += (c.splitk > 1 ? xch_us * cols + 0.6 : 0.04 * cols);  // epilogue columns (DSMEM reduction vs TMEM), cluster exchange
    return t * waves * over_mul + over_add;
  };
  static const int cand[] = {16, 32, 48, 64, 96, 128, 192, 256};
  bool found = false;
  double best = 0.0;
  UmmaPlan bestp = pl;
  for (int NT : cand) {
    if (NT > nt_cap) break;
    UmmaPlan c;
    if (!configure(NT, c)) continue;
    const double t = cost_us(c);
    if (!found || t < best - 1e-9) {
      found = true;
      best = t;
      bestp = c;
    }
  }
  if (!found) return pl;
  pl = bestp;
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
  cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  // clusters of 16 CTAs (non-portable size) if the device schedules them, else 8
  g_max_cluster = 8;
  const char* env = getenv("JEN1_MAX_CLUSTER");
  const int want = env ? atoi(env) : kMaxCluster;
  if (want >= 16 && cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(1, 8, 16);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 113 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 16;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, conv_umma_kernel, &cfg) == cudaSuccess && nclusters >= 4) g_max_cluster = 16;
  } else if (want >= 1 && want < 8) {
    g_max_cluster = want;
  }
  (void)cudaGetLastError();
  return cudaSuccess;
}
int conv_umma_max_cluster() { return g_max_cluster; }

cudaError_t launch_conv_umma(const ConvParams& p, const UmmaPlan& pl, const void* w0, const void* w1, bool out_f32,
                             bool pdl, cudaStream_t stream, long long* timeline) {
  UmmaArgs a;
  a.p = p;
  a.pl = pl;
  a.w0 = (const bf16*)w0;
  a.w1 = (const bf16*)w1;
  a.out_f32 = out_f32 ? 1 : 0;
  a.timeline = timeline;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pl.n_tiles, pl.m_tiles, p.nphase * pl.splitk);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pl.splitk > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = (unsigned)pl.splitk;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel, a);
}

}  // namespace jen1
