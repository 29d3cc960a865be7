// Key-tiled tcgen05 / TMEM attention with online softmax for bf16 storage on sm_100a (reference jen1/model/blocks.py:355-380
// AttentionBase.forward, causal mask :304-319) -- the general-length companion of attn_umma.cu (which holds a whole head's
// keys in one TMEM accumulator, <= 256 keys, the model's own shapes).  One CTA = (batch row, head, 256 queries = two query
// tiles of 128); the keys stream through in tiles of 128:
//
//   loaders (warps 5-7)   K_j, V_j tiles -> shared-memory rings (TMA: one thread per ring, or cp.async; K-major 128-byte-
//                         swizzled rows; V is consumed as an MN-major operand straight from its [key][channel] layout).
//                         The K and the V ring have their OWN full / empty barriers: a K stage is free as soon as the S_j of
//                         both query tiles have been computed, a whole tile period before the V stage (free after P_j V_j),
//                         so the load of the next K tile is never on the critical path of the next S
//   MMA issuer (warp 4)   S^g_j = Q^g K_j^T -> TMEM columns [g * 128, +128) for both query tiles g, issued one key tile AHEAD of
//                         the softmax;  O^g += P^g_j V_j -> TMEM columns [256 + g * 128, +d), accumulated over ALL key tiles.
//                         Warp-uniform control flow, one elect.sync lane issues (a divergent lane-0 issuer costs ~85 clocks
//                         per MMA: the compiler wraps every tcgen05 instruction in an elect-and-branch loop)
//   softmax (warps 0-3: query tile 0, warps 8-11: query tile 1)
//                         thread == query row == TMEM lane with the row's 128 logits of a key tile in registers: maximum, sum
//                         and scaling are private to the thread (no exchange).  P_j = 2^(s - m) rounded to bf16 into a
//                         swizzled A tile.  The reference maximum m may lag the row's true maximum by up to 2^8; the output
//                         accumulator stays in TMEM and is rescaled in place (tcgen05.ld / st) only when a maximum moves by
//                         more than that -- no per-tile read-back, no register accumulator, no correction pass
//
// Mbarrier pipelines: k_full/k_empty, v_full/v_empty (rings), s_full/s_empty, p_full, o_full (one each per query tile).  Causal
// tiles that lie completely above a query tile's diagonal are skipped for that tile.  Keys beyond the key count are zero-filled
// and excluded from the softmax.
#include <cuda.h>
#include <float.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace jen1 {

namespace {

constexpr int kFaThreads = 384;  // warps 0-3 softmax of query tile 0, 4 MMA issuer, 5-7 loaders, 8-11 softmax of query tile 1
constexpr int kFaLoaders = 96;   // warps 5-7
constexpr int kBN = 128;         // keys per tile
constexpr int kBM = 256;         // queries per CTA: two tiles of 128, one per softmax warpgroup

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: one [1][128 rows][64 channels] box of the packed [B][N][ld] bf16 tensor -> a 16 KB K-major tile in the 128-byte
// swizzle the UMMA descriptors expect; rows beyond N arrive as zeros
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* tm, int c0, int row0, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(row0), "r"(b), "r"(smem_u32(bar))
      : "memory");
}
// suspending wait (hardware time-limited sleep) for the warps whose waits are long: keeps them off the issue slots
#ifndef JEN1_FA_POLL
#define JEN1_FA_POLL 0  // A/B: 1 = the MMA issuer polls (test_wait) instead of the suspending try_wait (measured: -9 %)
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_wait(uint64_t* bar, uint32_t parity) {
#if JEN1_FA_POLL
  mbar_wait(bar, parity);
#else
  mbar_wait_sleep(bar, parity);
#endif
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
// one lane of a converged warp (elect.sync): the issue pattern the compiler recognises as uniform
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
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
// 64 consecutive accumulator columns of this thread's TMEM lane in ONE load (one wait instead of four)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
        "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
        "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
        "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
// 2^x with the single-instruction hardware approximation (2 ulp; the result is rounded to bf16 right away)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 -- one issue slot for two lanes of arithmetic)
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

// tcgen05.st: 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n\t"
      "tcgen05.wait::st.sync.aligned;" ::"r"(__float_as_uint(v[0])),
      "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])),
      "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(taddr)
      : "memory");
}

// The running maximum a row scales with may lag the true maximum by up to 2^kTau (P <= 2^kTau, exact in bf16's exponent
// range): the accumulator in TMEM is only rescaled when a row's maximum grows by more than that, which after the first
// few key tiles practically never happens.
constexpr float kTau = 8.0f;

// JEN1_FA_TRACE=1 (debug build, scripts/build_variant.py): CTA (0,0,0) records clock64 at the pipeline hand-offs of its
// first 36 key tiles and prints them at the end (MMA issuer: P^0_j seen / P^0_j V_j issued; thread 0: S seen, P written,
// P_{j-1} V_{j-1} seen)
#ifndef JEN1_FA_TRACE
#define JEN1_FA_TRACE 0
#endif
#if JEN1_FA_TRACE
#define FA_TR(arr, j) do { if (trace_on && (j) < 36) arr[j] = clock64(); } while (0)
#else
#define FA_TR(arr, j) do { } while (0)
#endif

__global__ void __launch_bounds__(kFaThreads, 1) attn_flash_kernel(const __grid_constant__ AttnParams p,
                                                                   const __grid_constant__ CUtensorMap tmap, const int use_tma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // causal: the query blocks with the most key tiles are scheduled first (longest job first)
  const int i0 = (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * kBM, h = blockIdx.y, r = blockIdx.z;
  const int d = p.d, M = p.M, N = p.N;
  const int DB = (d + 63) >> 6;                       // 64-channel blocks of the head dim
  const int sh = d == 16 ? 1 : (d == 32 ? 2 : (d == 64 ? 3 : 4));
  const int cpr = 1 << sh;                            // 16-byte chunks per head row
  const uint32_t tile_bytes = (uint32_t)DB * 128u * 128u;  // one Q / K / V tile of 128 rows
  // K / V ring depths (what fits next to two Q and two P tiles).  S_{j+1} is issued a whole tile period ahead of its use,
  // so K_{j+1} (loaded once S_j has completed) has that period to arrive even with ONE stage; V_j is needed at the END of
  // tile j, one stage would put its load (started after P_{j-1} V_{j-1}) on the critical path: d = 128 gives V the two stages
  const int NSK = d <= 64 ? 4 : 1, NSV = 2;  // (K2 / V1 at d = 128 measured the same)
  uint8_t* Qs = smem;                                 // [2 query tiles]
  uint8_t* Ks = Qs + 2 * tile_bytes;                  // [NSK stages]
  uint8_t* Vs = Ks + NSK * tile_bytes;                // [NSV stages]
  uint8_t* Ps = Vs + NSV * tile_bytes;                // [2 query tiles][2 key blocks][128 queries][128 B]
  constexpr uint32_t kPBytes = 2 * 128 * 128;         // one P tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 2 * kPBytes);
  uint64_t* q_full = bars;           // 1: producers of Q
  uint64_t* k_full = bars + 1;       // [4] loaders -> MMA
  uint64_t* k_empty = bars + 5;      // [4] MMA (commit after the S_j of both query tiles) -> loaders
  uint64_t* v_full = bars + 9;       // [4] loaders -> MMA
  uint64_t* v_empty = bars + 13;     // [4] MMA (commit after the P_j V_j of both query tiles) -> loaders
  uint64_t* s_full = bars + 17;      // [2 query tiles] MMA (commit) -> softmax
  uint64_t* s_empty = bars + 19;     // [2] softmax (logits are in registers) -> MMA
  uint64_t* p_full = bars + 21;      // [2] softmax -> MMA
  uint64_t* o_full = bars + 23;      // [2] MMA (commit) -> softmax
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  if (tid == 128) {
    mbar_init(q_full, use_tma ? 1 : kFaThreads);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&k_full[s], use_tma ? 1 : kFaLoaders);
      mbar_init(&v_full[s], use_tma ? 1 : kFaLoaders);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#if JEN1_FA_TRACE
  const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  long long tr0[36], tr1[36], tr2[36], tr3[36];
  const long long tr_base = clock64();
#endif
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // key tiles each query tile needs (causal: tiles completely above its diagonal are skipped; a query tile that starts
  // beyond N has none)
  const int nt_all = (M + kBN - 1) / kBN;
  int ntg[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    int n = (i0 + g * 128 < N) ? nt_all : 0;
    if (n > 0 && p.causal) {
      const int last_key = min(M - 1, i0 + g * 128 + 127 + (M - N));
      n = min(n, last_key / kBN + 1);
    }
    ntg[g] = n;
  }
  const int nt = max(ntg[0], ntg[1]);

  // ---- Q tiles (once): TMA (head dims that are multiples of 64; rows beyond N arrive as zeros) or all threads with
  //      cp.async; chunks beyond the head dim inside the last 64-channel block read as zero
  if (use_tma) {
    if (tid == 160) {
      mbar_expect_tx(q_full, 2 * tile_bytes);
      for (int g = 0; g < 2; ++g)
        for (int db = 0; db < DB; ++db)
          tma_load_tile(Qs + (size_t)g * tile_bytes + (size_t)db * 128 * 128, &tmap, p.q_off + h * d + db * 64, i0 + g * 128, r, q_full);
    }
  } else {
    const bf16* qsrc = (const bf16*)p.q + ((size_t)r * N + i0) * p.q_ld + p.q_off + h * d;
    for (int i = tid; i < (256 << sh); i += kFaThreads) {
      const int rw = i >> sh, part = i & (cpr - 1);  // rw: row of the 256-query block
      uint8_t* dst = Qs + (size_t)(rw >> 7) * tile_bytes + ((uint32_t)(part >> 3) * 128u + (uint32_t)(rw & 127)) * 128u +
                     (uint32_t)(((part & 7) ^ (rw & 7)) * 16);
      if (i0 + rw < N)
        cp_async16(dst, qsrc + (size_t)rw * p.q_ld + part * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);  // query rows beyond N: zeros (never stored)
    }
    if (cpr < 8) {
      const int zc = 8 - cpr;
      for (int it = tid; it < 256 * zc; it += kFaThreads) {
        const int rw = it / zc, ch = cpr + (it - rw * zc);
        *reinterpret_cast<uint4*>(Qs + (size_t)(rw >> 7) * tile_bytes + (uint32_t)(rw & 127) * 128u + (uint32_t)((ch ^ (rw & 7)) * 16)) =
            make_uint4(0u, 0u, 0u, 0u);
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
    fence_async_smem();
    mbar_arrive(q_full);
  }

  if (warp >= 5 && warp < 8) {
    // ======================================================================== K / V loaders
    if (use_tma) {  // one thread per ring (warp 5: K, warp 6: V): DB bulk tensor copies per key tile, completion counted in bytes
      if (tid == 160) {
        for (int j = 0; j < nt; ++j) {
          const int s = j % NSK;
          if (j >= NSK) mbar_wait_sleep(&k_empty[s], (uint32_t)((j / NSK - 1) & 1));
          mbar_expect_tx(&k_full[s], tile_bytes);
          for (int db = 0; db < DB; ++db)
            tma_load_tile(Ks + (size_t)s * tile_bytes + (size_t)db * 128 * 128, &tmap, p.k_off + h * d + db * 64, j * kBN, r, &k_full[s]);
        }
      } else if (tid == 192) {
        for (int j = 0; j < nt; ++j) {
          const int s = j % NSV;
          if (j >= NSV) mbar_wait_sleep(&v_empty[s], (uint32_t)((j / NSV - 1) & 1));
          mbar_expect_tx(&v_full[s], tile_bytes);
          for (int db = 0; db < DB; ++db)
            tma_load_tile(Vs + (size_t)s * tile_bytes + (size_t)db * 128 * 128, &tmap, p.v_off + h * d + db * 64, j * kBN, r, &v_full[s]);
        }
      }
      __syncwarp();
    } else {
      const int lt = tid - 160;
      const bf16* kbase = (const bf16*)p.kv + (size_t)r * N * p.kv_ld + h * d;  // self-attention layout: key rows r * N + j
      // one tile of K (v = 0) or V (v = 1) rows [j0, j0 + 128) -> `dst` (first == true: also zero the chunks beyond the head dim)
      auto load_tile = [&](uint8_t* dst, int j0, int v, bool first) {
        for (int q = lt; q < (kBN << sh); q += kFaLoaders) {
          const int rw = q >> sh, part = q & (cpr - 1);
          uint8_t* a = dst + ((uint32_t)(part >> 3) * 128u + (uint32_t)rw) * 128u + (uint32_t)(((part & 7) ^ (rw & 7)) * 16);
          if (j0 + rw < M)
            cp_async16(a, kbase + (size_t)(j0 + rw) * p.kv_ld + (v ? p.v_off : p.k_off) + part * 8);
          else
            *reinterpret_cast<uint4*>(a) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (cpr < 8 && first) {  // once per stage (never overwritten afterwards)
          const int zc = 8 - cpr;
          for (int q = lt; q < kBN * zc; q += kFaLoaders) {
            const int rw = q / zc, ch = cpr + (q - rw * zc);
            *reinterpret_cast<uint4*>(dst + (uint32_t)rw * 128u + (uint32_t)((ch ^ (rw & 7)) * 16)) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
        fence_async_smem();
      };
      for (int j = 0; j < nt; ++j) {
        const int ks = j % NSK, vs = j % NSV;
        if (j >= NSK) mbar_wait_sleep(&k_empty[ks], (uint32_t)((j / NSK - 1) & 1));
        load_tile(Ks + (size_t)ks * tile_bytes, j * kBN, 0, j < NSK);
        mbar_arrive(&k_full[ks]);
        if (j >= NSV) mbar_wait_sleep(&v_empty[vs], (uint32_t)((j / NSV - 1) & 1));
        load_tile(Vs + (size_t)vs * tile_bytes, j * kBN, 1, j < NSV);
        mbar_arrive(&v_full[vs]);
      }
    }
  } else if (warp == 4) {
    // ======================================================================== MMA issuer
    // The whole warp runs the (warp-uniform) control flow and one elected lane issues: with the issue code under a
    // divergent `lane == 0` branch the compiler wraps every tcgen05 instruction in an elect / branch loop and rebuilds
    // the descriptors per instruction (~85 clocks per MMA measured -- slower than the tensor pipe executes them).
    // Descriptors are built once; a K step only adds a constant to the 14-bit start-address field.
    {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(d >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t qd = make_desc_sw128(smem_u32(Qs), 16u, 1024u), kd = make_desc_sw128(smem_u32(Ks), 16u, 1024u);
      const uint64_t pd = make_desc_sw128(smem_u32(Ps), 16u, 1024u), vd = make_desc_sw128(smem_u32(Vs), 128u * 128u, 1024u);
      const uint32_t tile16 = tile_bytes >> 4;  // descriptor address units (16 bytes)
      const int nk = d >> 4;                    // K steps of S = Q K^T
      // S^g_j = Q^g K_j^T -> TMEM columns [g * 128, +128)
      auto issue_s = [&](int g, int j) {
        if (j > 0) mbar_wait(&s_empty[g], (uint32_t)((j - 1) & 1));  // the softmax warps hold S^g_{j-1} in registers
        tc_fence_after();
        const uint64_t ad = qd + (uint64_t)((uint32_t)g * tile16), bd = kd + (uint64_t)((uint32_t)(j % NSK) * tile16);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 64-channel block kk >> 2 (16 KB apart), 32 bytes per K step inside the swizzle atom
            const uint64_t off = (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2);
            if (kk < nk) umma_bf16(tmem_base + (uint32_t)(g * 128), ad + off, bd + off, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      // O^g += P^g_j V_j -> TMEM columns [256 + g * 128, +d): accumulates over ALL key tiles
      auto issue_o = [&](int g, int j) {
        mma_wait(&p_full[g], (uint32_t)(j & 1));  // P^g_j written (and O^g rescaled if the row maxima moved)
        if (g == 0) FA_TR(tr0, j); else FA_TR(tr2, j);
        tc_fence_after();
        const uint64_t ad = pd + (uint64_t)((uint32_t)g * (kPBytes >> 4)), bd = vd + (uint64_t)((uint32_t)(j % NSV) * tile16);
        if (elect_one()) {
#pragma unroll
          for (int k16 = 0; k16 < kBN / 16; ++k16)
            umma_bf16(tmem_base + (uint32_t)(256 + g * 128), ad + (uint64_t)((k16 >> 2) * 1024 + (k16 & 3) * 2), bd + (uint64_t)(k16 * 128),
                      idesc_o, (j > 0 || k16 > 0) ? 1u : 0u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
        if (g == 0) FA_TR(tr1, j); else FA_TR(tr3, j);
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mma_wait(&k_full[0], 0);
      if (ntg[0] > 0) issue_s(0, 0);
      if (ntg[1] > 0) issue_s(1, 0);
      commit(&k_empty[0]);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) {  // the next logits of both query tiles are computed while the softmax warps work on tile j
          const int ks = (j + 1) % NSK;
          mma_wait(&k_full[ks], (uint32_t)(((j + 1) / NSK) & 1));
          if (j + 1 < ntg[0]) issue_s(0, j + 1);
          if (j + 1 < ntg[1]) issue_s(1, j + 1);
          commit(&k_empty[ks]);  // K_{j+1} is free once both S_{j+1} are complete
        }
        const int vs = j % NSV;
        mma_wait(&v_full[vs], (uint32_t)((j / NSV) & 1));
        if (j < ntg[0]) issue_o(0, j);
        if (j < ntg[1]) issue_o(1, j);
        commit(&v_empty[vs]);  // V_j is free once both P_j V_j are complete
      }
#if JEN1_FA_TRACE
      if (trace_on && lane == 0)
        for (int j = 0; j < min(nt, 36); ++j)
          printf("[fa-mma] j %2d  P0 seen %7lld  PV0 issued %7lld  P1 seen %7lld  PV1 issued %7lld\n", j, tr0[j] - tr_base, tr1[j] - tr_base, tr2[j] - tr_base, tr3[j] - tr_base);
#endif
    }
  } else {
    // ======================================================================== online softmax
    // Warpgroup g (warps 0-3: g = 0, warps 8-11: g = 1) owns query tile g; thread == query row == TMEM lane, with the
    // row's 128 logits of a key tile in registers.  No exchange between threads: the row maximum, the row sum and the
    // scaling are private, and the two warpgroups drift apart by about one MMA so that the warps sharing a scheduler
    // are in different phases (one in the MUFU-bound exponentials while the other loads / reduces).  The output
    // accumulator stays in TMEM across all key tiles (tcgen05.mma accumulate) and is rescaled in place only when a row
    // maximum moves by more than 2^kTau.  The tile body is written for few instructions and short dependency chains:
    // packed fp32 pairs (FFMA2 / FADD2), four independent max chains, eight independent sum chains, and the key
    // masking of edge tiles (last tile, causal diagonal) in a code path of its own.
    const int g = warp >> 3;
    const int ntm = ntg[g];
    const int rowi = (warp & 3) * 32 + lane;          // query row of this thread inside its tile == TMEM lane
    const int i = i0 + g * 128 + rowi;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t ts = trow + (uint32_t)(g * 128), to = trow + (uint32_t)(256 + g * 128);
    const float sc = p.scale * 1.4426950408889634f;   // base-2 softmax
    const uint64_t sc2 = pk2(sc, sc);
    const int jmax = p.causal ? i + (M - N) : M - 1;  // last key this query may see
    float m = -FLT_MAX, l = 0.f;
    uint8_t* prow = Ps + (size_t)g * kPBytes + (size_t)rowi * 128;
    for (int j = 0; j < ntm; ++j) {
      FA_TR(tr3, j);
      mbar_wait(&s_full[g], (uint32_t)(j & 1));
      FA_TR(tr0, j);
      tc_fence_after();
      float v[128];
      tmem_ld64(ts, *reinterpret_cast<float(*)[64]>(&v[0]));
      tmem_ld64(ts + 64u, *reinterpret_cast<float(*)[64]>(&v[64]));
      tc_fence_before();
      mbar_arrive(&s_empty[g]);  // the logits are in registers: S^g may be overwritten
      const int kmaxv = min(jmax, M - 1) - j * kBN;  // last valid key of this row in the tile (may be < 0)
      const bool edge = kmaxv < 127;
      float mx;
      if (!edge) {
        float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int q = 4; q < 124; q += 8) {  // four independent chains of 3-input maxima
          m4[0] = fmaxf(m4[0], fmaxf(v[q], v[q + 1]));
          m4[1] = fmaxf(m4[1], fmaxf(v[q + 2], v[q + 3]));
          m4[2] = fmaxf(m4[2], fmaxf(v[q + 4], v[q + 5]));
          m4[3] = fmaxf(m4[3], fmaxf(v[q + 6], v[q + 7]));
        }
        m4[0] = fmaxf(m4[0], fmaxf(v[124], v[125]));
        m4[1] = fmaxf(m4[1], fmaxf(v[126], v[127]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
        mx = -FLT_MAX;
#pragma unroll
        for (int q = 0; q < 128; ++q)
          if (q <= kmaxv) mx = fmaxf(mx, v[q]);
      }
      mx = (mx == -FLT_MAX) ? mx : mx * sc;  // sc > 0: max commutes with the scaling
      const bool grow = mx > m + kTau || m == -FLT_MAX;
      if (__any_sync(0xffffffffu, grow)) {  // rare after the first tiles: move this warp's rows to their new maxima
        const float m_new = grow ? fmaxf(mx, m) : m;
        const float alpha = ex2_fast(m - m_new);  // 1 for the rows that stay; 0 for a row without any key so far
        if (j > 0) {
          mbar_wait(&o_full[g], (uint32_t)((j - 1) & 1));  // P^g_{j-1} V_{j-1} has completed: O^g is stable
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < d; c0 += 16) {
            float w[16];
            tmem_ld16(to + (uint32_t)c0, w);
#pragma unroll
            for (int q = 0; q < 16; ++q) w[q] *= alpha;
            tmem_st16(to + (uint32_t)c0, w);
          }
        }
        l *= alpha;
        m = m_new;
      }
      // The exponentials go to registers first (the logits die as they are consumed: no extra pressure) and the P tile is
      // only written afterwards: the wait for P^g_{j-1} V_{j-1}, which reads that tile, then comes a whole exponential
      // phase after its MMA was issued instead of right after the maximum
      const uint64_t nm2 = pk2(-m, -m);
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pw[64];
      if (!edge) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {  // 8 keys = one 16-byte chunk of the P row
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float e0, e1;
            upk2(ffma2(pk2(v[c * 8 + 2 * q], v[c * 8 + 2 * q + 1]), sc2, nm2), e0, e1);
            e0 = ex2_fast(e0);
            e1 = ex2_fast(e1);
            sum2[q] = fadd2(sum2[q], pk2(e0, e1));  // (the bf16 rounding of P happens in pack2; the row sum keeps the unrounded terms)
            pw[c * 4 + q] = pack2(e0, e1);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k0 = c * 8 + 2 * q;
            float e0 = ex2_fast(fmaf(v[k0], sc, -m)), e1 = ex2_fast(fmaf(v[k0 + 1], sc, -m));
            if (k0 > kmaxv) e0 = 0.0f;
            if (k0 + 1 > kmaxv) e1 = 0.0f;
            sum2[q] = fadd2(sum2[q], pk2(e0, e1));
            pw[c * 4 + q] = pack2(e0, e1);
          }
        }
      }
      FA_TR(tr1, j);
      if (j > 0) mbar_wait(&o_full[g], (uint32_t)((j - 1) & 1));  // P^g_{j-1} V_{j-1} has completed: the P tile is free
      FA_TR(tr2, j);
#pragma unroll
      for (int c = 0; c < 16; ++c)  // chunks 0-7: key block 0, 8-15: key block 1
        *reinterpret_cast<uint4*>(prow + (size_t)(c >> 3) * 128 * 128 + (((c & 7) ^ (rowi & 7)) * 16)) =
            make_uint4(pw[c * 4], pw[c * 4 + 1], pw[c * 4 + 2], pw[c * 4 + 3]);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      float a0, a1;
      upk2(fadd2(fadd2(sum2[0], sum2[1]), fadd2(sum2[2], sum2[3])), a0, a1);
      l += a0 + a1;
    }
#if JEN1_FA_TRACE
    if (trace_on && (tid == 0 || tid == 256))
      for (int j = 0; j < min(ntm, 36); ++j)
        printf("[fa-sm%d] j %2d  tile start %7lld  S seen %7lld  exps done %7lld  PV(j-1) seen %7lld\n", g, j, tr3[j] - tr_base, tr0[j] - tr_base, tr1[j] - tr_base, tr2[j] - tr_base);
#endif
    if (ntm > 0) {  // normalise and store this thread's row
      mbar_wait(&o_full[g], (uint32_t)((ntm - 1) & 1));
      tc_fence_after();
      const float inv = 1.0f / l;
      bf16* orow = (bf16*)p.out + ((size_t)r * N + (i < N ? i : 0)) * p.C + h * d;
#pragma unroll 1
      for (int c0 = 0; c0 < d; c0 += 16) {
        float w[16];
        tmem_ld16(to + (uint32_t)c0, w);
        if (i < N) {
          *reinterpret_cast<uint4*>(orow + c0) =
              make_uint4(pack2(w[0] * inv, w[1] * inv), pack2(w[2] * inv, w[3] * inv), pack2(w[4] * inv, w[5] * inv), pack2(w[6] * inv, w[7] * inv));
          *reinterpret_cast<uint4*>(orow + c0 + 8) =
              make_uint4(pack2(w[8] * inv, w[9] * inv), pack2(w[10] * inv, w[11] * inv), pack2(w[12] * inv, w[13] * inv), pack2(w[14] * inv, w[15] * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

bool attn_flash_supported(const AttnParams& p) {
  const int d = p.d;
  if (p.cross) return false;  // the context caches have <= 129 rows: attn_umma.cu's single-tile kernel covers them
  if (!(d == 16 || d == 32 || d == 64 || d == 128)) return false;
  if (p.M < 1 || p.N < 1 || p.M != p.N) return false;  // self-attention: keys are the rows of the same tensor
  if ((p.q_ld & 7) || (p.q_off & 7) || (p.C & 7) || (p.kv_ld & 7) || (p.k_off & 7) || (p.v_off & 7)) return false;
  return true;
}

static size_t attn_flash_smem(int d) {
  // two Q tiles + K / V rings + two P tiles + 25 barriers + the TMEM slot; the dynamic shared-memory window is 1024-byte
  // aligned (no manual round-up in the kernel, hence no slack here: d = 128 uses 224.2 of 227 KB)
  const int DB = (d + 63) / 64, NSK = d <= 64 ? 4 : 2, NSV = d <= 64 ? 2 : 1;  // (d = 128: three stages in all, see the kernel)
  return (size_t)(2 + NSK + NSV) * DB * 128 * 128 + 2 * (2 * 128 * 128) + 26 * 8;
}

cudaError_t attn_flash_init() {
  return cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_flash_smem(128));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    (void)cudaGetLastError();
  }
  return fn;
}

cudaError_t launch_attention_flash(const AttnParams& p, bool pdl, cudaStream_t stream) {
  if (!attn_flash_supported(p)) return cudaErrorInvalidValue;
  // TMA staging when q / k / v live in one packed tensor and the head dim fills whole 64-channel blocks
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int use_tma = 0;
  if (p.d % 64 == 0 && p.q == p.kv && p.q_ld == p.kv_ld && encode_tiled_fn() != nullptr) {
    const cuuint64_t dims[3] = {(cuuint64_t)p.kv_ld, (cuuint64_t)p.N, (cuuint64_t)p.B2};
    const cuuint64_t strides[2] = {(cuuint64_t)p.kv_ld * 2, (cuuint64_t)p.N * p.kv_ld * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode_tiled_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p.kv), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      use_tma = 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((p.N + kBM - 1) / kBM, p.H, p.B2);
  cfg.blockDim = dim3(kFaThreads);
  cfg.dynamicSmemBytes = attn_flash_smem(p.d);
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn_flash_kernel, p, tm, use_tma);
}

}  // namespace jen1
