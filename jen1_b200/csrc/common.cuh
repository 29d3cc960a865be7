// Shared device/host helpers for the JEN-1 B200 denoiser kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jen1 {

typedef __nv_bfloat16 bf16;

// ---- storage-type traits: activations/weights are stored as float (strict mode) or bf16 (fast mode);
//      all arithmetic is fp32.
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// torch.nn.SiLU / nn.GELU() (erf form) in fp32, accurate math (no fast-math intrinsics: parity first).
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// GroupNorm statistics side channel: (sum, sumsq) per (batch row, fine group) accumulated by every producer CTA with
// 64-bit FIXED-POINT atomics (scale 2^26).  Integer addition is associative, so the result is bit-identical whatever
// order the CTAs arrive in (deterministic without a reduction pass), the adds are fire-and-forget (no return value ->
// no latency on the producer side), and a consumer reads 16 bytes per fine group instead of reducing one entry per
// producer tile.  Range: |sum| < 2^63 / 2^26 = 1.4e11; resolution 1.5e-8 per add (below fp32 rounding of the partial).
constexpr float kStatScale = 67108864.0f;
constexpr double kStatInv = 1.0 / 67108864.0;
__device__ __forceinline__ void stat_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * kStatScale));
}
__device__ __forceinline__ float stat_get(long long v) { return (float)((double)v * kStatInv); }
__device__ __forceinline__ double stat_get_d(long long v) { return (double)v * kStatInv; }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace jen1
