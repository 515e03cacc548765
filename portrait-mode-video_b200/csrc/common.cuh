// Shared device/host helpers for the pmv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pmv_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// error plumbing (thread-local last-error string, returned by pmv_last_error)
// ---------------------------------------------------------------------------
void pmv_set_error(const char* fmt, ...);

#define PMV_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      pmv_set_error(__VA_ARGS__);         \
      return PMV_ERR_INVALID_ARGUMENT;    \
    }                                     \
  } while (0)

#define PMV_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      pmv_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PMV_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define PMV_CHECK_LAUNCH()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      pmv_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PMV_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------
// launches.  Every kernel of the library is launched through pmv_launch and calls pdl_wait() before it touches global
// memory.  With programmatic dependent launch enabled for its family (runtime.cu: pmv_set_pdl / PMV_PDL, default on)
// the CTAs of kernel N+1 are scheduled while kernel N drains - their prologue (barrier init, TMEM allocation,
// descriptor prefetch) overlaps its tail - and griddepcontrol.wait blocks until kernel N has completed and flushed;
// without the launch attribute the instruction is a no-op.  scripts/pdl_probe.cu measures the edge cost.
// ---------------------------------------------------------------------------
#ifndef PMV_PDL_FAMILY
#define PMV_PDL_FAMILY 256
#endif
bool pmv_pdl_enabled(int family);  // runtime.cu: reads PMV_PDL (bit mask of kernel families) once

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t pmv_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pmv_pdl_enabled(PMV_PDL_FAMILY) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Division by a run-time constant as multiply-high + shift (dividends < 2^31).  A hardware integer division
// costs ~25 instructions; the token decode of the memory-bound kernels did 5-25 of them per token and was
// instruction-bound on index arithmetic (profiles/r01_pool_relpos_ncu.md).
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t div) : d(div), mul(0), shr(0) {
    if (div > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < div) ++lg;  // ceil(log2(div))
      const uint32_t p = 31 + lg;
      mul = (uint32_t)(((1ull << p) + div - 1) / div);
      shr = p - 32;
    }
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return d == 1 ? n : (__umulhi(n, mul) >> shr);
#else
    return n / d;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

// ---------------------------------------------------------------------------
// scalar conversions
// ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// load / store 4 consecutive elements (16 B for float, 8 B for bf16); pointer must be so aligned
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

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
// reduce over aligned groups of G lanes (G power of two <= 32)
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float kInvSqrt2Pi = 0.39894228040143267794f;
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * kInvSqrt2Pi * expf(-0.5f * x * x);
}

// bf16-mode variants for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.25 (|err| <= 2.5e-5, far
// inside the 1e-2 bf16 tolerance), sharing one exp2 between the cdf and the pdf term: 2 MUFU + ~12 FMA per
// element instead of erff + expf (~50 instructions), which made the GELU epilogues latency-bound.
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& pdf_x) {
  const float ax = fabsf(x);
  const float e = exp2f(-0.72134752044448170368f * x * x);           // exp(-x^2 / 2)
  const float t = __fdividef(1.0f, fmaf(0.33267f, ax, 1.0f));         // 1 / (1 + 0.47047 |x| / sqrt(2))
  const float poly = t * fmaf(t, fmaf(t, 0.7478556f, -0.0958798f), 0.3480242f);
  const float half_tail = 0.5f * poly * e;                            // 0.5 * erfc(|x| / sqrt 2)
  cdf = x >= 0.f ? 1.0f - half_tail : half_tail;
  pdf_x = x * 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, px;
  gelu_fast_parts(x, cdf, px);
  return x * cdf;
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
  float cdf, px;
  gelu_fast_parts(x, cdf, px);
  return cdf + px;
}

#define PMV_DISPATCH_DTYPE(dtype, T, ...)                    \
  do {                                                       \
    if ((dtype) == PMV_F32) {                                \
      typedef float T;                                       \
      __VA_ARGS__;                                           \
    } else if ((dtype) == PMV_BF16) {                        \
      typedef bf16 T;                                        \
      __VA_ARGS__;                                           \
    } else {                                                 \
      pmv_set_error("unsupported dtype %d", (int)(dtype));   \
      return PMV_ERR_INVALID_ARGUMENT;                       \
    }                                                        \
  } while (0)
