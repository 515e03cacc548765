// Shared pieces of the pooling kernels (pool.cu: direct-load row marches; pool_tma.cu: TMA-staged t-marches).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace pool {

constexpr int HD = PMV_HEAD_DIM;  // 96
constexpr int TAPS = 27;
constexpr int NCP = HD / 2;        // 48 channel pairs
constexpr int ROWS = 4;            // output rows per CTA pass
constexpr int LNL = 8;             // lanes per token in the LayerNorm phase
constexpr int CPL = HD / LNL;      // 12 channels per lane
constexpr int NGRAD = (TAPS + 2) * HD;  // dW [96][27], dgamma [96], dbeta [96]
constexpr int NDW = TAPS * HD;
constexpr int MAX_JOBS = 3;

struct Job {
  const void* in;       // first channel of this tensor inside the QKV buffer
  const float* w;       // [96,1,3,3,3]
  const float* gamma;
  const float* beta;
  void* out;            // forward output [B, heads, 1+Lo, out_ld]
  int64_t out_ld;
  const void* dout;     // backward: gradient of `out`
  int64_t dout_ld;
  int dout_f32;         // dout is fp32 whatever T is (saved-statistics backward only)
  int onehot;           // forward: fill columns [96, out_ld) with the one-hot key coordinates (K' of the rel-pos scheme)
  void* din;            // backward: gradient wrt `in` (same strides)
  float* grads;         // backward: [NGRAD] fp32, overwritten
  void* dconv;          // backward: pre-LN gradient [B*heads*Lo*96] in the compute dtype
  void* xhat;           // optional: normalised pre-affine tokens [B*heads*(1+Lo)][96] (forward writes, backward reads)
  float* rstd;          // optional: [B*heads*(1+Lo)]
  float* part_ln;       // backward (i): [nblk + ncls_blk][2 * 96] per-CTA partials (dgamma, dbeta)
  float* part_dw;       // backward (ii): [nblk2][96 * 27] per-CTA partials
  int s, Ho, Wo;
  int blk_begin, nblk, ncls_blk;  // block range of this job in the current launch (march blocks, then cls blocks)
  int nblk_ln, nblk_dw;           // number of partial vectors behind part_ln / part_dw (for the reduce kernel)
};

// Pin a kernel-parameter field in a register.  Fields of the job table are addressed with a run-time index, so every use
// compiles to an indexed constant load (LDC / LDCU c[0x0][R + off]) that the compiler prefers to re-issue rather than
// keep in a register; in the store section of the LayerNorm warps those loads sat in front of every predicate and
// address (profiles/r02_pool_trace.md: 1.2 us for ~150 instructions).
template <typename P> __device__ __forceinline__ P* pin(P* p) {
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ int pin(int v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ int64_t pin(int64_t v) {
  asm volatile("" : "+l"(v));
  return v;
}
__device__ __forceinline__ float pin(float v) {
  asm volatile("" : "+f"(v));
  return v;
}

// ---- 2-channel loads / stores ---------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld2(const bf16* p) {
  const uint32_t u = __ldg(reinterpret_cast<const unsigned int*>(p));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(bf16* p, float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ void load12(const T* p, float (&v)[CPL]) {
  float a[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    load4(p + 4 * i, a);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[4 * i + j] = a[j];
  }
}
template <typename T> __device__ __forceinline__ void store12(T* p, const float (&v)[CPL]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
    store4(p + 4 * i, a);
  }
}

// the 27 taps of channels (2cp, 2cp+1); FLIP mirrors all three axes (transposed stencil)
template <bool FLIP> __device__ __forceinline__ void load_taps(const float* __restrict__ w, int cp, float2 (&wr)[TAPS]) {
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    const int src = FLIP ? (TAPS - 1 - tap) : tap;
    wr[tap] = make_float2(__ldg(w + (2 * cp) * TAPS + src), __ldg(w + (2 * cp + 1) * TAPS + src));
  }
}

// pool_tma.cu
int tma_items(int B, int heads, int Ho, int Wo);
bool tma_eligible(int stride_hw, int mode, int elem_bytes);
bool tma_fwd_writes_onehot();  // the warp-specialised forward kernel is in use (it fills Job::onehot columns itself)
int tma_launch(int mode, const Job* jobs, int njobs, int B, int heads, int T, int H, int W, int64_t bs, int64_t ts, int64_t hs,
               float eps, int dtype, cudaStream_t st, cudaStream_t st_strided);

// Fork / join of the independent pooling launches (pool.cu).  The kernels of one pooling call each fill 0.2-0.9 waves of the
// machine (profiles/r02_pool_ncu.md) and several of them do not depend on one another (dW vs the input-gradient kernels,
// dense-plane vs tap-tile job classes), so they are issued on two lazily created side streams between an event fork and
// an event join on the caller's stream: inside a captured CUDA graph these become parallel branches.  PMV_POOL_STREAMS=0
// keeps everything on the caller's stream.
struct PoolFork {
  cudaStream_t main, side[2];
  bool on;
};
PoolFork pool_fork(cudaStream_t main);
void pool_join(const PoolFork& f);

}  // namespace pool
