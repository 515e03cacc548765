// Row f3: multi-tensor AdamW with decoupled weight decay per tensor (the reference's zero-WD grouping of 1-D
// parameters, models/optimizer.py:42-57,124-131), global L2-norm gradient clipping (tools/train_net.py:196-199,
// torch.nn.utils.clip_grad_norm_) and the bf16 operand copy of every weight for the next forward, in ONE pass over
// parameters / gradients / moments.
//
// HBM-bound: 28 B per element (p, m, v read + write, g read) + 2 B for the bf16 copy + 4 B for the norm pass.
// torch.optim.AdamW(fused, capturable) needs 12 launches and 463 us for the 34.5 M parameters of MViTv2-S on B200
// (2.1 TB/s) and the forward then re-casts 68 weights with 68 more launches; here: two launches per <= 320 tensors.
//
// The tensor table travels by value in the kernel parameters (no device table to keep in sync; CUDA graphs capture
// it).  Work is cut into chunks of CHUNK elements; a block finds its tensor by binary search over the chunk prefix.
#include "common.cuh"

namespace {

constexpr int MAX_T = 320;            // tensors per launch (56 B each: well inside the 32 KB parameter space)
constexpr int THREADS = 256;
constexpr int CHUNK = THREADS * 4 * 4;  // 4096 elements: 4 float4 per thread

struct TensorDev {
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* shadow;
  int64_t n;
  float wd, lr_scale;
};

struct Table {
  TensorDev t[MAX_T];
  int chunk_begin[MAX_T + 1];
  int ntensors;
  int chunk_base;  // global index of this launch's first chunk (slot in the partial-sum buffer)
};

// ctl[0] = clip coefficient, ctl[1] = 1 - beta1^t, ctl[2] = sqrt(1 - beta2^t), ctl[3] = pre-clip gradient norm
constexpr int CTL_FLOATS = 4;

__device__ __forceinline__ int find_tensor(const Table& tab, int chunk) {
  int lo = 0, hi = tab.ntensors;  // chunk_begin[lo] <= chunk < chunk_begin[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.chunk_begin[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(THREADS) grad_sumsq_kernel(const __grid_constant__ Table tab, float* __restrict__ partials) {
  pdl_wait();
  __shared__ int s_t;
  __shared__ float red[THREADS / 32];
  if (threadIdx.x == 0) s_t = find_tensor(tab, blockIdx.x);
  __syncthreads();
  const TensorDev& T = tab.t[s_t];
  const int64_t base = (int64_t)(blockIdx.x - tab.chunk_begin[s_t]) * CHUNK;
  const int64_t left = T.n - base;
  const int cnt = left < CHUNK ? (int)left : CHUNK;
  const float* __restrict__ g = T.g + base;
  float s = 0.f;
  if (aligned16(g)) {
    for (int i = threadIdx.x * 4; i < cnt; i += THREADS * 4) {
      if (i + 4 <= cnt) {
        const float4 x = *reinterpret_cast<const float4*>(g + i);
        s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      } else {
        for (int j = i; j < cnt; ++j) s += g[j] * g[j];
      }
    }
  } else {
    for (int i = threadIdx.x; i < cnt; i += THREADS) s += g[i] * g[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) a += red[w];
    partials[tab.chunk_base + blockIdx.x] = a;
  }
}

// One block: folds the chunk partials into the global norm, derives the clip coefficient, advances the step counter
// and precomputes the bias corrections (double precision pow, like the Python-side scalars of torch.optim.AdamW).
__global__ void __launch_bounds__(1024) adamw_prep_kernel(const float* __restrict__ partials, int nchunks, float max_norm,
                                                          float beta1, float beta2, int32_t* __restrict__ step,
                                                          float* __restrict__ ctl) {
  pdl_wait();
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) s += (double)partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
    const float norm = (float)sqrt(a);
    float clip = 1.f;
    if (max_norm > 0.f) {
      clip = max_norm / (norm + 1e-6f);  // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
      if (clip > 1.f) clip = 1.f;
    }
    const int t = *step + 1;
    *step = t;
    ctl[0] = clip;
    ctl[1] = (float)(1.0 - pow((double)beta1, (double)t));
    ctl[2] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
    ctl[3] = norm;
  }
}

struct Hyper {
  float beta1, beta2, eps;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float lr, float wd, float clip, float bc1,
                                            float bc2s, const Hyper& h) {
  g *= clip;
  p *= 1.f - lr * wd;                        // decoupled weight decay
  m = m + (g - m) * (1.f - h.beta1);        // exp_avg.lerp_(grad, 1 - beta1)
  v = v * h.beta2 + g * g * (1.f - h.beta2);
  const float denom = sqrtf(v) / bc2s + h.eps;
  p -= (lr / bc1) * (m / denom);
}

__global__ void __launch_bounds__(THREADS) adamw_kernel(const __grid_constant__ Table tab, const float* __restrict__ lr_dev,
                                                        const float* __restrict__ ctl, Hyper h) {
  pdl_wait();
  __shared__ int s_t;
  if (threadIdx.x == 0) s_t = find_tensor(tab, blockIdx.x);
  __syncthreads();
  const TensorDev& T = tab.t[s_t];
  const int64_t base = (int64_t)(blockIdx.x - tab.chunk_begin[s_t]) * CHUNK;
  const int64_t left = T.n - base;
  const int cnt = left < CHUNK ? (int)left : CHUNK;
  const float clip = ctl[0], bc1 = ctl[1], bc2s = ctl[2];
  const float lr = lr_dev[0] * T.lr_scale, wd = T.wd;
  float* __restrict__ p = T.p + base;
  const float* __restrict__ g = T.g + base;
  float* __restrict__ m = T.m + base;
  float* __restrict__ v = T.v + base;
  bf16* __restrict__ sh = T.shadow != nullptr ? T.shadow + base : nullptr;
  const bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (sh == nullptr || (reinterpret_cast<uintptr_t>(sh) & 7) == 0);
  if (vec) {
    for (int i = threadIdx.x * 4; i < cnt; i += THREADS * 4) {
      if (i + 4 <= cnt) {
        float pp[4], gg[4], mm[4], vv[4];
        load4(p + i, pp); load4(g + i, gg); load4(m + i, mm); load4(v + i, vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) adam_update(pp[j], gg[j], mm[j], vv[j], lr, wd, clip, bc1, bc2s, h);
        store4(p + i, pp); store4(m + i, mm); store4(v + i, vv);
        if (sh != nullptr) store4(sh + i, pp);
      } else {
        for (int j = i; j < cnt; ++j) {
          float pp = p[j], mm = m[j], vv = v[j];
          adam_update(pp, g[j], mm, vv, lr, wd, clip, bc1, bc2s, h);
          p[j] = pp; m[j] = mm; v[j] = vv;
          if (sh != nullptr) sh[j] = __float2bfloat16_rn(pp);
        }
      }
    }
  } else {
    for (int j = threadIdx.x; j < cnt; j += THREADS) {
      float pp = p[j], mm = m[j], vv = v[j];
      adam_update(pp, g[j], mm, vv, lr, wd, clip, bc1, bc2s, h);
      p[j] = pp; m[j] = mm; v[j] = vv;
      if (sh != nullptr) sh[j] = __float2bfloat16_rn(pp);
    }
  }
}

int64_t count_chunks(const pmv_adamw_tensor* t, int n) {
  int64_t c = 0;
  for (int i = 0; i < n; ++i) c += ceil_div64(t[i].numel, CHUNK);
  return c;
}

}  // namespace

extern "C" int64_t pmv_adamw_workspace_bytes(const pmv_adamw_tensor* tensors, int ntensors) {
  return (count_chunks(tensors, ntensors) + CTL_FLOATS) * (int64_t)sizeof(float);
}

extern "C" int pmv_adamw_step(const pmv_adamw_tensor* tensors, int ntensors, const float* lr, float beta1, float beta2, float eps,
                              float max_grad_norm, int32_t* step, float* grad_norm_out, float* ws, void* stream) {
  PMV_CHECK_ARG(ntensors >= 0 && (ntensors == 0 || tensors != nullptr), "adamw: bad tensor list");
  PMV_CHECK_ARG(lr != nullptr && step != nullptr && ws != nullptr, "adamw: lr, step and ws must be device pointers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nchunks = count_chunks(tensors, ntensors);
  PMV_CHECK_ARG(nchunks < (1ll << 31), "adamw: too many elements");
  float* ctl = ws;
  float* partials = ws + CTL_FLOATS;
  const bool need_norm = max_grad_norm > 0.f || grad_norm_out != nullptr;
  static Table tab;  // 19 KB: filled per launch, passed by value (host calls into the library are single-threaded per process)
  for (int pass = need_norm ? 0 : 1; pass < 2; ++pass) {
    if (pass == 1) {
      pmv_launch(adamw_prep_kernel, dim3(1), dim3(1024), 0, st, (const float*)partials, need_norm ? (int)nchunks : 0, max_grad_norm, beta1, beta2,
                 step, ctl);
      PMV_CHECK_LAUNCH();
    }
    int chunk_base = 0;
    for (int first = 0; first < ntensors; first += MAX_T) {
      const int cnt = ntensors - first < MAX_T ? ntensors - first : MAX_T;
      int chunks = 0;
      for (int i = 0; i < cnt; ++i) {
        const pmv_adamw_tensor& s = tensors[first + i];
        PMV_CHECK_ARG(s.param != nullptr && s.grad != nullptr && s.exp_avg != nullptr && s.exp_avg_sq != nullptr && s.numel >= 0,
                      "adamw: tensor %d has a null pointer", first + i);
        tab.t[i] = TensorDev{s.param, s.grad, s.exp_avg, s.exp_avg_sq, reinterpret_cast<bf16*>(s.shadow), s.numel, s.weight_decay, s.lr_scale};
        tab.chunk_begin[i] = chunks;
        chunks += (int)ceil_div64(s.numel, CHUNK);
      }
      tab.chunk_begin[cnt] = chunks;
      tab.ntensors = cnt;
      tab.chunk_base = chunk_base;
      if (chunks > 0) {
        if (pass == 0) pmv_launch(grad_sumsq_kernel, dim3(chunks), dim3(THREADS), 0, st, tab, partials);
        else pmv_launch(adamw_kernel, dim3(chunks), dim3(THREADS), 0, st, tab, lr, (const float*)ctl, Hyper{beta1, beta2, eps});
        PMV_CHECK_LAUNCH();
      }
      chunk_base += chunks;
    }
  }
  if (grad_norm_out != nullptr) PMV_CHECK_CUDA(cudaMemcpyAsync(grad_norm_out, ctl + 3, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PMV_OK;
}
