// Row f2: the tail of the training step in two launches forward and four backward — final LayerNorm of the cls row
// (video_model_builder.py:2163-2165: norm(x)[:, 0]; only the selected row is normalised, which is the same value),
// TransformerBasicHead (head_helper.py:561-577: dropout -> Linear [-> softmax in eval]) and the soft-target
// cross-entropy of losses.py:69-71 (pytorchvideo SoftTargetCrossEntropyLoss, normalize_targets=False:
// mean_b sum_c -y[b,c] log_softmax(logits)[b,c]; integer labels are the one-hot case = nn.CrossEntropyLoss).
//
// Tiny work ([B,768] x [768,400]), latency-bound, everything fp32: two launches forward, four backward, each spread over
// enough CTAs that no thread walks more than ~50 dependent loads (a first version with one CTA per clip took 0.3 ms).
// It replaces ~20 eager launches (LN over all 393 tokens, index, dropout, addmm, log_softmax, nll, and their backward
// incl. a zero-filled [B,N,C] gradient).
#include "common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAX_C = 2048;     // channels of the cls row kept in shared memory

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) s += red[w];
  return s;
}

constexpr int CLS_CHUNKS = 16;  // forward: CTAs per clip, each owning a slice of the classes
constexpr int ROW_CHUNKS = 16;  // backward: CTAs per clip sharing the zero fill of dx

struct HeadFwd {
  const float* x;          // [B, N, C] tokens; row 0 of every clip is used
  int64_t batch_stride;    // N * C
  const float* gamma;
  const float* beta;
  const float* w;          // [ncls, C]
  const float* bias;       // [ncls] or NULL
  const uint8_t* keep;     // [B, C] dropout keep mask or NULL
  float keep_scale;        // 1 / (1 - p)
  float* logits;           // [B, ncls]
  float* xhat;             // [B, C] or NULL (saved for backward)
  float* rstd;             // [B] or NULL
  float* xd;               // [B, C] or NULL: the Linear's input (after dropout)
  int B, C, ncls;
  float eps;
};

// grid (CLS_CHUNKS, B): every CTA normalises the cls row of its clip (768 values: cheaper than a second launch) and
// computes its slice of the logits, one warp per class, lanes across the channels (coalesced weight rows).
__global__ void __launch_bounds__(THREADS) head_logits_kernel(const HeadFwd p) {
  pdl_wait();
  __shared__ float sx[MAX_C];
  __shared__ float red[WARPS];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = p.x + (int64_t)b * p.batch_stride;
  float s = 0.f;
  for (int k = tid; k < p.C; k += THREADS) { const float v = xr[k]; sx[k] = v; s += v; }
  const float mean = block_sum(s, red) / p.C;
  float q = 0.f;
  for (int k = tid; k < p.C; k += THREADS) { const float d = sx[k] - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q, red) / p.C + p.eps);
  const bool first = blockIdx.x == 0;
  for (int k = tid; k < p.C; k += THREADS) {
    const float xh = (sx[k] - mean) * rstd;
    float y = xh * p.gamma[k] + p.beta[k];
    if (p.keep != nullptr) y = p.keep[(int64_t)b * p.C + k] ? y * p.keep_scale : 0.f;
    sx[k] = y;
    if (first && p.xhat != nullptr) p.xhat[(int64_t)b * p.C + k] = xh;
    if (first && p.xd != nullptr) p.xd[(int64_t)b * p.C + k] = y;
  }
  if (first && tid == 0 && p.rstd != nullptr) p.rstd[b] = rstd;
  __syncthreads();
  const int per = (p.ncls + gridDim.x - 1) / gridDim.x;
  const int c_end = min(p.ncls, (int)(blockIdx.x + 1) * per);
  for (int c = blockIdx.x * per + warp; c < c_end; c += WARPS) {
    const float* wr = p.w + (int64_t)c * p.C;
    float a = 0.f;
    for (int k = lane; k < p.C; k += 32) a += sx[k] * __ldg(wr + k);
    a = warp_sum(a);
    if (lane == 0) p.logits[(int64_t)b * p.ncls + c] = a + (p.bias != nullptr ? p.bias[c] : 0.f);
  }
}

// One CTA, one warp per clip (looping when B > 8): log-sum-exp of the logits row, then
//   forward  (dlogits == NULL): probs (optional), loss row, and the mean over the clips;
//   backward (dlogits != NULL): dlogits[b, c] = (softmax_c * sum(y) - y_c) * dloss / B.
struct HeadSoftmax {
  const float* logits;     // [B, ncls]
  const int64_t* labels;   // [B] or NULL
  const float* soft;       // [B, ncls] or NULL
  float* probs;            // [B, ncls] or NULL
  float* loss;             // scalar or NULL
  const float* dloss;      // scalar (backward)
  float* dlogits;          // [B, ncls] (backward) or NULL
  int B, ncls;
};

__global__ void __launch_bounds__(THREADS) head_softmax_kernel(const HeadSoftmax p) {
  pdl_wait();
  __shared__ float rows[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;  // this warp's sum of loss rows
  for (int b = warp; b < p.B; b += WARPS) {
    const float* lg = p.logits + (int64_t)b * p.ncls;
    float m = -INFINITY;
    for (int c = lane; c < p.ncls; c += 32) m = fmaxf(m, lg[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e = 0.f, ysum = 0.f;
    for (int c = lane; c < p.ncls; c += 32) {
      e += expf(lg[c] - m);
      if (p.soft != nullptr) ysum += p.soft[(int64_t)b * p.ncls + c];
    }
    e = warp_sum(e);
    ysum = p.soft != nullptr ? warp_sum(ysum) : 1.f;
    const float lse = m + logf(e);
    const int64_t label = (p.soft == nullptr && p.labels != nullptr) ? p.labels[b] : -1;
    if (p.dlogits != nullptr) {
      const float g = p.dloss[0] / p.B;
      for (int c = lane; c < p.ncls; c += 32) {
        const float y = p.soft != nullptr ? p.soft[(int64_t)b * p.ncls + c] : (c == label ? 1.f : 0.f);
        p.dlogits[(int64_t)b * p.ncls + c] = (expf(lg[c] - lse) * ysum - y) * g;
      }
      continue;
    }
    if (p.probs != nullptr)
      for (int c = lane; c < p.ncls; c += 32) p.probs[(int64_t)b * p.ncls + c] = expf(lg[c] - lse);
    if (p.loss != nullptr) {
      float l = 0.f;
      if (p.soft != nullptr) {
        for (int c = lane; c < p.ncls; c += 32) l -= p.soft[(int64_t)b * p.ncls + c] * (lg[c] - lse);
        l = warp_sum(l);
      } else {
        l = -(lg[label] - lse);
      }
      acc += l;
    }
  }
  if (p.loss == nullptr || p.dlogits != nullptr) return;
  if (lane == 0) rows[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += rows[w];
    p.loss[0] = s / p.B;
  }
}

struct HeadBwd {
  const float* w;          // [ncls, C]
  const float* gamma;
  const uint8_t* keep;
  float keep_scale;
  const float* xhat;       // [B, C]
  const float* rstd;       // [B]
  const float* dlogits;    // [B, ncls]
  float* dxn;              // [B, C] gradient at the LayerNorm output
  float* dx;               // [B, N, C]: row 0 receives the gradient, rows 1.. are cleared
  int64_t batch_stride;
  int B, C, ncls, N;
};

// dxn[b, k] = dropout'(sum_c dlogits[b, c] W[c, k]).  grid (ceil(C / 32), B); block = 32 channels x 8 class slices.
__global__ void __launch_bounds__(THREADS) head_dinput_kernel(const HeadBwd p) {
  pdl_wait();
  __shared__ float part[WARPS][32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f;
  if (k < p.C) {
    const float* dl = p.dlogits + (int64_t)b * p.ncls;
    int c = slice;
    for (; c + WARPS < p.ncls; c += 2 * WARPS) {
      a0 += dl[c] * __ldg(p.w + (int64_t)c * p.C + k);
      a1 += dl[c + WARPS] * __ldg(p.w + (int64_t)(c + WARPS) * p.C + k);
    }
    if (c < p.ncls) a0 += dl[c] * __ldg(p.w + (int64_t)c * p.C + k);
  }
  part[slice][lane] = a0 + a1;
  __syncthreads();
  if (slice == 0 && k < p.C) {
    float a = 0.f;
#pragma unroll
    for (int s = 0; s < WARPS; ++s) a += part[s][lane];
    if (p.keep != nullptr) a = p.keep[(int64_t)b * p.C + k] ? a * p.keep_scale : 0.f;
    p.dxn[(int64_t)b * p.C + k] = a;
  }
}

// LayerNorm backward of the cls row + zero gradient for the other tokens.  grid (ROW_CHUNKS, B): every CTA clears its
// share of rows 1..N-1; CTA 0 of the clip also produces row 0.
__global__ void __launch_bounds__(THREADS) head_dx_kernel(const HeadBwd p) {
  pdl_wait();
  __shared__ float red[WARPS];
  const int b = blockIdx.y, tid = threadIdx.x;
  float* dxr = p.dx + (int64_t)b * p.batch_stride;
  const int64_t rest = (int64_t)(p.N - 1) * p.C;
  float* z = dxr + p.C;
  const int64_t per = ((rest + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
  const int64_t z0 = (int64_t)blockIdx.x * per, z1 = min(rest, z0 + per);
  if ((reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    for (int64_t i = z0 + tid * 4; i < z1; i += THREADS * 4) {
      if (i + 4 <= z1) *reinterpret_cast<float4*>(z + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      else for (int64_t j = i; j < z1; ++j) z[j] = 0.f;
    }
  } else {
    for (int64_t i = z0 + tid; i < z1; i += THREADS) z[i] = 0.f;
  }
  if (blockIdx.x != 0) return;
  float s1 = 0.f, s2 = 0.f;
  for (int k = tid; k < p.C; k += THREADS) {
    const float gg = p.dxn[(int64_t)b * p.C + k] * p.gamma[k];
    s1 += gg;
    s2 += gg * p.xhat[(int64_t)b * p.C + k];
  }
  s1 = block_sum(s1, red) / p.C;
  s2 = block_sum(s2, red) / p.C;
  const float rs = p.rstd[b];
  for (int k = tid; k < p.C; k += THREADS) {
    const float gg = p.dxn[(int64_t)b * p.C + k] * p.gamma[k];
    dxr[k] = rs * (gg - s1 - p.xhat[(int64_t)b * p.C + k] * s2);
  }
}

// dW[c, k] = sum_b dlogits[b, c] xd[b, k];  db[c] = sum_b dlogits[b, c];  dgamma[k] = sum_b dxn[b, k] xhat[b, k];
// dbeta[k] = sum_b dxn[b, k].  Thread per output element; the first ncls*C threads own dW.
__global__ void __launch_bounds__(THREADS) head_param_grad_kernel(const float* __restrict__ dlogits, const float* __restrict__ xd,
                                                                  const float* __restrict__ dxn, const float* __restrict__ xhat,
                                                                  float* __restrict__ dw, float* __restrict__ db,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C,
                                                                  int ncls) {
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int64_t nw = (int64_t)ncls * C;
  if (i < nw) {
    const int c = (int)(i / C), k = (int)(i - (int64_t)c * C);
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += dlogits[(int64_t)b * ncls + c] * xd[(int64_t)b * C + k];
    dw[i] = a;
  } else if (i < nw + ncls) {
    const int c = (int)(i - nw);
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += dlogits[(int64_t)b * ncls + c];
    if (db != nullptr) db[c] = a;
  } else if (i < nw + ncls + C) {
    const int k = (int)(i - nw - ncls);
    float a = 0.f, s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = dxn[(int64_t)b * C + k];
      a += d * xhat[(int64_t)b * C + k];
      s += d;
    }
    dgamma[k] = a;
    dbeta[k] = s;
  }
}

}  // namespace

extern "C" int pmv_head_loss_fwd(const float* x, int64_t batch_stride, const float* gamma, const float* beta, const float* w,
                                 const float* bias, const uint8_t* keep_mask, float dropout_p, const int64_t* labels,
                                 const float* soft_targets, float* logits, float* probs, float* loss, float* xhat,
                                 float* rstd, float* xd, int B, int C, int num_classes, float eps, void* stream) {
  PMV_CHECK_ARG(B > 0 && C > 0 && C <= MAX_C && num_classes > 0, "head: C must be in [1, %d]", MAX_C);
  PMV_CHECK_ARG(x != nullptr && gamma != nullptr && beta != nullptr && w != nullptr && logits != nullptr, "head: null pointer");
  PMV_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "head: dropout probability must be in [0, 1)");
  PMV_CHECK_ARG(loss == nullptr || ((labels != nullptr) != (soft_targets != nullptr)),
                "head: a loss needs exactly one of labels / soft_targets");
  HeadFwd p{x, batch_stride, gamma, beta, w, bias, keep_mask, 1.f / (1.f - dropout_p), logits, xhat, rstd, xd, B, C, num_classes, eps};
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = num_classes < CLS_CHUNKS * WARPS ? (num_classes + WARPS - 1) / WARPS : CLS_CHUNKS;
  pmv_launch(head_logits_kernel, dim3(chunks, B), dim3(THREADS), 0, st, p);
  PMV_CHECK_LAUNCH();
  if (loss != nullptr || probs != nullptr) {
    HeadSoftmax q{logits, labels, soft_targets, probs, loss, nullptr, nullptr, B, num_classes};
    pmv_launch(head_softmax_kernel, dim3(1), dim3(THREADS), 0, st, q);
    PMV_CHECK_LAUNCH();
  }
  return PMV_OK;
}

extern "C" int64_t pmv_head_loss_bwd_workspace_bytes(int B, int C, int num_classes) {
  return ((int64_t)B * num_classes + (int64_t)B * C) * (int64_t)sizeof(float);
}

extern "C" int pmv_head_loss_bwd(const float* dloss, const float* logits, const int64_t* labels, const float* soft_targets,
                                 const float* w, const float* gamma, const uint8_t* keep_mask, float dropout_p, const float* xhat,
                                 const float* rstd, const float* xd, float* dx, int64_t batch_stride, int N, float* dw, float* dbias,
                                 float* dgamma, float* dbeta, float* ws, int B, int C, int num_classes, void* stream) {
  PMV_CHECK_ARG(B > 0 && C > 0 && C <= MAX_C && num_classes > 0 && N >= 1, "head bwd: bad shape");
  PMV_CHECK_ARG((labels != nullptr) != (soft_targets != nullptr), "head bwd: exactly one of labels / soft_targets");
  PMV_CHECK_ARG(dloss != nullptr && dx != nullptr && dw != nullptr && dgamma != nullptr && dbeta != nullptr && ws != nullptr, "head bwd: null pointer");
  float* dlogits = ws;
  float* dxn = ws + (int64_t)B * num_classes;
  cudaStream_t st = (cudaStream_t)stream;
  HeadSoftmax q{logits, labels, soft_targets, nullptr, nullptr, dloss, dlogits, B, num_classes};
  pmv_launch(head_softmax_kernel, dim3(1), dim3(THREADS), 0, st, q);
  PMV_CHECK_LAUNCH();
  HeadBwd p{w, gamma, keep_mask, 1.f / (1.f - dropout_p), xhat, rstd, dlogits, dxn, dx, batch_stride, B, C, num_classes, N};
  pmv_launch(head_dinput_kernel, dim3((C + 31) / 32, B), dim3(THREADS), 0, st, p);
  PMV_CHECK_LAUNCH();
  pmv_launch(head_dx_kernel, dim3(N > 1 ? ROW_CHUNKS : 1, B), dim3(THREADS), 0, st, p);
  PMV_CHECK_LAUNCH();
  const int64_t outs = (int64_t)num_classes * C + num_classes + C;
  pmv_launch(head_param_grad_kernel, dim3((unsigned)ceil_div64(outs, THREADS)), dim3(THREADS), 0, st, (const float*)dlogits, xd,
             (const float*)dxn, xhat, dw, dbias, dgamma, dbeta, B, C, num_classes);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
