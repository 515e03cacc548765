// fp32-math pooling attention (forward + backward) on CUDA cores: the fp32 mode of the path (tolerance 1e-4)
// and the bring-up reference of the tcgen05 kernel.  softmax(scale * Q' K'^T) V with the relative-position
// bias carried by the augmented columns of Q'/K' (see relpos.cu), online softmax so the [Nq, Nk] score matrix
// the reference materialises (attention.py:412-446) never exists, residual pooling (attention.py:450-454)
// and the head merge (attention.py:456) fused into the epilogue.
//
// One CTA owns 64 query rows of one (batch, head) and walks the keys in tiles of 64.
#include "common.cuh"

namespace {

constexpr int HD = PMV_HEAD_DIM;
constexpr int BQ = 64, BKV = 64;
constexpr int THREADS = 256;
constexpr int VP = HD + 1;   // padded row of a 96-wide tile
constexpr int SP = BKV + 1;  // padded row of a score tile

struct AttnGeom {
  int B, heads, Nq, Nk, kd;
  int64_t ld_qk, ld_v;
  float scale;
  int residual;
};

// load rows [r0, r0+64) x cols [0, ncols) of a row-major matrix (row stride ld) into smem (row stride sp), zero fill
template <typename T>
__device__ __forceinline__ void load_rows(const T* __restrict__ src, int64_t ld, int r0, int nrows_total, int ncols,
                                          float* dst, int sp) {
  const int c4n = ncols >> 2;
  for (int i = threadIdx.x; i < 64 * c4n; i += THREADS) {
    const int r = i / c4n, c4 = i - r * c4n;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + r < nrows_total) load4(src + (int64_t)(r0 + r) * ld + c4 * 4, v);
    float* d = dst + r * sp + c4 * 4;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
  }
}

// ------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(THREADS) attn_fwd_kernel(const T* __restrict__ q_aug, const T* __restrict__ k_aug,
                                                           const T* __restrict__ v, T* __restrict__ out,
                                                           T* __restrict__ out_pre, float* __restrict__ lse, AttnGeom g) {
  pdl_wait();
  extern __shared__ float sm[];
  const int KP = g.kd + 1;
  float* Qs = sm;                 // [64][KP]
  float* Ks = Qs + BQ * KP;       // [64][KP]
  float* Vs = Ks + BKV * KP;      // [64][VP]
  float* Ss = Vs + BKV * VP;      // [64][SP]
  float* alpha_s = Ss + BQ * SP;  // [64]
  float* l_s = alpha_s + BQ;      // [64]
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * BQ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const T* qb = q_aug + (int64_t)bh * g.Nq * g.ld_qk;
  const T* kb = k_aug + (int64_t)bh * g.Nk * g.ld_qk;
  const T* vb = v + (int64_t)bh * g.Nk * g.ld_v;
  load_rows(qb, g.ld_qk, q0, g.Nq, g.kd, Qs, KP);

  // softmax bookkeeping: 4 threads per row (row = tid/4), each owns 16 columns of the score tile
  const int srow = threadIdx.x >> 2, spart = threadIdx.x & 3;
  float m_run = -INFINITY, l_run = 0.f;
  float o[4][6];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) o[i][j] = 0.f;

  for (int k0 = 0; k0 < g.Nk; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also covers the Q load on the first pass)
    load_rows(kb, g.ld_qk, k0, g.Nk, g.kd, Ks, KP);
    load_rows(vb, g.ld_v, k0, g.Nk, HD, Vs, VP);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int k = 0; k < g.kd; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(ty * 4 + i) * KP + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ks[(tx + 16 * j) * KP + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = (k0 + tx + 16 * j) < g.Nk;
        Ss[(ty * 4 + i) * SP + tx + 16 * j] = valid ? s[i][j] * g.scale : -INFINITY;
      }
    __syncthreads();
    {  // online softmax on row srow
      float* sr = Ss + srow * SP + spart * 16;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, sr[j]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);  // finite: every tile has at least one valid key
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p = expf(sr[j] - m_new);
        sr[j] = p;
        sum += p;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float alpha = expf(m_run - m_new);  // exp(-inf) = 0 on the first tile
      l_run = l_run * alpha + sum;
      m_run = m_new;
      if (spart == 0) alpha_s[srow] = alpha;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = alpha_s[ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 6; ++j) o[i][j] *= a;
    }
    for (int k = 0; k < BKV; ++k) {
      float p[4], vv[6];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Ss[(ty * 4 + i) * SP + k];
#pragma unroll
      for (int j = 0; j < 6; ++j) vv[j] = Vs[k * VP + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) o[i][j] = fmaf(p[i], vv[j], o[i][j]);
    }
  }
  if (spart == 0) l_s[srow] = l_run;
  if (spart == 0 && lse != nullptr && q0 + srow < g.Nq) lse[(int64_t)bh * g.Nq + q0 + srow] = m_run + logf(l_run);
  __syncthreads();
  const int b = bh / g.heads, head = bh - b * g.heads;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const int n = q0 + r;
    if (n >= g.Nq) continue;
    const float inv = 1.0f / l_s[r];
    const int64_t ooff = ((int64_t)b * g.Nq + n) * (g.heads * HD) + head * HD;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int c = tx + 16 * j;
      float val = o[i][j] * inv;
      if (out_pre != nullptr) out_pre[ooff + c] = from_f32<T>(val);  // pre-residual output, kept for backward
      if (g.residual && n >= 1) val += Qs[r * KP + c];  // residual pooling with the un-scaled q, cls row excluded
      out[ooff + c] = from_f32<T>(val);
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
// Q-tile owner: dQ' is exclusive to the CTA; dK / dV partial sums are atomically added into fp32 workspaces.
template <typename T>
__global__ void __launch_bounds__(THREADS) attn_bwd_kernel(const T* __restrict__ q_aug, const T* __restrict__ k_aug,
                                                           const T* __restrict__ v, const T* __restrict__ out,
                                                           const T* __restrict__ dout, const float* __restrict__ lse,
                                                           T* __restrict__ dq_aug, float* __restrict__ dk_ws,
                                                           float* __restrict__ dv_ws, AttnGeom g) {
  pdl_wait();
  extern __shared__ float sm[];
  const int KP = g.kd + 1;
  float* Qs = sm;                  // [64][KP]
  float* Ks = Qs + BQ * KP;        // [64][KP]
  float* Vs = Ks + BKV * KP;       // [64][VP]
  float* dOs = Vs + BKV * VP;      // [64][VP]
  float* Ps = dOs + BQ * VP;       // [64][SP]
  float* dSs = Ps + BQ * SP;       // [64][SP]
  float* delta_s = dSs + BQ * SP;  // [64]
  float* lse_s = delta_s + BQ;     // [64]
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * BQ;
  const int b = bh / g.heads, head = bh - b * g.heads;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const T* qb = q_aug + (int64_t)bh * g.Nq * g.ld_qk;
  const T* kb = k_aug + (int64_t)bh * g.Nk * g.ld_qk;
  const T* vb = v + (int64_t)bh * g.Nk * g.ld_v;
  const int64_t ld_o = (int64_t)g.heads * HD;
  const T* dob = dout + (int64_t)b * g.Nq * ld_o + head * HD;
  const T* ob = out + (int64_t)b * g.Nq * ld_o + head * HD;
  load_rows(qb, g.ld_qk, q0, g.Nq, g.kd, Qs, KP);
  load_rows(dob, ld_o, q0, g.Nq, HD, dOs, VP);
  __syncthreads();
  {  // delta[r] = sum_c dO[r][c] * O_attn[r][c]  (O_attn = attention output BEFORE the residual-pooling add)
    const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
    const int n = q0 + r;
    float d = 0.f;
    if (n < g.Nq) {
      for (int c = part * 24; c < part * 24 + 24; ++c) {
        d = fmaf(dOs[r * VP + c], to_f32(ob[(int64_t)n * ld_o + c]), d);
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    if (part == 0) {
      delta_s[r] = d;
      lse_s[r] = n < g.Nq ? lse[(int64_t)bh * g.Nq + n] : 0.f;
    }
  }
  constexpr int MAXC = 10;  // kd <= 160 -> at most 10 dQ columns per thread
  const int ncol = g.kd >> 4;
  float dq[4][MAXC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < MAXC; ++j) dq[i][j] = 0.f;

  for (int k0 = 0; k0 < g.Nk; k0 += BKV) {
    __syncthreads();
    load_rows(kb, g.ld_qk, k0, g.Nk, g.kd, Ks, KP);
    load_rows(vb, g.ld_v, k0, g.Nk, HD, Vs, VP);
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
    for (int k = 0; k < g.kd; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(ty * 4 + i) * KP + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Ks[(tx + 16 * j) * KP + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], bb[j], s[i][j]);
    }
    for (int k = 0; k < HD; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = dOs[(ty * 4 + i) * VP + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Vs[(tx + 16 * j) * VP + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dp[i][j] = fmaf(a[i], bb[j], dp[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const bool rvalid = (q0 + r) < g.Nq;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = rvalid && (k0 + tx + 16 * j) < g.Nk;
        const float p = valid ? expf(s[i][j] * g.scale - lse_s[r]) : 0.f;
        Ps[r * SP + tx + 16 * j] = p;
        dSs[r * SP + tx + 16 * j] = p * (dp[i][j] - delta_s[r]) * g.scale;
      }
    }
    __syncthreads();
    // dQ'[r][c] += sum_k dS[r][k] K'[k][c]
    for (int k = 0; k < BKV; ++k) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = dSs[(ty * 4 + i) * SP + k];
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        if (j < ncol) {
          const float kv = Ks[k * KP + tx + 16 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i) dq[i][j] = fmaf(a[i], kv, dq[i][j]);
        }
      }
    }
    // dK[k][c] += sum_r dS[r][k] Q[r][c] (c < 96) ; dV[k][c] += sum_r P[r][k] dO[r][c]
    float dk[4][6], dv[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
    for (int r = 0; r < BQ; ++r) {
      float ds4[4], p4[4], qv[6], dov[6];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ds4[i] = dSs[r * SP + ty * 4 + i]; p4[i] = Ps[r * SP + ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 6; ++j) { qv[j] = Qs[r * KP + tx + 16 * j]; dov[j] = dOs[r * VP + tx + 16 * j]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          dk[i][j] = fmaf(ds4[i], qv[j], dk[i][j]);
          dv[i][j] = fmaf(p4[i], dov[j], dv[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = k0 + ty * 4 + i;
      if (key >= g.Nk) continue;
      float* dkp = dk_ws + ((int64_t)bh * g.Nk + key) * HD;
      float* dvp = dv_ws + ((int64_t)bh * g.Nk + key) * HD;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        atomicAdd(dkp + tx + 16 * j, dk[i][j]);
        atomicAdd(dvp + tx + 16 * j, dv[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const int n = q0 + r;
    if (n >= g.Nq) continue;
    T* dqp = dq_aug + ((int64_t)bh * g.Nq + n) * g.ld_qk;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      if (j < ncol) {
        const int c = tx + 16 * j;
        float val = dq[i][j];
        if (g.residual && n >= 1 && c < HD) val += dOs[r * VP + c];  // residual-pooling path: d out / d q = I
        dqp[c] = from_f32<T>(val);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t rows, int64_t ld) {
  pdl_wait();
  const int64_t total = rows * (HD / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (HD / 4);
    const int c4 = (int)(i - r * (HD / 4));
    float v[4];
    load4(src + r * HD + c4 * 4, v);
    store4(dst + r * ld + c4 * 4, v);
  }
}

}  // namespace

int attn_simt_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, void* out,
                  void* out_pre, float* lse, int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, cudaStream_t stream) {
  AttnGeom g{B, heads, Nq, Nk, kd, ld_qk, ld_v, scale, residual};
  const size_t smem = ((size_t)(BQ + BKV) * (kd + 1) + (size_t)BKV * VP + (size_t)BQ * SP + 2 * BQ) * sizeof(float);
  dim3 grid((unsigned)ceil_div64(Nq, BQ), (unsigned)(B * heads));
  PMV_DISPATCH_DTYPE(dtype, T, {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    pmv_launch(attn_fwd_kernel<T>, grid, THREADS, smem, stream, (const T*)q_aug, (const T*)k_aug, (const T*)v, (T*)out, (T*)out_pre, lse, g);
  });
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

int attn_simt_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v,
                  const void* out, const void* dout, const float* lse,
                  void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                  int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, void* stream) {
  AttnGeom g{B, heads, Nq, Nk, kd, ld_qk, ld_v, scale, residual};
  const int64_t rows = (int64_t)B * heads * Nk;
  float* dk_ws = ws;
  float* dv_ws = ws + rows * HD;
  cudaStream_t st = (cudaStream_t)stream;
  PMV_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(2 * rows * HD) * sizeof(float), st));
  const size_t smem = ((size_t)(BQ + BKV) * (kd + 1) + (size_t)(BKV + BQ) * VP + (size_t)2 * BQ * SP + 2 * BQ) * sizeof(float);
  dim3 grid((unsigned)ceil_div64(Nq, BQ), (unsigned)(B * heads));
  int64_t cblocks = ceil_div64(rows * (HD / 4), 256);
  if (cblocks > 148 * 8) cblocks = 148 * 8;
  PMV_DISPATCH_DTYPE(dtype, T, {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    pmv_launch(attn_bwd_kernel<T>, grid, THREADS, smem, st, (const T*)q_aug, (const T*)k_aug, (const T*)v, (const T*)out,
                                                    (const T*)dout, lse, (T*)dq_aug, dk_ws, dv_ws, g);
    pmv_launch(cast_rows_kernel<T>, (unsigned)cblocks, 256, 0, st, dk_ws, (T*)dk, rows, ld_dk);
    pmv_launch(cast_rows_kernel<T>, (unsigned)cblocks, 256, 0, st, dv_ws, (T*)dv, rows, ld_dv);
  });
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
