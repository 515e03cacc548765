// Pooling kernels of the MViTv2 block.
//
//  * pool_ln_fwd / pool_ln_bwd : attention_pool() with the depthwise 3x3x3 Conv3d (stride (1,s,s),
//    pad 1, weights shared across heads) fused with the LayerNorm(96) that follows it
//    (attention.py:14-48, 241-282).  Channels-last: the kernel reads q / k / v straight out of the
//    QKV GEMM output [B, N, 3, heads, 96] (strided view) and writes [B, heads, 1+L', ld] — the three
//    permute().contiguous() copies, the cat and the separate LayerNorm of the reference disappear.
//  * maxpool_skip_fwd / bwd    : the residual-path MaxPool3d (1,3,3)/(1,2,2)/(0,1,1)
//    (attention.py:500-502, 558-564, 571-573).
//
// Layout of a thread group: 8 lanes own one output token, each lane 12 consecutive channels
// (24 B of bf16 / 48 B of fp32 -> 8- or 16-byte vector loads); a warp handles 4 tokens that are
// neighbours along w, so the 3-wide window overlap is served by L1.
#include "common.cuh"
#include "reduce.cuh"

namespace {

constexpr int HD = PMV_HEAD_DIM;  // 96
constexpr int TAPS = 27;
constexpr int CPL = 12;           // channels per lane
constexpr int LPT = 8;            // lanes per token
constexpr int POOL_THREADS = 256;
constexpr int TOK_PER_BLOCK = POOL_THREADS / LPT;

struct PoolGeom {
  int B, heads, T, H, W, Ho, Wo, s;
  int64_t in_bs, in_ts, in_hs;  // element strides of the input view
  int64_t out_ld;
};

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

// weights: reference layout [96][27] -> smem [27][96]
__device__ __forceinline__ void stage_weights(const float* __restrict__ w, float* sw) {
  for (int i = threadIdx.x; i < HD * TAPS; i += blockDim.x) {
    int c = i / TAPS, tap = i - c * TAPS;
    sw[tap * HD + c] = w[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_ln_fwd_kernel(
    const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ gamma,
    const float* __restrict__ beta, T* __restrict__ out, PoolGeom g, float eps) {
  __shared__ float sw[TAPS * HD];
  stage_weights(w, sw);
  __syncthreads();
  const int sub = threadIdx.x & (LPT - 1);
  const int c0 = sub * CPL;
  const int Lo = g.T * g.Ho * g.Wo;
  const int64_t ntok = (int64_t)g.B * g.heads * (Lo + 1);
  float gm[CPL], bt[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { gm[j] = gamma[c0 + j]; bt[j] = beta[c0 + j]; }

  for (int64_t tok = (int64_t)blockIdx.x * TOK_PER_BLOCK + threadIdx.x / LPT; tok < ntok;
       tok += (int64_t)gridDim.x * TOK_PER_BLOCK) {
    const int n = (int)(tok % (Lo + 1));
    const int64_t bh = tok / (Lo + 1);
    const int head = (int)(bh % g.heads);
    const int64_t b = bh / g.heads;
    const T* base = in + b * g.in_bs + head * g.in_hs + c0;
    float acc[CPL];
    if (n == 0) {
      load12(base, acc);  // cls token: no convolution (attention.py:25-26)
    } else {
      int l = n - 1;
      const int wo = l % g.Wo; l /= g.Wo;
      const int ho = l % g.Ho;
      const int t = l / g.Ho;
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ti = t + dt - 1;
        if (ti < 0 || ti >= g.T) continue;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int hi = ho * g.s + dh - 1;
          if (hi < 0 || hi >= g.H) continue;
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const int wi = wo * g.s + dw - 1;
            if (wi < 0 || wi >= g.W) continue;
            float xv[CPL];
            load12(base + (int64_t)(1 + (ti * g.H + hi) * g.W + wi) * g.in_ts, xv);
            const float* wt = sw + (dt * 9 + dh * 3 + dw) * HD + c0;
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[j] = fmaf(xv[j], wt[j], acc[j]);
          }
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) s += acc[j];
    const float mu = group_sum<LPT>(s) * (1.0f / HD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { float d = acc[j] - mu; q += d * d; }
    const float rs = rsqrtf(group_sum<LPT>(q) * (1.0f / HD) + eps);
    float o[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) o[j] = (acc[j] - mu) * rs * gm[j] + bt[j];
    store12(out + tok * g.out_ld + c0, o);
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: one warp per OUTPUT token, lane owns channels {lane, lane+32, lane+64}.
// Recomputes the convolution and the LN statistics, produces dconv (fp32 workspace), dgamma, dbeta
// and the 27x96 weight gradient (register accumulators, reduced through shared memory).
// The cls token's LN backward is written straight to din.
// ---------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(BWD_WARPS * 32) pool_ln_bwd_tokens_kernel(
    const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ gamma,
    const T* __restrict__ dout, int64_t dout_ld, T* __restrict__ din, float* __restrict__ partials,
    float* __restrict__ dconv, PoolGeom g, float eps) {
  __shared__ float sw[TAPS * HD];
  __shared__ float sred[(TAPS + 2) * HD];
  stage_weights(w, sw);
  for (int i = threadIdx.x; i < (TAPS + 2) * HD; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int Lo = g.T * g.Ho * g.Wo;
  const int64_t ntok = (int64_t)g.B * g.heads * (Lo + 1);
  float gm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gm[j] = gamma[lane + 32 * j];
  float adw[TAPS][3];
#pragma unroll
  for (int k = 0; k < TAPS; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) adw[k][j] = 0.f;
  float adg[3] = {0.f, 0.f, 0.f}, adb[3] = {0.f, 0.f, 0.f};

  for (int64_t tok = (int64_t)blockIdx.x * BWD_WARPS + (threadIdx.x >> 5); tok < ntok;
       tok += (int64_t)gridDim.x * BWD_WARPS) {
    const int n = (int)(tok % (Lo + 1));
    const int64_t bh = tok / (Lo + 1);
    const int head = (int)(bh % g.heads);
    const int64_t b = bh / g.heads;
    const int64_t base_off = b * g.in_bs + head * g.in_hs + lane;
    const T* base = in + base_off;
    float acc[3] = {0.f, 0.f, 0.f};
    int t = 0, ho = 0, wo = 0;
    if (n == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[j] = to_f32(base[32 * j]);
    } else {
      int l = n - 1;
      wo = l % g.Wo; l /= g.Wo;
      ho = l % g.Ho;
      t = l / g.Ho;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ti = t + dt - 1;
        if (ti < 0 || ti >= g.T) continue;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int hi = ho * g.s + dh - 1;
          if (hi < 0 || hi >= g.H) continue;
#pragma unroll
          for (int dwi = 0; dwi < 3; ++dwi) {
            const int wi = wo * g.s + dwi - 1;
            if (wi < 0 || wi >= g.W) continue;
            const T* p = base + (int64_t)(1 + (ti * g.H + hi) * g.W + wi) * g.in_ts;
            const float* wt = sw + (dt * 9 + dh * 3 + dwi) * HD + lane;
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[j] = fmaf(to_f32(p[32 * j]), wt[32 * j], acc[j]);
          }
        }
      }
    }
    const float mu = warp_sum(acc[0] + acc[1] + acc[2]) * (1.0f / HD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { float d = acc[j] - mu; q += d * d; }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / HD) + eps);
    const T* dyr = dout + tok * dout_ld + lane;
    float xh[3], gg[3], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float dy = to_f32(dyr[32 * j]);
      xh[j] = (acc[j] - mu) * rs;
      gg[j] = dy * gm[j];
      s1 += gg[j];
      s2 += gg[j] * xh[j];
      adg[j] += dy * xh[j];
      adb[j] += dy;
    }
    s1 = warp_sum(s1) * (1.0f / HD);
    s2 = warp_sum(s2) * (1.0f / HD);
    float dc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) dc[j] = rs * (gg[j] - s1 - xh[j] * s2);
    if (n == 0) {
      T* dp = din + base_off;
#pragma unroll
      for (int j = 0; j < 3; ++j) dp[32 * j] = from_f32<T>(dc[j]);
      continue;
    }
    float* dcr = dconv + (bh * Lo + (n - 1)) * HD + lane;
#pragma unroll
    for (int j = 0; j < 3; ++j) dcr[32 * j] = dc[j];
    // weight gradient: dw[tap][c] += x[neighbour(tap)][c] * dconv[c]   (inputs are L1-hot)
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int ti = t + dt - 1;
      if (ti < 0 || ti >= g.T) continue;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int hi = ho * g.s + dh - 1;
        if (hi < 0 || hi >= g.H) continue;
#pragma unroll
        for (int dwi = 0; dwi < 3; ++dwi) {
          const int wi = wo * g.s + dwi - 1;
          if (wi < 0 || wi >= g.W) continue;
          const T* p = base + (int64_t)(1 + (ti * g.H + hi) * g.W + wi) * g.in_ts;
#pragma unroll
          for (int j = 0; j < 3; ++j) adw[dt * 9 + dh * 3 + dwi][j] = fmaf(to_f32(p[32 * j]), dc[j], adw[dt * 9 + dh * 3 + dwi][j]);
        }
      }
    }
  }
  // block reduction through shared memory
#pragma unroll
  for (int k = 0; k < TAPS; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(&sred[k * HD + lane + 32 * j], adw[k][j]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    atomicAdd(&sred[TAPS * HD + lane + 32 * j], adg[j]);
    atomicAdd(&sred[(TAPS + 1) * HD + lane + 32 * j], adb[j]);
  }
  __syncthreads();
  // one partial vector per block, already in the destination order: dw in the reference layout [96][27], dgamma, dbeta
  float* pb = partials + (int64_t)blockIdx.x * ((TAPS + 2) * HD);
  for (int i = threadIdx.x; i < TAPS * HD; i += blockDim.x) {
    const int tap = i / HD, c = i - tap * HD;
    pb[c * TAPS + tap] = sred[i];
  }
  for (int i = threadIdx.x; i < 2 * HD; i += blockDim.x) pb[TAPS * HD + i] = sred[TAPS * HD + i];
}

// backward, pass 2: gather form of the transposed stencil — one 8-lane group per INPUT token,
// no atomics: din[ti,hi,wi][c] = sum over taps with (hi+1-dh) % s == 0 of w[c][tap] * dconv[to,ho,wo][c].
template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_ln_bwd_input_kernel(
    const float* __restrict__ w, const float* __restrict__ dconv, T* __restrict__ din, PoolGeom g) {
  __shared__ float sw[TAPS * HD];
  stage_weights(w, sw);
  __syncthreads();
  const int sub = threadIdx.x & (LPT - 1);
  const int c0 = sub * CPL;
  const int Li = g.T * g.H * g.W;
  const int Lo = g.T * g.Ho * g.Wo;
  const int64_t ntok = (int64_t)g.B * g.heads * Li;
  for (int64_t tok = (int64_t)blockIdx.x * TOK_PER_BLOCK + threadIdx.x / LPT; tok < ntok;
       tok += (int64_t)gridDim.x * TOK_PER_BLOCK) {
    int l = (int)(tok % Li);
    const int64_t bh = tok / Li;
    const int head = (int)(bh % g.heads);
    const int64_t b = bh / g.heads;
    const int wi = l % g.W; l /= g.W;
    const int hi = l % g.H;
    const int ti = l / g.H;
    float acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int to = ti + 1 - dt;
      if (to < 0 || to >= g.T) continue;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int nh = hi + 1 - dh;
        if (nh < 0 || nh % g.s != 0) continue;
        const int ho = nh / g.s;
        if (ho >= g.Ho) continue;
#pragma unroll
        for (int dwi = 0; dwi < 3; ++dwi) {
          const int nw = wi + 1 - dwi;
          if (nw < 0 || nw % g.s != 0) continue;
          const int wo = nw / g.s;
          if (wo >= g.Wo) continue;
          float dv[CPL];
          load12(dconv + (bh * Lo + (int64_t)(to * g.Ho + ho) * g.Wo + wo) * HD + c0, dv);
          const float* wt = sw + (dt * 9 + dh * 3 + dwi) * HD + c0;
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[j] = fmaf(dv[j], wt[j], acc[j]);
        }
      }
    }
    T* dp = din + b * g.in_bs + head * g.in_hs + (int64_t)(1 + (ti * g.H + hi) * g.W + wi) * g.in_ts + c0;
    store12(dp, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// skip-path max pool (fp32 residual stream), kernel (1,3,3) stride (1,2,2) pad (0,1,1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_skip_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               int B, int T, int H, int W, int Ho, int Wo, int C) {
  const int C4 = C >> 2;
  const int64_t Lo = (int64_t)T * Ho * Wo, Li = (int64_t)T * H * W;
  const int64_t total = (int64_t)B * (Lo + 1) * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    int64_t r = i / C4;
    const int64_t n = r % (Lo + 1);
    const int64_t b = r / (Lo + 1);
    const float* xb = x + b * (Li + 1) * C + c4 * 4;
    float m[4];
    if (n == 0) {
      load4(xb, m);
    } else {
      int64_t l = n - 1;
      const int wo = (int)(l % Wo); l /= Wo;
      const int ho = (int)(l % Ho);
      const int t = (int)(l / Ho);
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = -INFINITY;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int hi = ho * 2 + dh - 1;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int wi = wo * 2 + dw - 1;
          if (wi < 0 || wi >= W) continue;
          float v[4];
          load4(xb + (1 + ((int64_t)t * H + hi) * W + wi) * C, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) m[j] = fmaxf(m[j], v[j]);
        }
      }
    }
    store4(y + (b * (Lo + 1) + n) * C + c4 * 4, m);
  }
}

__global__ void __launch_bounds__(256) maxpool_skip_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int B, int T, int H, int W,
                                                               int Ho, int Wo, int C) {
  const int64_t Lo = (int64_t)T * Ho * Wo, Li = (int64_t)T * H * W;
  const int64_t total = (int64_t)B * (Lo + 1) * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int64_t n = r % (Lo + 1);
    const int64_t b = r / (Lo + 1);
    const float g = dy[i];
    const float* xb = x + b * (Li + 1) * C + c;
    float* dxb = dx + b * (Li + 1) * C + c;
    if (n == 0) {
      atomicAdd(dxb, g);
      continue;
    }
    int64_t l = n - 1;
    const int wo = (int)(l % Wo); l /= Wo;
    const int ho = (int)(l % Ho);
    const int t = (int)(l / Ho);
    float m = -INFINITY;
    int64_t arg = -1;
    for (int dh = 0; dh < 3; ++dh) {
      const int hi = ho * 2 + dh - 1;
      if (hi < 0 || hi >= H) continue;
      for (int dw = 0; dw < 3; ++dw) {
        const int wi = wo * 2 + dw - 1;
        if (wi < 0 || wi >= W) continue;
        const int64_t off = (1 + ((int64_t)t * H + hi) * W + wi) * C;
        const float v = xb[off];
        if (v > m || arg < 0) { m = v; arg = off; }  // first maximum in scan order (ATen max_pool3d)
      }
    }
    atomicAdd(dxb + arg, g);
  }
}

int check_geom(int B, int heads, int T, int H, int W, int s, int64_t ts, int64_t hs, int64_t ld, int dtype) {
  PMV_CHECK_ARG(B > 0 && heads > 0 && T > 0 && H > 0 && W > 0 && s >= 1, "pool: bad geometry");
  const int al = dtype == PMV_BF16 ? 4 : 4;  // 8-byte (bf16) / 16-byte (fp32) vector accesses
  PMV_CHECK_ARG(ts % al == 0 && hs % al == 0 && ld % al == 0 && ld >= HD, "pool: strides must be multiples of %d elements", al);
  return PMV_OK;
}

PoolGeom make_geom(int B, int heads, int T, int H, int W, int s, int64_t bs, int64_t ts, int64_t hs, int64_t ld) {
  PoolGeom g;
  g.B = B; g.heads = heads; g.T = T; g.H = H; g.W = W; g.s = s;
  g.Ho = (H - 1) / s + 1;  // (H + 2*1 - 3)/s + 1
  g.Wo = (W - 1) / s + 1;
  g.in_bs = bs; g.in_ts = ts; g.in_hs = hs; g.out_ld = ld;
  return g;
}

unsigned grid_for(int64_t items, int per_block, int max_blocks) {
  int64_t b = ceil_div64(items, per_block);
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" int pmv_pool_ln_fwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                               const float* w, const float* gamma, const float* beta, void* out, int64_t out_ld,
                               int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream) {
  int rc = check_geom(B, heads, T, H, W, stride_hw, in_token_stride, in_head_stride, out_ld, dtype);
  if (rc) return rc;
  PoolGeom g = make_geom(B, heads, T, H, W, stride_hw, in_batch_stride, in_token_stride, in_head_stride, out_ld);
  const int64_t ntok = (int64_t)B * heads * (1 + (int64_t)T * g.Ho * g.Wo);
  unsigned grid = grid_for(ntok, TOK_PER_BLOCK, 148 * 32);
  PMV_DISPATCH_DTYPE(dtype, TT, (pool_ln_fwd_kernel<TT><<<grid, POOL_THREADS, 0, (cudaStream_t)stream>>>(
                                    (const TT*)in, w, gamma, beta, (TT*)out, g, eps)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int64_t pmv_pool_ln_bwd_workspace_bytes(int B, int heads, int T, int H, int W, int stride_hw) {
  const int Ho = (H - 1) / stride_hw + 1, Wo = (W - 1) / stride_hw + 1;
  const int64_t ntok_out = (int64_t)B * heads * (1 + (int64_t)T * Ho * Wo);
  const int64_t dconv = (int64_t)B * heads * T * Ho * Wo * HD;
  return (dconv + (int64_t)grid_for(ntok_out, BWD_WARPS * 8, 148 * 2) * (TAPS + 2) * HD) * (int64_t)sizeof(float);
}

extern "C" int pmv_pool_ln_bwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                               const float* w, const float* gamma, const void* dout, int64_t dout_ld,
                               void* din, float* dw_dgamma_dbeta, float* ws,
                               int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream) {
  int rc = check_geom(B, heads, T, H, W, stride_hw, in_token_stride, in_head_stride, dout_ld, dtype);
  if (rc) return rc;
  PoolGeom g = make_geom(B, heads, T, H, W, stride_hw, in_batch_stride, in_token_stride, in_head_stride, dout_ld);
  const int64_t ntok_out = (int64_t)B * heads * (1 + (int64_t)T * g.Ho * g.Wo);
  const int64_t ntok_in = (int64_t)B * heads * T * H * W;
  unsigned grid1 = grid_for(ntok_out, BWD_WARPS * 8, 148 * 2);
  unsigned grid2 = grid_for(ntok_in, TOK_PER_BLOCK, 148 * 32);
  float* dconv_ws = ws;
  float* partials = ws + (int64_t)B * heads * T * g.Ho * g.Wo * HD;
  PMV_DISPATCH_DTYPE(dtype, TT, {
    pool_ln_bwd_tokens_kernel<TT><<<grid1, BWD_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const TT*)in, w, gamma, (const TT*)dout, dout_ld, (TT*)din, partials, dconv_ws, g, eps);
    pool_ln_bwd_input_kernel<TT><<<grid2, POOL_THREADS, 0, (cudaStream_t)stream>>>(w, dconv_ws, (TT*)din, g);
  });
  launch_reduce_partials(partials, (int)grid1, (TAPS + 2) * HD, dw_dgamma_dbeta, (cudaStream_t)stream);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int pmv_maxpool_skip_fwd(const float* x, float* y, int B, int T, int H, int W, int C, void* stream) {
  PMV_CHECK_ARG(C % 4 == 0, "maxpool: C must be a multiple of 4");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * Ho * Wo) * (C / 4);
  maxpool_skip_fwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, y, B, T, H, W, Ho, Wo, C);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int pmv_maxpool_skip_bwd(const float* x, const float* dy, float* dx, int B, int T, int H, int W, int C, void* stream) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * Ho * Wo) * C;
  maxpool_skip_bwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, B, T, H, W, Ho, Wo, C);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
