// attention_pool() with the depthwise 3x3x3 Conv3d (stride (1,s,s), pad 1, weights shared across heads) fused
// with the LayerNorm(96) that follows it (attention.py:14-48, 241-282), forward and backward, for q, k and v
// in ONE launch each.  Channels-last: the kernels read q / k / v straight out of the QKV GEMM output
// [B, N, 3, heads, 96] (strided view) and write [B, heads, 1+L', ld] — the three permute().contiguous()
// copies, the cat and the separate LayerNorm of the reference disappear.
//
// forward       : 8 lanes own one output token, each lane 12 consecutive channels (8/16-byte vector loads);
//                 a warp handles 4 tokens that are neighbours along w, so the window overlap is served by L1.
// backward (i)  : one warp per OUTPUT token, lane owns channels {lane, lane+32, lane+64}: the convolution and the
//                 LN statistics are recomputed, LN backward gives the pre-LN gradient (fp32 workspace) and the
//                 dgamma / dbeta partials.  Light on registers on purpose: the kernel is latency-bound (one token
//                 per warp at a time, three warp reductions on the critical path), so it lives on occupancy.
// backward (ii) : weight gradient dW[c][tap] = sum_tok x[nb(tok, tap)][c] * dconv[tok][c]: thread = (channel,
//                 temporal tap), 9 exclusive accumulators, no reductions and no dependences between tokens, so
//                 loads of several tokens are in flight -> one partial [96][27] per CTA.
// backward (iii): gather form of the transposed stencil per INPUT token (no atomics), written straight into the
//                 interleaved dQKV buffer.
// backward (iv) : reduce kernels fold the per-CTA partials (same-address global atomics from hundreds of CTAs
//                 serialise in L2).
#include "common.cuh"

namespace {

constexpr int HD = PMV_HEAD_DIM;  // 96
constexpr int TAPS = 27;
constexpr int CPL = 12;           // channels per lane (forward / input-gradient kernels)
constexpr int LPT = 8;            // lanes per token
constexpr int POOL_THREADS = 256;
constexpr int TOK_PER_BLOCK = POOL_THREADS / LPT;
constexpr int BWD_WARPS = 8;
constexpr int NGRAD = (TAPS + 2) * HD;  // dW [96][27], dgamma [96], dbeta [96]
constexpr int NDW = TAPS * HD;
constexpr int DW_THREADS = 3 * HD;     // (temporal tap, channel)
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
  void* din;            // backward: gradient wrt `in` (same strides)
  float* grads;         // backward: [NGRAD] fp32, added to
  float* dconv;         // backward: fp32 workspace [B*heads*Lo*96]
  int s, Ho, Wo;
  int blk_begin, nblk;  // block range of this job in the current launch
  int blk2_begin, nblk2;  // block range in the dW kernel
};

struct Launch {
  Job job[MAX_JOBS];
  int njobs;
  int B, heads, T, H, W;
  int64_t in_bs, in_ts, in_hs;  // element strides of the input views
  float eps;
  float* partials;              // token kernel: [total blocks][2 * 96] (dgamma, dbeta)
  float* partials_dw;           // dW kernel: [total blocks][96 * 27]
};

__device__ __forceinline__ int find_job(const Launch& L) {
  int j = 0;
  while (j + 1 < L.njobs && (int)blockIdx.x >= L.job[j + 1].blk_begin) ++j;
  return j;
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

// weights: reference layout [96][27] -> smem [27][96]
__device__ __forceinline__ void stage_weights(const float* __restrict__ w, float* sw) {
  for (int i = threadIdx.x; i < HD * TAPS; i += blockDim.x) {
    int c = i / TAPS, tap = i - c * TAPS;
    sw[tap * HD + c] = w[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_ln_fwd_kernel(const __grid_constant__ Launch L) {
  __shared__ float sw[TAPS * HD];
  const Job& J = L.job[find_job(L)];
  stage_weights(J.w, sw);
  __syncthreads();
  const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
  T* __restrict__ out = reinterpret_cast<T*>(J.out);
  const int sub = threadIdx.x & (LPT - 1);
  const int c0 = sub * CPL;
  const int Lo = L.T * J.Ho * J.Wo;
  const int64_t ntok = (int64_t)L.B * L.heads * (Lo + 1);
  float gm[CPL], bt[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { gm[j] = J.gamma[c0 + j]; bt[j] = J.beta[c0 + j]; }
  const int lb = blockIdx.x - J.blk_begin;
  for (int64_t tok = (int64_t)lb * TOK_PER_BLOCK + threadIdx.x / LPT; tok < ntok; tok += (int64_t)J.nblk * TOK_PER_BLOCK) {
    const int n = (int)(tok % (Lo + 1));
    const int64_t bh = tok / (Lo + 1);
    const int head = (int)(bh % L.heads);
    const int64_t b = bh / L.heads;
    const T* base = in + b * L.in_bs + head * L.in_hs + c0;
    float acc[CPL];
    if (n == 0) {
      load12(base, acc);  // cls token: no convolution (attention.py:25-26)
    } else {
      int l = n - 1;
      const int wo = l % J.Wo; l /= J.Wo;
      const int ho = l % J.Ho;
      const int t = l / J.Ho;
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ti = t + dt - 1;
        if (ti < 0 || ti >= L.T) continue;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int hi = ho * J.s + dh - 1;
          if (hi < 0 || hi >= L.H) continue;
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const int wi = wo * J.s + dw - 1;
            if (wi < 0 || wi >= L.W) continue;
            float xv[CPL];
            load12(base + (int64_t)(1 + (ti * L.H + hi) * L.W + wi) * L.in_ts, xv);
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
    const float rs = rsqrtf(group_sum<LPT>(q) * (1.0f / HD) + L.eps);
    float o[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) o[j] = (acc[j] - mu) * rs * gm[j] + bt[j];
    store12(out + tok * J.out_ld + c0, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(BWD_WARPS * 32, 4) pool_ln_bwd_tokens_kernel(const __grid_constant__ Launch L) {
  __shared__ float sw[TAPS * HD];
  __shared__ float sred[2 * HD];
  const Job& J = L.job[find_job(L)];
  stage_weights(J.w, sw);
  for (int i = threadIdx.x; i < 2 * HD; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
  const T* __restrict__ dout = reinterpret_cast<const T*>(J.dout);
  T* __restrict__ din = reinterpret_cast<T*>(J.din);
  const int lane = threadIdx.x & 31;
  const int Lo = L.T * J.Ho * J.Wo;
  const int64_t ntok = (int64_t)L.B * L.heads * (Lo + 1);
  float gm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gm[j] = J.gamma[lane + 32 * j];
  float adg[3] = {0.f, 0.f, 0.f}, adb[3] = {0.f, 0.f, 0.f};
  const int lb = blockIdx.x - J.blk_begin;

  for (int64_t tok = (int64_t)lb * BWD_WARPS + (threadIdx.x >> 5); tok < ntok; tok += (int64_t)J.nblk * BWD_WARPS) {
    const int n = (int)(tok % (Lo + 1));
    const int64_t bh = tok / (Lo + 1);
    const int head = (int)(bh % L.heads);
    const int64_t b = bh / L.heads;
    const int64_t base_off = b * L.in_bs + head * L.in_hs + lane;
    const T* base = in + base_off;
    const T* dyr = dout + tok * J.dout_ld + lane;
    float dy[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) dy[j] = to_f32(dyr[32 * j]);
    float acc[3] = {0.f, 0.f, 0.f};
    if (n == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[j] = to_f32(base[32 * j]);
    } else {
      int l = n - 1;
      const int wo = l % J.Wo; l /= J.Wo;
      const int ho = l % J.Ho;
      const int t = l / J.Ho;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int ti = t + dt - 1;
        if (ti < 0 || ti >= L.T) continue;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int hi = ho * J.s + dh - 1;
          if (hi < 0 || hi >= L.H) continue;
          float xv[3][3];
#pragma unroll
          for (int dwi = 0; dwi < 3; ++dwi) {  // the three loads of a window row are issued together
            const int wi = wo * J.s + dwi - 1;
            const bool ok = wi >= 0 && wi < L.W;
            const T* p = base + (int64_t)(1 + (ti * L.H + hi) * L.W + (ok ? wi : 0)) * L.in_ts;
#pragma unroll
            for (int j = 0; j < 3; ++j) xv[dwi][j] = ok ? to_f32(p[32 * j]) : 0.f;
          }
#pragma unroll
          for (int dwi = 0; dwi < 3; ++dwi) {
            const float* wt = sw + (dt * 9 + dh * 3 + dwi) * HD + lane;
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[j] = fmaf(xv[dwi][j], wt[32 * j], acc[j]);
          }
        }
      }
    }
    const float mu = warp_sum(acc[0] + acc[1] + acc[2]) * (1.0f / HD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { float d = acc[j] - mu; q += d * d; }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / HD) + L.eps);
    float xh[3], gg[3], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      xh[j] = (acc[j] - mu) * rs;
      gg[j] = dy[j] * gm[j];
      s1 += gg[j];
      s2 += gg[j] * xh[j];
      adg[j] += dy[j] * xh[j];
      adb[j] += dy[j];
    }
    // the two remaining reductions share their five shuffle rounds
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / HD);
    s2 *= (1.0f / HD);
    float dc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) dc[j] = rs * (gg[j] - s1 - xh[j] * s2);
    if (n == 0) {
      T* dp = din + base_off;
#pragma unroll
      for (int j = 0; j < 3; ++j) dp[32 * j] = from_f32<T>(dc[j]);
      continue;
    }
    float* dcr = J.dconv + (bh * Lo + (n - 1)) * HD + lane;
#pragma unroll
    for (int j = 0; j < 3; ++j) dcr[32 * j] = dc[j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    atomicAdd(&sred[lane + 32 * j], adg[j]);
    atomicAdd(&sred[HD + lane + 32 * j], adb[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * HD; i += blockDim.x) L.partials[(int64_t)blockIdx.x * 2 * HD + i] = sred[i];
}

// dW: thread = (temporal tap dt, channel c); the CTA walks its share of the output tokens, all threads on the same
// token (block-uniform index math), 9 exclusive accumulators per thread.
template <typename T>
__global__ void __launch_bounds__(DW_THREADS) pool_ln_bwd_dw_kernel(const __grid_constant__ Launch L) {
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk2_begin) ++jj;
  const Job& J = L.job[jj];
  const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
  const int c = threadIdx.x % HD, dt = threadIdx.x / HD;
  const int Lo = L.T * J.Ho * J.Wo;
  const int64_t ntok = (int64_t)L.B * L.heads * Lo;
  float a[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = 0.f;
  const int lb = blockIdx.x - J.blk2_begin;
#pragma unroll 2
  for (int64_t tok = lb; tok < ntok; tok += J.nblk2) {
    int l = (int)(tok % Lo);
    const int64_t bh = tok / Lo;
    const int head = (int)(bh % L.heads);
    const int64_t b = bh / L.heads;
    const int wo = l % J.Wo; l /= J.Wo;
    const int ho = l % J.Ho;
    const int t = l / J.Ho;
    const int ti = t + dt - 1;
    if (ti < 0 || ti >= L.T) continue;
    const float d = J.dconv[tok * HD + c];
    const T* base = in + b * L.in_bs + head * L.in_hs + c;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hi = ho * J.s + dh - 1;
      if (hi < 0 || hi >= L.H) continue;
#pragma unroll
      for (int dwi = 0; dwi < 3; ++dwi) {
        const int wi = wo * J.s + dwi - 1;
        if (wi < 0 || wi >= L.W) continue;
        a[dh * 3 + dwi] = fmaf(to_f32(base[(int64_t)(1 + (ti * L.H + hi) * L.W + wi) * L.in_ts]), d, a[dh * 3 + dwi]);
      }
    }
  }
  float* pb = L.partials_dw + (int64_t)blockIdx.x * NDW + c * TAPS + dt * 9;  // reference layout [96][27]
#pragma unroll
  for (int k = 0; k < 9; ++k) pb[k] = a[k];
}

// gather form of the transposed stencil — one 8-lane group per INPUT token:
// din[ti,hi,wi][c] = sum over taps with (hi+1-dh) % s == 0 of w[c][tap] * dconv[to,ho,wo][c].
template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_ln_bwd_input_kernel(const __grid_constant__ Launch L) {
  __shared__ float sw[TAPS * HD];
  const Job& J = L.job[find_job(L)];
  stage_weights(J.w, sw);
  __syncthreads();
  T* __restrict__ din = reinterpret_cast<T*>(J.din);
  const int sub = threadIdx.x & (LPT - 1);
  const int c0 = sub * CPL;
  const int Li = L.T * L.H * L.W;
  const int Lo = L.T * J.Ho * J.Wo;
  const int64_t ntok = (int64_t)L.B * L.heads * Li;
  const int lb = blockIdx.x - J.blk_begin;
  for (int64_t tok = (int64_t)lb * TOK_PER_BLOCK + threadIdx.x / LPT; tok < ntok; tok += (int64_t)J.nblk * TOK_PER_BLOCK) {
    int l = (int)(tok % Li);
    const int64_t bh = tok / Li;
    const int head = (int)(bh % L.heads);
    const int64_t b = bh / L.heads;
    const int wi = l % L.W; l /= L.W;
    const int hi = l % L.H;
    const int ti = l / L.H;
    float acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int to = ti + 1 - dt;
      if (to < 0 || to >= L.T) continue;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int nh = hi + 1 - dh;
        if (nh < 0 || nh % J.s != 0) continue;
        const int ho = nh / J.s;
        if (ho >= J.Ho) continue;
#pragma unroll
        for (int dwi = 0; dwi < 3; ++dwi) {
          const int nw = wi + 1 - dwi;
          if (nw < 0 || nw % J.s != 0) continue;
          const int wo = nw / J.s;
          if (wo >= J.Wo) continue;
          float dv[CPL];
          load12(J.dconv + (bh * Lo + (int64_t)(to * J.Ho + ho) * J.Wo + wo) * HD + c0, dv);
          const float* wt = sw + (dt * 9 + dh * 3 + dwi) * HD + c0;
#pragma unroll
          for (int j = 0; j < CPL; ++j) acc[j] = fmaf(dv[j], wt[j], acc[j]);
        }
      }
    }
    T* dp = din + b * L.in_bs + head * L.in_hs + (int64_t)(1 + (ti * L.H + hi) * L.W + wi) * L.in_ts + c0;
    store12(dp, acc);
  }
}

// grads_j[i] += sum over the job's blocks of the partial vectors: dW from the dW kernel, dgamma / dbeta from the
// token kernel      (grid.y = job, grid.z = slice of the blocks)
constexpr int RED_SLICES = 8;
__global__ void __launch_bounds__(256) reduce_jobs_kernel(const __grid_constant__ Launch L) {
  const Job& J = L.job[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NGRAD) return;
  const bool is_dw = i < NDW;
  const float* src = is_dw ? L.partials_dw + i : L.partials + (i - NDW);
  const int64_t stride = is_dw ? NDW : 2 * HD;
  const int begin = is_dw ? J.blk2_begin : J.blk_begin;
  const int end = begin + (is_dw ? J.nblk2 : J.nblk);
  float s0 = 0.f, s1 = 0.f;
  int b = begin + blockIdx.z;
  for (; b + RED_SLICES < end; b += 2 * RED_SLICES) {
    s0 += src[(int64_t)b * stride];
    s1 += src[(int64_t)(b + RED_SLICES) * stride];
  }
  if (b < end) s0 += src[(int64_t)b * stride];
  atomicAdd(J.grads + i, s0 + s1);
}

int nblocks_for(int64_t items, int per_block, int max_blocks) {
  int64_t b = ceil_div64(items, per_block);
  if (b > max_blocks) b = max_blocks;
  return b < 1 ? 1 : (int)b;
}

int out_hw(int n, int s) { return (n - 1) / s + 1; }  // (n + 2*1 - 3) / s + 1

int64_t ntok_out(int B, int heads, int T, int H, int W, int s) { return (int64_t)B * heads * (1 + (int64_t)T * out_hw(H, s) * out_hw(W, s)); }

// block budgets of the backward token / dW kernels per job (shared by the workspace query and the launcher)
int bwd_token_blocks(int B, int heads, int T, int H, int W, int s) { return nblocks_for(ntok_out(B, heads, T, H, W, s), BWD_WARPS * 4, 148 * 2); }
int bwd_dw_blocks(int B, int heads, int T, int H, int W, int s) { return nblocks_for(ntok_out(B, heads, T, H, W, s), 32, 148 * 2); }

int fill_launch(Launch& L, const void* qkv, int64_t bs, int64_t ts, int64_t ws_, int64_t hs, const pmv_pool_job* jobs, int njobs,
                int B, int heads, int T, int H, int W, float eps, int dtype) {
  PMV_CHECK_ARG(njobs >= 1 && njobs <= MAX_JOBS, "pool: 1..3 jobs");
  PMV_CHECK_ARG(B > 0 && heads > 0 && T > 0 && H > 0 && W > 0, "pool: bad geometry");
  PMV_CHECK_ARG(ts % 4 == 0 && hs % 4 == 0 && ws_ % 4 == 0, "pool: strides must be multiples of 4 elements");
  const int esz = dtype == PMV_BF16 ? 2 : 4;
  L.njobs = njobs; L.B = B; L.heads = heads; L.T = T; L.H = H; L.W = W;
  L.in_bs = bs; L.in_ts = ts; L.in_hs = hs; L.eps = eps; L.partials = nullptr; L.partials_dw = nullptr;
  for (int i = 0; i < njobs; ++i) {
    Job& J = L.job[i];
    const pmv_pool_job& p = jobs[i];
    PMV_CHECK_ARG(p.stride_hw >= 1, "pool: bad stride");
    J.in = reinterpret_cast<const char*>(qkv) + (int64_t)p.which * ws_ * esz;
    J.w = p.w; J.gamma = p.gamma; J.beta = p.beta; J.out = p.out; J.out_ld = p.out_ld;
    J.dout = p.dout; J.dout_ld = p.dout_ld; J.din = nullptr; J.grads = p.grads; J.dconv = nullptr;
    J.s = p.stride_hw; J.Ho = out_hw(H, p.stride_hw); J.Wo = out_hw(W, p.stride_hw);
    J.blk_begin = 0; J.nblk = 0; J.blk2_begin = 0; J.nblk2 = 0;
  }
  return PMV_OK;
}

}  // namespace

extern "C" int pmv_pool_ln_qkv_fwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride,
                                   int64_t head_stride, const pmv_pool_job* jobs, int njobs,
                                   int B, int heads, int T, int H, int W, float eps, int dtype, void* stream) {
  Launch L;
  int rc = fill_launch(L, qkv, batch_stride, token_stride, which_stride, head_stride, jobs, njobs, B, heads, T, H, W, eps, dtype);
  if (rc) return rc;
  int total = 0;
  for (int i = 0; i < njobs; ++i) {
    PMV_CHECK_ARG(jobs[i].out != nullptr && jobs[i].out_ld % 4 == 0 && jobs[i].out_ld >= HD, "pool: bad output");
    L.job[i].blk_begin = total;
    L.job[i].nblk = nblocks_for(ntok_out(B, heads, T, H, W, jobs[i].stride_hw), TOK_PER_BLOCK, 148 * 16);
    total += L.job[i].nblk;
  }
  PMV_DISPATCH_DTYPE(dtype, TT, (pool_ln_fwd_kernel<TT><<<(unsigned)total, POOL_THREADS, 0, (cudaStream_t)stream>>>(L)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int64_t pmv_pool_ln_qkv_bwd_workspace_bytes(int B, int heads, int T, int H, int W, const int* strides_hw, int njobs) {
  int64_t floats = 0;
  for (int i = 0; i < njobs; ++i) {
    floats += (ntok_out(B, heads, T, H, W, strides_hw[i]) - (int64_t)B * heads) * HD;  // pre-LN gradient, non-cls tokens
    floats += (int64_t)bwd_token_blocks(B, heads, T, H, W, strides_hw[i]) * 2 * HD;
    floats += (int64_t)bwd_dw_blocks(B, heads, T, H, W, strides_hw[i]) * NDW;
  }
  return floats * (int64_t)sizeof(float);
}

extern "C" int pmv_pool_ln_qkv_bwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride,
                                   int64_t head_stride, const pmv_pool_job* jobs, int njobs, void* dqkv, float* ws,
                                   int B, int heads, int T, int H, int W, float eps, int dtype, void* stream) {
  Launch L;
  int rc = fill_launch(L, qkv, batch_stride, token_stride, which_stride, head_stride, jobs, njobs, B, heads, T, H, W, eps, dtype);
  if (rc) return rc;
  const int esz = dtype == PMV_BF16 ? 2 : 4;
  float* cursor = ws;
  int total = 0, total_dw = 0;
  for (int i = 0; i < njobs; ++i) {
    PMV_CHECK_ARG(jobs[i].dout != nullptr && jobs[i].grads != nullptr && jobs[i].dout_ld % 4 == 0, "pool: bad backward job");
    Job& J = L.job[i];
    J.din = reinterpret_cast<char*>(dqkv) + (int64_t)jobs[i].which * which_stride * esz;
    J.dconv = cursor;
    cursor += (ntok_out(B, heads, T, H, W, J.s) - (int64_t)B * heads) * HD;
    J.blk_begin = total;
    J.nblk = bwd_token_blocks(B, heads, T, H, W, J.s);
    total += J.nblk;
    J.blk2_begin = total_dw;
    J.nblk2 = bwd_dw_blocks(B, heads, T, H, W, J.s);
    total_dw += J.nblk2;
  }
  L.partials = cursor;
  L.partials_dw = cursor + (int64_t)total * 2 * HD;
  cudaStream_t st = (cudaStream_t)stream;
  PMV_DISPATCH_DTYPE(dtype, TT, {
    pool_ln_bwd_tokens_kernel<TT><<<(unsigned)total, BWD_WARPS * 32, 0, st>>>(L);
    pool_ln_bwd_dw_kernel<TT><<<(unsigned)total_dw, DW_THREADS, 0, st>>>(L);
  });
  reduce_jobs_kernel<<<dim3((NGRAD + 255) / 256, njobs, RED_SLICES), 256, 0, st>>>(L);
  // second launch geometry: one block range per job over the INPUT tokens
  Launch L2 = L;
  int total2 = 0;
  const int64_t ntok_in = (int64_t)B * heads * T * H * W;
  for (int i = 0; i < njobs; ++i) {
    L2.job[i].blk_begin = total2;
    L2.job[i].nblk = nblocks_for(ntok_in, TOK_PER_BLOCK, 148 * 16);
    total2 += L2.job[i].nblk;
  }
  PMV_DISPATCH_DTYPE(dtype, TT, (pool_ln_bwd_input_kernel<TT><<<(unsigned)total2, POOL_THREADS, 0, st>>>(L2)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

// ---- single-tensor entry points (one job) -------------------------------------------------------------------------
extern "C" int pmv_pool_ln_fwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                               const float* w, const float* gamma, const float* beta, void* out, int64_t out_ld,
                               int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream) {
  pmv_pool_job j;
  memset(&j, 0, sizeof(j));
  j.w = w; j.gamma = gamma; j.beta = beta; j.out = out; j.out_ld = out_ld; j.stride_hw = stride_hw; j.which = 0;
  return pmv_pool_ln_qkv_fwd(in, in_batch_stride, in_token_stride, 0, in_head_stride, &j, 1, B, heads, T, H, W, eps, dtype, stream);
}

extern "C" int64_t pmv_pool_ln_bwd_workspace_bytes(int B, int heads, int T, int H, int W, int stride_hw) {
  return pmv_pool_ln_qkv_bwd_workspace_bytes(B, heads, T, H, W, &stride_hw, 1);
}

extern "C" int pmv_pool_ln_bwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                               const float* w, const float* gamma, const void* dout, int64_t dout_ld,
                               void* din, float* dw_dgamma_dbeta, float* ws,
                               int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream) {
  pmv_pool_job j;
  memset(&j, 0, sizeof(j));
  j.w = w; j.gamma = gamma; j.dout = dout; j.dout_ld = dout_ld; j.grads = dw_dgamma_dbeta; j.stride_hw = stride_hw; j.which = 0;
  return pmv_pool_ln_qkv_bwd(in, in_batch_stride, in_token_stride, 0, in_head_stride, &j, 1, din, ws, B, heads, T, H, W, eps, dtype,
                             stream);
}
