// Residual-path MaxPool3d (1,3,3) / (1,2,2) / (0,1,1) on the fp32 token stream (attention.py:500-502, 558-564,
// 571-573): channels-last, cls token copied; backward recomputes the arg-max (first maximum in scan order, like
// ATen) and scatters with atomics (overlapping windows).
#include "common.cuh"

namespace {

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

unsigned grid_for(int64_t items, int per_block, int max_blocks) {
  int64_t b = ceil_div64(items, per_block);
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

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
