// Residual-path MaxPool3d (1,3,3) / (1,2,2) / (0,1,1) on the fp32 token stream (attention.py:500-502, 558-564,
// 571-573): channels-last, cls token copied.  The forward records, per output element, which of the nine window
// positions won (first maximum in scan order, like ATen); the backward is a gather: every input element looks at the
// <= 4 windows that contain it and sums the gradients of those it won.  No atomics, no zero fill of dx.
// (The first version recomputed the arg-max and scattered with atomicAdd into a cleared buffer: 250 us at block 1.)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) maxpool_skip_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               uint8_t* __restrict__ win, int B, int T, int H, int W, int Ho, int Wo,
                                                               int C, FastDiv fdC4, FastDiv fdN, FastDiv fdW, FastDiv fdH) {
  pdl_wait();
  const int C4 = C >> 2;
  const int Lo = T * Ho * Wo, Li = T * H * W;
  const int64_t total = (int64_t)B * (Lo + 1) * C4;
  // 32-bit index arithmetic (the host checks total < 2^31) with multiply-shift divisions by the four loop-invariant
  // divisors: the eight hardware divisions per element (~20 instructions each) were half of the kernel's issue slots
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total; i += gridDim.x * blockDim.x) {
    uint32_t ru, c4u, bu, nu;
    fdC4.divmod(i, ru, c4u);
    fdN.divmod(ru, bu, nu);
    const int c4 = (int)c4u, n = (int)nu, b = (int)bu;
    const float* xb = x + (int64_t)b * (Li + 1) * C + c4 * 4;
    float m[4];
    uchar4 a = make_uchar4(0, 0, 0, 0);
    if (n == 0) {
      load4(xb, m);
    } else {
      uint32_t lq, wou, tu, hou;
      fdW.divmod((uint32_t)(n - 1), lq, wou);
      fdH.divmod(lq, tu, hou);
      const int wo = (int)wou, ho = (int)hou, t = (int)tu;
      bool first = true;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int hi = ho * 2 + dh - 1;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int wi = wo * 2 + dw - 1;
          if (wi < 0 || wi >= W) continue;
          float v[4];
          load4(xb + (int64_t)(1 + (t * H + hi) * W + wi) * C, v);
          const uint8_t p = (uint8_t)(dh * 3 + dw);
          if (first || v[0] > m[0] || v[0] != v[0]) { m[0] = v[0]; a.x = p; }  // NaN wins, like ATen
          if (first || v[1] > m[1] || v[1] != v[1]) { m[1] = v[1]; a.y = p; }  // NaN wins, like ATen
          if (first || v[2] > m[2] || v[2] != v[2]) { m[2] = v[2]; a.z = p; }  // NaN wins, like ATen
          if (first || v[3] > m[3] || v[3] != v[3]) { m[3] = v[3]; a.w = p; }  // NaN wins, like ATen
          first = false;
        }
      }
    }
    const int64_t o = ((int64_t)b * (Lo + 1) + n) * C + c4 * 4;
    store4(y + o, m);
    if (win != nullptr) *reinterpret_cast<uchar4*>(win + o) = a;
  }
}

__global__ void __launch_bounds__(256) maxpool_skip_bwd_kernel(const uint8_t* __restrict__ win, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int B, int T, int H, int W, int Ho, int Wo,
                                                               int C, FastDiv fdC4, FastDiv fdN, FastDiv fdW, FastDiv fdH) {
  pdl_wait();
  const int C4 = C >> 2;
  const int Lo = T * Ho * Wo, Li = T * H * W;
  const int64_t total = (int64_t)B * (Li + 1) * C4;
  // 32-bit index arithmetic (the host checks total < 2^31), multiply-shift divisions (see the forward kernel)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)total; i += gridDim.x * blockDim.x) {
    uint32_t ru, c4u, bu, nu;
    fdC4.divmod(i, ru, c4u);
    fdN.divmod(ru, bu, nu);
    const int c4 = (int)c4u, n = (int)nu, b = (int)bu;
    const int64_t ob = (int64_t)b * (Lo + 1) * C + c4 * 4;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (n == 0) {
      load4(dy + ob, g);
    } else {
      uint32_t lq, wu, tu, hu;
      fdW.divmod((uint32_t)(n - 1), lq, wu);
      fdH.divmod(lq, tu, hu);
      const int w = (int)wu, h = (int)hu, t = (int)tu;
      // windows containing row h: ho = h/2 with dh = 1 (h even), or ho = (h+1)/2 with dh = 0 and (h-1)/2 with dh = 2 (h odd).
      // All (up to four) winner words are requested before any is inspected, then the gradients of the windows this input
      // won: two dependent round trips per element instead of two per window.
      int ho2[2], dh2[2], wo2[2], dw2[2];
      ho2[0] = (h & 1) ? (h + 1) >> 1 : h >> 1; dh2[0] = (h & 1) ? 0 : 1;
      ho2[1] = (h - 1) >> 1;                    dh2[1] = 2;
      wo2[0] = (w & 1) ? (w + 1) >> 1 : w >> 1; dw2[0] = (w & 1) ? 0 : 1;
      wo2[1] = (w - 1) >> 1;                    dw2[1] = 2;
      const bool hv[2] = {ho2[0] < Ho, (h & 1) != 0};   // the second candidate exists only for odd coordinates
      const bool wv[2] = {wo2[0] < Wo, (w & 1) != 0};
      uchar4 a4[4];
      int64_t off[4];
      bool live[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = k >> 1, c = k & 1;
        live[k] = hv[a] && wv[c];
        off[k] = ob + (int64_t)(1 + (t * Ho + ho2[a]) * Wo + wo2[c]) * C;
        a4[k] = live[k] ? *reinterpret_cast<const uchar4*>(win + off[k]) : make_uchar4(255, 255, 255, 255);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint8_t p = (uint8_t)(dh2[k >> 1] * 3 + dw2[k & 1]);
        if (a4[k].x == p || a4[k].y == p || a4[k].z == p || a4[k].w == p) {
          float d[4];
          load4(dy + off[k], d);
          if (a4[k].x == p) g[0] += d[0];
          if (a4[k].y == p) g[1] += d[1];
          if (a4[k].z == p) g[2] += d[2];
          if (a4[k].w == p) g[3] += d[3];
        }
      }
    }
    store4(dx + ((int64_t)b * (Li + 1) + n) * C + c4 * 4, g);
  }
}

unsigned grid_for(int64_t items, int per_block, int max_blocks) {
  int64_t b = ceil_div64(items, per_block);
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

// `win` (uint8 [B, 1 + T*Ho*Wo, C], may be NULL for inference) receives the winning window position of every output.
extern "C" int pmv_maxpool_skip_fwd(const float* x, float* y, uint8_t* win, int B, int T, int H, int W, int C, void* stream) {
  PMV_CHECK_ARG(C % 4 == 0, "maxpool: C must be a multiple of 4");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * Ho * Wo) * (C / 4);
  PMV_CHECK_ARG((int64_t)B * (1 + (int64_t)T * H * W) * (C / 4) < (1ll << 31), "maxpool: too many elements");
  pmv_launch(maxpool_skip_fwd_kernel, grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream, x, y, win, B, T, H, W, Ho, Wo, C,
             FastDiv((uint32_t)(C / 4)), FastDiv((uint32_t)(1 + T * Ho * Wo)), FastDiv((uint32_t)Wo), FastDiv((uint32_t)Ho));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

// dx is OVERWRITTEN (every input element is produced exactly once).
extern "C" int pmv_maxpool_skip_bwd(const uint8_t* win, const float* dy, float* dx, int B, int T, int H, int W, int C, void* stream) {
  PMV_CHECK_ARG(C % 4 == 0, "maxpool: C must be a multiple of 4");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * (1 + (int64_t)T * H * W) * (C / 4);
  PMV_CHECK_ARG((int64_t)B * (1 + (int64_t)T * H * W) * (C / 4) < (1ll << 31), "maxpool: too many elements");
  pmv_launch(maxpool_skip_bwd_kernel, grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream, win, dy, dx, B, T, H, W, Ho, Wo, C,
             FastDiv((uint32_t)(C / 4)), FastDiv((uint32_t)(1 + T * H * W)), FastDiv((uint32_t)W), FastDiv((uint32_t)H));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
