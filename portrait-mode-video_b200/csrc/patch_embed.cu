// PatchEmbed (stem_helper.py:293-325): Conv3d(3 -> 96, k (3,7,7), s (2,4,4), p (1,3,3)) lowered to a GEMM.
// This kernel gathers each output token's receptive field (K = Cin*kt*kh*kw = 441 values, ordered like the
// reference weight [96, Cin, kt, kh, kw] flattened) into one row of `col`; the product with the weight, the
// bias and the scatter behind the cls slot run in pmv_gemm (row-remap epilogue), and the weight gradient is
// the same col matrix through the wgrad GEMM.
#include "common.cuh"

namespace {
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ clip, T* __restrict__ col, int64_t ld,
                                                     int B, int Cin, int Tn, int H, int W, int To, int Ho, int Wo,
                                                     int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw) {
  const int K = Cin * kt * kh * kw;
  const int64_t total = (int64_t)B * To * Ho * Wo * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % ld);
    int64_t r = i / ld;
    float v = 0.f;
    if (k < K) {
      const int wo = (int)(r % Wo); r /= Wo;
      const int ho = (int)(r % Ho); r /= Ho;
      const int to = (int)(r % To);
      const int64_t b = r / To;
      int kk = k;
      const int dw = kk % kw; kk /= kw;
      const int dh = kk % kh; kk /= kh;
      const int dt = kk % kt;
      const int c = kk / kt;
      const int ti = to * st + dt - pt, hi = ho * sh + dh - ph, wi = wo * sw + dw - pw;
      if (ti >= 0 && ti < Tn && hi >= 0 && hi < H && wi >= 0 && wi < W)
        v = clip[(((b * Cin + c) * Tn + ti) * H + hi) * (int64_t)W + wi];
    }
    col[i] = from_f32<T>(v);
  }
}
}  // namespace

extern "C" int pmv_patch_im2col(const float* clip, void* col, int64_t ld_col, int B, int Cin, int T, int H, int W,
                                int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                                int dtype, void* stream) {
  const int To = (T + 2 * pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  PMV_CHECK_ARG(ld_col >= Cin * kt * kh * kw, "im2col: ld_col too small");
  const int64_t total = (int64_t)B * To * Ho * Wo * ld_col;
  int64_t blocks = ceil_div64(total, 256 * 4);
  if (blocks > 148 * 32) blocks = 148 * 32;
  PMV_DISPATCH_DTYPE(dtype, TT, (im2col_kernel<TT><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
                                    clip, (TT*)col, ld_col, B, Cin, T, H, W, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
