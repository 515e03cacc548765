// PatchEmbed (stem_helper.py:293-325): Conv3d(3 -> 96, k (3,7,7), s (2,4,4), p (1,3,3)) lowered to a GEMM.
// This kernel gathers each output token's receptive field (K = Cin*kt*kh*kw = 441 values, ordered like the
// reference weight [96, Cin, kt, kh, kw] flattened) into one row of `col`; the product with the weight, the
// bias and the scatter behind the cls slot run in pmv_gemm (row-remap epilogue), and the weight gradient is
// the same col matrix through the wgrad GEMM.
#include "common.cuh"

namespace {
// One CTA per (clip, output frame, output row): the Cin*kt*kh input rows that row of tokens touches are staged in
// shared memory once (coalesced loads, zero rows / margins for the padding), then every thread owns two
// adjacent columns k of `col` — its (channel, dt, dh, dw) decode is done once — and walks the Wo tokens: one shared
// load per value, 4-byte coalesced stores.  (The first version decoded eight 64-bit div/mod per element: 0.3 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ clip, T* __restrict__ col, int64_t ld,
                                                     int B, int Cin, int Tn, int H, int W, int To, int Ho, int Wo,
                                                     int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                                                     int span, int pitch) {
  pdl_wait();
  extern __shared__ float rows_s[];  // [Cin*kt*kh][pitch]; element j of a row is input column j - pw
  const int K = Cin * kt * kh * kw;
  const int nrows = Cin * kt * kh;
  int r = blockIdx.x;
  const int ho = r % Ho; r /= Ho;
  const int to = r % To;
  const int b = r / To;
  // ---- stage
  if ((W & 3) == 0 && span >= W + pw) {
    // 16-byte loads over the flattened (row, float4) index: ~14 independent loads per thread instead of ~56 scalar ones in
    // a per-row lane loop (the staging phase, not the emit phase, was most of this kernel's 145 us); the left / right
    // padding columns of every row are cleared separately
    const int W4 = W >> 2;
    for (int idx = threadIdx.x; idx < nrows * W4; idx += blockDim.x) {
      const int rr = idx / W4, v4 = idx - rr * W4;
      const int dh = rr % kh;
      const int dt = (rr / kh) % kt;
      const int c = rr / (kh * kt);
      const int ti = to * st + dt - pt, hi = ho * sh + dh - ph;
      const bool ok = ti >= 0 && ti < Tn && hi >= 0 && hi < H;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = __ldg(reinterpret_cast<const float4*>(clip + (((int64_t)(b * Cin + c) * Tn + ti) * H + hi) * W) + v4);
      float* dst = rows_s + rr * pitch + pw + 4 * v4;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    const int npad = pw + (span - W - pw);  // columns left of 0 and right of W - 1
    for (int idx = threadIdx.x; idx < nrows * npad; idx += blockDim.x) {
      const int rr = idx / npad, j = idx - rr * npad;
      rows_s[rr * pitch + (j < pw ? j : W + j)] = 0.f;
    }
  } else {
    for (int rr = threadIdx.x >> 5; rr < nrows; rr += blockDim.x >> 5) {
      const int dh = rr % kh;
      const int dt = (rr / kh) % kt;
      const int c = rr / (kh * kt);
      const int ti = to * st + dt - pt, hi = ho * sh + dh - ph;
      const bool ok = ti >= 0 && ti < Tn && hi >= 0 && hi < H;
      const float* src = clip + (((int64_t)(b * Cin + c) * Tn + (ok ? ti : 0)) * H + (ok ? hi : 0)) * W;
      float* dst = rows_s + rr * pitch;
      for (int j = threadIdx.x & 31; j < span; j += 32) {
        const int wi = j - pw;
        dst[j] = (ok && wi >= 0 && wi < W) ? __ldg(src + wi) : 0.f;
      }
    }
  }
  __syncthreads();
  // ---- emit: thread = column pair (k, k+1)
  const int64_t tok0 = ((int64_t)(b * To + to) * Ho + ho) * Wo;
  for (int k = 2 * threadIdx.x; k < (int)ld; k += 2 * blockDim.x) {
    int off[2];
    bool live[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = k + h;
      live[h] = kk < K;
      const int dw = kk % kw;
      const int row = kk / kw;  // (c*kt + dt)*kh + dh: the staging order
      off[h] = live[h] ? row * pitch + dw : 0;
    }
    T* out = col + tok0 * ld + k;
    for (int wo = 0; wo < Wo; ++wo) {
      const float v0 = live[0] ? rows_s[off[0] + wo * sw] : 0.f;
      const float v1 = live[1] ? rows_s[off[1] + wo * sw] : 0.f;
      if (sizeof(T) == 2) {
        __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
        *reinterpret_cast<uint32_t*>(out + (int64_t)wo * ld) = *reinterpret_cast<uint32_t*>(&pk);
      } else {
        *reinterpret_cast<float2*>(out + (int64_t)wo * ld) = make_float2(v0, v1);
      }
    }
  }
}
}  // namespace

extern "C" int pmv_patch_im2col(const float* clip, void* col, int64_t ld_col, int B, int Cin, int T, int H, int W,
                                int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                                int dtype, void* stream) {
  const int To = (T + 2 * pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  PMV_CHECK_ARG(ld_col >= Cin * kt * kh * kw && ld_col % 2 == 0, "im2col: ld_col too small or odd");
  const int span = (Wo - 1) * sw + kw;       // input columns one row of tokens touches (starting at -pw)
  const int pitch = span | 1;                 // odd pitch: rows start in different banks
  const size_t smem = (size_t)Cin * kt * kh * pitch * sizeof(float);
  PMV_CHECK_ARG(smem <= 200 * 1024, "im2col: receptive rows do not fit in shared memory (%zu B)", smem);
  const int64_t blocks = (int64_t)B * To * Ho;
  PMV_DISPATCH_DTYPE(dtype, TT, {
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
      PMV_CHECK_CUDA(cudaFuncSetAttribute(im2col_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    pmv_launch(im2col_kernel<TT>, (unsigned)blocks, 256, smem, (cudaStream_t)stream, clip, (TT*)col, ld_col, B, Cin, T, H, W, To, Ho, Wo, kt,
                                                                             kh, kw, st, sh, sw, pt, ph, pw, span, pitch);
  });
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
