// LayerNorm forward / backward over the channel axis of the fp32 residual stream.
// Replaces native_layer_norm at attention.py:567,578 and video_model_builder.py:2163.
// Memory-bound: one warp per row, 16-byte loads, two-pass (mean, then centred variance) in fp32.
#include "common.cuh"
#include "reduce.cuh"

namespace {

constexpr int LN_WARPS = 8;

template <typename TOut>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    TOut* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  const int nvec = C >> 2;  // C % 4 == 0 checked on the host
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float* xr = x + r * C;
    // C <= 768 in MViTv2: at most 6 float4 per lane; keep the row in registers
    float v[8][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c4 = lane + i * 32;
      if (c4 < nvec) {
        load4(xr + c4 * 4, v[i]);
        s += v[i][0] + v[i][1] + v[i][2] + v[i][3];
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c4 = lane + i * 32;
      if (c4 < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float d = v[i][j] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0 && mean != nullptr) { mean[r] = mu; rstd[r] = rs; }
    TOut* yr = y + r * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c4 = lane + i * 32;
      if (c4 < nvec) {
        float g[4], b[4], o[4];
        load4(gamma + c4 * 4, g);
        load4(beta + c4 * 4, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
        store4(yr + c4 * 4, o);
      }
    }
  }
}

// Backward.  Per row:  xhat = (x-mu)*rstd, g = dy*gamma,
//   dx = rstd * (g - mean(g) - xhat*mean(g*xhat));  dgamma += dy*xhat; dbeta += dy.
// dgamma/dbeta: per-thread partials over the rows this warp visits -> smem reduce over warps ->
// one partial vector per block (folded by reduce_partials_kernel; dgamma and dbeta must be adjacent: [2][C]).
template <typename TDy>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(
    const TDy* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx, int accumulate,
    float* __restrict__ partials, int64_t rows, int C) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * LN_WARPS + warp;
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  const int nvec = C >> 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float dg[8][4], db[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dg[i][j] = db[i][j] = 0.f;

  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    const float* xr = x + r * C;
    const TDy* dyr = dy + r * C;
    float xh[8][4], g[8][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c4 = lane + i * 32;
      if (c4 < nvec) {
        float xv[4], dv[4], gm[4];
        load4(xr + c4 * 4, xv);
        load4(dyr + c4 * 4, dv);
        load4(gamma + c4 * 4, gm);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[i][j] = (xv[j] - mu) * rs;
          g[i][j] = dv[j] * gm[j];
          s1 += g[i][j];
          s2 += g[i][j] * xh[i][j];
          dg[i][j] += dv[j] * xh[i][j];
          db[i][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    float* dxr = dx + r * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c4 = lane + i * 32;
      if (c4 < nvec) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rs * (g[i][j] - s1 - xh[i][j] * s2);
        if (accumulate) {
          float p[4];
          load4(dxr + c4 * 4, p);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] += p[j];
        }
        store4(dxr + c4 * 4, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c4 = lane + i * 32;
    if (c4 < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&red[c4 * 4 + j], dg[i][j]);
        atomicAdd(&red[C + c4 * 4 + j], db[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partials[(int64_t)blockIdx.x * 2 * C + i] = red[i];
}

}  // namespace

extern "C" int pmv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                                 float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
  PMV_CHECK_ARG(C % 4 == 0 && C <= 1024 && C > 0, "layernorm: C=%d must be a multiple of 4 and <= 1024", C);
  PMV_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "layernorm: mean and rstd must both be given or both NULL");
  if (rows == 0) return PMV_OK;
  int64_t blocks = ceil_div64(rows, LN_WARPS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  PMV_DISPATCH_DTYPE(y_dtype, T, (layernorm_fwd_kernel<T><<<(unsigned)blocks, LN_WARPS * 32, 0, (cudaStream_t)stream>>>(
                                     x, gamma, beta, (T*)y, mean, rstd, rows, C, eps)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

static int64_t ln_bwd_blocks(int64_t rows) {
  int64_t blocks = ceil_div64(rows, LN_WARPS * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;
  return blocks < 1 ? 1 : blocks;
}

extern "C" int64_t pmv_layernorm_bwd_workspace_bytes(int64_t rows, int C) { return ln_bwd_blocks(rows) * 2 * C * (int64_t)sizeof(float); }

extern "C" int pmv_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                                 const float* mean, const float* rstd, float* dx, int accumulate,
                                 float* dgamma_dbeta, float* ws, int64_t rows, int C, void* stream) {
  PMV_CHECK_ARG(C % 4 == 0 && C <= 1024 && C > 0, "layernorm: C=%d must be a multiple of 4 and <= 1024", C);
  if (rows == 0) return PMV_OK;
  const int64_t blocks = ln_bwd_blocks(rows);
  size_t smem = 2 * (size_t)C * sizeof(float);
  PMV_DISPATCH_DTYPE(dy_dtype, T, (layernorm_bwd_kernel<T><<<(unsigned)blocks, LN_WARPS * 32, smem, (cudaStream_t)stream>>>(
                                      (const T*)dy, x, gamma, mean, rstd, dx, accumulate, ws, rows, C)));
  launch_reduce_partials(ws, (int)blocks, 2 * C, dgamma_dbeta, (cudaStream_t)stream);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
