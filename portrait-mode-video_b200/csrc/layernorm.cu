#define PMV_PDL_FAMILY 16
// LayerNorm forward / backward over the channel axis of the fp32 residual stream.
// Replaces native_layer_norm at attention.py:567,578 and video_model_builder.py:2163.
// Memory-bound: 8 / 16 / 32 lanes per row, 16-byte loads, two-pass (mean, then centred variance) in fp32.
#include "common.cuh"
#include "reduce.cuh"

namespace {

constexpr int LN_WARPS = 8;

// A row of C = 4 * LPR * VPL channels is owned by LPR lanes (LPR in {8, 16, 32}), each with VPL float4: a warp
// works on 32 / LPR rows at once, every lane has VPL independent 16-byte loads in flight and consecutive lanes read
// consecutive 16-byte pieces.  (The first version gave one row to one warp whatever C was: at C = 96 a warp had
// 384 bytes in flight and 8 idle lanes, 1.0 TB/s.)
template <int LPR, int VPL, typename TOut>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    TOut* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, float eps) {
  pdl_wait();
  constexpr int C = 4 * LPR * VPL;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rw = lane / LPR;
  const int64_t row0 = ((int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * RPW + rw;
  const int64_t rstep = (int64_t)gridDim.x * LN_WARPS * RPW;
  float g[VPL][4], b[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    load4(gamma + (sub + i * LPR) * 4, g[i]);
    load4(beta + (sub + i * LPR) * 4, b[i]);
  }
  for (int64_t r0 = row0 - rw; r0 < rows; r0 += rstep) {  // warp-uniform trip count (shuffles inside)
    const int64_t r = r0 + rw;
    const bool ok = r < rows;
    const float* xr = x + (ok ? r : 0) * C;
    float v[VPL][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      load4(xr + (sub + i * LPR) * 4, v[i]);
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mu = group_sum<LPR>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = v[i][j] - mu; q += d * d; }
    const float rs = rsqrtf(group_sum<LPR>(q) * (1.0f / C) + eps);
    if (!ok) continue;
    if (sub == 0 && mean != nullptr) { mean[r] = mu; rstd[r] = rs; }
    TOut* yr = y + r * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mu) * rs * g[i][j] + b[i][j];
      store4(yr + (sub + i * LPR) * 4, o);
    }
  }
}

// Backward.  Per row:  xhat = (x-mu)*rstd, g = dy*gamma,
//   dx = rstd * (g - mean(g) - xhat*mean(g*xhat));  dgamma += dy*xhat; dbeta += dy.
// dgamma/dbeta: per-lane partials over the rows this lane visits -> shuffle over the rows of the warp -> smem over
// the warps -> one partial vector [2][C] per block (folded by reduce_partials_kernel).
// The row loop is software pipelined (the loads of the next row group are issued before the current one is reduced:
// a warp otherwise sits through one DRAM round trip per row, 26 us for 12 552 x 384 where the traffic needs 10) and, for
// VPL <= 3, registers are capped so that two blocks fit an SM (two row groups in flight per warp).
// four gradient values as loaded: fp32 as they are, bf16 still packed (half the registers of the prefetched row group)
template <typename T> struct Raw4;
template <> struct Raw4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void get(float* o) const { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct Raw4<bf16> {
  uint2 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void get(float* o) const {
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
};

template <int VPL, typename TDy> struct LnRow {
  float xv[VPL][4];
  Raw4<TDy> dv[VPL];
  float mu, rs;
  int64_t r;
  bool ok;
};

constexpr int ln_bwd_min_blocks(int vpl) { return vpl <= 3 ? 2 : 1; }

template <int LPR, int VPL, typename TDy>
__global__ void __launch_bounds__(LN_WARPS * 32, ln_bwd_min_blocks(VPL)) layernorm_bwd_kernel(
    const TDy* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* dx, const float* dx_base,
    float* __restrict__ partials, float* __restrict__ dgb, int64_t rows) {
  pdl_wait();
  constexpr int C = 4 * LPR * VPL;
  constexpr int RPW = 32 / LPR;
  constexpr bool PIPE = VPL <= 3;  // wider rows keep one row group in registers only
  __shared__ float red[LN_WARPS][2 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rw = lane / LPR;
  const int64_t row0 = ((int64_t)blockIdx.x * LN_WARPS + warp) * RPW;
  const int64_t rstep = (int64_t)gridDim.x * LN_WARPS * RPW;
  // the fold kernel that follows ADDS the per-CTA partials into dgamma / dbeta: clear them here (saves a fill launch)
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) dgb[i] = 0.f;
  float gm[VPL][4], dg[VPL][4], db[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    load4(gamma + (sub + i * LPR) * 4, gm[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) dg[i][j] = db[i][j] = 0.f;
  }
  auto fetch = [&](int64_t r0, LnRow<VPL, TDy>& R) {
    R.r = r0 + rw;
    R.ok = R.r < rows;
    const int64_t rr = R.ok ? R.r : 0;
    R.mu = mean[rr];
    R.rs = R.ok ? rstd[rr] : 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      load4(x + rr * C + (sub + i * LPR) * 4, R.xv[i]);
      R.dv[i].load(dy + rr * C + (sub + i * LPR) * 4);
    }
  };
  auto consume = [&](const LnRow<VPL, TDy>& R) {
    // gradient arriving over the residual connection (may alias dx): requested first, used last - the row reduction
    // below and the other resident warps cover its latency (keeping it in the prefetched row group spilled)
    float pv[VPL][4];
    if (dx_base != nullptr && R.ok) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) load4(dx_base + R.r * C + (sub + i * LPR) * 4, pv[i]);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float d4[4];
      R.dv[i].get(d4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dv = R.ok ? d4[j] : 0.f;
        const float xh = (R.xv[i][j] - R.mu) * R.rs;
        const float g = dv * gm[i][j];
        s1 += g;
        s2 += g * xh;
        dg[i][j] += dv * xh;
        db[i][j] += dv;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / C);
    s2 *= (1.0f / C);
    if (!R.ok) return;
    float* dxr = dx + R.r * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float o[4], d4[4];
      R.dv[i].get(d4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // xhat and g are recomputed (2 FMAs) rather than kept: registers buy the second row group
        const float xh = (R.xv[i][j] - R.mu) * R.rs;
        o[j] = R.rs * (d4[j] * gm[i][j] - s1 - xh * s2);
        if (dx_base != nullptr) o[j] += pv[i][j];
      }
      store4(dxr + (sub + i * LPR) * 4, o);
    }
  };
  if constexpr (PIPE) {
    LnRow<VPL, TDy> A, B;
    if (row0 < rows) fetch(row0, A);
    for (int64_t r0 = row0; r0 < rows; r0 += 2 * rstep) {
      const bool has_b = r0 + rstep < rows;
      if (has_b) fetch(r0 + rstep, B);
      consume(A);
      if (r0 + 2 * rstep < rows) fetch(r0 + 2 * rstep, A);
      if (has_b) consume(B);
    }
  } else {
    LnRow<VPL, TDy> A;
    for (int64_t r0 = row0; r0 < rows; r0 += rstep) {
      fetch(r0, A);
      consume(A);
    }
  }
  // fold the RPW rows of the warp (lanes with the same `sub`), then the warps
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        dg[i][j] += __shfl_xor_sync(0xffffffffu, dg[i][j], o);
        db[i][j] += __shfl_xor_sync(0xffffffffu, db[i][j], o);
      }
    }
  if (rw == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[warp][(sub + i * LPR) * 4 + j] = dg[i][j];
        red[warp][C + (sub + i * LPR) * 4 + j] = db[i][j];
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[w][i];
    partials[(int64_t)blockIdx.x * 2 * C + i] = s;
  }
}

template <int LPR, int VPL>
int launch_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype, float* mean, float* rstd, int64_t rows,
               float eps, cudaStream_t st) {
  int64_t blocks = ceil_div64(rows, LN_WARPS * (32 / LPR) * 2);
  if (blocks > 148 * 8) blocks = 148 * 8;
  PMV_DISPATCH_DTYPE(y_dtype, T, (pmv_launch(layernorm_fwd_kernel<LPR, VPL, T>, (unsigned)blocks, LN_WARPS * 32, 0, st, 
                                     x, gamma, beta, (T*)y, mean, rstd, rows, eps)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

int64_t ln_bwd_blocks(int64_t rows) {
  int64_t blocks = ceil_div64(rows, LN_WARPS * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;  // one resident wave at two blocks per SM
  return blocks < 1 ? 1 : blocks;
}

template <int LPR, int VPL>
int launch_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx,
               const float* dx_base, float* ws, float* dgb, int64_t rows, cudaStream_t st) {
  const int64_t blocks = ln_bwd_blocks(rows);
  PMV_DISPATCH_DTYPE(dy_dtype, T, (pmv_launch(layernorm_bwd_kernel<LPR, VPL, T>, (unsigned)blocks, LN_WARPS * 32, 0, st, 
                                      (const T*)dy, x, gamma, mean, rstd, dx, dx_base, ws, dgb, rows)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

// C -> (lanes per row, float4 per lane)
#define PMV_LN_DISPATCH(C, CALL)                                   \
  switch (C) {                                                     \
    case 96: { constexpr int LPR = 8, VPL = 3; CALL; } break;      \
    case 192: { constexpr int LPR = 16, VPL = 3; CALL; } break;    \
    case 384: { constexpr int LPR = 32, VPL = 3; CALL; } break;    \
    case 768: { constexpr int LPR = 32, VPL = 6; CALL; } break;    \
    case 128: { constexpr int LPR = 32, VPL = 1; CALL; } break;    \
    case 256: { constexpr int LPR = 32, VPL = 2; CALL; } break;    \
    case 512: { constexpr int LPR = 32, VPL = 4; CALL; } break;    \
    case 32: { constexpr int LPR = 8, VPL = 1; CALL; } break;      \
    case 64: { constexpr int LPR = 16, VPL = 1; CALL; } break;     \
    default:                                                       \
      pmv_set_error("layernorm: unsupported channel count %d (supported: 32 64 96 128 192 256 384 512 768)", C); \
      return PMV_ERR_UNSUPPORTED;                                  \
  }

}  // namespace

extern "C" int pmv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                                 float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
  PMV_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "layernorm: mean and rstd must both be given or both NULL");
  if (rows == 0) return PMV_OK;
  int rc = PMV_OK;
  PMV_LN_DISPATCH(C, (rc = launch_fwd<LPR, VPL>(x, gamma, beta, y, y_dtype, mean, rstd, rows, eps, (cudaStream_t)stream)));
  return rc;
}

extern "C" int64_t pmv_layernorm_bwd_workspace_bytes(int64_t rows, int C) { return ln_bwd_blocks(rows) * 2 * C * (int64_t)sizeof(float); }

extern "C" int pmv_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                                 const float* mean, const float* rstd, float* dx, const float* dx_base,
                                 float* dgamma_dbeta, float* ws, int64_t rows, int C, void* stream) {
  if (rows == 0) return PMV_OK;
  int rc = PMV_OK;
  PMV_LN_DISPATCH(C, (rc = launch_bwd<LPR, VPL>(dy, dy_dtype, x, gamma, mean, rstd, dx, dx_base, ws, dgamma_dbeta, rows, (cudaStream_t)stream)));
  if (rc) return rc;
  launch_reduce_partials(ws, (int)ln_bwd_blocks(rows), 2 * C, dgamma_dbeta, (cudaStream_t)stream);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
