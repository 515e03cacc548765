#define PMV_PDL_FAMILY 32
// attention_pool() with the depthwise 3x3x3 Conv3d (stride (1,s,s), pad 1, weights shared across heads) fused
// with the LayerNorm(96) that follows it (attention.py:14-48, 241-282), forward and backward, for q, k and v
// in ONE launch each.  Channels-last: the kernels read q / k / v straight out of the QKV GEMM output
// [B, N, 3, heads, 96] (strided view) and write [B, heads, 1+L', ld] — the three permute().contiguous()
// copies, the cat and the separate LayerNorm of the reference disappear.
//
// All stencil kernels share one shape ("row march"): a CTA of 192 threads = 4 output rows x 48 channel PAIRS.
// A thread keeps the 27 taps of its two channels in registers (packed fp32 pairs -> FFMA2, 27 per output pair)
// and walks along w.  For stride 1 the 3x3 (t,h) x 3 (w) input window lives in registers and slides: 9 new
// 4-byte loads per output instead of 27, every load of a warp is a contiguous 128-byte run of channels.
// The convolution results of 4 rows x 6 positions are parked in shared memory; after one __syncthreads the
// CTA re-maps to 24 tokens x 8 lanes (12 channels per lane) for the LayerNorm (3 shuffle rounds per
// reduction) and writes 192-byte token rows.  The buffer is double-buffered: one barrier per 24 tokens.
//
// forward        : conv -> LN -> out.
// backward (i)   : conv recomputed -> LN statistics -> LN backward -> pre-LN gradient `dconv` (workspace, in the
//                  compute dtype) + dgamma / dbeta partials per CTA.  The cls token (no convolution) is handled
//                  by a few extra CTAs, one warp per token.
// backward (ii)  : dW[c][tap] += x[nb(tok, tap)][c] * dconv[tok][c]: the same march with the window of x and 27
//                  exclusive accumulator pairs per thread -> one partial [96][27] per CTA.
// backward (iii) : input gradient, gather form of the transposed stencil per INPUT row (no atomics), written
//                  straight into the interleaved dQKV buffer.  Stride 1: the forward march over dconv with
//                  flipped taps.  Stride >= 2: only taps with (h+1-dh) % s == 0 contribute (none for most rows at
//                  stride 4 / 8): guarded taps, zero rows written without arithmetic.
// backward (iv)  : a reduce kernel folds the per-CTA partials (same-address global atomics from hundreds of CTAs
//                  serialise in L2).
#include <cstdlib>

#include "pool_common.cuh"

namespace pool {
namespace {

constexpr int CW = 6;              // positions along w between two LayerNorm phases (multiple of 3: window rotation)
constexpr int THREADS = NCP * ROWS;  // 192
constexpr int CHUNK_TOK = ROWS * CW;  // 24 tokens per phase == THREADS / LNL
constexpr int CLS_WARPS = THREADS / 32;
static_assert(CHUNK_TOK * LNL == THREADS, "LayerNorm phase mapping");

struct Launch {
  Job job[MAX_JOBS];
  int njobs;
  int B, heads, T, H, W;
  int64_t in_bs, in_ts, in_hs;  // element strides of the input views
  float eps;
};

// One row of the march: where the 3x3 (t,h) neighbourhood of the row lives.
struct RowGeom {
  int off9[9];      // element offset of neighbour row k = dt*3+dh relative to `base`
  uint32_t mask9;   // bit k set: neighbour row k is inside the volume
};

// offsets / validity of the 3x3 (t,h) neighbour rows around (t, hc) in a [T][Hn][Wn] volume with token stride ts
__device__ __forceinline__ void make_geom(RowGeom& g, int t, int hc, int T, int Hn, int Wn, int ts) {
  g.mask9 = 0;
#pragma unroll
  for (int dt = 0; dt < 3; ++dt)
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int tt = t + dt - 1, hh = hc + dh - 1;
      const bool ok = tt >= 0 && tt < T && hh >= 0 && hh < Hn;
      g.off9[dt * 3 + dh] = ((dt - 1) * Hn + (dh - 1)) * Wn * ts;
      if (ok) g.mask9 |= 1u << (dt * 3 + dh);
    }
}

// load column `wi` of the window (9 neighbour rows) into x[.]: zero outside the volume
template <typename T>
__device__ __forceinline__ void load_col(float2 (&x)[9], const T* __restrict__ base, const RowGeom& g, int wi, int Wn, int ts) {
  const bool wok = wi >= 0 && wi < Wn;
  const int woff = wi * ts;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const bool ok = wok && ((g.mask9 >> k) & 1u);
    x[k] = ok ? ld2(base + (g.off9[k] + woff)) : make_float2(0.f, 0.f);
  }
}

// conv value from window columns (A, B, C) = (w-1, w, w+1): three independent FFMA2 chains
__device__ __forceinline__ float2 dot27(const float2 (&xa)[9], const float2 (&xb)[9], const float2 (&xc)[9], const float2 (&wr)[TAPS]) {
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    a0 = __ffma2_rn(xa[k], wr[3 * k + 0], a0);
    a1 = __ffma2_rn(xb[k], wr[3 * k + 1], a1);
    a2 = __ffma2_rn(xc[k], wr[3 * k + 2], a2);
  }
  return make_float2(a0.x + a1.x + a2.x, a0.y + a1.y + a2.y);
}

// One march step at output position wo.
//   stride 1: the window slides — slots (SA, SB, SC) hold columns (wo-1, wo, wo+1); only SC is loaded.
//   stride s: all three columns wo*s-1 .. wo*s+1 are loaded (slots 0, 1, 2).
template <typename T, bool SLIDE, int SA, int SB, int SC>
__device__ __forceinline__ void window_step(float2 (&x)[3][9], const T* __restrict__ base, const RowGeom& g, int wo, int s, int Wn,
                                            int ts) {
  if (SLIDE) {
    load_col(x[SC], base, g, wo + 1, Wn, ts);
  } else {
    load_col(x[SA], base, g, wo * s - 1, Wn, ts);
    load_col(x[SB], base, g, wo * s, Wn, ts);
    load_col(x[SC], base, g, wo * s + 1, Wn, ts);
  }
}

__device__ __forceinline__ int find_job(const Launch& L) {
  int j = 0;
  while (j + 1 < L.njobs && (int)blockIdx.x >= L.job[j + 1].blk_begin) ++j;
  return j;
}

// ---------------------------------------------------------------------------------------------------------------
// forward (BWD = false) and backward (i) (BWD = true)
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool BWD>
__global__ void __launch_bounds__(THREADS, 2) pool_ln_march_kernel(const __grid_constant__ Launch L) {
  pdl_wait();
  __shared__ __align__(16) float conv_s[2][CHUNK_TOK * HD];  // 18 KB; reused for the final partial reduction
  __shared__ float sgam[HD], sbet[HD];
  __shared__ int64_t row_out_s[ROWS];   // output token index of the row's first position (incl. cls slots)
  __shared__ int64_t row_dc_s[ROWS];    // dconv token index of the row's first position
  const Job& J = L.job[find_job(L)];
  const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
  const int tid = threadIdx.x;
  const int Lo = L.T * J.Ho * J.Wo;
  const int lb = blockIdx.x - J.blk_begin;

  if (tid < HD) {
    sgam[tid] = J.gamma[tid];
    sbet[tid] = BWD ? 0.f : J.beta[tid];
  }

  float adg[CPL], adb[CPL];  // backward: dgamma / dbeta of this lane's 12 channels (march blocks), 3 channels (cls blocks)
#pragma unroll
  for (int j = 0; j < CPL; ++j) { adg[j] = 0.f; adb[j] = 0.f; }

  if (lb >= J.nblk) {
    // ---------------------------------------------------------------- cls tokens: LayerNorm only, one warp per token
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int ncls = L.B * L.heads;
    for (int bh = (lb - J.nblk) * CLS_WARPS + warp; bh < ncls; bh += J.ncls_blk * CLS_WARPS) {
      const int head = bh % L.heads, b = bh / L.heads;
      const int64_t in_off = (int64_t)b * L.in_bs + (int64_t)head * L.in_hs + lane;
      const int64_t tok = (int64_t)bh * (Lo + 1);
      float v[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) v[j] = to_f32(in[in_off + 32 * j]);
      const float mu = warp_sum(v[0] + v[1] + v[2]) * (1.0f / HD);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) { const float d = v[j] - mu; q += d * d; }
      const float rs = rsqrtf(warp_sum(q) * (1.0f / HD) + L.eps);
      if (!BWD) {
        T* o = reinterpret_cast<T*>(J.out) + tok * J.out_ld + lane;
#pragma unroll
        for (int j = 0; j < 3; ++j) o[32 * j] = from_f32<T>((v[j] - mu) * rs * sgam[lane + 32 * j] + sbet[lane + 32 * j]);
        if (J.xhat) {
          T* xo = reinterpret_cast<T*>(J.xhat) + tok * HD + lane;
#pragma unroll
          for (int j = 0; j < 3; ++j) xo[32 * j] = from_f32<T>((v[j] - mu) * rs);
          if (lane == 0) J.rstd[tok] = rs;
        }
      } else {
        const T* dyr = reinterpret_cast<const T*>(J.dout) + tok * J.dout_ld + lane;
        float xh[3], gg[3], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float dy = to_f32(dyr[32 * j]);
          xh[j] = (v[j] - mu) * rs;
          gg[j] = dy * sgam[lane + 32 * j];
          s1 += gg[j];
          s2 += gg[j] * xh[j];
          adg[j] += dy * xh[j];
          adb[j] += dy;
        }
        s1 = warp_sum(s1) * (1.0f / HD);
        s2 = warp_sum(s2) * (1.0f / HD);
        T* dp = reinterpret_cast<T*>(J.din) + in_off;
#pragma unroll
        for (int j = 0; j < 3; ++j) dp[32 * j] = from_f32<T>(rs * (gg[j] - s1 - xh[j] * s2));
      }
    }
    if (BWD) {
      float* red = &conv_s[0][0];  // [CLS_WARPS][2*HD]
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        red[warp * 2 * HD + lane + 32 * j] = adg[j];
        red[warp * 2 * HD + HD + lane + 32 * j] = adb[j];
      }
      __syncthreads();
      if (tid < 2 * HD) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < CLS_WARPS; ++wv) s += red[wv * 2 * HD + tid];
        J.part_ln[(int64_t)lb * 2 * HD + tid] = s;
      }
    }
    return;
  }

  // ------------------------------------------------------------------ march blocks
  const int cp = tid % NCP, r = tid / NCP;
  float2 wr[TAPS];
  load_taps<false>(J.w, cp, wr);
  const int ts = (int)L.in_ts;
  const bool slide = J.s == 1;
  // LayerNorm-phase role of this thread
  const int g = tid >> 3, sub = tid & 7;
  const int g_row = g / CW, g_w = g - g_row * CW;
  const int nrows = L.B * L.heads * L.T * J.Ho;
  const int nitems = (nrows + ROWS - 1) / ROWS;
  int buf = 0;

  for (int item = lb; item < nitems; item += J.nblk) {
    const int row = item * ROWS + r;
    const bool rvalid = row < nrows;
    RowGeom geo;
    const T* base = in;
    {
      const int rr = rvalid ? row : 0;
      const int ho = rr % J.Ho;
      int q = rr / J.Ho;
      const int t = q % L.T; q /= L.T;
      const int head = q % L.heads, b = q / L.heads;
      make_geom(geo, t, ho * J.s, L.T, L.H, L.W, ts);
      if (!rvalid) geo.mask9 = 0;
      base = in + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs + (int64_t)(1 + (t * L.H + ho * J.s) * L.W) * L.in_ts + 2 * cp);
      if (cp == 0) {
        const int64_t bh = (int64_t)b * L.heads + head;
        const int64_t pos = (int64_t)(t * J.Ho + ho) * J.Wo;
        row_out_s[r] = rvalid ? bh * (Lo + 1) + 1 + pos : -1;
        row_dc_s[r] = bh * Lo + pos;
      }
    }
    float2 x[3][9];
    if (slide) {
#pragma unroll
      for (int k = 0; k < 9; ++k) x[0][k] = make_float2(0.f, 0.f);
      load_col(x[1], base, geo, 0, L.W, ts);
    }
    for (int c0 = 0; c0 < J.Wo; c0 += CW) {
      float* cb = &conv_s[buf][(r * CW) * HD + 2 * cp];
      if (slide) {
#pragma unroll
        for (int j = 0; j < CW; j += 3) {
          if (c0 + j < J.Wo) {
            window_step<T, true, 0, 1, 2>(x, base, geo, c0 + j, 1, L.W, ts);
            *reinterpret_cast<float2*>(cb + j * HD) = dot27(x[0], x[1], x[2], wr);
          }
          if (c0 + j + 1 < J.Wo) {
            window_step<T, true, 1, 2, 0>(x, base, geo, c0 + j + 1, 1, L.W, ts);
            *reinterpret_cast<float2*>(cb + (j + 1) * HD) = dot27(x[1], x[2], x[0], wr);
          }
          if (c0 + j + 2 < J.Wo) {
            window_step<T, true, 2, 0, 1>(x, base, geo, c0 + j + 2, 1, L.W, ts);
            *reinterpret_cast<float2*>(cb + (j + 2) * HD) = dot27(x[2], x[0], x[1], wr);
          }
        }
      } else {
#pragma unroll 2
        for (int j = 0; j < CW; ++j) {
          if (c0 + j < J.Wo) {
            window_step<T, false, 0, 1, 2>(x, base, geo, c0 + j, J.s, L.W, ts);
            *reinterpret_cast<float2*>(cb + j * HD) = dot27(x[0], x[1], x[2], wr);
          }
        }
      }
      __syncthreads();  // conv_s[buf] complete (and row_*_s of this item visible)
      // ---- LayerNorm phase: 8 lanes per token
      {
        const int64_t otok = row_out_s[g_row];
        const bool tvalid = otok >= 0 && c0 + g_w < J.Wo;
        float v[CPL];
        const float* src = &conv_s[buf][g * HD + sub * CPL];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * i);
          v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
        }
        if (!tvalid) {  // stale buffer contents must not reach the dgamma / dbeta accumulators
#pragma unroll
          for (int j = 0; j < CPL; ++j) v[j] = 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) s += v[j];
        const float mu = group_sum<LNL>(s) * (1.0f / HD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) { const float d = v[j] - mu; q += d * d; }
        const float rs = rsqrtf(group_sum<LNL>(q) * (1.0f / HD) + L.eps);
        if (!BWD) {
          float o[CPL];
#pragma unroll
          for (int j = 0; j < CPL; ++j) { v[j] = (v[j] - mu) * rs; o[j] = v[j] * sgam[sub * CPL + j] + sbet[sub * CPL + j]; }
          if (tvalid) {
            store12(reinterpret_cast<T*>(J.out) + (otok + c0 + g_w) * J.out_ld + sub * CPL, o);
            if (J.xhat) {
              store12(reinterpret_cast<T*>(J.xhat) + (otok + c0 + g_w) * HD + sub * CPL, v);
              if (sub == 0) J.rstd[otok + c0 + g_w] = rs;
            }
          }
        } else {
          float dy[CPL];
          if (tvalid) {
            load12(reinterpret_cast<const T*>(J.dout) + (otok + c0 + g_w) * J.dout_ld + sub * CPL, dy);
          } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) dy[j] = 0.f;
          }
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            v[j] = (v[j] - mu) * rs;                      // xhat
            const float gg = dy[j] * sgam[sub * CPL + j];
            s1 += gg;
            s2 += gg * v[j];
            adg[j] += dy[j] * v[j];
            adb[j] += dy[j];
            dy[j] = gg;
          }
#pragma unroll
          for (int o = LNL / 2; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          s1 *= (1.0f / HD);
          s2 *= (1.0f / HD);
          float dc[CPL];
#pragma unroll
          for (int j = 0; j < CPL; ++j) dc[j] = rs * (dy[j] - s1 - v[j] * s2);
          if (tvalid) store12(reinterpret_cast<T*>(J.dconv) + (row_dc_s[g_row] + c0 + g_w) * HD + sub * CPL, dc);
        }
      }
      buf ^= 1;
    }
    __syncthreads();  // row_*_s are rewritten by the next item
  }

  if (BWD) {
    __syncthreads();
    float* red = &conv_s[0][0];  // [CHUNK_TOK groups][2*HD] = 4608 floats = the whole buffer
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      red[g * 2 * HD + sub * CPL + j] = adg[j];
      red[g * 2 * HD + HD + sub * CPL + j] = adb[j];
    }
    __syncthreads();
    if (tid < 2 * HD) {
      float s = 0.f;
#pragma unroll 8
      for (int gg = 0; gg < CHUNK_TOK; ++gg) s += red[gg * 2 * HD + tid];
      J.part_ln[(int64_t)lb * 2 * HD + tid] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward (ii): dW partials
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dw_accum(float2 (&acc)[TAPS], const float2 (&xa)[9], const float2 (&xb)[9], const float2 (&xc)[9],
                                         float2 d) {
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    acc[3 * k + 0] = __ffma2_rn(xa[k], d, acc[3 * k + 0]);
    acc[3 * k + 1] = __ffma2_rn(xb[k], d, acc[3 * k + 1]);
    acc[3 * k + 2] = __ffma2_rn(xc[k], d, acc[3 * k + 2]);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS, 2) pool_ln_bwd_dw_kernel(const __grid_constant__ Launch L) {
  pdl_wait();
  __shared__ float dws[NDW];
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
  const T* __restrict__ dconv = reinterpret_cast<const T*>(J.dconv);
  const int tid = threadIdx.x;
  const int cp = tid % NCP, r = tid / NCP;
  const int Lo = L.T * J.Ho * J.Wo;
  const int ts = (int)L.in_ts;
  const bool slide = J.s == 1;
  const int nrows = L.B * L.heads * L.T * J.Ho;
  const int nitems = (nrows + ROWS - 1) / ROWS;
  const int lb = blockIdx.x - J.blk_begin;
  float2 acc[TAPS];
#pragma unroll
  for (int k = 0; k < TAPS; ++k) acc[k] = make_float2(0.f, 0.f);

  for (int item = lb; item < nitems; item += J.nblk) {
    const int row = item * ROWS + r;
    if (row >= nrows) continue;
    const int ho = row % J.Ho;
    int q = row / J.Ho;
    const int t = q % L.T; q /= L.T;
    const int head = q % L.heads, b = q / L.heads;
    RowGeom geo;
    make_geom(geo, t, ho * J.s, L.T, L.H, L.W, ts);
    const T* base = in + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs + (int64_t)(1 + (t * L.H + ho * J.s) * L.W) * L.in_ts + 2 * cp);
    const T* dc = dconv + (((int64_t)b * L.heads + head) * Lo + (int64_t)(t * J.Ho + ho) * J.Wo) * HD + 2 * cp;
    float2 x[3][9];
    if (slide) {
#pragma unroll
      for (int k = 0; k < 9; ++k) x[0][k] = make_float2(0.f, 0.f);
      load_col(x[1], base, geo, 0, L.W, ts);
      for (int wo = 0; wo < J.Wo; wo += 3) {
        {
          const float2 d = ld2(dc + wo * HD);
          window_step<T, true, 0, 1, 2>(x, base, geo, wo, 1, L.W, ts);
          dw_accum(acc, x[0], x[1], x[2], d);
        }
        if (wo + 1 < J.Wo) {
          const float2 d = ld2(dc + (wo + 1) * HD);
          window_step<T, true, 1, 2, 0>(x, base, geo, wo + 1, 1, L.W, ts);
          dw_accum(acc, x[1], x[2], x[0], d);
        }
        if (wo + 2 < J.Wo) {
          const float2 d = ld2(dc + (wo + 2) * HD);
          window_step<T, true, 2, 0, 1>(x, base, geo, wo + 2, 1, L.W, ts);
          dw_accum(acc, x[2], x[0], x[1], d);
        }
      }
    } else {
#pragma unroll 2
      for (int wo = 0; wo < J.Wo; ++wo) {
        const float2 d = ld2(dc + wo * HD);
        window_step<T, false, 0, 1, 2>(x, base, geo, wo, J.s, L.W, ts);
        dw_accum(acc, x[0], x[1], x[2], d);
      }
    }
  }
  // fold the 4 row groups of the CTA in shared memory ([c][tap], the reference layout), then one coalesced store
  for (int i = tid; i < NDW; i += THREADS) dws[i] = 0.f;
  __syncthreads();
#pragma unroll 1
  for (int rr = 0; rr < ROWS; ++rr) {
    if (r == rr) {
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        dws[(2 * cp) * TAPS + k] += acc[k].x;
        dws[(2 * cp + 1) * TAPS + k] += acc[k].y;
      }
    }
    __syncthreads();
  }
  float* pb = J.part_dw + (int64_t)lb * NDW;
  for (int i = tid; i < NDW; i += THREADS) pb[i] = dws[i];
}

// ---------------------------------------------------------------------------------------------------------------
// backward (iii): input gradient.  din[ti,hi,wi][c] = sum_taps w[c][dt,dh,dw] * dconv[ti+1-dt, (hi+1-dh)/s, (wi+1-dw)/s][c]
// over the taps whose (hi+1-dh), (wi+1-dw) are multiples of s inside the output grid.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int DH, int DW>
__device__ __forceinline__ void tap_t(float2& acc, const T* __restrict__ p, const int (&toff)[3], uint32_t tmask, const float2 (&wr)[TAPS]) {
#pragma unroll
  for (int dt = 0; dt < 3; ++dt) {
    if ((tmask >> dt) & 1u) acc = __ffma2_rn(ld2(p + toff[dt]), wr[dt * 9 + DH * 3 + DW], acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS, 2) pool_ln_bwd_input_kernel(const __grid_constant__ Launch L) {
  pdl_wait();
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const T* __restrict__ dconv = reinterpret_cast<const T*>(J.dconv);
  T* __restrict__ din = reinterpret_cast<T*>(J.din);
  const int tid = threadIdx.x;
  const int cp = tid % NCP, r = tid / NCP;
  const int Lo = L.T * J.Ho * J.Wo;
  const int S = J.s;
  const int nrows = L.B * L.heads * L.T * L.H;  // input rows
  const int nitems = (nrows + ROWS - 1) / ROWS;
  const int lb = blockIdx.x - J.blk_begin;
  float2 wr[TAPS];
  if (S == 1) load_taps<true>(J.w, cp, wr); else load_taps<false>(J.w, cp, wr);

  for (int item = lb; item < nitems; item += J.nblk) {
    const int row = item * ROWS + r;
    if (row >= nrows) continue;
    const int hi = row % L.H;
    int q = row / L.H;
    const int ti = q % L.T; q /= L.T;
    const int head = q % L.heads, b = q / L.heads;
    T* drow = din + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs + (int64_t)(1 + (ti * L.H + hi) * L.W) * L.in_ts + 2 * cp);
    const T* dcb = dconv + ((int64_t)b * L.heads + head) * Lo * HD + 2 * cp;
    if (S == 1) {
      // forward march over dconv (Ho == H, Wo == W) with mirrored taps
      RowGeom geo;
      make_geom(geo, ti, hi, L.T, L.H, L.W, HD);
      const T* base = dcb + (int64_t)((ti * L.H + hi) * L.W) * HD;
      float2 x[3][9];
#pragma unroll
      for (int k = 0; k < 9; ++k) x[0][k] = make_float2(0.f, 0.f);
      load_col(x[1], base, geo, 0, L.W, HD);
      for (int wi = 0; wi < L.W; wi += 3) {
        window_step<T, true, 0, 1, 2>(x, base, geo, wi, 1, L.W, HD);
        st2(drow + (int64_t)wi * L.in_ts, dot27(x[0], x[1], x[2], wr));
        if (wi + 1 < L.W) {
          window_step<T, true, 1, 2, 0>(x, base, geo, wi + 1, 1, L.W, HD);
          st2(drow + (int64_t)(wi + 1) * L.in_ts, dot27(x[1], x[2], x[0], wr));
        }
        if (wi + 2 < L.W) {
          window_step<T, true, 2, 0, 1>(x, base, geo, wi + 2, 1, L.W, HD);
          st2(drow + (int64_t)(wi + 2) * L.in_ts, dot27(x[2], x[0], x[1], wr));
        }
      }
      continue;
    }
    // stride >= 2: which (dh -> ho) and (dt -> to) contribute to this input row
    int toff[3];
    uint32_t tmask = 0;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int to = ti + 1 - dt;
      toff[dt] = to * J.Ho * J.Wo * HD;
      if (to >= 0 && to < L.T) tmask |= 1u << dt;
    }
    int hoff[3];
    uint32_t hmask = 0;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int nh = hi + 1 - dh;
      const int ho = nh / S;
      hoff[dh] = ho * J.Wo * HD;
      if (nh >= 0 && nh - ho * S == 0 && ho < J.Ho) hmask |= 1u << dh;
    }
    if (hmask == 0) {
      for (int wi = 0; wi < L.W; ++wi) st2(drow + (int64_t)wi * L.in_ts, make_float2(0.f, 0.f));
      continue;
    }
    // incremental (wi+1-dw) mod S / div S for dw = 0, 1, 2   (wi = 0: nw = 1, 0, -1)
    int wm[3], wd[3];
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int nw = 1 - dw;
      wm[dw] = nw < 0 ? S - 1 : nw % S;
      wd[dw] = nw < 0 ? -1 : nw / S;
    }
    for (int wi = 0; wi < L.W; ++wi) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        if (wm[dw] == 0 && wd[dw] >= 0 && wd[dw] < J.Wo) {
          const T* p = dcb + wd[dw] * HD;
          if (hmask & 1u) { if (dw == 0) tap_t<T, 0, 0>(acc, p + hoff[0], toff, tmask, wr); else if (dw == 1) tap_t<T, 0, 1>(acc, p + hoff[0], toff, tmask, wr); else tap_t<T, 0, 2>(acc, p + hoff[0], toff, tmask, wr); }
          if (hmask & 2u) { if (dw == 0) tap_t<T, 1, 0>(acc, p + hoff[1], toff, tmask, wr); else if (dw == 1) tap_t<T, 1, 1>(acc, p + hoff[1], toff, tmask, wr); else tap_t<T, 1, 2>(acc, p + hoff[1], toff, tmask, wr); }
          if (hmask & 4u) { if (dw == 0) tap_t<T, 2, 0>(acc, p + hoff[2], toff, tmask, wr); else if (dw == 1) tap_t<T, 2, 1>(acc, p + hoff[2], toff, tmask, wr); else tap_t<T, 2, 2>(acc, p + hoff[2], toff, tmask, wr); }
        }
        if (++wm[dw] == S) { wm[dw] = 0; ++wd[dw]; }
      }
      st2(drow + (int64_t)wi * L.in_ts, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward (i) from saved statistics: the forward kept xhat (normalised, pre-affine) and 1/sigma, so the LayerNorm
// backward is a plain token-wise pass (no convolution recompute): 8 lanes per token, 12 channels per lane.
//   g = dy * gamma;  dconv = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += dy * xhat;  dbeta += dy
// cls tokens write their gradient straight into din (no convolution in front of them).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SV_THREADS = 256;
constexpr int SV_TOK = SV_THREADS / LNL;  // 32 tokens per pass
template <typename T>
__global__ void __launch_bounds__(SV_THREADS) pool_ln_bwd_saved_kernel(const __grid_constant__ Launch L) {
  pdl_wait();
  __shared__ float red[SV_TOK * 2 * HD];
  __shared__ float sgam[HD];
  const Job& J = L.job[find_job(L)];
  const int tid = threadIdx.x;
  const int lb = blockIdx.x - J.blk_begin;
  if (tid < HD) sgam[tid] = J.gamma[tid];
  __syncthreads();
  const int g = tid >> 3, sub = tid & 7;
  // job-table fields of the token loop pinned in registers (indexed constant loads otherwise, see pin())
  const int heads = pin(L.heads);
  const int Lo = pin(L.T * J.Ho * J.Wo);
  const int64_t ntok = (int64_t)L.B * heads * (Lo + 1);
  const T* __restrict__ xhat = pin(reinterpret_cast<const T*>(J.xhat));
  const T* __restrict__ dout = pin(reinterpret_cast<const T*>(J.dout));
  const float* __restrict__ dout32 = reinterpret_cast<const float*>(dout);
  const bool f32 = J.dout_f32 != 0;
  const int64_t dout_ld = pin(J.dout_ld), in_bs = pin(L.in_bs), in_hs = pin(L.in_hs);
  const float* const rstdp = pin(J.rstd);
  T* const dinp = pin(reinterpret_cast<T*>(J.din));
  T* const dconvp = pin(reinterpret_cast<T*>(J.dconv));
  float gm[CPL], adg[CPL], adb[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { gm[j] = sgam[sub * CPL + j]; adg[j] = 0.f; adb[j] = 0.f; }
  // two tokens per lane group in flight: the loads of both are issued before either is reduced (one token per iteration
  // left the kernel at a third of the HBM rate: a DRAM round trip per token per group)
  struct Tok { float xh[CPL], dy[CPL]; float rs; int64_t tok; bool ok; };
  auto fetch = [&](int64_t t0, Tok& t) {
    t.tok = t0 + g;
    t.ok = t.tok < ntok;
    const int64_t tk = t.ok ? t.tok : 0;
    load12(xhat + tk * HD + sub * CPL, t.xh);
    if (f32) load12(dout32 + tk * dout_ld + sub * CPL, t.dy);
    else load12(dout + tk * dout_ld + sub * CPL, t.dy);
    t.rs = t.ok ? rstdp[tk] : 0.f;
  };
  auto consume = [&](Tok& t) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      if (!t.ok) t.dy[j] = 0.f;
      const float gg = t.dy[j] * gm[j];
      s1 += gg;
      s2 += gg * t.xh[j];
      adg[j] += t.dy[j] * t.xh[j];
      adb[j] += t.dy[j];
      t.dy[j] = gg;
    }
#pragma unroll
    for (int o = LNL / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / HD);
    s2 *= (1.0f / HD);
    if (!t.ok) return;
    float dc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) dc[j] = t.rs * (t.dy[j] - s1 - t.xh[j] * s2);
    const int64_t bh = t.tok / (Lo + 1);
    const int n = (int)(t.tok - bh * (Lo + 1));
    if (n == 0) {
      const int head = (int)(bh % heads);
      const int64_t b = bh / heads;
      store12(dinp + b * in_bs + head * in_hs + sub * CPL, dc);
    } else {
      store12(dconvp + (bh * Lo + (n - 1)) * HD + sub * CPL, dc);
    }
  };
  const int64_t tstep = (int64_t)J.nblk * SV_TOK;
  for (int64_t t0 = (int64_t)lb * SV_TOK; t0 < ntok; t0 += 2 * tstep) {
    Tok A, B;
    fetch(t0, A);
    const bool has_b = t0 + tstep < ntok;
    if (has_b) fetch(t0 + tstep, B);
    consume(A);
    if (has_b) consume(B);
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    red[g * 2 * HD + sub * CPL + j] = adg[j];
    red[g * 2 * HD + HD + sub * CPL + j] = adb[j];
  }
  __syncthreads();
  if (tid < 2 * HD) {
    float s = 0.f;
#pragma unroll 8
    for (int gg = 0; gg < SV_TOK; ++gg) s += red[gg * 2 * HD + tid];
    J.part_ln[(int64_t)lb * 2 * HD + tid] = s;
  }
}

// grads_j[i] = sum over the job's partial vectors: dW from backward (ii), dgamma / dbeta from backward (i).
// Block = 32 gradient entries x 8 slices of the partial vectors, combined through shared memory: grads is
// OVERWRITTEN (no zero fill by the caller, no atomics).  grid.y = job.
constexpr int RED_SLICES = 8;
__global__ void __launch_bounds__(256) reduce_jobs_kernel(const __grid_constant__ Launch L) {
  pdl_wait();
  __shared__ float part[RED_SLICES][32];
  const Job& J = L.job[blockIdx.y];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < NGRAD) {
    const bool is_dw = i < NDW;
    const float* src = is_dw ? J.part_dw + i : J.part_ln + (i - NDW);
    const int64_t stride = is_dw ? NDW : 2 * HD;
    const int end = is_dw ? J.nblk_dw : J.nblk_ln;
    int b = slice;
    // up to 1200 partial vectors: eight independent loads in flight per thread (two per iteration left the kernel
    // waiting on one L2 round trip per pair: 14.6 us for 10 MB)
    for (; b + 7 * RED_SLICES < end; b += 8 * RED_SLICES) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(int64_t)(b + u * RED_SLICES) * stride];
      s0 += v[0] + v[4]; s1 += v[1] + v[5]; s2 += v[2] + v[6]; s3 += v[3] + v[7];
    }
    for (; b < end; b += RED_SLICES) s0 += src[(int64_t)b * stride];
  }
  part[slice][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice == 0 && i < NGRAD) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < RED_SLICES; ++k) s += part[k][lane];
    J.grads[i] = s;
  }
}

int nblocks_for(int64_t items, int max_blocks) {
  int64_t b = items;
  if (b > max_blocks) b = max_blocks;
  return b < 1 ? 1 : (int)b;
}

int out_hw(int n, int s) { return (n - 1) / s + 1; }  // (n + 2*1 - 3) / s + 1

int64_t out_rows(int B, int heads, int T, int H, int s) { return (int64_t)B * heads * T * out_hw(H, s); }
int64_t ntok_conv(int B, int heads, int T, int H, int W, int s) { return (int64_t)B * heads * T * out_hw(H, s) * out_hw(W, s); }

// block budgets.  The partial-vector buffers in the workspace are sized for the caps.
constexpr int MAX_LN_BLOCKS = 148 * 8;
constexpr int MAX_CLS_BLOCKS = 16;
constexpr int MAX_DW_BLOCKS = 148 * 2;
int march_blocks(int B, int heads, int T, int H, int s) { return nblocks_for(ceil_div64(out_rows(B, heads, T, H, s), ROWS), MAX_LN_BLOCKS); }
int cls_blocks(int B, int heads) { return nblocks_for(ceil_div64((int64_t)B * heads, CLS_WARPS), MAX_CLS_BLOCKS); }
int dw_blocks(int B, int heads, int T, int H, int s) { return nblocks_for(ceil_div64(out_rows(B, heads, T, H, s), ROWS), MAX_DW_BLOCKS); }
int input_blocks(int B, int heads, int T, int H) { return nblocks_for(ceil_div64((int64_t)B * heads * T * H, ROWS), 148 * 8); }

int64_t align16(int64_t bytes) { return (bytes + 15) / 16 * 16; }

struct Geom {
  int B, heads, T, H, W;
  int64_t bs, ts, hs;
  float eps;
  int dtype;
};

int fill_jobs(Job* J, const void* qkv, int64_t ws_, const pmv_pool_job* jobs, int njobs, const Geom& g) {
  PMV_CHECK_ARG(njobs >= 1 && njobs <= MAX_JOBS, "pool: 1..3 jobs");
  PMV_CHECK_ARG(g.B > 0 && g.heads > 0 && g.T > 0 && g.H > 0 && g.W > 0, "pool: bad geometry");
  PMV_CHECK_ARG(g.ts % 8 == 0 && g.hs % 8 == 0 && ws_ % 8 == 0 && g.bs % 8 == 0, "pool: strides must be multiples of 8 elements");
  PMV_CHECK_ARG((int64_t)3 * g.H * g.W * g.ts < (1ll << 31) && (int64_t)g.B * g.heads * g.T * g.H < (1ll << 29),
                "pool: volume too large for 32-bit row offsets");
  const int esz = g.dtype == PMV_BF16 ? 2 : 4;
  for (int i = 0; i < njobs; ++i) {
    const pmv_pool_job& p = jobs[i];
    PMV_CHECK_ARG(p.stride_hw >= 1, "pool: bad stride");
    memset(&J[i], 0, sizeof(Job));
    J[i].in = reinterpret_cast<const char*>(qkv) + (int64_t)p.which * ws_ * esz;
    J[i].w = p.w; J[i].gamma = p.gamma; J[i].beta = p.beta; J[i].out = p.out; J[i].out_ld = p.out_ld;
    J[i].dout = p.dout; J[i].dout_ld = p.dout_ld; J[i].dout_f32 = p.dout_f32; J[i].onehot = p.onehot; J[i].grads = p.grads; J[i].xhat = p.xhat; J[i].rstd = p.rstd;
    J[i].s = p.stride_hw; J[i].Ho = out_hw(g.H, p.stride_hw); J[i].Wo = out_hw(g.W, p.stride_hw);
  }
  return PMV_OK;
}

// Runs one mode over the jobs: those the TMA t-march can take go to pool_tma.cu, the rest to the direct kernels here.
// mode 0 forward, 1 backward (i), 2 backward (ii), 3 backward (iii).
int run_mode(int mode, Job* all, int njobs, const Geom& g, cudaStream_t st, cudaStream_t st_b) {
  // st_b: stream of the second job class of the launch (tap tiles; mode 3: the strided input-gradient kernels)
  Job tj[MAX_JOBS], dj[MAX_JOBS];
  int ti[MAX_JOBS], di[MAX_JOBS];
  int nt = 0, nd = 0;
  const bool tma_ok = pmv_has_tcgen05() && std::getenv("PMV_POOL_DIRECT") == nullptr;
  for (int i = 0; i < njobs; ++i) {
    if (tma_ok && tma_eligible(all[i].s, mode, g.dtype == PMV_BF16 ? 2 : 4)) { ti[nt] = i; tj[nt++] = all[i]; } else { di[nd] = i; dj[nd++] = all[i]; }
  }
  if (nt > 0) {
    // One launch per job class, so that a class with large shared-memory planes (stride >= 3 tap tiles) does not
    // take the second CTA per SM away from the big stride-1 / stride-2 jobs, and the block budget of a launch is
    // computed from the jobs that are really in it (profiles/r01_pool_ncu.md: 0.3 waves on the stride-1 input gradient).
    //   class 0: dense planes (stride 1, 2; mode 3: stride 1 only)   class 1: everything else of this mode
    for (int cls = 0; cls < 2; ++cls) {
      Job cj[MAX_JOBS];
      int ci[MAX_JOBS];
      int nc = 0;
      for (int i = 0; i < nt; ++i) {
        const bool dense = mode == 3 ? tj[i].s == 1 : tj[i].s <= 2;
        if (dense == (cls == 0)) { ci[nc] = ti[i]; cj[nc++] = tj[i]; }
      }
      if (nc == 0) continue;
      int64_t items_total = 0;
      int items[MAX_JOBS];
      for (int i = 0; i < nc; ++i) {
        items[i] = tma_items(g.B, g.heads, mode == 3 ? g.H : cj[i].Ho, mode == 3 ? g.W : cj[i].Wo);
        items_total += items[i];
      }
      // balanced rounds over the CTA slots of the machine (2 per SM for dense planes, 1 for the tap-tile class)
      const int slots = 148 * (cls == 0 ? 2 : 1);
      const int64_t rounds = ceil_div64(items_total, slots);
      int total = 0;
      for (int i = 0; i < nc; ++i) {
        cj[i].blk_begin = total;
        cj[i].nblk = nblocks_for(ceil_div64(items[i], rounds), mode == 2 ? MAX_DW_BLOCKS : MAX_LN_BLOCKS);
        cj[i].ncls_blk = (mode == 0 || mode == 1) ? cls_blocks(g.B, g.heads) : 0;
        total += cj[i].nblk + cj[i].ncls_blk;
        if (mode == 1) all[ci[i]].nblk_ln = cj[i].nblk + cj[i].ncls_blk;
        if (mode == 2) all[ci[i]].nblk_dw = cj[i].nblk;
      }
      int rc = tma_launch(mode, cj, nc, g.B, g.heads, g.T, g.H, g.W, g.bs, g.ts, g.hs, g.eps, g.dtype,
                          (mode != 3 && cls == 1) ? st_b : st, st_b);
      if (rc) return rc;
    }
  }
  if (nd > 0) {
    Launch L;
    L.njobs = nd; L.B = g.B; L.heads = g.heads; L.T = g.T; L.H = g.H; L.W = g.W;
    L.in_bs = g.bs; L.in_ts = g.ts; L.in_hs = g.hs; L.eps = g.eps;
    int total = 0;
    for (int i = 0; i < nd; ++i) {
      Job& J = dj[i];
      J.blk_begin = total;
      J.nblk = mode == 2 ? dw_blocks(g.B, g.heads, g.T, g.H, J.s)
               : mode == 3 ? input_blocks(g.B, g.heads, g.T, g.H) : march_blocks(g.B, g.heads, g.T, g.H, J.s);
      J.ncls_blk = (mode == 0 || mode == 1) ? cls_blocks(g.B, g.heads) : 0;
      total += J.nblk + J.ncls_blk;
      if (mode == 1) all[di[i]].nblk_ln = J.nblk + J.ncls_blk;
      if (mode == 2) all[di[i]].nblk_dw = J.nblk;
      L.job[i] = J;
    }
    PMV_DISPATCH_DTYPE(g.dtype, TT, {
      if (mode == 0) pmv_launch(pool_ln_march_kernel<TT, false>, (unsigned)total, THREADS, 0, st, L);
      else if (mode == 1) pmv_launch(pool_ln_march_kernel<TT, true>, (unsigned)total, THREADS, 0, st, L);
      else if (mode == 2) pmv_launch(pool_ln_bwd_dw_kernel<TT>, (unsigned)total, THREADS, 0, st, L);
      else pmv_launch(pool_ln_bwd_input_kernel<TT>, (unsigned)total, THREADS, 0, st, L);
    });
    PMV_CHECK_LAUNCH();
  }
  return PMV_OK;
}

}  // namespace

namespace {
struct AuxStreams {
  cudaStream_t s[2];
  cudaEvent_t fork, join[2];
  int state;  // 0 = not created, 1 = ready, -1 = unavailable
};
AuxStreams g_aux[64];
bool streams_enabled() {
  static const bool on = [] { const char* e = std::getenv("PMV_POOL_STREAMS"); return e == nullptr || e[0] != '0'; }();
  return on;
}
AuxStreams* aux_for_device() {
  int dev = 0;
  if (!streams_enabled() || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  AuxStreams& a = g_aux[dev];
  if (a.state == 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    // creating streams / events is legal during capture, but keep the first use outside of it simple: any failure
    // disables the fork for this device
    bool ok = cudaStreamCreateWithFlags(&a.s[0], cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&a.s[1], cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&a.join[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&a.join[1], cudaEventDisableTiming) == cudaSuccess;
    (void)cs;
    a.state = ok ? 1 : -1;
    if (!ok) cudaGetLastError();
  }
  return a.state == 1 ? &a : nullptr;
}
}  // namespace

PoolFork pool_fork(cudaStream_t main) {
  PoolFork f;
  f.main = main; f.side[0] = main; f.side[1] = main; f.on = false;
  AuxStreams* a = aux_for_device();
  if (a == nullptr) return f;
  if (cudaEventRecord(a->fork, main) != cudaSuccess || cudaStreamWaitEvent(a->s[0], a->fork, 0) != cudaSuccess ||
      cudaStreamWaitEvent(a->s[1], a->fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return f;
  }
  f.side[0] = a->s[0]; f.side[1] = a->s[1]; f.on = true;
  return f;
}

void pool_join(const PoolFork& f) {
  if (!f.on) return;
  AuxStreams* a = aux_for_device();
  for (int i = 0; i < 2; ++i) {
    cudaEventRecord(a->join[i], f.side[i]);
    cudaStreamWaitEvent(f.main, a->join[i], 0);
  }
}
}  // namespace pool

using namespace pool;

extern "C" int pmv_pool_ln_qkv_fwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride,
                                   int64_t head_stride, const pmv_pool_job* jobs, int njobs,
                                   int B, int heads, int T, int H, int W, float eps, int dtype, void* stream) {
  PMV_CHECK_ARG(dtype == PMV_BF16 || dtype == PMV_F32, "pool: bad dtype");
  Geom g{B, heads, T, H, W, batch_stride, token_stride, head_stride, eps, dtype};
  Job J[MAX_JOBS];
  int rc = fill_jobs(J, qkv, which_stride, jobs, njobs, g);
  if (rc) return rc;
  for (int i = 0; i < njobs; ++i)
    PMV_CHECK_ARG(jobs[i].out != nullptr && jobs[i].out_ld % 4 == 0 && jobs[i].out_ld >= HD, "pool: bad output");
  // one-hot key-coordinate columns: written by the warp-specialised forward kernel itself; every other path (direct
  // kernels, PMV_POOL_WS=0) gets the separate launch
  const bool tma_ok = pmv_has_tcgen05() && std::getenv("PMV_POOL_DIRECT") == nullptr;
  bool fused[MAX_JOBS];
  for (int i = 0; i < njobs; ++i) {
    fused[i] = J[i].onehot && tma_ok && tma_eligible(J[i].s, 0, dtype == PMV_BF16 ? 2 : 4) && tma_fwd_writes_onehot() &&
               (J[i].out_ld - HD) % 32 == 0;
    if (!fused[i]) J[i].onehot = 0;
  }
  {
    // dense-plane jobs on the caller's stream, tap-tile jobs (stride >= 3: 32-64 CTAs) beside them
    bool two_classes = false, has_dense = false, has_tap = false;
    for (int i = 0; i < njobs; ++i) (J[i].s <= 2 ? has_dense : has_tap) = true;
    two_classes = has_dense && has_tap && tma_ok;
    PoolFork fk = two_classes ? pool_fork((cudaStream_t)stream) : PoolFork{(cudaStream_t)stream, {(cudaStream_t)stream, (cudaStream_t)stream}, false};
    rc = run_mode(0, J, njobs, g, (cudaStream_t)stream, fk.side[0]);
    pool_join(fk);
  }
  if (rc) return rc;
  for (int i = 0; i < njobs; ++i) {
    if (jobs[i].onehot && !fused[i]) {
      rc = pmv_relpos_augment_k(jobs[i].out, jobs[i].out_ld, B * heads, T, J[i].Ho, J[i].Wo, dtype, stream);
      if (rc) return rc;
    }
  }
  return PMV_OK;
}

extern "C" int64_t pmv_pool_ln_qkv_bwd_workspace_bytes(int B, int heads, int T, int H, int W, const int* strides_hw, int njobs) {
  int64_t bytes = 0;
  for (int i = 0; i < njobs; ++i) {
    bytes += align16(ntok_conv(B, heads, T, H, W, strides_hw[i]) * HD * 4);  // pre-LN gradient (sized for fp32)
    bytes += (int64_t)(MAX_LN_BLOCKS + MAX_CLS_BLOCKS) * 2 * HD * 4;
    bytes += (int64_t)MAX_DW_BLOCKS * NDW * 4;
  }
  return bytes;
}

extern "C" int pmv_pool_ln_qkv_bwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride,
                                   int64_t head_stride, const pmv_pool_job* jobs, int njobs, void* dqkv, float* ws,
                                   int B, int heads, int T, int H, int W, float eps, int dtype, void* stream) {
  PMV_CHECK_ARG(dtype == PMV_BF16 || dtype == PMV_F32, "pool: bad dtype");
  Geom g{B, heads, T, H, W, batch_stride, token_stride, head_stride, eps, dtype};
  Job J[MAX_JOBS];
  int rc = fill_jobs(J, qkv, which_stride, jobs, njobs, g);
  if (rc) return rc;
  const int esz = dtype == PMV_BF16 ? 2 : 4;
  char* cursor = reinterpret_cast<char*>(ws);
  for (int i = 0; i < njobs; ++i) {
    PMV_CHECK_ARG(jobs[i].dout != nullptr && jobs[i].grads != nullptr && jobs[i].dout_ld % 4 == 0, "pool: bad backward job");
    PMV_CHECK_ARG(!jobs[i].dout_f32 || (jobs[i].xhat != nullptr && jobs[i].rstd != nullptr),
                  "pool: fp32 dout needs the saved-statistics backward (xhat / rstd)");
    J[i].din = reinterpret_cast<char*>(dqkv) + (int64_t)jobs[i].which * which_stride * esz;
    J[i].dconv = cursor;
    cursor += align16(ntok_conv(B, heads, T, H, W, J[i].s) * HD * 4);
    J[i].part_ln = reinterpret_cast<float*>(cursor);
    cursor += (int64_t)(MAX_LN_BLOCKS + MAX_CLS_BLOCKS) * 2 * HD * 4;
    J[i].part_dw = reinterpret_cast<float*>(cursor);
    cursor += (int64_t)MAX_DW_BLOCKS * NDW * 4;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // backward (i): token-wise from the saved statistics where the forward kept them, otherwise with the recompute
  {
    Job rj[MAX_JOBS];
    int ri[MAX_JOBS];
    int nr = 0;
    Launch LS;
    LS.njobs = 0; LS.B = B; LS.heads = heads; LS.T = T; LS.H = H; LS.W = W;
    LS.in_bs = batch_stride; LS.in_ts = token_stride; LS.in_hs = head_stride; LS.eps = eps;
    int total = 0;
    for (int i = 0; i < njobs; ++i) {
      if (J[i].xhat != nullptr && J[i].rstd != nullptr) {
        const int64_t ntok = (int64_t)B * heads * (1 + (int64_t)T * J[i].Ho * J[i].Wo);
        J[i].blk_begin = total;
        J[i].nblk = nblocks_for(ceil_div64(ntok, SV_TOK * 4), 148 * 4);
        J[i].nblk_ln = J[i].nblk;
        total += J[i].nblk;
        LS.job[LS.njobs++] = J[i];
      } else {
        ri[nr] = i;
        rj[nr++] = J[i];
      }
    }
    if (LS.njobs > 0) {
      PMV_DISPATCH_DTYPE(dtype, TT, (pmv_launch(pool_ln_bwd_saved_kernel<TT>, (unsigned)total, SV_THREADS, 0, st, LS)));
      PMV_CHECK_LAUNCH();
    }
    if (nr > 0) {
      rc = run_mode(1, rj, nr, g, st, st);
      if (rc) return rc;
      for (int k = 0; k < nr; ++k) J[ri[k]].nblk_ln = rj[k].nblk_ln;
    }
  }
  // dW (caller's stream; its tap-tile class on side stream 1), the stride-1 input gradient (side stream 0) and the strided
  // input-gradient kernels (side stream 1) only depend on the pre-LN gradient written above
  {
    PoolFork fk = pool_fork(st);
    rc = run_mode(2, J, njobs, g, st, fk.side[1]);
    if (!rc) rc = run_mode(3, J, njobs, g, fk.side[0], fk.side[1]);
    pool_join(fk);
    if (rc) return rc;
  }
  Launch L;
  L.njobs = njobs; L.B = B; L.heads = heads; L.T = T; L.H = H; L.W = W;
  L.in_bs = batch_stride; L.in_ts = token_stride; L.in_hs = head_stride; L.eps = eps;
  for (int i = 0; i < njobs; ++i) L.job[i] = J[i];
  pmv_launch(reduce_jobs_kernel, dim3((NGRAD + 31) / 32, njobs), 256, 0, st, L);
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
