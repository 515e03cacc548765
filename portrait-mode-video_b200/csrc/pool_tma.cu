#define PMV_PDL_FAMILY 64
// TMA-staged "t-march" version of the pooling stencils for stride 1 and 2 (attention.py:14-48, 241-282).
//
// A CTA owns a spatial tile of 4 output rows x 7 output columns of one (batch, head) and walks the T frames.
// Per frame ONE input plane ((3 + 3 s) x (3 + 6 s) tokens x 96 channels, halo included, zero padding supplied by
// the TMA out-of-bounds fill, negative start coordinates included) is fetched by a single 5-D
// cp.async.bulk.tensor into a 3-deep shared-memory ring, two planes ahead of the arithmetic.  The plane is used in
// "scatter" form along t: it contributes to the output frames t-1, t, t+1 at once (three accumulator sets of 7
// positions per thread), so every input element leaves DRAM/L2 once per tile and is read from shared memory once
// per output row: 27 shared loads feed 189 FFMA2 (packed fp32 pairs) per thread and frame.  Thread = (output row,
// channel pair) as in pool.cu; the frame that completes is parked in shared memory and the CTA re-maps to 28
// tokens x 8 lanes for the LayerNorm (forward) or the LayerNorm backward (backward (i)).
//
// Modes: forward, backward (i) (conv recompute + LN backward -> dconv, dgamma, dbeta), backward (ii) (dW: the same
// march with the pre-LN gradient of the three frames in registers and 27 exclusive accumulator pairs), and
// backward (iii) for stride 1 (input gradient = the forward march over dconv with mirrored taps).
// The direct-load kernels of pool.cu remain for strides >= 3 (mostly-zero gathers) and for backward (iii) at
// stride 2.  Before this file the register-window kernels were latency-bound on L2 (9 dependent 4-byte loads per
// output, 12 warps per SM): 0.2-0.8 TB/s (profiles/r01_kernels_before.txt).
#include "pool_common.cuh"
#include "tc_common.cuh"

namespace pool {
namespace {

constexpr int CW = 7;                 // output columns per tile (56, 28, 14, 7 are multiples)
constexpr int TOK = ROWS * CW;        // 28 tokens per LayerNorm phase
constexpr int THREADS = TOK * LNL;    // 224: conv phase uses the first 192 (4 rows x 48 channel pairs)
constexpr int NST = 3;                // plane ring depth
constexpr int CLS_WARPS = THREADS / 32;
enum { M_FWD = 0, M_BWD_LN = 1, M_BWD_DW = 2, M_BWD_IN = 3 };

struct TLaunch {
  CUtensorMap tm[MAX_JOBS];  // 5-D (channel, w, h, t, batch) views of q / k / v (or of dconv for M_BWD_IN)
  Job job[MAX_JOBS];
  int njobs;
  int B, heads, T, H, W;
  int64_t in_bs, in_ts, in_hs;
  float eps;
};

// S = 1, 2: one dense halo plane per frame.  S = 0 stands for "any stride >= 3": the 3 x 3 windows do not overlap, so
// a plane is fetched as nine tap tiles of [ROWS][CW] tokens, each by one TMA load that walks the tensor with element
// strides (1, s, s, 1, 1) from the tap's own origin — exactly the tokens the stencil touches, zero-filled by
// coordinate outside the frame.
template <int S> struct Geo {
  static constexpr int BH = 3 + (ROWS - 1) * S;
  static constexpr int BW = 3 + (CW - 1) * S;
  static constexpr int PLANE_ELEMS = BH * BW * HD;
};
template <> struct Geo<0> {
  static constexpr int BH = ROWS, BW = CW;
  static constexpr int TILE_ELEMS = ROWS * CW * HD;
  static constexpr int PLANE_ELEMS = 9 * TILE_ELEMS;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 lds2(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

struct Smem {
  uint8_t* planes;   // NST x plane bytes
  float* conv;       // [2][TOK * HD]
  uint64_t* full;    // [NST]
};

// ---------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, int S>
__device__ __forceinline__ void march(const TLaunch& L, const Job& J, const CUtensorMap* tm, const Smem& sm, int lb, int nblk,
                                      const float* sgam, const float* sbet, float* partial_out) {
  using G = Geo<S>;
  constexpr uint32_t PLANE_BYTES = G::PLANE_ELEMS * sizeof(T);
  constexpr uint32_t PLANE_STRIDE = (PLANE_BYTES + 127u) & ~127u;  // TMA destinations are 128-byte aligned
  const int tid = threadIdx.x;
  const bool conv_thread = tid < ROWS * NCP;
  const int cp = tid % NCP, r = conv_thread ? tid / NCP : 0;
  const int Ho = J.Ho, Wo = J.Wo, Tn = L.T;
  const int Lo = Tn * Ho * Wo;
  const int n_th = (Ho + ROWS - 1) / ROWS, n_tw = (Wo + CW - 1) / CW;
  const int per_bh = n_th * n_tw;
  const int nitems = L.B * L.heads * per_bh;
  const int n_my = lb < nitems ? (nitems - lb + nblk - 1) / nblk : 0;
  const int total = n_my * Tn;
  // LayerNorm-phase role
  const int g = tid >> 3, sub = tid & 7;
  const int g_row = g / CW, g_w = g - g_row * CW;

  auto decode = [&](int item, int& b, int& head, int& th, int& tw) {
    const int bh = item / per_bh;
    const int rem = item - bh * per_bh;
    th = rem / n_tw; tw = rem - th * n_tw;
    b = bh / L.heads; head = bh - b * L.heads;
  };
  auto issue = [&](int k) {  // thread 0
    const int item = lb + (k / Tn) * nblk, tin = k % Tn;
    int b, head, th, tw;
    decode(item, b, head, th, tw);
    uint64_t* bar = &sm.full[k % NST];
    tc::mbar_expect_tx(bar, PLANE_BYTES);
    // M_BWD_IN reads dconv [bh][T][H][W][96]: channel coordinate 0, outermost coordinate bh
    const int c0 = MODE == M_BWD_IN ? 0 : head * HD;
    const int c4 = MODE == M_BWD_IN ? b * L.heads + head : b;
    if constexpr (S == 0) {
      constexpr uint32_t TILE_BYTES = Geo<0>::TILE_ELEMS * sizeof(T);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        tma_load_5d(sm.planes + (size_t)(k % NST) * PLANE_STRIDE + tap * TILE_BYTES, tm, c0, tw * CW * J.s + tap % 3 - 1,
                    th * ROWS * J.s + tap / 3 - 1, tin, c4, bar);
    } else {
      tma_load_5d(sm.planes + (size_t)(k % NST) * PLANE_STRIDE, tm, c0, tw * CW * S - 1, th * ROWS * S - 1, tin, c4, bar);
    }
  };

  if (tid == 0) {
    for (int k = 0; k < NST && k < total; ++k) issue(k);
  }

  float2 wr[TAPS];
  if (MODE != M_BWD_DW) {
    if (MODE == M_BWD_IN) load_taps<true>(J.w, cp, wr); else load_taps<false>(J.w, cp, wr);
  }
  float2 accw[MODE == M_BWD_DW ? TAPS : 1];
  if (MODE == M_BWD_DW) {
#pragma unroll
    for (int k = 0; k < TAPS; ++k) accw[k] = make_float2(0.f, 0.f);
  }
  float adg[MODE == M_BWD_LN ? CPL : 1], adb[MODE == M_BWD_LN ? CPL : 1];
  if (MODE == M_BWD_LN) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) { adg[j] = 0.f; adb[j] = 0.f; }
  }
  int buf = 0;

  for (int i = 0; i < n_my; ++i) {
    const int item = lb + i * nblk;
    int b, head, th, tw;
    decode(item, b, head, th, tw);
    const int64_t bh = (int64_t)b * L.heads + head;
    const int row = th * ROWS + r;             // output row of this conv thread
    const bool row_ok = conv_thread && row < Ho;
    const int col0 = tw * CW;
    // acc[slot][j]: output frame (slot == tout % 3), column col0 + j.  M_BWD_DW: the pre-LN gradient instead.
    float2 acc[3][CW];
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
      for (int j = 0; j < CW; ++j) acc[s3][j] = make_float2(0.f, 0.f);

    auto load_dc = [&](float2 (&dst)[CW], int tout) {  // M_BWD_DW
      const T* dc = reinterpret_cast<const T*>(J.dconv) + ((bh * Lo + (int64_t)(tout * Ho + row) * Wo + col0) * HD + 2 * cp);
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const bool ok = row_ok && tout >= 0 && tout < Tn && col0 + j < Wo;
        dst[j] = ok ? ld2(dc + j * HD) : make_float2(0.f, 0.f);
      }
    };
    // M_BWD_DW: step tin needs the pre-LN gradient of frames tin-1, tin, tin+1.  Frame tin+2 is requested at the END of
    // step tin, into the register slot frame tin-1 just vacated, so the loads fly during the barrier and the wait for the
    // next plane (they used to be consumed right after issue: 37 % of the samples, profiles/r01_pool_ncu.md)
    if constexpr (MODE == M_BWD_DW) {
      load_dc(acc[0], 0);
      load_dc(acc[1], 1);
    }

    auto step = [&](auto ptag, int tin) {
      constexpr int P = decltype(ptag)::value;
      constexpr int SLOT_M1 = (P + 2) % 3, SLOT_0 = P, SLOT_P1 = (P + 1) % 3;  // output frames tin-1, tin, tin+1
      if (tin > Tn) return;
      const int k = i * Tn + tin;
      if (tin < Tn) {
        tc::mbar_wait(&sm.full[k % NST], (uint32_t)((k / NST) & 1));
        if (conv_thread) {
          const T* plane = reinterpret_cast<const T*>(sm.planes + (size_t)(k % NST) * PLANE_STRIDE) + 2 * cp;
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            constexpr int NX = S == 0 ? 3 * CW : G::BW;
            float2 x[NX];
            if constexpr (S == 0) {  // tap tiles (dh, dw): [ROWS][CW] tokens each
#pragma unroll
              for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                for (int j = 0; j < CW; ++j) x[dw * CW + j] = lds2(plane + (((dh * 3 + dw) * ROWS + r) * CW + j) * HD);
            } else {
              const T* prow = plane + (r * S + dh) * (G::BW * HD);
#pragma unroll
              for (int c = 0; c < G::BW; ++c) x[c] = lds2(prow + c * HD);
            }
            // dw outside j: consecutive FFMA2 of one accumulator are 21 instructions apart (7 positions x 3 frames)
#pragma unroll
            for (int dw = 0; dw < 3; ++dw)
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                const float2 xv = S == 0 ? x[dw * CW + j] : x[j * S + dw];
                if (MODE == M_BWD_DW) {
                  accw[0 * 9 + dh * 3 + dw] = __ffma2_rn(xv, acc[SLOT_P1][j], accw[0 * 9 + dh * 3 + dw]);
                  accw[1 * 9 + dh * 3 + dw] = __ffma2_rn(xv, acc[SLOT_0][j], accw[1 * 9 + dh * 3 + dw]);
                  accw[2 * 9 + dh * 3 + dw] = __ffma2_rn(xv, acc[SLOT_M1][j], accw[2 * 9 + dh * 3 + dw]);
                } else {
                  acc[SLOT_P1][j] = __ffma2_rn(xv, wr[0 * 9 + dh * 3 + dw], acc[SLOT_P1][j]);
                  acc[SLOT_0][j] = __ffma2_rn(xv, wr[1 * 9 + dh * 3 + dw], acc[SLOT_0][j]);
                  acc[SLOT_M1][j] = __ffma2_rn(xv, wr[2 * 9 + dh * 3 + dw], acc[SLOT_M1][j]);
                }
              }
          }
        }
      }
      const int tout = tin - 1;
      if constexpr (MODE == M_BWD_DW) {
        if (tin < Tn) load_dc(acc[SLOT_M1], tin + 2);
        __syncthreads();  // everyone is done with plane k: its ring slot may be refilled
        if (tid == 0 && tin < Tn && k + NST < total) issue(k + NST);
        return;
      }
      if (MODE == M_BWD_IN) {
        if (tout >= 0 && row_ok) {
          T* drow = reinterpret_cast<T*>(J.din) + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs +
                                                   (int64_t)(1 + (tout * Ho + row) * Wo + col0) * L.in_ts + 2 * cp);
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (col0 + j < Wo) st2(drow + (int64_t)j * L.in_ts, acc[SLOT_M1][j]);
        }
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[SLOT_M1][j] = make_float2(0.f, 0.f);
        __syncthreads();
        if (tid == 0 && tin < Tn && k + NST < total) issue(k + NST);
        return;
      }
      // forward / backward (i): park the completed frame, LayerNorm phase
      float* cbuf = sm.conv + buf * (TOK * HD);
      if (tout >= 0 && conv_thread) {
#pragma unroll
        for (int j = 0; j < CW; ++j) *reinterpret_cast<float2*>(cbuf + (r * CW + j) * HD + 2 * cp) = acc[SLOT_M1][j];
      }
#pragma unroll
      for (int j = 0; j < CW; ++j) acc[SLOT_M1][j] = make_float2(0.f, 0.f);
      __syncthreads();
      if (tid == 0 && tin < Tn && k + NST < total) issue(k + NST);
      if (tout < 0) return;
      {
        const int orow = th * ROWS + g_row, ocol = col0 + g_w;
        const bool tvalid = orow < Ho && ocol < Wo;
        const int64_t pos = (int64_t)(tout * Ho + orow) * Wo + ocol;
        float v[CPL];
        const float* src = cbuf + g * HD + sub * CPL;
#pragma unroll
        for (int q4 = 0; q4 < 3; ++q4) {
          const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * q4);
          v[4 * q4] = t4.x; v[4 * q4 + 1] = t4.y; v[4 * q4 + 2] = t4.z; v[4 * q4 + 3] = t4.w;
        }
        if (!tvalid) {  // rows / columns outside the grid hold junk-free zeros anyway, but keep the statistics finite
#pragma unroll
          for (int j = 0; j < CPL; ++j) v[j] = 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) s += v[j];
        const float mu = group_sum<LNL>(s) * (1.0f / HD);
        float qv = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) { const float d = v[j] - mu; qv += d * d; }
        const float rs = rsqrtf(group_sum<LNL>(qv) * (1.0f / HD) + L.eps);
        if (MODE == M_FWD) {
          float o[CPL];
#pragma unroll
          for (int j = 0; j < CPL; ++j) { v[j] = (v[j] - mu) * rs; o[j] = v[j] * sgam[sub * CPL + j] + sbet[sub * CPL + j]; }
          if (tvalid) {
            const int64_t otok = bh * (Lo + 1) + 1 + pos;
            store12(reinterpret_cast<T*>(J.out) + otok * J.out_ld + sub * CPL, o);
            if (J.xhat) {
              store12(reinterpret_cast<T*>(J.xhat) + otok * HD + sub * CPL, v);
              if (sub == 0) J.rstd[otok] = rs;
            }
          }
        } else {
          float dy[CPL];
          if (tvalid) {
            load12(reinterpret_cast<const T*>(J.dout) + (bh * (Lo + 1) + 1 + pos) * J.dout_ld + sub * CPL, dy);
          } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) dy[j] = 0.f;
          }
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            v[j] = (v[j] - mu) * rs;  // xhat
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
          if (tvalid) store12(reinterpret_cast<T*>(J.dconv) + (bh * Lo + pos) * HD + sub * CPL, dc);
        }
      }
      buf ^= 1;
    };

    for (int t3 = 0; t3 <= Tn; t3 += 3) {
      step(std::integral_constant<int, 0>{}, t3);
      step(std::integral_constant<int, 1>{}, t3 + 1);
      step(std::integral_constant<int, 2>{}, t3 + 2);
    }
  }

  // ---- per-CTA partial results
  if (MODE == M_BWD_LN) {
    __syncthreads();
    float* red = sm.conv;  // [TOK groups][2*HD] = 5376 floats = the whole double buffer
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      red[g * 2 * HD + sub * CPL + j] = adg[j];
      red[g * 2 * HD + HD + sub * CPL + j] = adb[j];
    }
    __syncthreads();
    if (tid < 2 * HD) {
      float s = 0.f;
#pragma unroll 7
      for (int gg = 0; gg < TOK; ++gg) s += red[gg * 2 * HD + tid];
      partial_out[tid] = s;
    }
  }
  if (MODE == M_BWD_DW) {
    __syncthreads();
    float* dws = sm.conv;  // NDW = 2592 floats
    for (int q = tid; q < NDW; q += THREADS) dws[q] = 0.f;
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < ROWS; ++rr) {
      if (conv_thread && r == rr) {
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
          dws[(2 * cp) * TAPS + k] += accw[k].x;
          dws[(2 * cp + 1) * TAPS + k] += accw[k].y;
        }
      }
      __syncthreads();
    }
    for (int q = tid; q < NDW; q += THREADS) partial_out[q] = dws[q];
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(THREADS, 2) pool_tma_kernel(const __grid_constant__ TLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ float sgam[HD], sbet[HD];
  __shared__ __align__(8) uint64_t full_bar[NST];
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const int tid = threadIdx.x;
  const int lb = blockIdx.x - J.blk_begin;
  Smem sm;
  sm.conv = reinterpret_cast<float*>(base);
  sm.planes = base + 2 * TOK * HD * sizeof(float);
  sm.full = full_bar;
  if (tid == 0) {
    tc::tma_prefetch_desc(&L.tm[jj]);
    for (int i = 0; i < NST; ++i) tc::mbar_init(&full_bar[i], 1);
    tc::fence_barrier_init();
  }
  pdl_wait();  // nothing above reads or writes global memory
  if (MODE == M_FWD || MODE == M_BWD_LN) {
    if (tid < HD) {
      sgam[tid] = J.gamma[tid];
      sbet[tid] = MODE == M_FWD ? J.beta[tid] : 0.f;
    }
  }
  __syncthreads();

  if ((MODE == M_FWD || MODE == M_BWD_LN) && lb >= J.nblk) {
    // ---------------------------------------------------------------- cls tokens: LayerNorm only, one warp per token
    const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
    const int Lo = L.T * J.Ho * J.Wo;
    const int warp = tid >> 5, lane = tid & 31;
    const int ncls = L.B * L.heads;
    float adg[3] = {0.f, 0.f, 0.f}, adb[3] = {0.f, 0.f, 0.f};
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
      if (MODE == M_FWD) {
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
    if (MODE == M_BWD_LN) {
      float* red = sm.conv;  // [CLS_WARPS][2*HD]
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

  float* partial = nullptr;
  if (MODE == M_BWD_LN) partial = J.part_ln + (int64_t)lb * 2 * HD;
  if (MODE == M_BWD_DW) partial = J.part_dw + (int64_t)lb * NDW;
  if (J.s == 1) march<T, MODE, 1>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
  else if (MODE != M_BWD_IN && J.s == 2) march<T, MODE, 2>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
  else if (MODE != M_BWD_IN) march<T, MODE, 0>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
}


// ---------------------------------------------------------------------------------------------------------------
// backward (iii), stride 2: gather form over an INPUT tile of 8 rows x 14 columns (= the 4 x 7 output tile it feeds,
// plus one halo output row / column).  The pre-LN gradient planes [5 x 8 tokens] of the three frames t-1, t, t+1 sit
// in a 4-deep TMA ring; a thread owns one even and one odd input row and a channel pair, so the tap pattern of every
// (row parity, column parity) is static: 1, 2, 2 or 4 (dh, dw) taps x 3 frames, 6.75 FFMA2 per input element.
// ---------------------------------------------------------------------------------------------------------------
constexpr int S2_NST = 4;
constexpr int S2_BH = ROWS + 1, S2_BW = CW + 1;
constexpr int S2_THREADS = ROWS * NCP;  // 192

template <typename T>
__global__ void __launch_bounds__(S2_THREADS, 3) pool_din_s2_kernel(const __grid_constant__ TLaunch L) {
  constexpr uint32_t PLANE_BYTES = S2_BH * S2_BW * HD * sizeof(T);
  constexpr uint32_t PLANE_STRIDE = (PLANE_BYTES + 127u) & ~127u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* planes = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ __align__(8) uint64_t full_bar[S2_NST];
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const CUtensorMap* tm = &L.tm[jj];
  const int tid = threadIdx.x;
  const int lb = blockIdx.x - J.blk_begin;
  const int cp = tid % NCP, r = tid / NCP;
  const int Tn = L.T, Ho = J.Ho, Wo = J.Wo;
  if (tid == 0) {
    tc::tma_prefetch_desc(tm);
    for (int i = 0; i < S2_NST; ++i) tc::mbar_init(&full_bar[i], 1);
    tc::fence_barrier_init();
  }
  pdl_wait();  // nothing above reads or writes global memory
  __syncthreads();
  float2 wr[TAPS];
  load_taps<false>(J.w, cp, wr);
  const int n_th = (L.H + 2 * ROWS - 1) / (2 * ROWS), n_tw = (L.W + 2 * CW - 1) / (2 * CW);
  const int per_bh = n_th * n_tw;
  const int nitems = L.B * L.heads * per_bh;
  int nload = 0;  // planes requested so far by this CTA (ring slot / parity bookkeeping, uniform)
  for (int item = lb; item < nitems; item += J.nblk) {
    const int bh = item / per_bh;
    const int rem = item - bh * per_bh;
    const int th = rem / n_tw, tw = rem - th * n_tw;
    const int b = bh / L.heads, head = bh - b * L.heads;
    const int base_load = nload;  // plane `to` of this item is request base_load + to
    auto issue = [&](int to) {
      if (tid == 0) {
        uint64_t* bar = &full_bar[(base_load + to) % S2_NST];
        tc::mbar_expect_tx(bar, PLANE_BYTES);
        tma_load_5d(planes + (size_t)((base_load + to) % S2_NST) * PLANE_STRIDE, tm, 0, tw * CW, th * ROWS, to, bh, bar);
      }
    };
    issue(0);
    if (Tn > 1) issue(1);
    const int hi0 = th * 2 * ROWS + 2 * r;  // even input row of this thread; hi0 + 1 is its odd row
    T* dbase = reinterpret_cast<T*>(J.din) + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs + 2 * cp);
    for (int ti = 0; ti < Tn; ++ti) {
      // frames ti-1, ti, ti+1 are needed; ti+1 arrives now, ti+2 is requested after the barrier below
      if (ti + 1 < Tn) tc::mbar_wait(&full_bar[(base_load + ti + 1) % S2_NST], (uint32_t)(((base_load + ti + 1) / S2_NST) & 1));
      else if (Tn == 1 || ti == 0) { /* single frame: plane 0 waited below */ }
      if (ti == 0) tc::mbar_wait(&full_bar[base_load % S2_NST], (uint32_t)((base_load / S2_NST) & 1));
      const T* pl[3];  // dt = 0, 1, 2  <->  output frame to = ti + 1 - dt
      bool pv[3];
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int to = ti + 1 - dt;
        pv[dt] = to >= 0 && to < Tn;
        pl[dt] = reinterpret_cast<const T*>(planes + (size_t)((base_load + (pv[dt] ? to : 0)) % S2_NST) * PLANE_STRIDE) + 2 * cp;
      }
      // tile-local output rows: even input row -> (dh = 1, row r); odd input row -> (dh = 0, row r + 1), (dh = 2, row r)
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int hi = hi0 + par;
        if (hi >= L.H) continue;
        T* drow = dbase + (int64_t)(1 + (ti * L.H + hi) * L.W + tw * 2 * CW) * L.in_ts;
#pragma unroll
        for (int c = 0; c < 2 * CW; ++c) {
          if (tw * 2 * CW + c >= L.W) continue;
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int dt = 0; dt < 3; ++dt) {
            if (!pv[dt]) continue;
#pragma unroll
            for (int a = 0; a < 2; ++a) {      // row taps
              if (par == 0 && a == 1) continue;
              const int dh = par == 0 ? 1 : (a == 0 ? 0 : 2);
              const int orow = par == 0 ? r : (a == 0 ? r + 1 : r);
#pragma unroll
              for (int e = 0; e < 2; ++e) {    // column taps
                if ((c & 1) == 0 && e == 1) continue;
                const int dw = (c & 1) == 0 ? 1 : (e == 0 ? 0 : 2);
                const int ocol = (c & 1) == 0 ? c / 2 : (e == 0 ? (c + 1) / 2 : (c - 1) / 2);
                acc = __ffma2_rn(lds2(pl[dt] + (orow * S2_BW + ocol) * HD), wr[dt * 9 + dh * 3 + dw], acc);
              }
            }
          }
          st2(drow + (int64_t)c * L.in_ts, acc);
        }
      }
      __syncthreads();  // frame ti-1 is no longer needed: its ring slot takes frame ti+2
      if (ti + 2 < Tn) issue(ti + 2);
    }
    nload += Tn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward (iii), stride >= 3: the 3 x 3 windows do not overlap, so every input position is fed by at most one
// output position: din[t, ho s + dh - 1, wo s + dw - 1] = sum_dt w[dt,dh,dw] dconv[t + 1 - dt, ho, wo], all other
// positions are zero.  One kernel zero-fills the job's slice of dQKV (16-byte stores, cls token skipped), a second
// one walks the frames per (output position, channel pair) with the three dconv values in registers.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_din_zero_kernel(const __grid_constant__ TLaunch L) {
  pdl_wait();
  constexpr int PIECES = HD * sizeof(T) / 16;  // 16-byte pieces per (token, head)
  const Job& J = L.job[blockIdx.y];
  const int64_t ntok = (int64_t)L.T * L.H * L.W;
  const int64_t total = (int64_t)L.B * ntok * L.heads * PIECES;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int piece = (int)(i % PIECES);
    int64_t q = i / PIECES;
    const int head = (int)(q % L.heads); q /= L.heads;
    const int64_t n = q % ntok;
    const int64_t b = q / ntok;
    T* p = reinterpret_cast<T*>(J.din) + b * L.in_bs + (1 + n) * L.in_ts + head * L.in_hs;
    reinterpret_cast<uint4*>(p)[piece] = make_uint4(0u, 0u, 0u, 0u);
  }
}

template <typename T>
__global__ void __launch_bounds__(S2_THREADS) pool_din_scatter_kernel(const __grid_constant__ TLaunch L) {
  pdl_wait();
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const int tid = threadIdx.x;
  const int cp = tid % NCP, r = tid / NCP;
  const int Tn = L.T, Ho = J.Ho, Wo = J.Wo, S = J.s;
  const int Lo = Tn * Ho * Wo;
  float2 wr[TAPS];
  load_taps<false>(J.w, cp, wr);
  const int npos = L.B * L.heads * Ho * Wo;
  for (int pos = (blockIdx.x - J.blk_begin) * ROWS + r; pos < npos; pos += J.nblk * ROWS) {
    const int wo = pos % Wo;
    int q = pos / Wo;
    const int ho = q % Ho; q /= Ho;
    const int head = q % L.heads, b = q / L.heads;
    const T* dc = reinterpret_cast<const T*>(J.dconv) + (((int64_t)b * L.heads + head) * Lo + (int64_t)ho * Wo + wo) * HD + 2 * cp;
    T* dbase = reinterpret_cast<T*>(J.din) + ((int64_t)b * L.in_bs + (int64_t)head * L.in_hs + 2 * cp);
    float2 d_m1 = make_float2(0.f, 0.f), d_0 = ld2(dc), d_p1;
    for (int ti = 0; ti < Tn; ++ti) {
      d_p1 = ti + 1 < Tn ? ld2(dc + (int64_t)(ti + 1) * Ho * Wo * HD) : make_float2(0.f, 0.f);
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int hi = ho * S + dh - 1;
        if (hi < 0 || hi >= L.H) continue;
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int wi = wo * S + dw - 1;
          if (wi < 0 || wi >= L.W) continue;
          float2 v = __fmul2_rn(d_p1, wr[0 * 9 + dh * 3 + dw]);
          v = __ffma2_rn(d_0, wr[1 * 9 + dh * 3 + dw], v);
          v = __ffma2_rn(d_m1, wr[2 * 9 + dh * 3 + dw], v);
          st2(dbase + (int64_t)(1 + (ti * L.H + hi) * L.W + wi) * L.in_ts, v);
        }
      }
      d_m1 = d_0;
      d_0 = d_p1;
    }
  }
}

size_t plane_bytes(int s, int esz) {
  if (s >= 3) return (size_t)9 * ROWS * CW * HD * esz;
  return (size_t)(3 + (ROWS - 1) * s) * (3 + (CW - 1) * s) * HD * esz;
}
size_t smem_bytes(size_t max_plane) { return 128 + 2 * TOK * HD * sizeof(float) + NST * ((max_plane + 127) / 128 * 128); }

int make_map(CUtensorMap* tm, const void* base, int esz, int64_t dims[5], int64_t strides_elems[4], int box_w, int box_h,
             int walk = 1) {
  return pmv_make_tensor_map_5d(tm, base, esz, (uint64_t)dims[0], (uint64_t)dims[1], (uint64_t)dims[2], (uint64_t)dims[3],
                                (uint64_t)dims[4], (uint64_t)strides_elems[0], (uint64_t)strides_elems[1],
                                (uint64_t)strides_elems[2], (uint64_t)strides_elems[3], HD, (uint32_t)box_w, (uint32_t)box_h, 1, 1,
                                (uint32_t)walk);
}

template <typename T, int MODE> int launch_mode(const TLaunch& L, int total_blocks, size_t max_plane, cudaStream_t st) {
  const size_t smem = smem_bytes(max_plane);
  auto kern = pool_tma_kernel<T, MODE>;
  static size_t attr = 0;
  if (smem > attr) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  pmv_launch(kern, (unsigned)total_blocks, THREADS, smem, st, L);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

}  // namespace

int tma_items(int B, int heads, int Ho, int Wo) { return B * heads * ((Ho + ROWS - 1) / ROWS) * ((Wo + CW - 1) / CW); }

bool tma_eligible(int stride_hw, int mode, int elem_bytes) {
  if (stride_hw < 1 || stride_hw > 8) return false;  // element strides of a tensor map go up to 8
  // nine fp32 tap tiles x 3 ring slots exceed the shared memory of an SM: fp32 mode keeps the direct kernels there
  if (mode != 3 && stride_hw >= 3 && elem_bytes != 2) return false;
  return true;
}

namespace {
int nblk_for(int64_t items, int64_t cap) { return (int)(items < 1 ? 1 : (items > cap ? cap : items)); }

// backward (iii) for the stride-2 jobs and the stride >= 3 jobs of a launch (stride 1 goes through the t-march)
template <typename T>
int launch_din_strided(const Job* jobs, int njobs, TLaunch& L, int esz, cudaStream_t st) {
  const int B = L.B, heads = L.heads, Tn = L.T, H = L.H, W = L.W;
  // ---- stride 2
  TLaunch L2 = L;
  L2.njobs = 0;
  int total = 0;
  for (int i = 0; i < njobs; ++i) {
    if (jobs[i].s != 2) continue;
    Job J = jobs[i];
    int64_t dims[5] = {HD, J.Wo, J.Ho, Tn, (int64_t)B * heads};
    int64_t str[4] = {HD, (int64_t)J.Wo * HD, (int64_t)J.Ho * J.Wo * HD, (int64_t)Tn * J.Ho * J.Wo * HD};
    int rc = make_map(&L2.tm[L2.njobs], J.dconv, esz, dims, str, S2_BW, S2_BH);
    if (rc) return rc;
    const int64_t items = (int64_t)B * heads * ((H + 2 * ROWS - 1) / (2 * ROWS)) * ((W + 2 * CW - 1) / (2 * CW));
    J.blk_begin = total;
    J.nblk = nblk_for(items, 148 * 6);
    total += J.nblk;
    L2.job[L2.njobs++] = J;
  }
  if (L2.njobs > 0) {
    const size_t smem = 128 + (size_t)S2_NST * ((S2_BH * S2_BW * HD * sizeof(T) + 127) / 128 * 128);
    auto kern = pool_din_s2_kernel<T>;
    static bool attr_set = false;
    if (!attr_set) {
      PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    pmv_launch(kern, (unsigned)total, S2_THREADS, smem, st, L2);
    PMV_CHECK_LAUNCH();
  }
  // ---- stride >= 3
  TLaunch L3 = L;
  L3.njobs = 0;
  total = 0;
  for (int i = 0; i < njobs; ++i) {
    if (jobs[i].s < 3) continue;
    Job J = jobs[i];
    J.blk_begin = total;
    J.nblk = nblk_for(ceil_div64((int64_t)B * heads * J.Ho * J.Wo, ROWS), 148 * 8);
    total += J.nblk;
    L3.job[L3.njobs++] = J;
  }
  if (L3.njobs > 0) {
    pmv_launch(pool_din_zero_kernel<T>, dim3(148 * 8, (unsigned)L3.njobs), 256, 0, st, L3);
    pmv_launch(pool_din_scatter_kernel<T>, (unsigned)total, S2_THREADS, 0, st, L3);
    PMV_CHECK_LAUNCH();
  }
  return PMV_OK;
}
}  // namespace

// mode: 0 forward, 1 backward (i), 2 backward (ii), 3 backward (iii).  `jobs` hold their block ranges for this launch
// (blk_begin / nblk / ncls_blk) and, for the backward modes, part_ln / part_dw / dconv / din.
int tma_launch(int mode, const Job* jobs, int njobs, int B, int heads, int T, int H, int W, int64_t bs, int64_t ts, int64_t hs,
               float eps, int dtype, cudaStream_t st) {
  TLaunch L;
  memset(&L, 0, sizeof(L));
  L.njobs = njobs; L.B = B; L.heads = heads; L.T = T; L.H = H; L.W = W;
  L.in_bs = bs; L.in_ts = ts; L.in_hs = hs; L.eps = eps;
  const int esz = dtype == PMV_BF16 ? 2 : 4;
  if (mode == 3) {
    // stride 1 jobs: t-march over dconv below; stride 2 / >= 3: dedicated kernels
    Job s1[MAX_JOBS];
    int n1 = 0, nother = 0, tot1 = 0;
    for (int i = 0; i < njobs; ++i) {
      if (jobs[i].s == 1) {
        s1[n1] = jobs[i];
        s1[n1].blk_begin = tot1;
        tot1 += s1[n1].nblk;
        ++n1;
      } else {
        ++nother;
      }
    }
    if (nother > 0) {
      int rc = dtype == PMV_BF16 ? launch_din_strided<bf16>(jobs, njobs, L, esz, st) : launch_din_strided<float>(jobs, njobs, L, esz, st);
      if (rc) return rc;
    }
    if (n1 == 0) return PMV_OK;
    if (n1 < njobs) return tma_launch(3, s1, n1, B, heads, T, H, W, bs, ts, hs, eps, dtype, st);
  }
  int total = 0;
  size_t max_plane = 0;
  for (int i = 0; i < njobs; ++i) {
    L.job[i] = jobs[i];
    const Job& J = jobs[i];
    const int S = J.s;
    const size_t pb = plane_bytes(mode == 3 ? 1 : S, esz);
    if (pb > max_plane) max_plane = pb;
    int rc;
    if (mode == 3) {
      int64_t dims[5] = {HD, W, H, T, (int64_t)B * heads};
      int64_t str[4] = {HD, (int64_t)W * HD, (int64_t)H * W * HD, (int64_t)T * H * W * HD};
      rc = make_map(&L.tm[i], J.dconv, esz, dims, str, 3 + (CW - 1), 3 + (ROWS - 1));
    } else {
      int64_t dims[5] = {(int64_t)heads * HD, W, H, T, B};
      int64_t str[4] = {ts, (int64_t)W * ts, (int64_t)H * W * ts, bs};
      // token 0 is the cls token: the spatial volume starts one token in
      const char* vol = reinterpret_cast<const char*>(J.in) + ts * esz;
      if (S >= 3) rc = make_map(&L.tm[i], vol, esz, dims, str, (CW - 1) * S + 1, (ROWS - 1) * S + 1, S);  // walks every S-th token
      else rc = make_map(&L.tm[i], vol, esz, dims, str, 3 + (CW - 1) * S, 3 + (ROWS - 1) * S);
    }
    if (rc) return rc;
    total = J.blk_begin + J.nblk + ((mode == 0 || mode == 1) ? J.ncls_blk : 0);
  }
  if (dtype == PMV_BF16) {
    switch (mode) {
      case 0: return launch_mode<bf16, M_FWD>(L, total, max_plane, st);
      case 1: return launch_mode<bf16, M_BWD_LN>(L, total, max_plane, st);
      case 2: return launch_mode<bf16, M_BWD_DW>(L, total, max_plane, st);
      default: return launch_mode<bf16, M_BWD_IN>(L, total, max_plane, st);
    }
  }
  switch (mode) {
    case 0: return launch_mode<float, M_FWD>(L, total, max_plane, st);
    case 1: return launch_mode<float, M_BWD_LN>(L, total, max_plane, st);
    case 2: return launch_mode<float, M_BWD_DW>(L, total, max_plane, st);
    default: return launch_mode<float, M_BWD_IN>(L, total, max_plane, st);
  }
}

}  // namespace pool
