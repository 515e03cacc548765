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
#include <cstdlib>
#include "pool_common.cuh"
#include "tc_common.cuh"

namespace pool {
#ifdef PMV_ATTN_TRACE
// debug builds only (scripts/build_trace_lib.py): per-step phase stamps of thread 0 of the first 1024 CTAs
constexpr int PTRACE_CTAS = 1024, PTRACE_SLOTS = 64;
__device__ long long pmv_pool_trace_buf[4 * PTRACE_CTAS * PTRACE_SLOTS];  // [MODE][cta][slot]
#define PTRACE(slot)                                                                                              \
  do {                                                                                                            \
    if (threadIdx.x == 0 && blockIdx.x < PTRACE_CTAS && (slot) < PTRACE_SLOTS)                                    \
      pmv_pool_trace_buf[(MODE * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + (slot)] = clock64();                                         \
  } while (0)
#define WTRACE(cond, slot)                                                                                        \
  do {                                                                                                            \
    if ((cond) && blockIdx.x < PTRACE_CTAS && (slot) < PTRACE_SLOTS)                                              \
      pmv_pool_trace_buf[(1 * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + (slot)] = clock64();                     \
  } while (0)
#else
#define PTRACE(slot) do { } while (0)
#define WTRACE(cond, slot) do { } while (0)
#endif
namespace {

constexpr int CW = 7;                 // output columns per tile (56, 28, 14, 7 are multiples)
constexpr int TOK = ROWS * CW;        // 28 tokens per LayerNorm phase
constexpr int THREADS = TOK * LNL;    // 224: conv phase uses the first 192 (4 rows x 48 channel pairs)
constexpr int NST = 3;                // plane ring depth
constexpr int CLS_WARPS = THREADS / 32;
// the dW and stride-1 input-gradient marches have no LayerNorm phase: 192 threads, so 2 CTAs per SM get 168 registers
constexpr int threads_for(int mode) { return (mode == 2 || mode == 3) ? ROWS * NCP : THREADS; }
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

__device__ __forceinline__ void tma_load_5d_a(uint32_t smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 lds2(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

__device__ const uint32_t pool_zero_words[4] = {0u, 0u, 0u, 0u};
__device__ __forceinline__ uint32_t ld_raw(const bf16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ float2 ld_raw(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 raw_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
__device__ __forceinline__ float2 raw_to_f2(float2 v) { return v; }

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
  // job-table fields used inside the frame loop are pinned in registers for the dW / input-gradient marches (indexed
  // constant loads otherwise, see pin()); the forward / recompute modes of this kernel are fallbacks and short of registers
  constexpr bool PIN = MODE == M_BWD_DW || MODE == M_BWD_IN;
  const int Ho = PIN ? pin(J.Ho) : J.Ho, Wo = PIN ? pin(J.Wo) : J.Wo, Tn = PIN ? pin(L.T) : L.T;
  T* const dinp = PIN ? pin(reinterpret_cast<T*>(J.din)) : reinterpret_cast<T*>(J.din);
  const T* const dconvp = PIN ? pin(reinterpret_cast<const T*>(J.dconv)) : reinterpret_cast<const T*>(J.dconv);
  const int64_t in_bs = PIN ? pin(L.in_bs) : L.in_bs, in_ts = PIN ? pin(L.in_ts) : L.in_ts, in_hs = PIN ? pin(L.in_hs) : L.in_hs;
  const int Lo = Tn * Ho * Wo;
  const int n_th = (Ho + ROWS - 1) / ROWS, n_tw = (Wo + CW - 1) / CW;
  const int per_bh = n_th * n_tw;
  const int nitems = L.B * L.heads * per_bh;
  const int n_my = lb < nitems ? (nitems - lb + nblk - 1) / nblk : 0;
  const int total = n_my * Tn;
  // LayerNorm-phase role
  const int g = tid >> 3, sub = tid & 7;
  const int g_row = g / CW, g_w = g - g_row * CW;
  using Raw = typename std::conditional<std::is_same<T, bf16>::value, uint32_t, float2>::type;

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
      const T* dc = dconvp + ((bh * Lo + (int64_t)(tout * Ho + row) * Wo + col0) * HD + 2 * cp);
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const bool ok = row_ok && tout >= 0 && tout < Tn && col0 + j < Wo;
        dst[j] = ok ? ld2(dc + j * HD) : make_float2(0.f, 0.f);
      }
    };
    // M_BWD_DW: step tin needs the pre-LN gradient of frames tin-1, tin, tin+1.  The phase stamps of round 2
    // (profiles/r02_pool_trace.md) showed 0.7-1.5 us per step between the FFMA2 section and the barrier: `ok ? load : 0`
    // and the bf16 unpack consume the loaded word right after the request, so the warp sat out a full L2 / DRAM round trip
    // every step.  Now the raw words of frame tin+3 are requested at the end of step tin (out-of-range positions read a
    // zero word instead of being selected afterwards) and unpacked one step later, into the slot frame tin-1 vacated.
    Raw raw[MODE == M_BWD_DW ? CW : 1];
    auto fetch_dc = [&](int tout) {
      const T* dc = dconvp + ((bh * Lo + (int64_t)(tout * Ho + row) * Wo + col0) * HD + 2 * cp);
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const bool ok = row_ok && tout >= 0 && tout < Tn && col0 + j < Wo;
        raw[j] = ld_raw(ok ? dc + j * HD : reinterpret_cast<const T*>(pool_zero_words));
      }
    };
    if constexpr (MODE == M_BWD_DW) {
      load_dc(acc[0], 0);
      load_dc(acc[1], 1);
      fetch_dc(2);
    }

    auto step = [&](auto ptag, int tin) {
      constexpr int P = decltype(ptag)::value;
      constexpr int SLOT_M1 = (P + 2) % 3, SLOT_0 = P, SLOT_P1 = (P + 1) % 3;  // output frames tin-1, tin, tin+1
      if (tin > Tn) return;
      const int k = i * Tn + tin;
      if (i == 0) PTRACE(8 + 6 * tin);
      if (tin < Tn) {
        tc::mbar_wait(&sm.full[k % NST], (uint32_t)((k / NST) & 1));
        if (i == 0) PTRACE(8 + 6 * tin + 1);
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
      if (i == 0) PTRACE(8 + 6 * tin + 2);
      if constexpr (MODE == M_BWD_DW) {
        if (tin < Tn) {
#pragma unroll
          for (int j = 0; j < CW; ++j) acc[SLOT_M1][j] = raw_to_f2(raw[j]);  // frame tin + 2, requested a step ago
          fetch_dc(tin + 3);
        }
        __syncthreads();  // everyone is done with plane k: its ring slot may be refilled
        if (i == 0) PTRACE(8 + 6 * tin + 3);
        if (tid == 0 && tin < Tn && k + NST < total) issue(k + NST);
        return;
      }
      if (MODE == M_BWD_IN) {
        if (tout >= 0 && row_ok) {
          T* drow = dinp + ((int64_t)b * in_bs + (int64_t)head * in_hs + (int64_t)(1 + (tout * Ho + row) * Wo + col0) * in_ts + 2 * cp);
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (col0 + j < Wo) st2(drow + (int64_t)j * in_ts, acc[SLOT_M1][j]);
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
      if (i == 0) PTRACE(8 + 6 * tin + 4);
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
    for (int q = tid; q < NDW; q += threads_for(MODE)) dws[q] = 0.f;
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
    for (int q = tid; q < NDW; q += threads_for(MODE)) partial_out[q] = dws[q];
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(threads_for(MODE), 2) pool_tma_kernel(const __grid_constant__ TLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ float sgam[HD], sbet[HD];
  __shared__ __align__(8) uint64_t full_bar[NST];
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const int tid = threadIdx.x;
  const int lb = blockIdx.x - J.blk_begin;
#ifdef PMV_ATTN_TRACE
  if (tid == 0 && blockIdx.x < PTRACE_CTAS) {
    PTRACE(0);
    unsigned smid; unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    pmv_pool_trace_buf[(MODE * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + 1] = smid;
    pmv_pool_trace_buf[(MODE * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + 2] = (long long)gt;
    pmv_pool_trace_buf[(MODE * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + 4] = J.s * 10000 + jj * 1000 + (lb >= J.nblk ? 999 : 0);
  }
#endif
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
  PTRACE(5);
  if (J.s == 1) march<T, MODE, 1>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
  else if (MODE != M_BWD_IN && J.s == 2) march<T, MODE, 2>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
  else if (MODE != M_BWD_IN) march<T, MODE, 0>(L, J, &L.tm[jj], sm, lb, J.nblk, sgam, sbet, partial);
#ifdef PMV_ATTN_TRACE
  if (tid == 0 && blockIdx.x < PTRACE_CTAS) {
    PTRACE(6);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    pmv_pool_trace_buf[(MODE * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS + 3] = (long long)gt;
  }
#endif
}


// ---------------------------------------------------------------------------------------------------------------
// Forward, warp-specialised (round 2).  The phase stamps of the kernel above (scripts/pool_trace.py,
// profiles/r02_pool_trace.md) showed where a frame step of 1.7-1.9 us went: 0.1 us waiting for the plane, 0.55-0.7 us in
// the FFMA2 section (= the FMA-pipe rate of the 12 conv warps of an SM) and 1.1 us in park -> __syncthreads -> LayerNorm
// -> stores, during which the FMA pipe idles because all warps of the CTA move through the phases together.
// Here the two halves run as a producer / consumer pipeline inside the CTA, with no CTA-wide barrier in the march:
//   warps 0-5  (192 threads = 4 rows x 48 channel pairs): wait full[k] -> 27 LDS + 189 FFMA2 -> arrive empty[k] -> park
//              the completed frame in a 2-deep ring (wait cfree / arrive parked) -> next plane;
//   warps 6-7  LayerNorm + stores of the parked frames: 4 lanes per token x 24 channels, 14 tokens per warp in two
//              interleaved passes, packed fp32 math, 16-byte stores; warp 6 also refills the plane ring (all lanes wait
//              on empty[k], lane 0 issues the TMA request) as soon as every conv warp released a slot.
// All hand-offs are mbarriers (one elected arrive per warp after __syncwarp).  Lessons of the first versions (same trace
// file): (a) the role branch must use a provably warp-uniform warp index and the wait loops must not be lane-divergent,
// otherwise every __shfl_xor_sync / __syncwarp compiles to its WARPSYNC.COLLECTIVE form and a LayerNorm pass takes 3.5-5.5
// us; (b) doing the LayerNorm in the conv warps themselves one frame later ("deferred", no role split) keeps all warps in
// phase through the shared barriers - 0.5 us FFMA2 + 0.75 us LayerNorm per step, no overlap; (c) shared loads / stores
// are explicit ld.shared / st.shared (the pointer-in-struct form above compiled to generic LD / ST).
// ---------------------------------------------------------------------------------------------------------------
constexpr int WS_THREADS = 256;
constexpr int WS_CONV_WARPS = 6, WS_LN_WARPS = 2;
constexpr int NCB = 2;          // parked-frame ring depth
constexpr int LN4 = 4;          // lanes per token in the LayerNorm warps
constexpr int CPL4 = HD / LN4;  // 24 channels per lane
constexpr int TPW = TOK / WS_LN_WARPS;  // 14 tokens per LayerNorm warp and frame: two passes of 8 lane groups
static_assert(WS_CONV_WARPS * 32 == ROWS * NCP && TPW * WS_LN_WARPS == TOK && TPW <= 16, "role mapping");

__device__ __forceinline__ float2 lds_pair(uint32_t addr, const bf16*) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ float2 lds_pair(uint32_t addr, const float*) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 8 consecutive channels: one 16-byte store (bf16) / two (fp32)
__device__ __forceinline__ void store8(bf16* p, float2 a, float2 b, float2 c, float2 d) {
  uint4 t;
  t.x = pack2(a); t.y = pack2(b); t.z = pack2(c); t.w = pack2(d);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void store8(float* p, float2 a, float2 b, float2 c, float2 d) {
  *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
  *reinterpret_cast<float4*>(p + 4) = make_float4(c.x, c.y, d.x, d.y);
}

struct WsBars {
  uint64_t full[NST], empty[NST], parked[NCB], cfree[NCB];
};

struct WsItem {
  int b, head, th, tw;
};
__device__ __forceinline__ WsItem ws_decode(int item, int per_bh, int n_tw, int heads) {
  WsItem it;
  const int bh = item / per_bh;
  const int rem = item - bh * per_bh;
  it.th = rem / n_tw;
  it.tw = rem - it.th * n_tw;
  it.b = bh / heads;
  it.head = bh - it.b * heads;
  return it;
}

// ---- conv warps -----------------------------------------------------------------------------------------------------
template <typename T, int S>
__device__ __forceinline__ void ws_conv(const TLaunch& L, const Job& J, uint32_t planes_a, uint32_t cbuf_a, WsBars* bars, int n_my,
                                        int tid /* 0..191: conv thread index (compact over the conv warps) */) {
  using G = Geo<S>;
  constexpr uint32_t PLANE_BYTES = G::PLANE_ELEMS * sizeof(T);
  constexpr uint32_t PLANE_STRIDE = (PLANE_BYTES + 127u) & ~127u;
  constexpr uint32_t ESZ = sizeof(T);
  const int lane = tid & 31;
  const int cp = tid % NCP, r = tid / NCP;
  const int Tn = pin(L.T);
  float2 wr[TAPS];
  load_taps<false>(J.w, cp, wr);
  const uint32_t park_off = (uint32_t)((r * CW) * HD + 2 * cp) * 4u;
  int f = 0;  // parked frames so far
  auto park = [&](float2 (&a)[CW]) {
    const int slot = f % NCB;
    tc::mbar_wait(&bars->cfree[slot], (uint32_t)(((f / NCB) & 1) ^ 1));
    const uint32_t dst = cbuf_a + (uint32_t)slot * (TOK * HD * 4u) + park_off;
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      sts_f2(dst + (uint32_t)j * (HD * 4u), a[j]);
      a[j] = make_float2(0.f, 0.f);
    }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->parked[slot]);
    ++f;
  };
  for (int i = 0; i < n_my; ++i) {
    float2 acc[3][CW];
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
      for (int j = 0; j < CW; ++j) acc[s3][j] = make_float2(0.f, 0.f);
    auto step = [&](auto ptag, int tin) {
      constexpr int P = decltype(ptag)::value;
      constexpr int SLOT_M1 = (P + 2) % 3, SLOT_0 = P, SLOT_P1 = (P + 1) % 3;  // output frames tin-1, tin, tin+1
      if (tin >= Tn) return;
      const int k = i * Tn + tin;
      WTRACE(tid == 0 && i == 0, 8 + 4 * tin);
      tc::mbar_wait(&bars->full[k % NST], (uint32_t)((k / NST) & 1));
      WTRACE(tid == 0 && i == 0, 8 + 4 * tin + 1);
      const uint32_t plane = planes_a + (uint32_t)(k % NST) * PLANE_STRIDE + (uint32_t)(2 * cp) * ESZ;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        constexpr int NX = S == 0 ? 3 * CW : G::BW;
        float2 x[NX];
        if constexpr (S == 0) {  // tap tiles (dh, dw): [ROWS][CW] tokens each
#pragma unroll
          for (int dw = 0; dw < 3; ++dw)
#pragma unroll
            for (int j = 0; j < CW; ++j)
              x[dw * CW + j] = lds_pair(plane + (uint32_t)((((dh * 3 + dw) * ROWS + r) * CW + j) * HD) * ESZ, (const T*)nullptr);
        } else {
          const uint32_t prow = plane + (uint32_t)((r * S + dh) * (G::BW * HD)) * ESZ;
#pragma unroll
          for (int c = 0; c < G::BW; ++c) x[c] = lds_pair(prow + (uint32_t)(c * HD) * ESZ, (const T*)nullptr);
        }
#pragma unroll
        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float2 xv = S == 0 ? x[dw * CW + j] : x[j * S + dw];
            acc[SLOT_P1][j] = __ffma2_rn(xv, wr[0 * 9 + dh * 3 + dw], acc[SLOT_P1][j]);
            acc[SLOT_0][j] = __ffma2_rn(xv, wr[1 * 9 + dh * 3 + dw], acc[SLOT_0][j]);
            acc[SLOT_M1][j] = __ffma2_rn(xv, wr[2 * 9 + dh * 3 + dw], acc[SLOT_M1][j]);
          }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->empty[k % NST]);  // this warp is done with plane k
      WTRACE(tid == 0 && i == 0, 8 + 4 * tin + 2);
      if (tin >= 1) {
        park(acc[SLOT_M1]);
      } else {  // what plane 0 contributed to "frame -1" is discarded; the slot becomes frame 2 in the next step
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[SLOT_M1][j] = make_float2(0.f, 0.f);
      }
      if (tin == Tn - 1) park(acc[SLOT_0]);  // frame Tn would be zero padding: the last frame is complete as well
      WTRACE(tid == 0 && i == 0, 8 + 4 * tin + 3);
    };
    for (int t3 = 0; t3 < Tn; t3 += 3) {
      step(std::integral_constant<int, 0>{}, t3);
      step(std::integral_constant<int, 1>{}, t3 + 1);
      step(std::integral_constant<int, 2>{}, t3 + 2);
    }
  }
}

// ---- LayerNorm warps (+ the TMA producer in the first one) -----------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ws_ln(const TLaunch& L, const Job& J, const CUtensorMap* tm, uint32_t planes_a, uint32_t cbuf_a,
                                      WsBars* bars, const float* sgam, const float* sbet, int w, int lb, int nblk, int n_my,
                                      int per_bh, int n_tw) {
  // one instance for all strides (the plane geometry only matters to the TMA requests): keeps the kernel's code small
  const int S = pin(J.s);
  const uint32_t PLANE_BYTES = (uint32_t)(S == 1 ? Geo<1>::PLANE_ELEMS : S == 2 ? Geo<2>::PLANE_ELEMS : Geo<0>::PLANE_ELEMS) * sizeof(T);
  const uint32_t PLANE_STRIDE = (PLANE_BYTES + 127u) & ~127u;
  const int lane = threadIdx.x & 31;
  const int g8 = lane >> 2, sub = lane & (LN4 - 1);  // lane group (token) 0..7, channel quarter
  const int Tn = pin(L.T), Ho = pin(J.Ho), Wo = pin(J.Wo), heads = pin(L.heads);
  const int Lo = Tn * Ho * Wo;
  const int total = n_my * Tn;
  // ---- producer state (warp 0): next plane to request = (item index pi, frame pt), decoded tile of that item.  (Moving
  // the requests into conv warp 0 was tried: it then waits for the slowest conv warp every step, 0.3-0.5 us.)
  int pk = 0, pi = 0, pt = 0;
  WsItem pit = ws_decode(lb, per_bh, n_tw, heads);
  auto issue_next = [&]() {  // all lanes keep the bookkeeping, lane 0 talks to the TMA unit
    if (lane == 0) {
      uint64_t* bar = &bars->full[pk % NST];
      tc::mbar_expect_tx(bar, PLANE_BYTES);
      const uint32_t dst = planes_a + (uint32_t)(pk % NST) * PLANE_STRIDE;
      if (S >= 3) {
        constexpr uint32_t TILE_BYTES = Geo<0>::TILE_ELEMS * sizeof(T);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
          tma_load_5d_a(dst + tap * TILE_BYTES, tm, pit.head * HD, pit.tw * CW * S + tap % 3 - 1, pit.th * ROWS * S + tap / 3 - 1, pt,
                        pit.b, bar);
      } else {
        tma_load_5d_a(dst, tm, pit.head * HD, pit.tw * CW * S - 1, pit.th * ROWS * S - 1, pt, pit.b, bar);
      }
    }
    ++pk;
    if (++pt == Tn) {
      pt = 0;
      ++pi;
      pit = ws_decode(lb + pi * nblk, per_bh, n_tw, heads);
    }
  };
  if (w == 0) {
    while (pk < NST && pk < total) issue_next();
    __syncwarp();
  }
  int kr = 0;  // planes released so far (warp 0)
  // gamma / beta of this lane's 24 channels come from shared memory (two LDS.128 per 8 channels): 48 registers less
  const uint32_t gam_a = tc::smem_u32(sgam) + (uint32_t)(sub * CPL4) * 4u, bet_a = tc::smem_u32(sbet) + (uint32_t)(sub * CPL4) * 4u;
  T* const outp = pin(reinterpret_cast<T*>(J.out));
  T* const xhp = pin(reinterpret_cast<T*>(J.xhat));
  float* const rsp = pin(J.rstd);
  const int64_t out_ld = pin(J.out_ld);
  const float eps = pin(L.eps);
  const int onehot = pin(J.onehot);
  const int aug4 = ((int)out_ld - HD) / LN4;  // one-hot columns per lane (a multiple of 8 when J.onehot is set)
  // this lane's two tokens of a frame (pass 0: all 8 groups, pass 1: groups 0..5)
  int trow[2], tcol[2];
  uint32_t toff[2];
  bool tlane[2];
#pragma unroll
  for (int p2 = 0; p2 < 2; ++p2) {
    const int tw_ = p2 * 8 + g8;
    tlane[p2] = tw_ < TPW;
    const int tok = w * TPW + (tlane[p2] ? tw_ : TPW - 1);
    trow[p2] = tok / CW;
    tcol[p2] = tok - trow[p2] * CW;
    toff[p2] = (uint32_t)(tok * HD + sub * CPL4) * 4u;
  }
  int i = 0, tout = 0;
  WsItem it = ws_decode(lb, per_bh, n_tw, heads);
  for (int f = 0; f < total; ++f) {
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4, 40 + 6 * tout);
    if (w == 0) {  // refill the plane ring: every plane the conv warps have finished by the time frame f is parked
      int target = i * Tn + tout + 2;
      if (target > (i + 1) * Tn) target = (i + 1) * Tn;
      for (; kr < target; ++kr) {
        tc::mbar_wait(&bars->empty[kr % NST], (uint32_t)((kr / NST) & 1));
        if (pk < total) issue_next();  // plane kr + NST
      }
      __syncwarp();
    }
    const int slot = f % NCB;
    tc::mbar_wait(&bars->parked[slot], (uint32_t)((f / NCB) & 1));
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4, 40 + 6 * tout + 1);
    const uint32_t src = cbuf_a + (uint32_t)slot * (TOK * HD * 4u);
    float2 v[2][CPL4 / 2];
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2)
#pragma unroll
      for (int q4 = 0; q4 < CPL4 / 4; ++q4) {
        const float4 t4 = lds_f4(src + toff[p2] + 16u * q4);
        v[p2][2 * q4] = make_float2(t4.x, t4.y);
        v[p2][2 * q4 + 1] = make_float2(t4.z, t4.w);
      }
    float mu[2], var[2];
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2) {
      float2 s2 = __fadd2_rn(v[p2][0], v[p2][1]), s3 = __fadd2_rn(v[p2][2], v[p2][3]);
#pragma unroll
      for (int j = 4; j < CPL4 / 2; j += 2) { s2 = __fadd2_rn(s2, v[p2][j]); s3 = __fadd2_rn(s3, v[p2][j + 1]); }
      s2 = __fadd2_rn(s2, s3);
      mu[p2] = s2.x + s2.y;
    }
    // every value of the frame has gone into the sums above, so the loads have landed: release the slot
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4 && mu[0] != 12345.f, 40 + 6 * tout + 2);
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->cfree[slot]);
#pragma unroll
    for (int o = 1; o < LN4; o <<= 1)
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) mu[p2] += __shfl_xor_sync(0xffffffffu, mu[p2], o);
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4 && mu[0] != 12345.f, 40 + 6 * tout + 3);
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2) {
      mu[p2] *= (1.0f / HD);
      const float2 nm = make_float2(-mu[p2], -mu[p2]);
      float2 q2 = make_float2(0.f, 0.f), q3 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < CPL4 / 2; j += 2) {
        v[p2][j] = __fadd2_rn(v[p2][j], nm);
        v[p2][j + 1] = __fadd2_rn(v[p2][j + 1], nm);
        q2 = __ffma2_rn(v[p2][j], v[p2][j], q2);
        q3 = __ffma2_rn(v[p2][j + 1], v[p2][j + 1], q3);
      }
      q2 = __fadd2_rn(q2, q3);
      var[p2] = q2.x + q2.y;
    }
#pragma unroll
    for (int o = 1; o < LN4; o <<= 1)
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) var[p2] += __shfl_xor_sync(0xffffffffu, var[p2], o);
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4 && var[0] != 12345.f, 40 + 6 * tout + 4);
    const int64_t tok0 = ((int64_t)it.b * heads + it.head) * (Lo + 1) + 1 + (int64_t)tout * Ho * Wo;
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2) {
      const float rs = rsqrtf(var[p2] * (1.0f / HD) + eps);
      const float2 r2 = make_float2(rs, rs);
      const int orow = it.th * ROWS + trow[p2], ocol = it.tw * CW + tcol[p2];
      const bool tvalid = tlane[p2] && orow < Ho && ocol < Wo;
      const int64_t otok = tok0 + orow * Wo + ocol;
      T* const op = outp + otok * out_ld + sub * CPL4;
      T* const xp = xhp + otok * HD + sub * CPL4;
#pragma unroll
      for (int q8 = 0; q8 < CPL4 / 8; ++q8) {
        float2 o[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 g4 = lds_f4(gam_a + 32u * q8 + 16u * e), b4 = lds_f4(bet_a + 32u * q8 + 16u * e);
          v[p2][4 * q8 + 2 * e] = __fmul2_rn(v[p2][4 * q8 + 2 * e], r2);
          v[p2][4 * q8 + 2 * e + 1] = __fmul2_rn(v[p2][4 * q8 + 2 * e + 1], r2);
          o[2 * e] = __ffma2_rn(v[p2][4 * q8 + 2 * e], make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
          o[2 * e + 1] = __ffma2_rn(v[p2][4 * q8 + 2 * e + 1], make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
        }
        if (tvalid) {
          store8(op + 8 * q8, o[0], o[1], o[2], o[3]);
          if (xhp) store8(xp + 8 * q8, v[p2][4 * q8], v[p2][4 * q8 + 1], v[p2][4 * q8 + 2], v[p2][4 * q8 + 3]);
        }
      }
      if (tvalid && xhp && sub == 0) rsp[otok] = rs;
      if (onehot && tvalid) {
        // K' of the rel-pos scheme: columns [96, ld) = onehot(k_h) | onehot(k_w) | onehot(k_t) (csrc/relpos.cu); this lane
        // writes its quarter of them
        const int c0 = orow, c1 = Ho + ocol, c2 = Ho + Wo + tout;
        T* const ap = outp + otok * out_ld + HD + sub * aug4;
        for (int e = 0; e < aug4; e += 8) {
          float2 h[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = sub * aug4 + e + 2 * u;
            h[u].x = (j == c0 || j == c1 || j == c2) ? 1.f : 0.f;
            h[u].y = (j + 1 == c0 || j + 1 == c1 || j + 1 == c2) ? 1.f : 0.f;
          }
          store8(ap + e, h[0], h[1], h[2], h[3]);
        }
      }
    }
    WTRACE(w == 0 && lane == 0 && i == 0 && tout < 4, 40 + 6 * tout + 5);
    if (++tout == Tn) {
      tout = 0;
      ++i;
      it = ws_decode(lb + i * nblk, per_bh, n_tw, heads);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(WS_THREADS, 2) pool_ws_fwd_kernel(const __grid_constant__ TLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float sgam[HD];
  __shared__ __align__(16) float sbet[HD];
  __shared__ __align__(8) WsBars bars;
  int jj = 0;
  while (jj + 1 < L.njobs && (int)blockIdx.x >= L.job[jj + 1].blk_begin) ++jj;
  const Job& J = L.job[jj];
  const int tid = threadIdx.x;
  const int lb = blockIdx.x - J.blk_begin;
  const uint32_t base_a = (tc::smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t cbuf_a = base_a, planes_a = base_a + NCB * TOK * HD * 4u;
#ifdef PMV_ATTN_TRACE
  if (tid == 0 && blockIdx.x < PTRACE_CTAS) {
    long long* tb = pmv_pool_trace_buf + (1 * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS;
    unsigned smid; unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tb[0] = clock64(); tb[1] = smid; tb[2] = (long long)gt; tb[4] = J.s * 10000 + jj * 1000 + (lb >= J.nblk ? 999 : 0);
  }
#endif
  if (tid == 0) {
    tc::tma_prefetch_desc(&L.tm[jj]);
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&bars.full[i], 1); tc::mbar_init(&bars.empty[i], WS_CONV_WARPS); }
    for (int i = 0; i < NCB; ++i) { tc::mbar_init(&bars.parked[i], WS_CONV_WARPS); tc::mbar_init(&bars.cfree[i], WS_LN_WARPS); }
    tc::fence_barrier_init();
  }
  pdl_wait();  // nothing above reads or writes global memory
  if (tid < HD) { sgam[tid] = J.gamma[tid]; sbet[tid] = J.beta[tid]; }
  __syncthreads();  // barrier initialisation, gamma / beta visible to every warp
  if (lb >= J.nblk) {
    // ---------------------------------------------------------------- cls tokens: LayerNorm only, one warp per token
    const T* __restrict__ in = reinterpret_cast<const T*>(J.in);
    const int Lo = L.T * J.Ho * J.Wo;
    const int warp = tid >> 5, lane = tid & 31;
    const int ncls = L.B * L.heads;
    for (int bh = (lb - J.nblk) * (WS_THREADS / 32) + warp; bh < ncls; bh += J.ncls_blk * (WS_THREADS / 32)) {
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
      T* o = reinterpret_cast<T*>(J.out) + tok * J.out_ld + lane;
#pragma unroll
      for (int j = 0; j < 3; ++j) o[32 * j] = from_f32<T>((v[j] - mu) * rs * sgam[lane + 32 * j] + sbet[lane + 32 * j]);
      if (J.onehot) {  // the cls key carries no coordinates
        for (int j = HD + lane; j < (int)J.out_ld; j += 32) o[j - lane] = from_f32<T>(0.f);
      }
      if (J.xhat) {
        T* xo = reinterpret_cast<T*>(J.xhat) + tok * HD + lane;
#pragma unroll
        for (int j = 0; j < 3; ++j) xo[32 * j] = from_f32<T>((v[j] - mu) * rs);
        if (lane == 0) J.rstd[tok] = rs;
      }
    }
    return;
  }
  const int n_th = (J.Ho + ROWS - 1) / ROWS, n_tw = (J.Wo + CW - 1) / CW;
  const int per_bh = n_th * n_tw;
  const int nitems = L.B * L.heads * per_bh;
  const int n_my = lb < nitems ? (nitems - lb + J.nblk - 1) / J.nblk : 0;
  // warp index through a shuffle: provably warp-uniform, so the role branch below is not treated as divergent
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // roles by warp: 3 and 7 are the LayerNorm warps, so that they sit together on one scheduler (warp id mod 4) instead
  // of sharing schedulers with FFMA2-saturated conv warps (profiles/r02_pool_trace.md: their ~350 instructions per frame
  // took 1800 cycles there); the conv warps 0,1,2,4,5,6 fill the other three schedulers evenly
  if ((warp_u & 3) != 3) {
    const int ctid = (warp_u - (warp_u >> 2)) * 32 + (tid & 31);
    if (J.s == 1) ws_conv<T, 1>(L, J, planes_a, cbuf_a, &bars, n_my, ctid);
    else if (J.s == 2) ws_conv<T, 2>(L, J, planes_a, cbuf_a, &bars, n_my, ctid);
    else ws_conv<T, 0>(L, J, planes_a, cbuf_a, &bars, n_my, ctid);
  } else {
    ws_ln<T>(L, J, &L.tm[jj], planes_a, cbuf_a, &bars, sgam, sbet, warp_u >> 2, lb, J.nblk, n_my, per_bh, n_tw);
  }
#ifdef PMV_ATTN_TRACE
  if (tid == 0 && blockIdx.x < PTRACE_CTAS) {
    long long* tb = pmv_pool_trace_buf + (1 * PTRACE_CTAS + blockIdx.x) * PTRACE_SLOTS;
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tb[6] = clock64(); tb[3] = (long long)gt;
  }
#endif
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
  const int Tn = pin(L.T), Ho = pin(J.Ho), Wo = pin(J.Wo), LH = pin(L.H), LW = pin(L.W);
  const int64_t in_ts = pin(L.in_ts);
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
        if (hi >= LH) continue;
        T* drow = dbase + (int64_t)(1 + (ti * LH + hi) * LW + tw * 2 * CW) * in_ts;
#pragma unroll
        for (int c = 0; c < 2 * CW; ++c) {
          if (tw * 2 * CW + c >= LW) continue;
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
          st2(drow + (int64_t)c * in_ts, acc);
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

template <typename T> int launch_ws_fwd(const TLaunch& L, int total_blocks, size_t max_plane, cudaStream_t st) {
  const size_t smem = 128 + NCB * TOK * HD * sizeof(float) + NST * ((max_plane + 127) / 128 * 128);
  auto kern = pool_ws_fwd_kernel<T>;
  static size_t attr = 0;
  if (smem > attr) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  pmv_launch(kern, (unsigned)total_blocks, WS_THREADS, smem, st, L);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

// PMV_POOL_WS=0 selects the forward kernel of round 1 (kept as the A/B reference of the warp-specialised one)
bool ws_enabled() {
  static const bool on = [] { const char* e = std::getenv("PMV_POOL_WS"); return e == nullptr || e[0] != '0'; }();
  return on;
}

template <typename T, int MODE> int launch_mode(const TLaunch& L, int total_blocks, size_t max_plane, cudaStream_t st) {
  if (MODE == M_FWD && ws_enabled()) return launch_ws_fwd<T>(L, total_blocks, max_plane, st);
  const size_t smem = smem_bytes(max_plane);
  auto kern = pool_tma_kernel<T, MODE>;
  static size_t attr = 0;
  if (smem > attr) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  pmv_launch(kern, (unsigned)total_blocks, threads_for(MODE), smem, st, L);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

}  // namespace

bool tma_fwd_writes_onehot() { return ws_enabled(); }

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
               float eps, int dtype, cudaStream_t st, cudaStream_t st_strided) {
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
      int rc = dtype == PMV_BF16 ? launch_din_strided<bf16>(jobs, njobs, L, esz, st_strided)
                                 : launch_din_strided<float>(jobs, njobs, L, esz, st_strided);
      if (rc) return rc;
    }
    if (n1 == 0) return PMV_OK;
    if (n1 < njobs) return tma_launch(3, s1, n1, B, heads, T, H, W, bs, ts, hs, eps, dtype, st, st);
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

#ifdef PMV_ATTN_TRACE
extern "C" int pmv_debug_pool_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, pool::pmv_pool_trace_buf, sizeof(long long) * 4 * pool::PTRACE_CTAS * pool::PTRACE_SLOTS) == cudaSuccess ? 0 : 1;
}
#endif
