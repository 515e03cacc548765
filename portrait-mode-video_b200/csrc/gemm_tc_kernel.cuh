// tcgen05 / TMEM / TMA GEMM for the Linear layers of the MViTv2 block (bf16 operands, fp32 accumulate):
//   TN        y = x W^T            (qkv, proj, skip-proj, fc1, fc2, PatchEmbed:  attention.py:328,457,570; common.py:27-31)
//   NN        dx = dy W            (dgrad)
//   REDUCE_M  dW = dy^T x          (wgrad; reduction over the token rows, optional split-K with fp32 atomics)
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, 128-byte swizzle, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma 128 x BN x 16; accumulators double-buffered in TMEM)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b -> smem transpose -> bias / GELU / DropPath scale / residual -> global)
//
// Operand staging.  A "K-major" operand (reduction axis contiguous in global memory) is one TMA box of
// [rows x 64 elements] -> rows of 128 B, 8-row swizzle atoms of 1024 B (SBO = 1024).  An "MN-major" operand
// (row axis contiguous: W in dgrad, both operands in wgrad) is staged as 64-wide groups, each one TMA box of
// [64 reduction rows x 64 elements]; within a group 8 reduction rows form a 1024 B atom (SBO = 1024) and the
// groups are 8192 B apart (LBO = 8192).  The UMMA reads either through its matrix descriptor; the a_major /
// b_major bits of the instruction descriptor select the transposed read.
//
// Epilogue kinds (compile time, so that each kernel carries only its own epilogue code: the all-in-one epilogue
// was 116 KB of SASS, three times the instruction cache, and its warps stalled on instruction fetch):
//   EK_PLAIN     out = acc (+ bias)
//   EK_GELU      out = gelu_erf(acc + bias), optionally the pre-activation to aux_out
//   EK_GELU_BWD  out = acc * gelu_erf'(aux_in)
//   EK_ACCUM     out(bf16) += acc            (rel-pos bias gradient added into dQ'; the previous tile arrives by TMA load)
//   EK_RES       out(fp32) = residual + [row_scale *] (acc + bias)
//   EK_ATOMIC    out(fp32) += acc   (split-K wgrad partials: one 16-byte vector reduction per lane)
//   EK_GENERIC   every pmv_epilogue term at run time (accumulate, row remap, ...)
#pragma once
#include "gemm.h"
#include "tc_common.cuh"

namespace gemm_tc {

enum { EK_PLAIN = 0, EK_GELU = 1, EK_GELU_BWD = 2, EK_RES = 3, EK_GENERIC = 4, EK_ATOMIC = 5, EK_ACCUM = 6 };

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 4;
constexpr int EPI_WARPS = 8;                        // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;

struct TcParams {
  int64_t M, N, K;          // logical output rows / cols and reduction length
  int tiles_m, tiles_n, splits;
  int64_t k_per_split;      // multiple of BK
  EpiDev e;
#ifdef PMV_ATTN_TRACE
  long long* trace;  // debug build (scripts/gemm_trace.py): [CTA][GT_ITEMS][GT_SLOTS] clock64 stamps
#endif
};
#ifdef PMV_ATTN_TRACE
constexpr int GT_CTAS = 148, GT_ITEMS = 12, GT_SLOTS = 8;
#define GTRACE(item, slot)                                                                                             \
  do {                                                                                                                 \
    if (p.trace != nullptr && (item) < GT_ITEMS && blockIdx.x < GT_CTAS)                                               \
      p.trace[((int64_t)blockIdx.x * GT_ITEMS + (item)) * GT_SLOTS + (slot)] = clock64();                              \
  } while (0)
#else
#define GTRACE(item, slot) do { } while (0)
#endif

// kinds whose epilogue goes registers -> swizzled smem tile -> TMA store (thread = accumulator row)
__host__ __device__ constexpr bool kind_tma(int kind) { return kind == EK_PLAIN || kind == EK_GELU || kind == EK_GELU_BWD || kind == EK_ACCUM; }

template <int BN, int KIND = EK_GENERIC, bool PAIR = false> struct TileCfg {
  static constexpr int BN_GROUPS = (BN + 63) / 64;
  static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB either layout
  static constexpr int B_BYTES_K = BN * BK * 2;               // K-major box
  static constexpr int B_BYTES_MN = BN_GROUPS * 64 * BK * 2;  // MN-major groups
  static constexpr int B_BYTES = B_BYTES_MN;                  // reserve the larger of the two
  // a CTA of a pair stages only half of the B tile: smaller stages, more of them in flight
  static constexpr int STAGE_BYTES = A_BYTES + (((PAIR ? B_BYTES_K / 2 : B_BYTES) + 1023) / 1024) * 1024;
  static constexpr int ACC_COLS = BN <= 128 ? 128 : 256;      // TMEM columns per accumulator stage
  // per-epilogue-warp staging: 4 KB transpose buffer / one fp32 or two bf16 32x32 output tiles; the GELU kinds keep two
  // tiles per chunk (output + pre-activation), double-buffered: 8 KB
  static constexpr int EPI_WARP_BYTES = (KIND == EK_GELU || KIND == EK_GELU_BWD || KIND == EK_ACCUM) ? 8192 : 4096;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int FIT_STAGES = (227 * 1024 - 2048 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = PAIR ? (FIT_STAGES > 8 ? 8 : FIT_STAGES) : (FIT_STAGES >= MAX_STAGES ? MAX_STAGES : MAX_STAGES - 1);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + EPI_BYTES;
};

// ---- packed-fp32 epilogue math -----------------------------------------------------------------------------------
// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) = 0.5 + xc * Q(xc^2), xc = clamp(x, +-3 sqrt 2): degree-8 polynomial in x^2,
// |error| <= 1.1e-5 (oracle/fit_gelu.py), no MUFU: the exp/rcp form cost 2 MUFU + 12 FMA per element and made the
// GELU epilogues the longest stage of the fc1 GEMMs.
__device__ __forceinline__ float2 phi2(float2 x) {
  const float Z = 4.242640687f;
  float2 xc = make_float2(fminf(fmaxf(x.x, -Z), Z), fminf(fmaxf(x.y, -Z), Z));
  const float2 s = __fmul2_rn(xc, xc);
  float2 q = make_float2(5.6236895431e-11f, 5.6236895431e-11f);
  q = __ffma2_rn(q, s, make_float2(-5.3744284878e-09f, -5.3744284878e-09f));
  q = __ffma2_rn(q, s, make_float2(2.2710010238e-07f, 2.2710010238e-07f));
  q = __ffma2_rn(q, s, make_float2(-5.6547267380e-06f, -5.6547267380e-06f));
  q = __ffma2_rn(q, s, make_float2(9.3721011908e-05f, 9.3721011908e-05f));
  q = __ffma2_rn(q, s, make_float2(-1.1104664642e-03f, -1.1104664642e-03f));
  q = __ffma2_rn(q, s, make_float2(9.8226745766e-03f, 9.8226745766e-03f));
  q = __ffma2_rn(q, s, make_float2(-6.6355885986e-02f, -6.6355885986e-02f));
  q = __ffma2_rn(q, s, make_float2(3.9890877892e-01f, 3.9890877892e-01f));
  return __ffma2_rn(xc, q, make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ float2 gelu2(float2 x) { return __fmul2_rn(x, phi2(x)); }
// gelu'(x) = Phi(x) + x * pdf(x)
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  const float2 cdf = phi2(x);
  const float2 t = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  const float2 e = make_float2(tc::fast_ex2(t.x), tc::fast_ex2(t.y));  // exp(-x^2 / 2); exp2f() adds a 3-instruction range fix-up per call
  const float2 xp = __fmul2_rn(x, make_float2(0.39894228040143267794f, 0.39894228040143267794f));
  return __ffma2_rn(xp, e, cdf);
}
__device__ __forceinline__ void st4(bf16* p, float2 a, float2 b) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(b.x, b.y);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&lo);
  t.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ void st4(float* p, float2 a, float2 b) { *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

// PAIR: launched as clusters of two CTAs; a work item is a 256 x BN tile computed by ONE cta_group::2 MMA stream issued
// by the leader CTA (rank 0).  Each CTA stages its own 128 rows of A and half of the B tile, owns the accumulator of its
// 128 rows in its own TMEM and runs its own epilogue.  Implemented for the TN (forward) layout.
template <int BN, bool A_MN, bool B_MN, typename TOut, int KIND, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, TcParams p) {
  // tmC: output (32 x 32 boxes, swizzled); tmD: aux_out (EK_GELU) / aux_in (EK_GELU_BWD); unused by the other kinds
  using Cfg = TileCfg<BN, KIND, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* aux_bar = bars + 2 * STAGES + 6;  // [EPI_WARPS][2]  EK_GELU_BWD: pre-activation tiles landed
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 1024;  // 1024-aligned: the swizzle patterns are address based
  float* epi_buf = reinterpret_cast<float*>(epi_smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(!PAIR || (!A_MN && !B_MN), "CTA pairs: TN layout only");
  const uint32_t cta_rank = PAIR ? tc::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int BM_TILE = PAIR ? 2 * BM : BM;                 // rows of a work item
  // work items are indexed in 32 bits (the host checks the count): the tile decode runs once per 32-column chunk in
  // every epilogue thread and once per tile in the single TMA / MMA threads - 64-bit div / mod there cost more than
  // the chunk's own arithmetic
  const uint32_t work0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;  // first work item / stride of this CTA (pair)
  const uint32_t work_step = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // PAIR: both CTAs' epilogue warps arrive at the leader
    }
    for (int i = 0; i < 2 * EPI_WARPS; ++i) tc::mbar_init(&aux_bar[i], 1);
    if (kind_tma(KIND)) {
      tc::tma_prefetch_desc(&tmC);
      if (KIND != EK_PLAIN) tc::tma_prefetch_desc(&tmD);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tc::tmem_alloc2(tmem_slot, 2 * Cfg::ACC_COLS);
      tc::tmem_relinquish2();
    } else {
      tc::tmem_alloc(tmem_slot, 2 * Cfg::ACC_COLS);
      tc::tmem_relinquish();
    }
  }
  tc::tc_fence_before();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();  // PAIR: the peer's barriers must exist before anyone signals them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlaps the tail of the previous kernel; nothing global has been touched yet

  const uint32_t num_work = (uint32_t)p.tiles_m * (uint32_t)p.tiles_n * (uint32_t)p.splits;
  const uint32_t tiles_n = (uint32_t)p.tiles_n, tiles_m = (uint32_t)p.tiles_m, tiles_mn = tiles_n * tiles_m;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_bar0_cl = PAIR ? tc::mapa_u32(&full_bar[0], 0) : 0u;  // the leader's full barriers
      for (uint32_t wi = work0; wi < num_work; wi += work_step) {
        const int tn = (int)(wi % tiles_n);
        const int tm = (int)((wi / tiles_n) % tiles_m);
        const int sp = (int)(wi / tiles_mn);
        const int m0 = tm * BM_TILE + (int)cta_rank * BM, n0 = tn * BN;
        const int64_t kbeg = (int64_t)sp * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        for (int64_t kb = kbeg; kb < kend; kb += BK) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (kb == kbeg) GTRACE((wi - work0) / work_step, 3);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if constexpr (PAIR) {
            // both CTAs load into their own smem and credit the LEADER's barrier; the leader expects both halves
            if (leader) tc::mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES_K / 2));
            const uint32_t fb = full_bar0_cl + stage * 8;
            tc::tma_load_2d_2sm(sa, &tmA, (int)kb, m0, fb);
            tc::tma_load_2d_2sm(sb, &tmB, (int)kb, n0 + (int)cta_rank * (BN / 2), fb);  // box of BN / 2 rows
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          tc::mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + (B_MN ? Cfg::B_BYTES_MN : Cfg::B_BYTES_K));
          if (!A_MN) {
            tc::tma_load_2d(sa, &tmA, (int)kb, m0, &full_bar[stage]);
          } else {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) tc::tma_load_2d(sa + g * 8192, &tmA, m0 + g * 64, (int)kb, &full_bar[stage]);
          }
          if (!B_MN) {
            tc::tma_load_2d(sb, &tmB, (int)kb, n0, &full_bar[stage]);
          } else {
#pragma unroll
            for (int g = 0; g < Cfg::BN_GROUPS; ++g) tc::tma_load_2d(sb + g * 8192, &tmB, n0 + g * 64, (int)kb, &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        GTRACE((wi - work0) / work_step, 4);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (leader && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(BM_TILE, BN, A_MN, B_MN);
      // One thread issues every MMA; at BN = 96 an MMA is ~48 tensor cycles, so the issue loop itself must be short.
      // A descriptor's high word depends only on the layout and its low word is additive in the address: both operands'
      // low words are made once for stage 0, a stage adds STAGE_BYTES >> 4 and a 16-wide k-step 2 (K-major) or 128
      // (MN-major: 2048 B).
      const uint64_t proto_a = A_MN ? tc::make_smem_desc(tc::smem_u32(smem), 8192, 1024, tc::SWIZZLE_128B)
                                    : tc::make_smem_desc(tc::smem_u32(smem), 16, 1024, tc::SWIZZLE_128B);
      const uint64_t proto_b = B_MN ? tc::make_smem_desc(tc::smem_u32(smem) + Cfg::A_BYTES, 8192, 1024, tc::SWIZZLE_128B)
                                    : tc::make_smem_desc(tc::smem_u32(smem) + Cfg::A_BYTES, 16, 1024, tc::SWIZZLE_128B);
      const uint32_t hi_a = (uint32_t)(proto_a >> 32), hi_b = (uint32_t)(proto_b >> 32);
      const uint32_t lo_a0 = (uint32_t)proto_a, lo_b0 = (uint32_t)proto_b;
      constexpr uint32_t STEP_A = A_MN ? 128u : 2u, STEP_B = B_MN ? 128u : 2u, STAGE16 = (uint32_t)Cfg::STAGE_BYTES >> 4;
      auto mma = [&](uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t accumulate) {
        const uint64_t da = ((uint64_t)hi_a << 32) | lo_a, db = ((uint64_t)hi_b << 32) | lo_b;
        if (PAIR) tc::umma_ss2(tmem_d, da, db, idesc, accumulate);
        else tc::umma_ss(tmem_d, da, db, idesc, accumulate);
      };
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (uint32_t wi = work0; wi < num_work; wi += work_step, ++it) {
        const int sp = (int)(wi / tiles_mn);
        const int64_t kbeg = (int64_t)sp * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        const int as = (int)(it & 1);
        const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&acc_empty[as], aphase ^ 1);
        tc::tc_fence_after();
        GTRACE(it, 0);
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_COLS;
        const int nkb = (int)((kend - kbeg + BK - 1) / BK);
        const int tail = (int)(kend - kbeg) - (nkb - 1) * BK;  // reduction length of the last block (1..BK)
        for (int kbi = 0; kbi < nkb; ++kbi) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          if (kbi == 0) GTRACE(it, 1);
          const uint32_t lo_a = lo_a0 + (uint32_t)stage * STAGE16, lo_b = lo_b0 + (uint32_t)stage * STAGE16;
          if (kbi + 1 < nkb || tail == BK) {
            mma(tmem_d, lo_a, lo_b, kbi > 0 ? 1u : 0u);
            mma(tmem_d, lo_a + STEP_A, lo_b + STEP_B, 1u);
            mma(tmem_d, lo_a + 2 * STEP_A, lo_b + 2 * STEP_B, 1u);
            mma(tmem_d, lo_a + 3 * STEP_A, lo_b + 3 * STEP_B, 1u);
          } else {
            const int nk = (tail + 15) / 16;
            for (int k = 0; k < nk; ++k) mma(tmem_d, lo_a + k * STEP_A, lo_b + k * STEP_B, (kbi > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage (PAIR: in both CTAs) once these MMAs have read it
          if (PAIR) tc::umma_commit2_mc(&empty_bar[stage], 3); else tc::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc::umma_commit2_mc(&acc_full[as], 3); else tc::umma_commit(&acc_full[as]);  // accumulator complete -> epilogue(s)
        GTRACE(it, 2);
      }
    }
  } else if constexpr (kind_tma(KIND)) {
    // ------------------------------------------------------------ epilogue warps, TMA-store form
    // thread = accumulator row (the TMEM lane), 32 consecutive columns per chunk in registers: bias / GELU in packed
    // fp32, 16-byte stores into a 32 x 32 tile in the swizzle pattern of the output tensor map (conflict-free:
    // 64B swizzle for bf16 rows of 64 B, 128B swizzle for fp32 rows of 128 B), then ONE bulk tensor store per
    // tile issued by lane 0.  No shared-memory read-back, no per-row global addressing, no bounds predicates (the
    // TMA clips rows >= M and columns >= N).  ~60 instructions per 32 x 32 chunk instead of ~450.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const EpiDev& e = p.e;
    uint8_t* wbuf = epi_smem + (warp - 2) * Cfg::EPI_WARP_BYTES;
    constexpr int OUT_TILE = 32 * 32 * (int)sizeof(TOut);           // 2 KB (bf16) / 4 KB (fp32)
    constexpr int OUT_BUFS = (KIND == EK_PLAIN) ? (4096 / OUT_TILE) : 2;
    // EK_GELU: out tiles at [0, 4 KB), aux tiles at [4 KB, 8 KB).  EK_GELU_BWD: out tiles at [0, 4 KB), aux_in at [4 KB, 8 KB)
    uint64_t* my_aux_bar = aux_bar + (warp - 2) * 2;
    const int row = lane;  // row inside the warp's 32-row slab
    // The chunk loop below ran 237 instructions per 32 x 32 chunk for ~55 of work (ncu source page of the qkv GEMM): 64-bit
    // bounds compares, predicated bias loads with zero fills, register copies around the optional bias add, 64-bit
    // shared-memory addressing.  With two epilogue warps per scheduler that IS the tile time of the K <= 384 GEMMs
    // (scripts/gemm_trace.py: the MMA warp waits for an accumulator stage), so everything here is 32-bit and branch-lean.
    const int Mi = (int)p.M, Ni = (int)p.N;  // the host checks that both fit 31 bits
    const uint32_t wbuf32 = tc::smem_u32(wbuf);
    auto sw16 = [&](int j) -> int {  // byte offset of 16-byte piece j of this thread's row inside a tile
      if (sizeof(TOut) == 2) return row * 64 + ((j ^ ((row >> 1) & 3)) << 4);
      return row * 128 + ((j ^ (row & 7)) << 4);
    };
    auto sw16_bf = [&](int j) -> int { return row * 64 + ((j ^ ((row >> 1) & 3)) << 4); };
    uint32_t so[sizeof(TOut) == 2 ? 4 : 8];  // this lane's 16-byte pieces inside an output tile (constant for the kernel)
#pragma unroll
    for (int j = 0; j < (sizeof(TOut) == 2 ? 4 : 8); ++j) so[j] = (uint32_t)sw16(j);

    // the tile decode (two 32-bit divisions) runs once per tile, not once per 32-column chunk
    struct Cursor {
      uint32_t wi, it;
      int c;
      int row0, colt;  // first row of this warp's slab, first column of the tile
    };
    auto valid = [&](const Cursor& cu) { return cu.wi < num_work; };
    auto decode = [&](Cursor& cu) {
      const int tn = (int)(cu.wi % tiles_n);
      const int tm = (int)((cu.wi / tiles_n) % tiles_m);
      cu.row0 = tm * BM_TILE + (int)cta_rank * BM + q * 32;
      cu.colt = tn * BN;
    };
    auto advance = [&](Cursor& cu) {
      cu.c += 64;
      if (cu.c >= BN) { cu.c = half * 32; cu.wi += work_step; ++cu.it; decode(cu); }
    };
    auto coords = [&](const Cursor& cu, int& row0, int& col0) {
      row0 = cu.row0;
      col0 = cu.colt + cu.c;
    };
    const uint32_t acc_empty0_cl = PAIR ? tc::mapa_u32(&acc_empty[0], 0) : 0u;
    auto release_acc = [&](int as) {  // lane 0: this warp has its part of accumulator stage `as` in registers
      if (PAIR) tc::mbar_arrive_cluster(acc_empty0_cl + as * 8); else tc::mbar_arrive(&acc_empty[as]);
    };
    int n_item = 0;    // live chunks stored so far (staging-buffer parity)
    int n_loaded = 0;  // EK_GELU_BWD: pre-activation tiles requested so far (the k-th request feeds the k-th live chunk)
    auto load_aux = [&](const Cursor& cu) {  // EK_GELU_BWD, warp-uniform: pre-activation tile of a live chunk -> aux buffer
      int row0, col0;
      coords(cu, row0, col0);
      if (col0 >= Ni || row0 >= Mi) return;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&my_aux_bar[n_loaded & 1], 2048);
        tc::tma_load_2d(wbuf + 4096 + (n_loaded & 1) * 2048, &tmD, col0, row0, &my_aux_bar[n_loaded & 1]);
      }
      ++n_loaded;
    };

    Cursor cur{work0, 0, half * 32, 0, 0};
    decode(cur);
    const bool has_bias = KIND != EK_GELU_BWD && KIND != EK_ACCUM && e.bias != nullptr;
    float4 bq[8];
    auto load_bias = [&](const Cursor& cu) {
      const int col0 = cu.colt + cu.c;
      const float4* bp = reinterpret_cast<const float4*>(e.bias + col0);
      if (col0 + 32 <= Ni) {  // whole slice inside the row (every GEMM of the models): eight plain loads
#pragma unroll
        for (int g = 0; g < 8; ++g) bq[g] = __ldg(bp + g);
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) bq[g] = (col0 + 4 * g < Ni) ? __ldg(bp + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (has_bias && cur.c < BN && valid(cur)) load_bias(cur);
    if (cur.c < BN) {
      if ((KIND == EK_GELU_BWD || KIND == EK_ACCUM) && valid(cur)) load_aux(cur);
      while (valid(cur)) {
        Cursor nxt = cur;
        advance(nxt);
        int row0, col0;
        coords(cur, row0, col0);
        const int as = (int)(cur.it & 1);
        const bool first_chunk = cur.c == half * 32, last_chunk = cur.c + 64 >= BN;
        const bool live = col0 < Ni && row0 < Mi;
        if (KIND == EK_GELU_BWD || KIND == EK_ACCUM) {
          __syncwarp();  // everyone has consumed the aux buffer the next request overwrites (two requests back)
          if (valid(nxt)) load_aux(nxt);
        }
        // the chunk's bias slice was requested one chunk ago, right after the previous slice had been consumed (requested
        // next to its use, the FADD2 behind it waited ~0.1 us per chunk: scripts/gemm_trace.py)
        const bool use_bias = has_bias && live;
        if (first_chunk) {
          tc::mbar_wait(&acc_full[as], (uint32_t)((cur.it >> 1) & 1));
          tc::tc_fence_after();
          if (warp == 2 && lane == 0) GTRACE(cur.it, 5);
        }
        uint32_t r[32];
        if (live) {
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::ACC_COLS + cur.c, r);
          tc::tmem_ld_wait();
        }
        if (last_chunk) {  // the accumulator stage is in registers: hand it back to the MMA warp
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(as);
          if (warp == 2 && lane == 0) GTRACE(cur.it, 6);
        }
        if (live) {
          if (use_bias) {  // in place: a separate result array costs 32 register copies on the path without bias
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float2 a = __fadd2_rn(make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])), make_float2(bq[g].x, bq[g].y));
              const float2 b = __fadd2_rn(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), make_float2(bq[g].z, bq[g].w));
              r[4 * g] = __float_as_uint(a.x); r[4 * g + 1] = __float_as_uint(a.y);
              r[4 * g + 2] = __float_as_uint(b.x); r[4 * g + 3] = __float_as_uint(b.y);
            }
          }
          if (has_bias && valid(nxt)) load_bias(nxt);
          float2 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          const int ob = n_item % OUT_BUFS;
          uint8_t* otile = wbuf + ob * OUT_TILE;
          // the bulk store that last read this buffer must be done with it
          if (tc::elect_one()) bulk_wait_read<OUT_BUFS - 1>();
          __syncwarp();
          if (KIND == EK_GELU && e.aux_out) {
            const uint32_t atile32 = wbuf32 + 4096u + (uint32_t)(ob * 2048);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts16(atile32 + (sizeof(TOut) == 2 ? so[j] : (uint32_t)sw16_bf(j)), pack_bf2(v[4 * j]), pack_bf2(v[4 * j + 1]), pack_bf2(v[4 * j + 2]),
                    pack_bf2(v[4 * j + 3]));
          }
          if (KIND == EK_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu2(v[i]);
          }
          if (KIND == EK_ACCUM) {
            tc::mbar_wait(&my_aux_bar[n_item & 1], (uint32_t)((n_item >> 1) & 1));
            const uint8_t* atile = wbuf + 4096 + (n_item & 1) * 2048;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = *reinterpret_cast<const uint4*>(atile + sw16_bf(j));
              v[4 * j] = __fadd2_rn(v[4 * j], bf2_to_f2(u.x));
              v[4 * j + 1] = __fadd2_rn(v[4 * j + 1], bf2_to_f2(u.y));
              v[4 * j + 2] = __fadd2_rn(v[4 * j + 2], bf2_to_f2(u.z));
              v[4 * j + 3] = __fadd2_rn(v[4 * j + 3], bf2_to_f2(u.w));
            }
          }
          if (KIND == EK_GELU_BWD) {
            tc::mbar_wait(&my_aux_bar[n_item & 1], (uint32_t)((n_item >> 1) & 1));
            const uint8_t* atile = wbuf + 4096 + (n_item & 1) * 2048;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = *reinterpret_cast<const uint4*>(atile + sw16_bf(j));
              v[4 * j] = __fmul2_rn(v[4 * j], gelu_grad2(bf2_to_f2(u.x)));
              v[4 * j + 1] = __fmul2_rn(v[4 * j + 1], gelu_grad2(bf2_to_f2(u.y)));
              v[4 * j + 2] = __fmul2_rn(v[4 * j + 2], gelu_grad2(bf2_to_f2(u.z)));
              v[4 * j + 3] = __fmul2_rn(v[4 * j + 3], gelu_grad2(bf2_to_f2(u.w)));
            }
          }
          const uint32_t otile32 = wbuf32 + (uint32_t)(ob * OUT_TILE);
          if (sizeof(TOut) == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts16(otile32 + so[j], pack_bf2(v[4 * j]), pack_bf2(v[4 * j + 1]), pack_bf2(v[4 * j + 2]), pack_bf2(v[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts16(otile32 + so[j], __float_as_uint(v[2 * j].x), __float_as_uint(v[2 * j].y), __float_as_uint(v[2 * j + 1].x),
                    __float_as_uint(v[2 * j + 1].y));
          }
          tc::fence_proxy_async();  // generic-proxy writes -> visible to the bulk store
          __syncwarp();
          if (tc::elect_one()) {
            tma_store_2d(&tmC, otile, col0, row0);
            if (KIND == EK_GELU && e.aux_out) tma_store_2d(&tmD, wbuf + 4096 + ob * 2048, col0, row0);
            bulk_commit();
          }
          ++n_item;
        } else if (has_bias && valid(nxt)) {
          load_bias(nxt);
        }
        if (last_chunk && warp == 2 && lane == 0) GTRACE(cur.it, 7);
        cur = nxt;
      }
      if (tc::elect_one()) bulk_wait_all();  // global writes complete before the CTA retires its shared memory
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    // Work items are (tile, 32-column chunk); the warp walks them as one flat sequence.  While the accumulator of
    // item i is on its way out of TMEM, the global operands of item i+1 (residual, GELU' pre-activation, DropPath
    // scale) are already being fetched into a second register set, so the epilogue is not exposed to a DRAM
    // round trip per chunk.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // which of the two warps of this lane quarter
    float* stage_buf = epi_buf + (warp - 2) * (32 * 32);
    const int lr = lane >> 3, lc = lane & 7;
    const EpiDev& e = p.e;
    constexpr bool PRE_RES = KIND == EK_RES, PRE_AUX = KIND == EK_GELU_BWD;
    const bool has_scale = (KIND == EK_RES || KIND == EK_GENERIC) && e.row_scale != nullptr;

    // 32-bit indexing, and the tile decode (two divisions) once per tile: the loop below used to spend most of its ~600
    // instructions per chunk on 64-bit row / column arithmetic, re-read constant-bank fields and per-chunk decodes
    const int Mi = (int)p.M, Ni = (int)p.N;  // the host checks M, N < 2^31 and rows x leading dimension < 2^31 elements
    const uint32_t ld_res = (uint32_t)e.ld_residual, ld_aux = (uint32_t)e.ld_aux, ldo = (uint32_t)e.ldo;
    struct Cursor {
      uint32_t wi, it;
      int c;
      int row0, colt;  // first row this lane handles in the tile (q, lr included), first column of the tile
    };
    struct Pre {     // operands of one chunk fetched ahead: fp32 residual (16 B) or bf16 pre-activation (8 B) per itr
      uint4 buf[(PRE_RES || PRE_AUX) ? 8 : 1];
      float sc[KIND == EK_RES ? 8 : 1];
    };
    auto valid = [&](const Cursor& cu) { return cu.wi < num_work; };
    auto decode = [&](Cursor& cu) {
      const int tn = (int)(cu.wi % tiles_n);
      const int tm = (int)((cu.wi / tiles_n) % tiles_m);
      cu.row0 = tm * BM_TILE + (int)cta_rank * BM + q * 32 + lr;
      cu.colt = tn * BN;
    };
    auto advance = [&](Cursor& cu) {
      cu.c += 64;
      if (cu.c >= BN) { cu.c = half * 32; cu.wi += work_step; ++cu.it; decode(cu); }
    };
    auto coords = [&](const Cursor& cu, int& row0, int& col) {
      row0 = cu.row0;
      col = cu.colt + cu.c + lc * 4;
    };
    const uint32_t acc_empty0_cl = PAIR ? tc::mapa_u32(&acc_empty[0], 0) : 0u;
    auto prefetch = [&](const Cursor& cu, Pre& pre) {
      if constexpr (PRE_RES || PRE_AUX) {
        if (!valid(cu)) return;
        int row0, col;
        coords(cu, row0, col);
        const bool colok = col < Ni;
        if constexpr (PRE_RES) {
          const float* rp = e.residual + (size_t)((uint32_t)row0 * ld_res + (uint32_t)col);
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int row = row0 + itr * 4;
            const bool ok = colok && row < Mi;
            pre.buf[itr] = ok ? __ldg(reinterpret_cast<const uint4*>(rp + (size_t)((uint32_t)(itr * 4) * ld_res))) : make_uint4(0u, 0u, 0u, 0u);
            pre.sc[itr] = (ok && has_scale) ? __ldg(e.row_scale + e.fd_scale.div((uint32_t)row)) : 1.f;
          }
        } else {
          const bf16* ap = reinterpret_cast<const bf16*>(e.aux_in) + (size_t)((uint32_t)row0 * ld_aux + (uint32_t)col);
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const bool ok = colok && row0 + itr * 4 < Mi;
            const uint2 t2 = ok ? __ldg(reinterpret_cast<const uint2*>(ap + (size_t)((uint32_t)(itr * 4) * ld_aux))) : make_uint2(0u, 0u);
            pre.buf[itr].x = t2.x; pre.buf[itr].y = t2.y;
          }
        }
      }
    };
    auto process = [&](const Cursor& cu, const Pre& pre) {
      int row0, col;
      coords(cu, row0, col);
      const int as = (int)(cu.it & 1);
      const bool first_chunk = cu.c == half * 32, last_chunk = cu.c + 64 >= BN;
      if (first_chunk) {
        tc::mbar_wait(&acc_full[as], (uint32_t)((cu.it >> 1) & 1));
        tc::tc_fence_after();
        if (warp == 2 && lane == 0) GTRACE(cu.it, 5);
      }
      uint32_t r[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::ACC_COLS + cu.c, r);
      tc::tmem_ld_wait();
      if (last_chunk) {  // the accumulator stage is in registers: hand it back to the MMA warp
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) tc::mbar_arrive_cluster(acc_empty0_cl + as * 8); else tc::mbar_arrive(&acc_empty[as]);
        }
        if (warp == 2 && lane == 0) GTRACE(cu.it, 6);
      }
      // transpose through shared memory (XOR-swizzled 16-byte groups: conflict-free both ways) so that 8
      // consecutive lanes cover one 32-column row segment: every global access of the epilogue is a full
      // 64/128-byte run per row
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(stage_buf + lane * 32 + ((g ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
      __syncwarp();
      const bool colok = col < Ni;
      const uint32_t off0 = (uint32_t)row0 * ldo + (uint32_t)col;  // element offset of (row0, col) in the output
      float2 b01 = make_float2(0.f, 0.f), b23 = b01;
      if (KIND != EK_GELU_BWD && KIND != EK_ATOMIC && e.bias && colok) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
        b01 = make_float2(b4.x, b4.y); b23 = make_float2(b4.z, b4.w);
      }
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int rr = itr * 4 + lr;
        const int row = row0 + itr * 4;
        const float4 a4 = *reinterpret_cast<const float4*>(stage_buf + rr * 32 + ((lc ^ (rr & 7)) << 2));
        if (!(colok && row < Mi)) continue;
        const size_t off = (size_t)(off0 + (uint32_t)(itr * 4) * ldo);  // KIND != GENERIC: plain row-major output
        float2 v01 = make_float2(a4.x, a4.y), v23 = make_float2(a4.z, a4.w);
        if constexpr (KIND == EK_ATOMIC) {
          red_add4(reinterpret_cast<float*>(e.out) + off, a4);
        } else if constexpr (KIND == EK_PLAIN) {
          v01 = __fadd2_rn(v01, b01); v23 = __fadd2_rn(v23, b23);
          st4(reinterpret_cast<TOut*>(e.out) + off, v01, v23);
        } else if constexpr (KIND == EK_GELU) {
          v01 = __fadd2_rn(v01, b01); v23 = __fadd2_rn(v23, b23);
          if (e.aux_out) st4(reinterpret_cast<bf16*>(e.aux_out) + (size_t)((uint32_t)row * ld_aux + (uint32_t)col), v01, v23);
          st4(reinterpret_cast<TOut*>(e.out) + off, gelu2(v01), gelu2(v23));
        } else if constexpr (KIND == EK_GELU_BWD) {
          v01 = __fmul2_rn(v01, gelu_grad2(bf2_to_f2(pre.buf[itr].x)));
          v23 = __fmul2_rn(v23, gelu_grad2(bf2_to_f2(pre.buf[itr].y)));
          st4(reinterpret_cast<TOut*>(e.out) + off, v01, v23);
        } else if constexpr (KIND == EK_RES) {
          const float2 sc = make_float2(pre.sc[itr], pre.sc[itr]);
          const float2 r01 = make_float2(__uint_as_float(pre.buf[itr].x), __uint_as_float(pre.buf[itr].y));
          const float2 r23 = make_float2(__uint_as_float(pre.buf[itr].z), __uint_as_float(pre.buf[itr].w));
          v01 = __ffma2_rn(__fadd2_rn(v01, b01), sc, r01);
          v23 = __ffma2_rn(__fadd2_rn(v23, b23), sc, r23);
          st4(reinterpret_cast<TOut*>(e.out) + off, v01, v23);
        } else {
          float v[4] = {a4.x + b01.x, a4.y + b01.y, a4.z + b23.x, a4.w + b23.y};
          if (e.atomic) {
            float* o = reinterpret_cast<float*>(e.out) + row * e.ldo + col;
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(o + j, v[j]);
            continue;
          }
          if (e.aux_out) store4(reinterpret_cast<bf16*>(e.aux_out) + row * e.ld_aux + col, v);
          if (e.act == PMV_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_fast(v[j]);
          } else if (e.act == PMV_ACT_GELU_BWD) {
            float u[4];
            load4(reinterpret_cast<const bf16*>(e.aux_in) + row * e.ld_aux + col, u);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] *= gelu_fast_grad(u[j]);
          }
          if (has_scale) {
            const float s = e.row_scale[e.fd_scale.div((uint32_t)row)];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] *= s;
          }
          if (e.residual) {
            float rv[4];
            load4(e.residual + row * e.ld_residual + col, rv);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += rv[j];
          }
          const int64_t orow = e.out_group > 0 ? (int64_t)row + (int64_t)(e.fd_group.div((uint32_t)row) + 1) * e.out_skip : (int64_t)row;
          TOut* o = reinterpret_cast<TOut*>(e.out) + orow * e.ldo + col;
          if (e.accumulate) {
            float pv[4];
            load4(o, pv);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += pv[j];
          }
          store4(o, v);
        }
      }
      __syncwarp();
      if (last_chunk && warp == 2 && lane == 0) GTRACE(cu.it, 7);
    };

    Cursor cur{work0, 0, half * 32, 0, 0};
    decode(cur);
    if (cur.c < BN) {
      Pre pa, pb;
      prefetch(cur, pa);
      while (valid(cur)) {
        Cursor nxt = cur;
        advance(nxt);
        prefetch(nxt, pb);
        process(cur, pa);
        pa = pb;
        cur = nxt;
      }
    }
  }

  tc::tc_fence_before();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();  // PAIR: nobody leaves while the peer may still signal its barriers
  if (warp == 1) {
    tc::tc_fence_after();
    if (PAIR) tc::tmem_dealloc2(tmem_base, 2 * Cfg::ACC_COLS); else tc::tmem_dealloc(tmem_base, 2 * Cfg::ACC_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN, typename TOut, int KIND, bool PAIR = false>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmD, const TcParams& p,
               int num_sms, cudaStream_t stream) {
  using Cfg = TileCfg<BN, KIND, PAIR>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TOut, KIND, PAIR>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int64_t work = (int64_t)p.tiles_m * p.tiles_n * p.splits;
  if (PAIR) {
    const int pairs = num_sms / 2;
    const unsigned grid = 2u * (unsigned)(work < pairs ? work : pairs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PMV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmD, p));
    return PMV_OK;
  }
  unsigned grid = (unsigned)(work < num_sms ? work : num_sms);
  pmv_launch(kern, grid, NUM_THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, tmC, tmD, p);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

// CTA-pair kernels (TN layout): tmB must have been encoded with a box of BN / 2 rows, p.tiles_m counts 256-row tiles
template <int BN>
int launch_bn_pair(int out_dtype, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                   const CUtensorMap& tmD, const TcParams& p, int num_sms, cudaStream_t s) {
  const bool f32 = out_dtype == PMV_F32;
  if (kind == EK_PLAIN && !f32) return launch_cfg<BN, false, false, bf16, EK_PLAIN, true>(tmA, tmB, tmC, tmD, p, num_sms, s);
  if (kind == EK_GELU && !f32) return launch_cfg<BN, false, false, bf16, EK_GELU, true>(tmA, tmB, tmC, tmD, p, num_sms, s);
  if (kind == EK_RES && f32) return launch_cfg<BN, false, false, float, EK_RES, true>(tmA, tmB, tmC, tmD, p, num_sms, s);
  pmv_set_error("gemm(tc): no CTA-pair kernel for this epilogue");
  return PMV_ERR_UNSUPPORTED;
}

// (layout, output type, epilogue kind) combinations that exist as kernels; anything else runs EK_GENERIC
template <int BN>
int launch_bn(int layout, int out_dtype, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
              const CUtensorMap& tmD, const TcParams& p, int num_sms, cudaStream_t s) {
  const bool f32 = out_dtype == PMV_F32;
  if (layout == PMV_GEMM_TN) {
    if (kind == EK_PLAIN) return f32 ? launch_cfg<BN, false, false, float, EK_PLAIN>(tmA, tmB, tmC, tmD, p, num_sms, s)
                                     : launch_cfg<BN, false, false, bf16, EK_PLAIN>(tmA, tmB, tmC, tmD, p, num_sms, s);
    if (kind == EK_GELU && !f32) return launch_cfg<BN, false, false, bf16, EK_GELU>(tmA, tmB, tmC, tmD, p, num_sms, s);
    if (kind == EK_RES && f32) return launch_cfg<BN, false, false, float, EK_RES>(tmA, tmB, tmC, tmD, p, num_sms, s);
    return f32 ? launch_cfg<BN, false, false, float, EK_GENERIC>(tmA, tmB, tmC, tmD, p, num_sms, s)
               : launch_cfg<BN, false, false, bf16, EK_GENERIC>(tmA, tmB, tmC, tmD, p, num_sms, s);
  }
  if (layout == PMV_GEMM_NN) {
    if (kind == EK_PLAIN && !f32) return launch_cfg<BN, false, true, bf16, EK_PLAIN>(tmA, tmB, tmC, tmD, p, num_sms, s);
    if (kind == EK_GELU_BWD && !f32) return launch_cfg<BN, false, true, bf16, EK_GELU_BWD>(tmA, tmB, tmC, tmD, p, num_sms, s);
    if (kind == EK_ACCUM && !f32) return launch_cfg<BN, false, true, bf16, EK_ACCUM>(tmA, tmB, tmC, tmD, p, num_sms, s);
    return f32 ? launch_cfg<BN, false, true, float, EK_GENERIC>(tmA, tmB, tmC, tmD, p, num_sms, s)
               : launch_cfg<BN, false, true, bf16, EK_GENERIC>(tmA, tmB, tmC, tmD, p, num_sms, s);
  }
  if (kind == EK_PLAIN) return launch_cfg<BN, true, true, float, EK_PLAIN>(tmA, tmB, tmC, tmD, p, num_sms, s);
  if (kind == EK_ATOMIC) return launch_cfg<BN, true, true, float, EK_ATOMIC>(tmA, tmB, tmC, tmD, p, num_sms, s);
  return launch_cfg<BN, true, true, float, EK_GENERIC>(tmA, tmB, tmC, tmD, p, num_sms, s);
}

}  // namespace gemm_tc
