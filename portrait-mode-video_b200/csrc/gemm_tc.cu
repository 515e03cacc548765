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
#include "gemm.h"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;                        // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;

struct TcParams {
  int64_t M, N, K;          // logical output rows / cols and reduction length
  int tiles_m, tiles_n, splits;
  int64_t k_per_split;      // multiple of BK
  EpiDev e;
};

template <int BN> struct TileCfg {
  static constexpr int BN_GROUPS = (BN + 63) / 64;
  static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB either layout
  static constexpr int B_BYTES_K = BN * BK * 2;               // K-major box
  static constexpr int B_BYTES_MN = BN_GROUPS * 64 * BK * 2;  // MN-major groups
  static constexpr int B_BYTES = B_BYTES_MN;                  // reserve the larger of the two
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int ACC_COLS = BN <= 128 ? 128 : 256;      // TMEM columns per accumulator stage
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * 33 * 4;           // per-epilogue-warp transpose buffer (padded rows)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
};

template <int BN, bool A_MN, bool B_MN, typename TOut>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* epi_buf = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 2 * Cfg::ACC_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t num_work = (int64_t)p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        const int tn = (int)(wi % p.tiles_n);
        const int tm = (int)((wi / p.tiles_n) % p.tiles_m);
        const int sp = (int)(wi / ((int64_t)p.tiles_n * p.tiles_m));
        const int m0 = tm * BM, n0 = tn * BN;
        const int64_t kbeg = (int64_t)sp * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        for (int64_t kb = kbeg; kb < kend; kb += BK) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
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
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t wi = blockIdx.x; wi < num_work; wi += gridDim.x, ++it) {
        const int sp = (int)(wi / ((int64_t)p.tiles_n * p.tiles_m));
        const int64_t kbeg = (int64_t)sp * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        const int as = (int)(it & 1);
        const uint32_t aphase = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&acc_empty[as], aphase ^ 1);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_COLS;
        for (int64_t kb = kbeg; kb < kend; kb += BK) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const int64_t rem = kend - kb;
          const int nk = rem >= BK ? BK / 16 : (int)((rem + 15) / 16);
#pragma unroll 4
          for (int k = 0; k < nk; ++k) {
            const uint64_t da = A_MN ? tc::make_smem_desc(sa + k * 2048, 8192, 1024, tc::SWIZZLE_128B)
                                     : tc::make_smem_desc(sa + k * 32, 16, 1024, tc::SWIZZLE_128B);
            const uint64_t db = B_MN ? tc::make_smem_desc(sb + k * 2048, 8192, 1024, tc::SWIZZLE_128B)
                                     : tc::make_smem_desc(sb + k * 32, 16, 1024, tc::SWIZZLE_128B);
            tc::umma_ss(tmem_d, da, db, idesc, (kb > kbeg || k > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&acc_full[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // which of the two warps of this lane quarter
    float* stage_buf = epi_buf + (warp - 2) * (32 * 33);
    int64_t it = 0;
    for (int64_t wi = blockIdx.x; wi < num_work; wi += gridDim.x, ++it) {
      const int tn = (int)(wi % p.tiles_n);
      const int tm = (int)((wi / p.tiles_n) % p.tiles_m);
      const int64_t row0 = (int64_t)tm * BM + q * 32;
      const int64_t n0 = (int64_t)tn * BN;
      const int as = (int)(it & 1);
      const uint32_t aphase = (uint32_t)((it >> 1) & 1);
      tc::mbar_wait(&acc_full[as], aphase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::ACC_COLS;
#pragma unroll 1
      for (int c = half * 32; c < BN; c += 64) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + c, r);
        tc::tmem_ld_wait();
        // transpose through shared memory so that 8 consecutive lanes cover one 32-column row segment:
        // every global access of the epilogue (out, residual, aux) is then a full 64/128-byte run per row
#pragma unroll
        for (int j = 0; j < 32; ++j) stage_buf[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int64_t col = n0 + c + (lane & 7) * 4;
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + (lane >> 3);
          const int64_t row = row0 + rr;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = stage_buf[rr * 33 + (lane & 7) * 4 + j];
          if (row < p.M && col < p.N) epi_store4<bf16, TOut, true>(p.e, row, col, v);
        }
        __syncwarp();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 2 * Cfg::ACC_COLS);
  }
}

int g_num_sms = 0;

template <int BN, bool A_MN, bool B_MN, typename TOut>
int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TOut>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    PMV_CHECK_CUDA(cudaGetDevice(&dev));
    PMV_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int64_t work = (int64_t)p.tiles_m * p.tiles_n * p.splits;
  unsigned grid = (unsigned)(work < g_num_sms ? work : g_num_sms);
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

template <int BN>
int launch_bn(int layout, int out_dtype, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  const bool f32 = out_dtype == PMV_F32;
  if (layout == PMV_GEMM_TN) return f32 ? launch_cfg<BN, false, false, float>(tmA, tmB, p, s) : launch_cfg<BN, false, false, bf16>(tmA, tmB, p, s);
  if (layout == PMV_GEMM_NN) return f32 ? launch_cfg<BN, false, true, float>(tmA, tmB, p, s) : launch_cfg<BN, false, true, bf16>(tmA, tmB, p, s);
  return launch_cfg<BN, true, true, float>(tmA, tmB, p, s);
}

int pick_bn(int64_t M, int64_t N, int splits) {
  // wave-quantisation model: cost = waves over 148 SMs x (tile width + fixed per-tile overhead in columns);
  // ties go to the wider tile (better operand reuse per byte staged)
  const int cands[4] = {256, 192, 128, 96};
  int best = 96;
  double best_cost = -1.0;
  const int64_t tiles_m = ceil_div64(M, BM);
  for (int i = 0; i < 4; ++i) {
    const int64_t tiles = tiles_m * ceil_div64(N, cands[i]) * splits;
    const int64_t waves = ceil_div64(tiles, 148);
    const double cost = (double)waves * (cands[i] + 48);
    if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best = cands[i]; }
  }
  return best;
}

}  // namespace

int gemm_tc_launch(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                   int out_dtype, const EpiDev& e, int split_k, cudaStream_t stream) {
  // logical output [MM x NN], reduction KK
  int64_t MM, NN, KK;
  if (layout == PMV_GEMM_NT_REDUCE_M) { MM = N; NN = K; KK = M; } else { MM = M; NN = N; KK = K; }
  PMV_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm(tc): leading dimensions must be multiples of 8 elements (16-byte TMA strides)");
  PMV_CHECK_ARG(NN % 4 == 0 && e.ldo % 4 == 0, "gemm(tc): N and ldo must be multiples of 4");
  PMV_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm(tc): operands must be 16-byte aligned");
  PMV_CHECK_ARG(layout != PMV_GEMM_NT_REDUCE_M || out_dtype == PMV_F32, "gemm(tc): wgrad output must be fp32");
  if (split_k < 1) split_k = 1;
  const int BN = pick_bn(MM, NN, split_k);
  CUtensorMap tmA, tmB;
  int rc;
  if (layout == PMV_GEMM_TN) {
    rc = pmv_make_tensor_map_2d(&tmA, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, 128);
    if (rc) return rc;
    rc = pmv_make_tensor_map_2d(&tmB, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)BN, 128);
  } else if (layout == PMV_GEMM_NN) {
    rc = pmv_make_tensor_map_2d(&tmA, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM, 128);
    if (rc) return rc;
    rc = pmv_make_tensor_map_2d(&tmB, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK, 128);  // B [K, N], N contiguous
  } else {
    rc = pmv_make_tensor_map_2d(&tmA, A, 2, (uint64_t)N, (uint64_t)M, (uint64_t)lda, 64, BK, 128);  // A [M, N]
    if (rc) return rc;
    rc = pmv_make_tensor_map_2d(&tmB, B, 2, (uint64_t)K, (uint64_t)M, (uint64_t)ldb, 64, BK, 128);  // B [M, K]
  }
  if (rc) return rc;
  TcParams p;
  p.M = MM; p.N = NN; p.K = KK;
  p.tiles_m = (int)ceil_div64(MM, BM);
  p.tiles_n = (int)ceil_div64(NN, BN);
  if (split_k < 1) split_k = 1;
  p.k_per_split = ceil_div64(ceil_div64(KK, split_k), BK) * BK;
  p.splits = (int)ceil_div64(KK, p.k_per_split);
  p.e = e;
  if (p.splits > 1) {
    PMV_CHECK_ARG(out_dtype == PMV_F32 && e.atomic, "gemm(tc): split-K needs the atomic fp32 epilogue");
  }
  switch (BN) {
    case 96: return launch_bn<96>(layout, out_dtype, tmA, tmB, p, stream);
    case 128: return launch_bn<128>(layout, out_dtype, tmA, tmB, p, stream);
    case 192: return launch_bn<192>(layout, out_dtype, tmA, tmB, p, stream);
    default: return launch_bn<256>(layout, out_dtype, tmA, tmB, p, stream);
  }
}
