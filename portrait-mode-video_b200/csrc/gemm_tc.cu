// Host side of the tcgen05 GEMM: tensor maps, tile-width choice, epilogue-kind choice.  The kernel template lives in
// gemm_tc_kernel.cuh and is instantiated per tile width in gemm_tc_bn*.cu.
#include <cstdlib>

#include "gemm_tc_kernel.cuh"

#ifdef PMV_ATTN_TRACE
// Debug build only (scripts/gemm_trace.py): per-CTA, per-work-item cycle stamps of the tcgen05 GEMM (GTRACE points).
static long long* g_gemm_trace = nullptr;
extern "C" int pmv_debug_gemm_trace(long long* host_out, int reset) {
  const size_t n = (size_t)gemm_tc::GT_CTAS * gemm_tc::GT_ITEMS * gemm_tc::GT_SLOTS * sizeof(long long);
  if (g_gemm_trace == nullptr && cudaMalloc(&g_gemm_trace, n) != cudaSuccess) return 1;
  if (host_out != nullptr && cudaMemcpy(host_out, g_gemm_trace, n, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  if (reset && cudaMemset(g_gemm_trace, 0, n) != cudaSuccess) return 1;
  return 0;
}
#endif


using namespace gemm_tc;

int gemm_tc_launch_bn96(int, int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                         const TcParams&, int, cudaStream_t);
int gemm_tc_launch_bn128(int, int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                         const TcParams&, int, cudaStream_t);
int gemm_tc_launch_bn192(int, int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                         const TcParams&, int, cudaStream_t);
int gemm_tc_launch_bn256(int, int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                         const TcParams&, int, cudaStream_t);

int gemm_tc_launch_pair_bn128(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const TcParams&,
                              int, cudaStream_t);
int gemm_tc_launch_pair_bn192(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const TcParams&,
                              int, cudaStream_t);
int gemm_tc_launch_pair_bn256(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const TcParams&,
                              int, cudaStream_t);

static int g_num_sms = 0;

static int pick_kind(const EpiDev& e) {
  if (e.atomic) return EK_ATOMIC;  // only instantiated for the wgrad layout; other layouts fall back to EK_GENERIC
  if (e.accumulate && e.out_group == 0 && !e.bias && !e.row_scale && !e.residual && !e.aux_out && e.act == PMV_ACT_NONE)
    return EK_ACCUM;  // only instantiated for NN / bf16 (the rel-pos dQ product); everything else falls back to EK_GENERIC
  if (e.accumulate || e.out_group > 0) return EK_GENERIC;
  const bool scale = e.row_scale != nullptr, res = e.residual != nullptr;
  if (e.act == PMV_ACT_GELU) return (!scale && !res) ? EK_GELU : EK_GENERIC;
  if (e.act == PMV_ACT_GELU_BWD) return (!scale && !res && !e.bias && !e.aux_out) ? EK_GELU_BWD : EK_GENERIC;
  if (e.aux_out) return EK_GENERIC;
  if (res) return EK_RES;
  return scale ? EK_GENERIC : EK_PLAIN;
}

static int pick_bn(int64_t M, int64_t N, int64_t K, int splits) {
  // Cost model in SM cycles, max of two rates plus a per-wave constant:
  //   tensor : waves over 148 SMs x k-blocks x 4 MMAs x BN / 2 cycles (a 128 x BN x 16 MMA)
  //   L2     : every column tile re-reads A, every row tile re-reads B; the GEMMs of this model saturate at ~9 TB/s of
  //            L2 -> SM traffic (4580 B per cycle): 128 x 96 tiles of the K = 1152 / 1536 dgrad and fc2 GEMMs measured exactly
  //            bytes / 9 TB/s (22 and 30 us), which the wave count alone (the previous model) could not see.
  // Ties go to the wider tile.
  const int cands[4] = {256, 192, 128, 96};
  int best = 96;
  double best_cost = -1.0;
  const int64_t tiles_m = ceil_div64(M, BM);
  const int64_t kb = ceil_div64(ceil_div64(K, splits), BK);
  for (int i = 0; i < 4; ++i) {
    const int64_t tiles_n = ceil_div64(N, cands[i]);
    const int64_t tiles = tiles_m * tiles_n * splits;
    const int64_t waves = ceil_div64(tiles, 148);
    const double tensor = (double)waves * (double)kb * 4.0 * (cands[i] / 2.0);
    const double bytes = 2.0 * (double)K * ((double)tiles_n * (double)M + (double)tiles_m * (double)N);
    const double l2 = bytes / 4580.0;
    const double cost = (tensor > l2 ? tensor : l2) + (double)waves * (cands[i] * 6.0 + 600.0);
    if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best = cands[i]; }
  }
  if (const char* ev = std::getenv("PMV_GEMM_BN")) {  // experiments: force a tile width
    const int f = atoi(ev);
    if (f == 96 || f == 128 || f == 192 || f == 256) best = f;
  }
  return best;
}

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
  const int BN = pick_bn(MM, NN, KK, split_k);
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
#ifdef PMV_ATTN_TRACE
  p.trace = g_gemm_trace;
#endif
  PMV_CHECK_ARG((int64_t)p.tiles_m * p.tiles_n * p.splits < (1ll << 31), "gemm(tc): too many tiles");
  PMV_CHECK_ARG(MM < (1ll << 31) - 512 && NN < (1ll << 31) - 512, "gemm(tc): M and N must fit 31 bits (the epilogues index in 32 bits)");
  {
    int64_t ld_max = e.ldo > e.ld_residual ? e.ldo : e.ld_residual;
    if (e.ld_aux > ld_max) ld_max = e.ld_aux;
    PMV_CHECK_ARG((MM + 256) * ld_max < (1ll << 32), "gemm(tc): rows x leading dimension must fit 32 bits (%lld x %lld)", (long long)MM, (long long)ld_max);
  }
  if (p.splits > 1) {
    PMV_CHECK_ARG(out_dtype == PMV_F32 && e.atomic, "gemm(tc): split-K needs the atomic fp32 epilogue");
  }
  if (g_num_sms == 0) {
    int dev = 0;
    PMV_CHECK_CUDA(cudaGetDevice(&dev));
    PMV_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int kind = pick_kind(e);
  // TMA-store epilogues: output (and pre-activation) tensor maps with 32 x 32 boxes in the swizzle the epilogue writes
  CUtensorMap tmC = tmA, tmD = tmA;
  if (kind_tma(kind)) {
    const int osz = out_dtype == PMV_F32 ? 4 : 2;
    const bool ok_out = ((uintptr_t)e.out & 15) == 0 && (e.ldo * osz) % 16 == 0;
    const void* aux = kind == EK_GELU ? e.aux_out : kind == EK_GELU_BWD ? e.aux_in : kind == EK_ACCUM ? e.out : nullptr;
    const int64_t ld_auxmap = kind == EK_ACCUM ? e.ldo : e.ld_aux;
    const bool ok_aux = aux == nullptr || (((uintptr_t)aux & 15) == 0 && (ld_auxmap * 2) % 16 == 0);
    // kernels exist for: TN PLAIN bf16/f32, TN GELU bf16, NN PLAIN bf16, NN GELU_BWD bf16, wgrad PLAIN f32
    if (!ok_out || !ok_aux || (kind == EK_ACCUM && (layout != PMV_GEMM_NN || out_dtype != PMV_BF16))) {
      kind = EK_GENERIC;
    } else {
      rc = pmv_make_tensor_map_2d(&tmC, e.out, osz, (uint64_t)NN, (uint64_t)MM, (uint64_t)e.ldo, 32, 32, osz == 2 ? 64 : 128);
      if (rc) return rc;
      if (aux) {
        rc = pmv_make_tensor_map_2d(&tmD, aux, 2, (uint64_t)NN, (uint64_t)MM, (uint64_t)ld_auxmap, 32, 32, 64);
        if (rc) return rc;
      }
    }
  }
  // CTA pairs (cta_group::2, 256-row tiles, the B tile crosses L2 -> SM once per pair): forward layout, when the problem
  // has at least one full round of pair tiles
  {
    static int pair_mode = -1;  // PMV_GEMM_PAIR: 0 = never, 1 = when eligible (default 0 until validated on the model)
    if (pair_mode < 0) {
      const char* ev = std::getenv("PMV_GEMM_PAIR");
      pair_mode = ev ? atoi(ev) : 0;
    }
    const bool kind_ok = (kind == EK_PLAIN && out_dtype == PMV_BF16) || (kind == EK_GELU && out_dtype == PMV_BF16) ||
                         (kind == EK_RES && out_dtype == PMV_F32);
    if (pair_mode && layout == PMV_GEMM_TN && p.splits == 1 && kind_ok && BN >= 128 && MM >= 256 * 37) {
      TcParams pp = p;
      pp.tiles_m = (int)ceil_div64(MM, 2 * BM);
      CUtensorMap tmBh;
      rc = pmv_make_tensor_map_2d(&tmBh, B, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)(BN / 2), 128);
      if (rc) return rc;
      switch (BN) {
        case 128: return gemm_tc_launch_pair_bn128(out_dtype, kind, tmA, tmBh, tmC, tmD, pp, g_num_sms, stream);
        case 192: return gemm_tc_launch_pair_bn192(out_dtype, kind, tmA, tmBh, tmC, tmD, pp, g_num_sms, stream);
        default: return gemm_tc_launch_pair_bn256(out_dtype, kind, tmA, tmBh, tmC, tmD, pp, g_num_sms, stream);
      }
    }
  }
  switch (BN) {
    case 96: return gemm_tc_launch_bn96(layout, out_dtype, kind, tmA, tmB, tmC, tmD, p, g_num_sms, stream);
    case 128: return gemm_tc_launch_bn128(layout, out_dtype, kind, tmA, tmB, tmC, tmD, p, g_num_sms, stream);
    case 192: return gemm_tc_launch_bn192(layout, out_dtype, kind, tmA, tmB, tmC, tmD, p, g_num_sms, stream);
    default: return gemm_tc_launch_bn256(layout, out_dtype, kind, tmA, tmB, tmC, tmD, p, g_num_sms, stream);
  }
}
