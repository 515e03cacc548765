// pmv_gemm dispatcher (fp32 FFMA kernel vs tcgen05 kernel) and the column-sum / cast helper.
#include "gemm.h"
#include "reduce.cuh"

extern "C" int pmv_gemm(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
                        int64_t M, int64_t N, int64_t K, int io_dtype, int out_dtype, const pmv_epilogue* epi,
                        int tc, int split_k, void* stream) {
  PMV_CHECK_ARG(layout >= PMV_GEMM_TN && layout <= PMV_GEMM_NT_REDUCE_M, "gemm: bad layout %d", layout);
  PMV_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad shape %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
  PMV_CHECK_ARG(A && B && out, "gemm: null operand");
  if (M == 0) return PMV_OK;
  EpiDev e;
  memset(&e, 0, sizeof(e));
  e.out = out;
  e.ldo = ldo;
  if (epi) {
    e.bias = epi->bias; e.act = epi->act; e.aux_in = epi->aux_in; e.aux_out = epi->aux_out; e.ld_aux = epi->ld_aux;
    e.row_scale = epi->row_scale; e.rows_per_scale = epi->rows_per_scale > 0 ? epi->rows_per_scale : 1;
    e.residual = epi->residual; e.ld_residual = epi->ld_residual; e.accumulate = epi->accumulate;
    e.out_group = epi->out_group; e.out_skip = epi->out_skip;
    PMV_CHECK_ARG(e.act != PMV_ACT_GELU_BWD || e.aux_in, "gemm: GELU_BWD needs aux_in");
  }
  PMV_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimensions must fit 31 bits");
  e.fd_scale = FastDiv((uint32_t)(e.rows_per_scale > 0 ? e.rows_per_scale : 1));
  e.fd_group = FastDiv((uint32_t)(e.out_group > 0 ? e.out_group : 1));
  if (split_k > 1) {
    PMV_CHECK_ARG(layout == PMV_GEMM_NT_REDUCE_M && out_dtype == PMV_F32 && !epi, "gemm: split_k only for fp32 wgrad without epilogue");
    e.atomic = 1;
  }
  if (tc) {
    PMV_CHECK_ARG(io_dtype == PMV_BF16, "gemm: the tcgen05 kernel takes bf16 operands");
    if (!pmv_has_tcgen05()) {
      pmv_set_error("gemm: tcgen05 kernel requested on a device that is not sm_100");
      return PMV_ERR_UNSUPPORTED;
    }
    return gemm_tc_launch(layout, A, lda, B, ldb, M, N, K, out_dtype, e, split_k, (cudaStream_t)stream);
  }
  return gemm_simt_launch(layout, A, lda, B, ldb, M, N, K, io_dtype, out_dtype, e, split_k, (cudaStream_t)stream);
}

namespace {
// rows are split over blockIdx.y; each thread owns 4 consecutive columns
template <typename TIn, typename TCast>
__global__ void __launch_bounds__(256) colsum_cast_kernel(const TIn* __restrict__ in, int64_t ld_in, int64_t rows, int64_t cols,
                                                          const float* __restrict__ row_scale, int64_t rows_per_scale,
                                                          float* __restrict__ out_sum, TCast* __restrict__ cast_out, int64_t ld_cast,
                                                          int64_t rows_per_block) {
  const int64_t c4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = r0; r < r1; ++r) {
    float v[4];
    load4(in + r * ld_in + c4, v);
    if (row_scale) {
      const float sc = row_scale[r / rows_per_scale];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= sc;
    }
    if (cast_out) store4(cast_out + r * ld_cast + c4, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] += v[j];
  }
  if (out_sum) store4(out_sum + (int64_t)blockIdx.y * cols + c4, s);  // one partial row per row slice
}
}  // namespace

static void colsum_grid(int64_t rows, int64_t cols, unsigned* bx, int64_t* rpb, unsigned* by) {
  const int64_t col_threads = cols / 4;
  *bx = (unsigned)ceil_div64(col_threads, 256);
  int64_t slices = ceil_div64(148 * 4, *bx);  // enough row slices to fill the machine
  if (slices > ceil_div64(rows, 16)) slices = ceil_div64(rows, 16);
  if (slices < 1) slices = 1;
  *rpb = ceil_div64(rows, slices);
  *by = (unsigned)ceil_div64(rows, *rpb);
}

extern "C" int64_t pmv_colsum_workspace_bytes(int64_t rows, int64_t cols) {
  unsigned bx, by;
  int64_t rpb;
  colsum_grid(rows, cols, &bx, &rpb, &by);
  return (int64_t)by * cols * (int64_t)sizeof(float);
}

extern "C" int pmv_colsum_cast(const void* in, int in_dtype, int64_t ld_in, int64_t rows, int64_t cols,
                               const float* row_scale, int64_t rows_per_scale, float* out_sum, float* ws,
                               void* cast_out, int cast_dtype, int64_t ld_cast, void* stream) {
  PMV_CHECK_ARG(cols % 4 == 0 && ld_in % 4 == 0 && (cast_out == nullptr || ld_cast % 4 == 0), "colsum: cols / ld must be multiples of 4");
  if (rows == 0) return PMV_OK;
  if (rows_per_scale <= 0) rows_per_scale = 1;
  const int64_t col_threads = cols / 4;
  unsigned bx, by;
  int64_t rpb;
  colsum_grid(rows, cols, &bx, &rpb, &by);
  dim3 grid(bx, by);
  PMV_CHECK_ARG(out_sum == nullptr || ws != nullptr, "colsum: workspace required when out_sum is given");
  const int threads = col_threads < 256 ? (int)((col_threads + 31) / 32 * 32) : 256;
#define LAUNCH(TI, TC) colsum_cast_kernel<TI, TC><<<grid, threads, 0, (cudaStream_t)stream>>>( \
      (const TI*)in, ld_in, rows, cols, row_scale, rows_per_scale, out_sum ? ws : nullptr, (TC*)cast_out, ld_cast, rpb)
  if (in_dtype == PMV_F32 && cast_dtype == PMV_F32) LAUNCH(float, float);
  else if (in_dtype == PMV_F32 && cast_dtype == PMV_BF16) LAUNCH(float, bf16);
  else if (in_dtype == PMV_BF16 && cast_dtype == PMV_BF16) LAUNCH(bf16, bf16);
  else if (in_dtype == PMV_BF16 && cast_dtype == PMV_F32) LAUNCH(bf16, float);
  else { pmv_set_error("colsum: bad dtype"); return PMV_ERR_INVALID_ARGUMENT; }
#undef LAUNCH
  if (out_sum) launch_reduce_partials(ws, (int)by, (int)cols, out_sum, (cudaStream_t)stream);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
