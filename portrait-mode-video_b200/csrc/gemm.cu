#define PMV_PDL_FAMILY 8
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
// Thread block = CX column quads x RY rows (CX * RY = 256): consecutive threads read consecutive 16-byte pieces of
// a row, RY rows are in flight per pass and each thread has 4 independent row loads per iteration; the RY partial
// sums are folded through shared memory and every block writes one partial row.  (The first version walked its
// row slice serially with a single warp per block at C = 96: 0.8 TB/s.)
constexpr int CS_THREADS = 256;
template <typename TIn, typename TCast>
__global__ void __launch_bounds__(CS_THREADS) colsum_cast_kernel(const TIn* __restrict__ in, int64_t ld_in, int64_t rows, int64_t cols,
                                                                 const float* __restrict__ row_scale, FastDiv fd_scale,
                                                                 float* __restrict__ out_sum, TCast* __restrict__ cast_out, int64_t ld_cast,
                                                                 int64_t rows_per_block, int cx, float* __restrict__ final_sum) {
  pdl_wait();
  __shared__ float red[CS_THREADS * 4];
  // the fold kernel that follows ADDS the per-slice partials into final_sum: clear it here (saves a fill launch)
  if (final_sum != nullptr && blockIdx.y == 0) {
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    for (int64_t i = c; i < cols; i += (int64_t)gridDim.x * blockDim.x) final_sum[i] = 0.f;
  }
  const int ry = CS_THREADS / cx;
  const int tx = threadIdx.x % cx, ty = threadIdx.x / cx;
  const int64_t c4 = ((int64_t)blockIdx.x * cx + tx) * 4;
  const bool colok = c4 < cols && ty < ry;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (colok) {
    int64_t r = r0 + ty;
    for (; r + 3 * ry < r1; r += 4 * ry) {
      float v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) load4(in + (r + u * ry) * ld_in + c4, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (row_scale) {
          const float sc = row_scale[fd_scale.div((uint32_t)(r + u * ry))];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[u][j] *= sc;
        }
        if (cast_out) store4(cast_out + (r + u * ry) * ld_cast + c4, v[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] += v[u][j];
      }
    }
    for (; r < r1; r += ry) {
      float v[4];
      load4(in + r * ld_in + c4, v);
      if (row_scale) {
        const float sc = row_scale[fd_scale.div((uint32_t)r)];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= sc;
      }
      if (cast_out) store4(cast_out + r * ld_cast + c4, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += v[j];
    }
  }
  if (out_sum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = s[j];
  __syncthreads();
  if (ty == 0 && c4 < cols) {
    for (int y = 1; y < ry; ++y) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += red[(y * cx + tx) * 4 + j];
    }
    store4(out_sum + (int64_t)blockIdx.y * cols + c4, s);  // one partial row per row slice
  }
}
}  // namespace

// column quads per block (power-of-two-free: any cx <= 256), row slices
static void colsum_grid(int64_t rows, int64_t cols, int* cx, unsigned* bx, int64_t* rpb, unsigned* by) {
  const int64_t col_threads = cols / 4;
  // one warp across 32 column quads (512 contiguous bytes of fp32 per row) whenever that tiles the columns exactly:
  // min(col_threads, 256) left half of the second block idle at 1536 columns and a quarter of every block at 384
  *cx = col_threads % 32 == 0 ? 32 : (col_threads < CS_THREADS ? (int)col_threads : CS_THREADS);
  *bx = (unsigned)ceil_div64(col_threads, *cx);
  const int ry = CS_THREADS / *cx;
  int64_t slices = ceil_div64(148 * 6, *bx);  // enough row slices to fill the machine
  const int64_t min_rows = (int64_t)ry * 8;
  if (slices > ceil_div64(rows, min_rows)) slices = ceil_div64(rows, min_rows);
  if (slices < 1) slices = 1;
  *rpb = ceil_div64(rows, slices);
  *by = (unsigned)ceil_div64(rows, *rpb);
}

extern "C" int64_t pmv_colsum_workspace_bytes(int64_t rows, int64_t cols) {
  unsigned bx, by;
  int64_t rpb;
  int cx;
  colsum_grid(rows, cols, &cx, &bx, &rpb, &by);
  return (int64_t)by * cols * (int64_t)sizeof(float);
}

extern "C" int pmv_colsum_cast(const void* in, int in_dtype, int64_t ld_in, int64_t rows, int64_t cols,
                               const float* row_scale, int64_t rows_per_scale, float* out_sum, float* ws,
                               void* cast_out, int cast_dtype, int64_t ld_cast, void* stream) {
  PMV_CHECK_ARG(cols % 4 == 0 && ld_in % 4 == 0 && (cast_out == nullptr || ld_cast % 4 == 0), "colsum: cols / ld must be multiples of 4");
  if (rows == 0) return PMV_OK;
  if (rows_per_scale <= 0) rows_per_scale = 1;
  PMV_CHECK_ARG(rows < (1ll << 31), "colsum: too many rows");
  unsigned bx, by;
  int64_t rpb;
  int cx;
  colsum_grid(rows, cols, &cx, &bx, &rpb, &by);
  dim3 grid(bx, by);
  const FastDiv fd((uint32_t)rows_per_scale);
  PMV_CHECK_ARG(out_sum == nullptr || ws != nullptr, "colsum: workspace required when out_sum is given");
#define LAUNCH(TI, TC) pmv_launch(colsum_cast_kernel<TI, TC>, grid, CS_THREADS, 0, (cudaStream_t)stream,  \
      (const TI*)in, ld_in, rows, cols, row_scale, fd, out_sum ? ws : nullptr, (TC*)cast_out, ld_cast, rpb, cx, out_sum)
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
