// fp32-accumulate FFMA GEMM used by the fp32 mode of the path (tolerance 1e-4 rules out a single
// TF32/bf16 tensor-core pass, SURVEY.md §7 hard part 1) and as the bring-up reference for the
// tcgen05 kernel.  128x128x16 tiles, 256 threads, 8x8 register micro-tiles, generic operand strides
// so the same kernel serves y = x W^T (TN), dx = dy W (NN) and dW = dy^T x (reduce over rows).
#include "common.cuh"
#include "epilogue.cuh"
#include "gemm.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

// element (i, k) of an operand lives at p[i * si + k * sk]
template <typename T>
__device__ __forceinline__ void load_tile(const T* __restrict__ p, int64_t si, int64_t sk, int64_t i0, int64_t k0,
                                          int64_t I, int64_t Kend, float (*s)[BM + PAD]) {
  const int t = threadIdx.x;
  if (sk == 1) {  // reduction axis contiguous: 16 consecutive lanes walk k
    const int kk = t & 15, r0 = t >> 4;
#pragma unroll
    for (int it = 0; it < BM / 16; ++it) {
      const int r = r0 + it * 16;
      const int64_t gi = i0 + r, gk = k0 + kk;
      s[kk][r] = (gi < I && gk < Kend) ? to_f32(p[gi * si + gk]) : 0.f;
    }
  } else {  // row axis contiguous (si == 1 in practice): 128 consecutive threads walk i
    const int r = t & 127, kk0 = t >> 7;
#pragma unroll
    for (int it = 0; it < BK / 2; ++it) {
      const int kk = kk0 + it * 2;
      const int64_t gi = i0 + r, gk = k0 + kk;
      s[kk][r] = (gi < I && gk < Kend) ? to_f32(p[gi * si + gk * sk]) : 0.f;
    }
  }
}

template <typename TIO, typename TOut>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TIO* __restrict__ A, int64_t sam, int64_t sak,
                                                        const TIO* __restrict__ B, int64_t sbn, int64_t sbk,
                                                        int64_t M, int64_t N, int64_t K, int64_t k_per_split, EpiDev e) {
  pdl_wait();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    load_tile(A, sam, sak, m0, k0, M, kend, As);
    load_tile(B, sbn, sbk, n0, k0, N, kend, Bs);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool vec_ok = (N % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + ty * 8 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j4 = 0; j4 < 2; ++j4) {
      const int64_t col = n0 + tx * 8 + j4 * 4;
      if (vec_ok && col + 3 < N) {
        float v[4] = {acc[i][j4 * 4], acc[i][j4 * 4 + 1], acc[i][j4 * 4 + 2], acc[i][j4 * 4 + 3]};
        epi_store4<TIO, TOut>(e, row, col, v);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) epi_store<TIO, TOut>(e, row, col + j, acc[i][j4 * 4 + j]);
      }
    }
  }
}

}  // namespace

int gemm_simt_launch(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                     int io_dtype, int out_dtype, const EpiDev& e, int split_k, cudaStream_t stream) {
  // logical problem: C[MM, NN] = sum_k A(m,k) B(n,k)
  int64_t MM, NN, KK, sam, sak, sbn, sbk;
  if (layout == PMV_GEMM_TN) {
    MM = M; NN = N; KK = K; sam = lda; sak = 1; sbn = ldb; sbk = 1;
  } else if (layout == PMV_GEMM_NN) {
    MM = M; NN = N; KK = K; sam = lda; sak = 1; sbn = 1; sbk = ldb;
  } else {  // REDUCE_M: C[N, K] = A[M, N]^T B[M, K]; reduction over the M rows
    MM = N; NN = K; KK = M; sam = 1; sak = lda; sbn = 1; sbk = ldb;
  }
  if (split_k < 1) split_k = 1;
  int64_t kps = ceil_div64(ceil_div64(KK, split_k), BK) * BK;
  split_k = (int)ceil_div64(KK, kps);
  dim3 grid((unsigned)ceil_div64(NN, BN), (unsigned)ceil_div64(MM, BM), (unsigned)split_k);
#define LAUNCH(TIO, TOUT)                                                                                  \
  pmv_launch(gemm_simt_kernel<TIO, TOUT>, grid, 256, 0, stream, (const TIO*)A, sam, sak, (const TIO*)B, sbn, sbk, MM, NN, KK, kps, e)
  if (io_dtype == PMV_F32 && out_dtype == PMV_F32) LAUNCH(float, float);
  else if (io_dtype == PMV_BF16 && out_dtype == PMV_BF16) LAUNCH(bf16, bf16);
  else if (io_dtype == PMV_BF16 && out_dtype == PMV_F32) LAUNCH(bf16, float);
  else if (io_dtype == PMV_F32 && out_dtype == PMV_BF16) LAUNCH(float, bf16);
  else { pmv_set_error("gemm: bad dtype"); return PMV_ERR_INVALID_ARGUMENT; }
#undef LAUNCH
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
