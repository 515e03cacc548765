// Decomposed relative-position bias folded into the score GEMM (cal_rel_pos_spatial / _temporal,
// attention.py:67-159).  The reference materialises the [B,heads,Nq,Nk] score matrix and makes three
// read-modify-write passes over it; here the bias is produced by the tensor cores themselves:
//
//     bias[q,(kt,kh,kw)] = rq[q, kh] + rq[q, KH + kw] + rq[q, KH + KW + kt],   rq[q, j] = q . R_j(q)
//     scale * [q | rq/scale] . [k | onehot(kh) onehot(kw) onehot(kt)] = scale * q.k + bias
//
// pmv_relpos_augment_q fills columns [96, ld) of Q' with rq/scale (zeros for the cls row: the cls query gets
// no bias, attention.py:111,154), pmv_relpos_augment_k fills columns [96, ld) of K' with the one-hot key
// coordinates (zeros for the cls key).  R_j(q) are rows of rel_pos_h/w/t selected by the reference's index
// arithmetic (attention.py:80-99,132-139), passed in as small int32 tables.
#include "common.cuh"
#include "reduce.cuh"

namespace {

constexpr int HD = PMV_HEAD_DIM;
constexpr int AUG_WARPS = 8;
constexpr int TPAD = HD + 1;

struct RelGeom {
  int qt, qh, qw, kt, kh, kw;
  int rows_h, rows_w, rows_t;  // table lengths
};

// dynamic smem: tables [(rows_h + rows_w + rows_t)][97] fp32, then per-warp q rows [AUG_WARPS][96]
template <typename T>
__global__ void __launch_bounds__(AUG_WARPS * 32) relpos_augment_q_kernel(
    T* __restrict__ q_aug, int64_t ld, const float* __restrict__ rel_h, const float* __restrict__ rel_w,
    const float* __restrict__ rel_t, const int32_t* __restrict__ idx_h, const int32_t* __restrict__ idx_w,
    const int32_t* __restrict__ idx_t, int64_t BH, RelGeom g, float inv_scale) {
  extern __shared__ float sm[];
  const int rows = g.rows_h + g.rows_w + g.rows_t;
  float* tab = sm;
  float* qrow = sm + (size_t)rows * TPAD;
  for (int i = threadIdx.x; i < rows * HD; i += blockDim.x) {
    const int r = i / HD, c = i - r * HD;
    const float v = r < g.rows_h ? rel_h[r * HD + c]
                    : r < g.rows_h + g.rows_w ? rel_w[(r - g.rows_h) * HD + c]
                                              : rel_t[(r - g.rows_h - g.rows_w) * HD + c];
    tab[r * TPAD + c] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* myq = qrow + warp * HD;
  const int Lq = g.qt * g.qh * g.qw;
  const int Nq = Lq + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int aug = (int)ld - HD;
  const int64_t total = BH * Nq;
  for (int64_t row = (int64_t)blockIdx.x * AUG_WARPS + warp; row < total; row += (int64_t)gridDim.x * AUG_WARPS) {
    const int n = (int)(row % Nq);
    T* qp = q_aug + row * ld;
    if (n == 0) {
      for (int j = lane; j < aug; j += 32) qp[HD + j] = from_f32<T>(0.f);
      continue;
    }
    int l = n - 1;
    const int iw = l % g.qw; l /= g.qw;
    const int ih = l % g.qh;
    const int it = l / g.qh;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) myq[lane + 32 * j] = to_f32(qp[lane + 32 * j]);
    __syncwarp();
    for (int j0 = 0; j0 < aug; j0 += 32) {
      const int j = j0 + lane;
      float acc = 0.f;
      if (j < RK) {
        int trow;
        if (j < g.kh) trow = idx_h[ih * g.kh + j];
        else if (j < g.kh + g.kw) trow = g.rows_h + idx_w[iw * g.kw + (j - g.kh)];
        else trow = g.rows_h + g.rows_w + idx_t[it * g.kt + (j - g.kh - g.kw)];
        const float* tr = tab + trow * TPAD;
#pragma unroll 8
        for (int c = 0; c < HD; ++c) acc = fmaf(myq[c], tr[c], acc);
      }
      if (j < aug) qp[HD + j] = from_f32<T>(acc * inv_scale);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) relpos_augment_k_kernel(T* __restrict__ k_aug, int64_t ld, int64_t BH, int kt, int kh, int kw) {
  const int aug = (int)ld - HD;
  const int Nk = kt * kh * kw + 1;
  const int64_t total = BH * Nk * aug;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % aug);
    const int64_t row = i / aug;
    const int n = (int)(row % Nk);
    float v = 0.f;
    if (n > 0) {
      int l = n - 1;
      const int iw = l % kw; l /= kw;
      const int ih = l % kh;
      const int it = l / kh;
      v = (j == ih || j == kh + iw || j == kh + kw + it) ? 1.f : 0.f;
    }
    k_aug[row * ld + HD + j] = from_f32<T>(v);
  }
}

// backward of augment_q, two kernels that keep the table gradients in registers instead of hammering shared
// memory with one atomic per (query, column, channel):
//   row walk    : one warp per (bh, t, h) row of queries.  All queries of a row share their rel_pos_h and rel_pos_t
//                 rows, so d rel_h / d rel_t accumulate in registers over the row and are flushed once per row;
//                 the bias path's contribution to dq (all three parts) is added in place.
//   column walk : one warp per (bh, t, w) column of queries; d rel_w accumulates in registers over the column.
// Lane owns channels {lane, lane+32, lane+64}.  Both kernels write one partial table per CTA.
constexpr int KMAX = 16;  // max k_h, k_w, k_t (MViTv2-S: 14 / 14 / 8, MViTv2-B: 14 / 14 / 16)
constexpr int RW_WARPS = 16;

template <typename T>
__global__ void __launch_bounds__(RW_WARPS * 32, 1) relpos_bwd_rows_kernel(
    T* __restrict__ dq_aug, const T* __restrict__ q_aug, int64_t ld, const float* __restrict__ rel_h,
    const float* __restrict__ rel_w, const float* __restrict__ rel_t, const int32_t* __restrict__ idx_h,
    const int32_t* __restrict__ idx_w, const int32_t* __restrict__ idx_t, float* __restrict__ partials,
    int64_t BH, RelGeom g, float inv_scale) {
  extern __shared__ float sm[];
  const int rows = g.rows_h + g.rows_w + g.rows_t;
  float* tab = sm;                    // the three tables stacked [rows][96]
  float* dtab = sm + rows * HD;       // gradients of the h and t tables [rows_h + rows_t][96]
  const int nht = g.rows_h + g.rows_t;
  for (int i = threadIdx.x; i < rows * HD; i += blockDim.x) {
    const int r = i / HD, c = i - r * HD;
    tab[i] = r < g.rows_h ? rel_h[r * HD + c]
             : r < g.rows_h + g.rows_w ? rel_w[(r - g.rows_h) * HD + c]
                                       : rel_t[(r - g.rows_h - g.rows_w) * HD + c];
  }
  for (int i = threadIdx.x; i < nht * HD; i += blockDim.x) dtab[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int64_t nrows = BH * g.qt * g.qh;
  for (int64_t rr = (int64_t)blockIdx.x * RW_WARPS + warp; rr < nrows; rr += (int64_t)gridDim.x * RW_WARPS) {
    const int ih = (int)(rr % g.qh);
    const int it = (int)((rr / g.qh) % g.qt);
    const int64_t bh = rr / ((int64_t)g.qh * g.qt);
    // table rows shared by the whole query row (lane j holds the row of column j)
    const int my_h = lane < g.kh ? idx_h[ih * g.kh + lane] : 0;
    const int my_t = lane < g.kt ? g.rows_h + g.rows_w + idx_t[it * g.kt + lane] : 0;
    float gh[KMAX][3], gt[KMAX][3];
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) { gh[j][c] = 0.f; gt[j][c] = 0.f; }
    const int64_t row0 = bh * Nq + 1 + ((int64_t)it * g.qh + ih) * g.qw;
    for (int iw = 0; iw < g.qw; ++iw) {
      T* dqp = dq_aug + (row0 + iw) * ld;
      const T* qp = q_aug + (row0 + iw) * ld;
      float qv[3], acc[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { qv[c] = to_f32(qp[lane + 32 * c]); acc[c] = to_f32(dqp[lane + 32 * c]); }
      const float d_lo = lane < RK ? to_f32(dqp[HD + lane]) * inv_scale : 0.f;
      const float d_hi = lane + 32 < RK ? to_f32(dqp[HD + 32 + lane]) * inv_scale : 0.f;
      const int my_w = lane < g.kw ? g.rows_h + idx_w[iw * g.kw + lane] : 0;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < g.kh) {
          const float d = __shfl_sync(0xffffffffu, d_lo, j);
          const float* src = tab + __shfl_sync(0xffffffffu, my_h, j) * HD + lane;
#pragma unroll
          for (int c = 0; c < 3; ++c) { acc[c] = fmaf(d, src[32 * c], acc[c]); gh[j][c] = fmaf(d, qv[c], gh[j][c]); }
        }
      }
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < g.kw) {
          const int col = g.kh + j;
          const float d = __shfl_sync(0xffffffffu, col < 32 ? d_lo : d_hi, col & 31);
          const float* src = tab + __shfl_sync(0xffffffffu, my_w, j) * HD + lane;
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] = fmaf(d, src[32 * c], acc[c]);
        }
      }
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < g.kt) {
          const int col = g.kh + g.kw + j;
          const float d = __shfl_sync(0xffffffffu, col < 32 ? d_lo : d_hi, col & 31);
          const float* src = tab + __shfl_sync(0xffffffffu, my_t, j) * HD + lane;
#pragma unroll
          for (int c = 0; c < 3; ++c) { acc[c] = fmaf(d, src[32 * c], acc[c]); gt[j][c] = fmaf(d, qv[c], gt[j][c]); }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) dqp[lane + 32 * c] = from_f32<T>(acc[c]);
    }
    // flush the row's table gradients (h rows first, then t rows, in the partial layout)
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (j < g.kh) {
        float* dst = dtab + __shfl_sync(0xffffffffu, my_h, j) * HD + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(dst + 32 * c, gh[j][c]);
      }
      if (j < g.kt) {
        float* dst = dtab + (__shfl_sync(0xffffffffu, my_t, j) - g.rows_w) * HD + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(dst + 32 * c, gt[j][c]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nht * HD; i += blockDim.x) partials[(int64_t)blockIdx.x * nht * HD + i] = dtab[i];
}

template <typename T>
__global__ void __launch_bounds__(RW_WARPS * 32) relpos_bwd_cols_kernel(
    const T* __restrict__ dq_aug, const T* __restrict__ q_aug, int64_t ld, const int32_t* __restrict__ idx_w,
    float* __restrict__ partials, int64_t BH, RelGeom g, float inv_scale) {
  extern __shared__ float dtab[];  // [rows_w][96]
  for (int i = threadIdx.x; i < g.rows_w * HD; i += blockDim.x) dtab[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int64_t ncols = BH * g.qt * g.qw;
  for (int64_t cc = (int64_t)blockIdx.x * RW_WARPS + warp; cc < ncols; cc += (int64_t)gridDim.x * RW_WARPS) {
    const int iw = (int)(cc % g.qw);
    const int it = (int)((cc / g.qw) % g.qt);
    const int64_t bh = cc / ((int64_t)g.qw * g.qt);
    const int my_w = lane < g.kw ? idx_w[iw * g.kw + lane] : 0;
    float gw[KMAX][3];
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) gw[j][c] = 0.f;
    for (int ih = 0; ih < g.qh; ++ih) {
      const int64_t row = bh * Nq + 1 + ((int64_t)it * g.qh + ih) * g.qw + iw;
      const T* qp = q_aug + row * ld;
      const T* dqp = dq_aug + row * ld;
      float qv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) qv[c] = to_f32(qp[lane + 32 * c]);
      const float dw_ = lane < g.kw ? to_f32(dqp[HD + g.kh + lane]) * inv_scale : 0.f;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < g.kw) {
          const float d = __shfl_sync(0xffffffffu, dw_, j);
#pragma unroll
          for (int c = 0; c < 3; ++c) gw[j][c] = fmaf(d, qv[c], gw[j][c]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      if (j < g.kw) {
        float* dst = dtab + __shfl_sync(0xffffffffu, my_w, j) * HD + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(dst + 32 * c, gw[j][c]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.rows_w * HD; i += blockDim.x) partials[(int64_t)blockIdx.x * g.rows_w * HD + i] = dtab[i];
}

RelGeom make_rel(int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g;
  g.qt = qt; g.qh = qh; g.qw = qw; g.kt = kt; g.kh = kh; g.kw = kw;
  g.rows_h = 2 * (qh > kh ? qh : kh) - 1;
  g.rows_w = 2 * (qw > kw ? qw : kw) - 1;
  g.rows_t = 2 * (qt > kt ? qt : kt) - 1;
  return g;
}

}  // namespace

extern "C" int pmv_relpos_augment_q(void* q_aug, int64_t ld, const float* rel_h, const float* rel_w, const float* rel_t,
                                    const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t,
                                    int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                    float inv_scale, int dtype, void* stream) {
  PMV_CHECK_ARG(ld >= HD + kh + kw + kt && ld % 8 == 0, "relpos: ld=%lld too small for 96+%d bias columns", (long long)ld, kh + kw + kt);
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const size_t smem = ((size_t)(g.rows_h + g.rows_w + g.rows_t) * TPAD + AUG_WARPS * HD) * sizeof(float);
  PMV_CHECK_ARG(smem <= 200 * 1024, "relpos: tables too large for shared memory (%zu B)", smem);
  const int64_t total = (int64_t)BH * (qt * qh * qw + 1);
  int64_t blocks = ceil_div64(total, AUG_WARPS * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;
  PMV_DISPATCH_DTYPE(dtype, T, {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(relpos_augment_q_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    relpos_augment_q_kernel<T><<<(unsigned)blocks, AUG_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (T*)q_aug, ld, rel_h, rel_w, rel_t, idx_h, idx_w, idx_t, BH, g, inv_scale);
  });
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int pmv_relpos_augment_k(void* k_aug, int64_t ld, int BH, int kt, int kh, int kw, int dtype, void* stream) {
  PMV_CHECK_ARG(ld >= HD + kh + kw + kt, "relpos: ld too small");
  if (ld == HD) return PMV_OK;
  const int64_t total = (int64_t)BH * (kt * kh * kw + 1) * (ld - HD);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  PMV_DISPATCH_DTYPE(dtype, T, (relpos_augment_k_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((T*)k_aug, ld, BH, kt, kh, kw)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

static int rw_blocks(int64_t units) {
  int64_t b = ceil_div64(units, RW_WARPS * 2);
  if (b > 148) b = 148;
  return b < 1 ? 1 : (int)b;
}

extern "C" int64_t pmv_relpos_bwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const int64_t r = (int64_t)rw_blocks((int64_t)BH * qt * qh) * (g.rows_h + g.rows_t) * HD;
  const int64_t c = (int64_t)rw_blocks((int64_t)BH * qt * qw) * g.rows_w * HD;
  return (r + c) * (int64_t)sizeof(float);
}

/* d_rel: [rows_h + rows_w + rows_t][96] fp32 (the three tables stacked), added to. */
extern "C" int pmv_relpos_augment_q_bwd(void* dq_aug, const void* q_aug, int64_t ld, const float* rel_h, const float* rel_w,
                                        const float* rel_t, const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t,
                                        float* d_rel, float* ws,
                                        int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                        float inv_scale, int dtype, void* stream) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  PMV_CHECK_ARG(kh <= KMAX && kw <= KMAX && kt <= KMAX && kh + kw + kt <= 64, "relpos: k_h, k_w, k_t must be <= %d", KMAX);
  const int rows = g.rows_h + g.rows_w + g.rows_t, nht = g.rows_h + g.rows_t;
  const size_t smem_r = (size_t)(rows + nht) * HD * sizeof(float);
  const size_t smem_c = (size_t)g.rows_w * HD * sizeof(float);
  PMV_CHECK_ARG(smem_r <= 200 * 1024, "relpos: tables too large for shared memory");
  const int nb_r = rw_blocks((int64_t)BH * qt * qh), nb_c = rw_blocks((int64_t)BH * qt * qw);
  float* part_r = ws;
  float* part_c = ws + (int64_t)nb_r * nht * HD;
  cudaStream_t st = (cudaStream_t)stream;
  PMV_DISPATCH_DTYPE(dtype, T, {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(relpos_bwd_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PMV_CHECK_CUDA(cudaFuncSetAttribute(relpos_bwd_cols_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // the column walk reads the d rq columns, which the row walk leaves untouched (it rewrites columns [0, 96) only)
    relpos_bwd_cols_kernel<T><<<nb_c, RW_WARPS * 32, smem_c, st>>>((const T*)dq_aug, (const T*)q_aug, ld, idx_w, part_c, BH, g, inv_scale);
    relpos_bwd_rows_kernel<T><<<nb_r, RW_WARPS * 32, smem_r, st>>>((T*)dq_aug, (const T*)q_aug, ld, rel_h, rel_w, rel_t, idx_h, idx_w,
                                                                    idx_t, part_r, BH, g, inv_scale);
  });
  launch_reduce_partials(part_r, nb_r, g.rows_h * HD, d_rel, st, (int64_t)nht * HD);
  launch_reduce_partials(part_r + g.rows_h * HD, nb_r, g.rows_t * HD, d_rel + (int64_t)(g.rows_h + g.rows_w) * HD, st, (int64_t)nht * HD);
  launch_reduce_partials(part_c, nb_c, g.rows_w * HD, d_rel + (int64_t)g.rows_h * HD, st);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
