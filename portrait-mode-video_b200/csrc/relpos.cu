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

// backward of augment_q.  One warp per query, lane owns channels {lane, lane+32, lane+64}.
// dynamic smem: fp32 gradient tables [(rows)][96] accumulated with shared atomics, written out once per block.
template <typename T>
__global__ void __launch_bounds__(AUG_WARPS * 32) relpos_augment_q_bwd_kernel(
    T* __restrict__ dq_aug, const T* __restrict__ q_aug, int64_t ld, const float* __restrict__ rel_h,
    const float* __restrict__ rel_w, const float* __restrict__ rel_t, const int32_t* __restrict__ idx_h,
    const int32_t* __restrict__ idx_w, const int32_t* __restrict__ idx_t, float* __restrict__ partials,
    int64_t BH, RelGeom g, float inv_scale) {
  extern __shared__ float dtab[];
  const int rows = g.rows_h + g.rows_w + g.rows_t;
  float* tab = dtab + rows * HD;  // the three tables stacked, staged once per block
  for (int i = threadIdx.x; i < rows * HD; i += blockDim.x) {
    dtab[i] = 0.f;
    const int r = i / HD, c = i - r * HD;
    tab[i] = r < g.rows_h ? rel_h[r * HD + c]
             : r < g.rows_h + g.rows_w ? rel_w[(r - g.rows_h) * HD + c]
                                       : rel_t[(r - g.rows_h - g.rows_w) * HD + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Lq = g.qt * g.qh * g.qw;
  const int Nq = Lq + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int64_t total = BH * Nq;
  for (int64_t row = (int64_t)blockIdx.x * AUG_WARPS + warp; row < total; row += (int64_t)gridDim.x * AUG_WARPS) {
    const int n = (int)(row % Nq);
    if (n == 0) continue;
    int l = n - 1;
    const int iw = l % g.qw; l /= g.qw;
    const int ih = l % g.qh;
    const int it = l / g.qh;
    T* dqp = dq_aug + row * ld;
    const T* qp = q_aug + row * ld;
    float qv[3], acc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { qv[j] = to_f32(qp[lane + 32 * j]); acc[j] = to_f32(dqp[lane + 32 * j]); }
    // the d rq values of this query: one coalesced load, broadcast by shuffle inside the loop
    const float d_lo = lane < RK ? to_f32(dqp[HD + lane]) * inv_scale : 0.f;
    const float d_hi = lane + 32 < RK ? to_f32(dqp[HD + 32 + lane]) * inv_scale : 0.f;
    // table row of column (j0 + lane): resolved once per query by the lanes, broadcast by shuffle
    int my_row_lo = 0, my_row_hi = 0;
    {
      const int j = lane;
      if (j < RK) my_row_lo = j < g.kh ? idx_h[ih * g.kh + j]
                              : j < g.kh + g.kw ? g.rows_h + idx_w[iw * g.kw + (j - g.kh)]
                                                : g.rows_h + g.rows_w + idx_t[it * g.kt + (j - g.kh - g.kw)];
      const int j2 = lane + 32;
      if (j2 < RK) my_row_hi = j2 < g.kh ? idx_h[ih * g.kh + j2]
                               : j2 < g.kh + g.kw ? g.rows_h + idx_w[iw * g.kw + (j2 - g.kh)]
                                                  : g.rows_h + g.rows_w + idx_t[it * g.kt + (j2 - g.kh - g.kw)];
    }
#pragma unroll 4
    for (int j = 0; j < RK; ++j) {
      const float d = __shfl_sync(0xffffffffu, j < 32 ? d_lo : d_hi, j & 31);
      const int trow = __shfl_sync(0xffffffffu, j < 32 ? my_row_lo : my_row_hi, j & 31);
      const float* src = tab + trow * HD;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] = fmaf(d, src[lane + 32 * c], acc[c]);
        atomicAdd(&dtab[trow * HD + lane + 32 * c], d * qv[c]);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dqp[lane + 32 * j] = from_f32<T>(acc[j]);
  }
  __syncthreads();
  // one partial table per block ([rows_h + rows_w + rows_t][96], folded by reduce_partials_kernel)
  for (int i = threadIdx.x; i < rows * HD; i += blockDim.x) partials[(int64_t)blockIdx.x * rows * HD + i] = dtab[i];
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

static int64_t relpos_bwd_blocks(int64_t total_rows) {
  int64_t blocks = ceil_div64(total_rows, AUG_WARPS * 8);
  if (blocks > 148 * 2) blocks = 148 * 2;
  return blocks < 1 ? 1 : blocks;
}

extern "C" int64_t pmv_relpos_bwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  return relpos_bwd_blocks((int64_t)BH * (qt * qh * qw + 1)) * (g.rows_h + g.rows_w + g.rows_t) * HD * (int64_t)sizeof(float);
}

/* d_rel: [rows_h + rows_w + rows_t][96] fp32 (the three tables stacked), added to. */
extern "C" int pmv_relpos_augment_q_bwd(void* dq_aug, const void* q_aug, int64_t ld, const float* rel_h, const float* rel_w,
                                        const float* rel_t, const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t,
                                        float* d_rel, float* ws,
                                        int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                        float inv_scale, int dtype, void* stream) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const size_t smem = (size_t)2 * (g.rows_h + g.rows_w + g.rows_t) * HD * sizeof(float);
  PMV_CHECK_ARG(smem <= 200 * 1024, "relpos: tables too large for shared memory");
  PMV_CHECK_ARG(kh + kw + kt <= 64, "relpos: at most 64 bias columns");
  const int64_t total = (int64_t)BH * (qt * qh * qw + 1);
  const int64_t blocks = relpos_bwd_blocks(total);
  PMV_DISPATCH_DTYPE(dtype, T, {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(relpos_augment_q_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    relpos_augment_q_bwd_kernel<T><<<(unsigned)blocks, AUG_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (T*)dq_aug, (const T*)q_aug, ld, rel_h, rel_w, rel_t, idx_h, idx_w, idx_t, ws, BH, g, inv_scale);
  });
  launch_reduce_partials(ws, (int)blocks, (g.rows_h + g.rows_w + g.rows_t) * HD, d_rel, (cudaStream_t)stream);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
