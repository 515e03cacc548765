#define PMV_PDL_FAMILY 128
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
//
// The per-query dot products are themselves a GEMM against the three stacked tables:
//     RQ[M, rows_h + rows_w + rows_t] = Q[M, 96] x (Rcat / scale)^T                     (pmv_gemm, tensor cores in bf16)
//     Q'[q, 96 + j] = RQ[q, row_j(q)]                                                   (gather kernel)
// and the backward is the mirror image: the d rq columns are scattered into a dense dRQ (zero elsewhere), then
//     dQ[:, :96] += dRQ x (Rcat / scale)      (dgrad GEMM, accumulating epilogue)
//     dRcat       = dRQ^T x Q / scale         (wgrad GEMM, split over the query rows)
// The dense detour costs (rows_h + rows_w + rows_t) instead of (k_h + k_w + k_t) columns per query, all of it on
// the tensor cores; the previous CUDA-core kernels (one warp per query, 96-long dot products out of shared
// memory) ran at 3 TFLOP/s and were 10 % of the training step (profiles/r01_kernels_before.txt).  Round 2 retried a
// CUDA-core version that computes only the k_h + k_w + k_t products a query needs in ONE launch (tables in shared memory, 8
// lanes per query, halving-butterfly reduction, no intermediate in HBM): correct, but 131 / 42 us against 101 / 34 us of this
// detour at block 0 / the mid-stage blocks (profiles/r02_relpos_fused_rejected.txt) — every query re-reads 22-36 table rows
// (4-7 KB) from shared memory, i.e. 1.4 MB per SM at the mid-stage: shared-memory bandwidth again.  The products are a
// small GEMM and stay on the tensor cores.
#include "common.cuh"

namespace {

constexpr int HD = PMV_HEAD_DIM;

struct RelGeom {
  int qt, qh, qw, kt, kh, kw;
  int rows_h, rows_w, rows_t;  // table lengths
};

// cat[r][c] = inv_scale * table[r][c] for the three tables stacked, zero rows up to ncat_pad
template <typename T>
__global__ void __launch_bounds__(256) relpos_cat_kernel(T* __restrict__ cat, const float* __restrict__ rel_h,
                                                         const float* __restrict__ rel_w, const float* __restrict__ rel_t,
                                                         RelGeom g, int ncat_pad, float inv_scale) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncat_pad * HD) return;
  const int r = i / HD, c = i - r * HD;
  float v = 0.f;
  if (r < g.rows_h) v = rel_h[r * HD + c];
  else if (r < g.rows_h + g.rows_w) v = rel_w[(r - g.rows_h) * HD + c];
  else if (r < g.rows_h + g.rows_w + g.rows_t) v = rel_t[(r - g.rows_h - g.rows_w) * HD + c];
  cat[i] = from_f32<T>(v * inv_scale);
}

// column of the dense RQ that bias column j of query position (it, ih, iw) reads
__device__ __forceinline__ int rel_col(const RelGeom& g, const int32_t* __restrict__ idx_h, const int32_t* __restrict__ idx_w,
                                       const int32_t* __restrict__ idx_t, int it, int ih, int iw, int j) {
  if (j < g.kh) return idx_h[ih * g.kh + j];
  if (j < g.kh + g.kw) return g.rows_h + idx_w[iw * g.kw + (j - g.kh)];
  return g.rows_h + g.rows_w + idx_t[it * g.kt + (j - g.kh - g.kw)];
}

// Q'[row, 96 + j] = RQ[row, col_j(row)]  (zero for the cls row and the padding columns).  One warp per row: the
// position decode is warp-uniform 32-bit arithmetic, lanes are the bias columns (the first version decoded every
// element with 64-bit div / mod: 23 us for 1.6 M elements).
constexpr int GA_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(GA_WARPS * 32) relpos_gather_kernel(T* __restrict__ q_aug, int ld, const T* __restrict__ rq, int ld_rq,
                                                                      const int32_t* __restrict__ idx_h, const int32_t* __restrict__ idx_w,
                                                                      const int32_t* __restrict__ idx_t, int rows, RelGeom g) {
  pdl_wait();
  const int aug = ld - HD;
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * GA_WARPS + warp; row < rows; row += gridDim.x * GA_WARPS) {
    const int n = row % Nq;
    int it = 0, ih = 0, iw = 0;
    if (n > 0) {
      int l = n - 1;
      iw = l % g.qw; l /= g.qw;
      ih = l % g.qh;
      it = l / g.qh;
    }
    const T* src = rq + (int64_t)row * ld_rq;
    T* dst = q_aug + (int64_t)row * ld + HD;
    for (int j = lane; j < aug; j += 32) {
      T v = from_f32<T>(0.f);
      if (n > 0 && j < RK) v = src[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j)];
      dst[j] = v;
    }
  }
}

// dRQ[row, :] = 0 except dRQ[row, col_j(row)] = dQ'[row, 96 + j]; one warp per row, the row is assembled in shared
// memory and written with full-width stores
constexpr int SC_WARPS = 8;
constexpr int SC_MAXCOLS = 512;
template <typename T>
__global__ void __launch_bounds__(SC_WARPS * 32) relpos_scatter_kernel(const T* __restrict__ dq_aug, int ld, T* __restrict__ drq,
                                                                       int ncat_pad, const int32_t* __restrict__ idx_h,
                                                                       const int32_t* __restrict__ idx_w,
                                                                       const int32_t* __restrict__ idx_t, int64_t rows, RelGeom g) {
  pdl_wait();
  __shared__ __align__(16) T buf[SC_WARPS][SC_MAXCOLS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int RK = g.kh + g.kw + g.kt;
  T* mine = buf[warp];
  for (int64_t row = (int64_t)blockIdx.x * SC_WARPS + warp; row < rows; row += (int64_t)gridDim.x * SC_WARPS) {
    for (int c = lane; c < ncat_pad; c += 32) mine[c] = from_f32<T>(0.f);
    __syncwarp();
    const int n = (int)((uint32_t)row % (uint32_t)Nq);
    if (n > 0) {
      int l = n - 1;
      const int iw = l % g.qw; l /= g.qw;
      const int ih = l % g.qh;
      const int it = l / g.qh;
      for (int j = lane; j < RK; j += 32) mine[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j)] = dq_aug[row * ld + HD + j];
    }
    __syncwarp();
    // ncat_pad % 8 == 0: 16-byte pieces
    const int pieces = ncat_pad * (int)sizeof(T) / 16;
    uint4* dst = reinterpret_cast<uint4*>(drq + row * ncat_pad);
    const uint4* src = reinterpret_cast<const uint4*>(mine);
    for (int p = lane; p < pieces; p += 32) dst[p] = src[p];
    __syncwarp();
  }
}

// 16-bit variants of the two kernels above with FOUR lanes per row (eight rows per warp, 64 per CTA, no row loop): the
// one-warp-per-row forms are a chain of dependent latencies per row (position decode, index lookup, one 2-byte load per lane)
// with 2-3 rows per warp in sequence: 241 + 339 us per training step for 0.35 GB of traffic.  Here every lane assembles 8 (or
// 16) neighbouring bias columns and moves them with 16-byte accesses, and eight rows are in flight per warp.
constexpr int R4_ROWS = 64;  // rows per CTA of 256 threads
template <typename T>
__global__ void __launch_bounds__(256) relpos_gather4_kernel(T* __restrict__ q_aug, int ld, const T* __restrict__ rq, int ld_rq,
                                                             const int32_t* __restrict__ idx_h, const int32_t* __restrict__ idx_w,
                                                             const int32_t* __restrict__ idx_t, int rows, RelGeom g) {
  static_assert(sizeof(T) == 2, "16-bit elements");
  pdl_wait();
  const int aug = ld - HD, cpl = aug >> 2;  // bias columns per row (multiple of 8 .. 32), per lane (8 or 16)
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int sub = threadIdx.x & 3;
  const int row = (int)((blockIdx.x * 256u + threadIdx.x) >> 2);
  if (row >= rows) return;
  const int n = row % Nq;
  int it = 0, ih = 0, iw = 0;
  if (n > 0) {
    int l = n - 1;
    iw = l % g.qw; l /= g.qw;
    ih = l % g.qh;
    it = l / g.qh;
  }
  const uint16_t* src = reinterpret_cast<const uint16_t*>(rq) + (int64_t)row * ld_rq;
  T* dst = q_aug + (int64_t)row * ld + HD + sub * cpl;
  for (int p = 0; p < cpl; p += 8) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j0 = sub * cpl + p + 2 * i;
      const uint32_t a = (n > 0 && j0 < RK) ? src[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j0)] : 0u;
      const uint32_t b = (n > 0 && j0 + 1 < RK) ? src[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j0 + 1)] : 0u;
      w[i] = a | (b << 16);
    }
    *reinterpret_cast<uint4*>(dst + p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) relpos_scatter4_kernel(const T* __restrict__ dq_aug, int ld, T* __restrict__ drq, int ncat_pad,
                                                              const int32_t* __restrict__ idx_h, const int32_t* __restrict__ idx_w,
                                                              const int32_t* __restrict__ idx_t, int rows, RelGeom g) {
  static_assert(sizeof(T) == 2, "16-bit elements");
  pdl_wait();
  extern __shared__ __align__(16) uint8_t r4_smem[];  // [R4_ROWS][ncat_pad] elements: the dense rows being assembled
  const int aug = ld - HD, cpl = aug >> 2;
  const int Nq = g.qt * g.qh * g.qw + 1;
  const int RK = g.kh + g.kw + g.kt;
  const int sub = threadIdx.x & 3, r_in = threadIdx.x >> 2;
  const int row = (int)blockIdx.x * R4_ROWS + r_in;
  uint16_t* mine = reinterpret_cast<uint16_t*>(r4_smem) + r_in * ncat_pad;
  const int pieces = ncat_pad >> 3;  // 16-byte pieces of a dense row (ncat_pad % 8 == 0)
  for (int p = sub; p < pieces; p += 4) reinterpret_cast<uint4*>(mine)[p] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();  // the four lanes of a row sit in one warp
  const int n = row < rows ? row % Nq : 0;
  if (n > 0) {
    int l = n - 1;
    const int iw = l % g.qw; l /= g.qw;
    const int ih = l % g.qh;
    const int it = l / g.qh;
    const uint16_t* s = reinterpret_cast<const uint16_t*>(dq_aug) + (int64_t)row * ld + HD + sub * cpl;
    for (int p = 0; p < cpl; p += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(s + p);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j0 = sub * cpl + p + 2 * i;
        if (j0 < RK) mine[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j0)] = (uint16_t)(w[i] & 0xffffu);
        if (j0 + 1 < RK) mine[rel_col(g, idx_h, idx_w, idx_t, it, ih, iw, j0 + 1)] = (uint16_t)(w[i] >> 16);
      }
    }
  }
  __syncwarp();
  if (row < rows) {
    uint4* dst = reinterpret_cast<uint4*>(drq + (int64_t)row * ncat_pad);
    for (int p = sub; p < pieces; p += 4) dst[p] = reinterpret_cast<const uint4*>(mine)[p];
  }
}

// d_rel[i] = inv_scale * dcat[i]
__global__ void __launch_bounds__(256) relpos_dtab_kernel(float* __restrict__ d_rel, const float* __restrict__ dcat, int n, float inv_scale) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d_rel[i] = inv_scale * dcat[i];
}

// K'[row, 96 + j] = 1 where j is the key's h, KH + w or KH + KW + t coordinate, else 0 (zero row for the cls key).
// One warp per key row, lanes = the augmented columns (the element-per-thread form paid a 64-bit div / mod per element:
// 7 us for 0.4 M elements).
template <typename T>
__global__ void __launch_bounds__(256) relpos_augment_k_kernel(T* __restrict__ k_aug, int64_t ld, int64_t BH, int kt, int kh, int kw) {
  pdl_wait();
  const int aug = (int)ld - HD;
  const int Nk = kt * kh * kw + 1;
  const int rows = (int)(BH * Nk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const int n = row % Nk;
    int c0 = -1, c1 = -1, c2 = -1;
    if (n > 0) {
      int l = n - 1;
      const int iw = l % kw; l /= kw;
      const int ih = l % kh;
      const int it = l / kh;
      c0 = ih; c1 = kh + iw; c2 = kh + kw + it;
    }
    T* dst = k_aug + (int64_t)row * ld + HD;
    for (int j = lane; j < aug; j += 32) dst[j] = from_f32<T>((j == c0 || j == c1 || j == c2) ? 1.f : 0.f);
  }
}

RelGeom make_rel(int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g;
  g.qt = qt; g.qh = qh; g.qw = qw; g.kt = kt; g.kh = kh; g.kw = kw;
  g.rows_h = 2 * (qh > kh ? qh : kh) - 1;
  g.rows_w = 2 * (qw > kw ? qw : kw) - 1;
  g.rows_t = 2 * (qt > kt ? qt : kt) - 1;
  return g;
}


int ncat_padded(const RelGeom& g) { return (g.rows_h + g.rows_w + g.rows_t + 7) / 8 * 8; }
int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

}  // namespace

extern "C" int64_t pmv_relpos_fwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const int64_t np = ncat_padded(g), M = (int64_t)BH * (qt * qh * qw + 1);
  return align256(np * HD * 4) + align256(M * np * 4);  // sized for fp32
}

extern "C" int pmv_relpos_augment_q(void* q_aug, int64_t ld, const float* rel_h, const float* rel_w, const float* rel_t,
                                    const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t, float* ws,
                                    int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                    float inv_scale, int dtype, int tc, void* stream) {
  PMV_CHECK_ARG(ld >= HD + kh + kw + kt && ld % 8 == 0, "relpos: ld=%lld too small for 96+%d bias columns", (long long)ld, kh + kw + kt);
  PMV_CHECK_ARG(ws != nullptr, "relpos: workspace required");
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const int np = ncat_padded(g);
  const int64_t M = (int64_t)BH * (qt * qh * qw + 1);
  char* cat = reinterpret_cast<char*>(ws);
  char* rq = cat + align256((int64_t)np * HD * 4);
  cudaStream_t st = (cudaStream_t)stream;
  PMV_DISPATCH_DTYPE(dtype, T, (pmv_launch(relpos_cat_kernel<T>, (np * HD + 255) / 256, 256, 0, st, (T*)cat, rel_h, rel_w, rel_t, g, np, inv_scale)));
  int rc = pmv_gemm(PMV_GEMM_TN, q_aug, ld, cat, HD, rq, np, M, np, HD, dtype, dtype, nullptr, tc, 1, stream);
  if (rc) return rc;
  PMV_CHECK_ARG(M < (1ll << 31), "relpos: too many query rows");
  if (dtype == PMV_BF16 && (ld - HD) % 32 == 0 && (ld - HD) <= 64) {  // 8 or 16 bias columns per lane, 16-byte stores
    pmv_launch(relpos_gather4_kernel<bf16>, (unsigned)ceil_div64(M, R4_ROWS), 256, 0, st, (bf16*)q_aug, (int)ld, (const bf16*)rq, np, idx_h, idx_w,
               idx_t, (int)M, g);
    PMV_CHECK_LAUNCH();
    return PMV_OK;
  }
  int64_t blocks = ceil_div64(M, GA_WARPS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  PMV_DISPATCH_DTYPE(dtype, T, (pmv_launch(relpos_gather_kernel<T>, (unsigned)blocks, GA_WARPS * 32, 0, st, (T*)q_aug, (int)ld, (const T*)rq, np, idx_h,
                                                                                                      idx_w, idx_t, (int)M, g)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int pmv_relpos_augment_k(void* k_aug, int64_t ld, int BH, int kt, int kh, int kw, int dtype, void* stream) {
  PMV_CHECK_ARG(ld >= HD + kh + kw + kt, "relpos: ld too small");
  if (ld == HD) return PMV_OK;
  const int64_t rows = (int64_t)BH * (kt * kh * kw + 1);
  PMV_CHECK_ARG(rows < (1ll << 31), "relpos: too many key rows");
  int64_t blocks = ceil_div64(rows, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  PMV_DISPATCH_DTYPE(dtype, T, (pmv_launch(relpos_augment_k_kernel<T>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, (T*)k_aug, ld, BH, kt, kh, kw)));
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}


extern "C" int64_t pmv_relpos_bwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const int64_t np = ncat_padded(g), M = (int64_t)BH * (qt * qh * qw + 1);
  return 2 * align256(np * HD * 4) + align256(M * np * 4);
}

/* d_rel: [rows_h + rows_w + rows_t][96] fp32 (the three tables stacked), OVERWRITTEN. */
extern "C" int pmv_relpos_augment_q_bwd(void* dq_aug, const void* q_aug, int64_t ld, const float* rel_h, const float* rel_w,
                                        const float* rel_t, const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t,
                                        float* d_rel, float* ws,
                                        int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                                        float inv_scale, int dtype, int tc, void* stream) {
  RelGeom g = make_rel(qt, qh, qw, kt, kh, kw);
  const int np = ncat_padded(g);
  PMV_CHECK_ARG(np <= SC_MAXCOLS, "relpos: stacked tables too long (%d rows)", np);
  PMV_CHECK_ARG(ws != nullptr && d_rel != nullptr, "relpos: workspace / d_rel required");
  const int64_t M = (int64_t)BH * (qt * qh * qw + 1);
  const int ncat = g.rows_h + g.rows_w + g.rows_t;
  char* cat = reinterpret_cast<char*>(ws);
  float* dcat = reinterpret_cast<float*>(cat + align256((int64_t)np * HD * 4));
  char* drq = reinterpret_cast<char*>(dcat) + align256((int64_t)np * HD * 4);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ceil_div64(M, SC_WARPS);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == PMV_BF16 && (ld - HD) % 32 == 0 && (ld - HD) <= 64 && M < (1ll << 31) && R4_ROWS * np * 2 <= 48 * 1024) {
    pmv_launch(relpos_cat_kernel<bf16>, (np * HD + 255) / 256, 256, 0, st, (bf16*)cat, rel_h, rel_w, rel_t, g, np, inv_scale);
    pmv_launch(relpos_scatter4_kernel<bf16>, (unsigned)ceil_div64(M, R4_ROWS), 256, (size_t)R4_ROWS * np * 2, st, (const bf16*)dq_aug, (int)ld,
               (bf16*)drq, np, idx_h, idx_w, idx_t, (int)M, g);
  } else {
    PMV_DISPATCH_DTYPE(dtype, T, {
      pmv_launch(relpos_cat_kernel<T>, (np * HD + 255) / 256, 256, 0, st, (T*)cat, rel_h, rel_w, rel_t, g, np, inv_scale);
      pmv_launch(relpos_scatter_kernel<T>, (unsigned)blocks, SC_WARPS * 32, 0, st, (const T*)dq_aug, (int)ld, (T*)drq, np, idx_h, idx_w, idx_t, M, g);
    });
  }
  PMV_CHECK_LAUNCH();
  // dQ[:, :96] += dRQ x cat      (cat already carries 1/scale)
  pmv_epilogue e;
  memset(&e, 0, sizeof(e));
  e.accumulate = 1;
  int rc = pmv_gemm(PMV_GEMM_NN, drq, np, cat, HD, dq_aug, ld, M, HD, np, dtype, dtype, &e, tc, 1, stream);
  if (rc) return rc;
  // dcat = dRQ^T x Q, split over the query rows
  int split = 1;
  {
    const int tiles = ((np + 127) / 128) * 1;
    int want = (148 * 2) / tiles;
    if (want < 1) want = 1;
    const int64_t cap = M >= 512 ? M / 512 : 1;
    split = (int)(want < cap ? want : cap);
    if (split < 1) split = 1;
  }
  if (split > 1) PMV_CHECK_CUDA(cudaMemsetAsync(dcat, 0, (size_t)np * HD * 4, st));
  rc = pmv_gemm(PMV_GEMM_NT_REDUCE_M, drq, np, q_aug, ld, dcat, HD, M, np, HD, dtype, PMV_F32, nullptr, tc, split, stream);
  if (rc) return rc;
  pmv_launch(relpos_dtab_kernel, (ncat * HD + 255) / 256, 256, 0, st, d_rel, dcat, ncat * HD, inv_scale);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
