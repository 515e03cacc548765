// Library plumbing: last-error string, version, device capability query, TMA tensor-map creation.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "tc_common.cuh"

static thread_local char g_err[512] = "";

void pmv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pmv_last_error(void) { return g_err; }

// Programmatic dependent launch (common.cuh: pmv_launch / pdl_wait).  Bit mask of kernel families (1 attention fwd,
// 2 attention bwd, 4 tcgen05 GEMM, 8 column sums, 16 LayerNorm, 32 / 64 pooling, 128 rel-pos, 256 the rest).
// Default: every family.  Measured on B200 inside the step's CUDA graph at the end of round 1: inference +2 % (4.78 -> 4.69 ms),
// training neutral (14.65 ms either way).  (Earlier in the round, with torch's optimizer and ~80 fill kernels in the graph,
// the training graph LOST 3 % as soon as some families used it; that penalty went away with those nodes.)  pmv_set_pdl(0) or
// PMV_PDL=0 (read once, before the first launch) turns it off.
static int g_pdl_mask = -1;
extern "C" void pmv_set_pdl(int family_mask) { g_pdl_mask = family_mask & 0x7fffffff; }
bool pmv_pdl_enabled(int family) {
  if (g_pdl_mask < 0) {
    const char* ev = getenv("PMV_PDL");
    g_pdl_mask = ev != nullptr ? (atoi(ev) & 0x7fffffff) : 0x7fffffff;
  }
  return (g_pdl_mask & family) != 0;
}
extern "C" int pmv_version(void) { return 100; }

extern "C" int pmv_has_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int encode(CUtensorMap* out, const void* base, int elem_bytes, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle, const cuuint32_t* walk = nullptr) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    pmv_set_error("cuTensorMapEncodeTiled entry point not available");
    return PMV_ERR_CUDA;
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (walk) {
    for (int i = 0; i < rank; ++i) estr[i] = walk[i];
  }
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pmv_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu] stride1 %llu box [%u,%u] swizzle %d",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)strides_bytes[0], box[0], box[1], swizzle);
    return PMV_ERR_CUDA;
  }
  return PMV_OK;
}

int pmv_make_tensor_map_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0, uint64_t dim1,
                           uint64_t stride1_elems, uint32_t box0, uint32_t box1, int swizzle) {
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box0, box1};
  return encode(out, base, elem_bytes, 2, dims, strides, box, swizzle);
}

int pmv_make_tensor_map_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                           uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t box2,
                           int swizzle) {
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_elems * (uint64_t)elem_bytes, stride2_elems * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  return encode(out, base, elem_bytes, 3, dims, strides, box, swizzle);
}

int pmv_make_tensor_map_5d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                           uint64_t d4, uint64_t s1_elems, uint64_t s2_elems, uint64_t s3_elems, uint64_t s4_elems, uint32_t b0,
                           uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4, uint32_t walk_hw) {
  cuuint64_t dims[5] = {d0, d1, d2, d3, d4};
  cuuint32_t walk[5] = {1, walk_hw, walk_hw, 1, 1};  // element strides along w and h (strided pooling windows)
  cuuint64_t strides[4] = {s1_elems * (uint64_t)elem_bytes, s2_elems * (uint64_t)elem_bytes, s3_elems * (uint64_t)elem_bytes,
                           s4_elems * (uint64_t)elem_bytes};
  cuuint32_t box[5] = {b0, b1, b2, b3, b4};
  return encode(out, base, elem_bytes, 5, dims, strides, box, 0, walk_hw > 1 ? walk : nullptr);
}
