// Bring-up probes: a single-tile tcgen05.mma driven entirely by host-provided shared-memory images and
// descriptors, and a single TMA box load dumped back to global memory.  tests/test_tcgen05_probe.py uses
// them to pin the operand layouts (K-major / MN-major, 128B / 64B swizzle, A-from-TMEM packing) that
// gemm_tc.cu and attn_tc.cu rely on.  Not on the product path.
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const uint8_t* __restrict__ image, int image_bytes, uint64_t desc_a_tmpl, uint64_t desc_b_tmpl,
                  uint32_t a_off, uint32_t b_off, uint32_t idesc, int nk, uint32_t a_step, uint32_t b_step, int a_from_tmem,
                  const uint32_t* __restrict__ tmem_a_image, int tmem_a_cols, float* __restrict__ d_out, int n_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  tc::fence_proxy_async();
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(&tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_a = tmem_base + 256;
  if (a_from_tmem) {
    const int row = warp * 32 + lane;
    for (int c = 0; c < tmem_a_cols; c += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = tmem_a_image[row * tmem_a_cols + c + j];
      tc::tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + c, r);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (threadIdx.x == 0) {
    const uint32_t base = tc::smem_u32(smem);
    for (int k = 0; k < nk; ++k) {
      const uint64_t db = desc_b_tmpl | (uint64_t)(((base + b_off + k * b_step) >> 4) & 0x3FFF);
      if (a_from_tmem) {
        tc::umma_ts(tmem_base, tmem_a + k * a_step, db, idesc, k > 0);
      } else {
        const uint64_t da = desc_a_tmpl | (uint64_t)(((base + a_off + k * a_step) >> 4) & 0x3FFF);
        tc::umma_ss(tmem_base, da, db, idesc, k > 0);
      }
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < n_cols; c += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c + j < n_cols) d_out[row * n_cols + c + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(128, 1)
probe_tma_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, uint32_t tx_bytes, uint8_t* __restrict__ dump, int dump_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < dump_bytes; i += blockDim.x) smem[i] = 0xEE;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  tc::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&bar, tx_bytes);
    tc::tma_load_2d(smem, &tm, c0, c1, &bar);
  }
  tc::mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < dump_bytes; i += blockDim.x) dump[i] = smem[i];
}

}  // namespace

extern "C" int pmv_probe_umma(const void* smem_image, int smem_bytes, uint64_t desc_a, uint64_t desc_b, uint32_t a_off,
                              uint32_t b_off, uint32_t idesc, int num_k_steps, uint32_t a_step, uint32_t b_step, int a_from_tmem,
                              const uint32_t* tmem_a_image, int tmem_a_cols, float* d_out, int n_cols, void* stream) {
  PMV_CHECK_ARG(smem_bytes % 16 == 0 && smem_bytes <= 200 * 1024, "probe: bad image size");
  PMV_CHECK_ARG(n_cols > 0 && n_cols <= 256 && tmem_a_cols % 16 == 0 && tmem_a_cols <= 256, "probe: bad column counts");
  const int smem = smem_bytes + 1024;
  PMV_CHECK_CUDA(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024));
  probe_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)smem_image, smem_bytes, desc_a, desc_b, a_off, b_off,
                                                            idesc, num_k_steps, a_step, b_step, a_from_tmem, tmem_a_image,
                                                            tmem_a_cols, d_out, n_cols);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

extern "C" int pmv_probe_tma(const void* gsrc, int dtype_bytes, uint64_t dim0, uint64_t dim1, uint64_t stride1_elems,
                             uint32_t box0, uint32_t box1, int swizzle_mode, int c0, int c1, void* smem_dump, int dump_bytes,
                             void* stream) {
  CUtensorMap tm;
  int rc = pmv_make_tensor_map_2d(&tm, gsrc, dtype_bytes, dim0, dim1, stride1_elems, box0, box1, swizzle_mode);
  if (rc) return rc;
  PMV_CHECK_ARG(dump_bytes <= 64 * 1024, "probe: dump too large");
  PMV_CHECK_CUDA(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
  probe_tma_kernel<<<1, 128, dump_bytes + 1024, (cudaStream_t)stream>>>(tm, c0, c1, box0 * box1 * dtype_bytes, (uint8_t*)smem_dump, dump_bytes);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}
