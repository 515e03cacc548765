// Second stage of the block-partial reductions (parameter gradients): dst[i] += sum_b partials[b * n + i].
// Same-address global atomics from hundreds of CTAs serialise in L2 (~0.5 us each on B200: 100-450 us per
// launch measured for 2.8 K addresses x 200-600 CTAs), so every kernel that reduces over all tokens writes one
// partial vector per CTA and this kernel folds them, deterministically.
#pragma once
#include "common.cuh"

static __global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nblocks, int n,
                                                                       int64_t stride, float* __restrict__ dst) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // grid.y slices the partial rows; 4 independent accumulators keep 4 loads in flight per thread
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int step = gridDim.y;
  int b = blockIdx.y;
  for (; b + 7 * step < nblocks; b += 8 * step) {  // eight loads in flight
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = partials[(int64_t)(b + u * step) * stride + i];
    s0 += v[0] + v[4]; s1 += v[1] + v[5]; s2 += v[2] + v[6]; s3 += v[3] + v[7];
  }
  for (; b + 3 * step < nblocks; b += 4 * step) {
    s0 += partials[(int64_t)b * stride + i];
    s1 += partials[(int64_t)(b + step) * stride + i];
    s2 += partials[(int64_t)(b + 2 * step) * stride + i];
    s3 += partials[(int64_t)(b + 3 * step) * stride + i];
  }
  for (; b < nblocks; b += step) s0 += partials[(int64_t)b * stride + i];
  atomicAdd(dst + i, (s0 + s1) + (s2 + s3));  // at most gridDim.y (<= 16) adds per address
}

static inline void launch_reduce_partials(const float* partials, int nblocks, int n, float* dst, cudaStream_t stream,
                                          int64_t stride = 0) {
  int slices = nblocks / 8;
  if (slices < 1) slices = 1;
  if (slices > 16) slices = 16;
  pmv_launch(reduce_partials_kernel, dim3((n + 255) / 256, slices), 256, 0, stream, partials, nblocks, n, stride ? stride : n, dst);
}
