// Second stage of the block-partial reductions (parameter gradients): dst[i] += sum_b partials[b * n + i].
// Same-address global atomics from hundreds of CTAs serialise in L2 (~0.5 us each on B200: 100-450 us per
// launch measured for 2.8 K addresses x 200-600 CTAs), so every kernel that reduces over all tokens writes one
// partial vector per CTA and this kernel folds them, deterministically.
#pragma once
#include "common.cuh"

static __global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nblocks, int n,
                                                                       float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partials[(int64_t)b * n + i];
  dst[i] += s;
}

static inline void launch_reduce_partials(const float* partials, int nblocks, int n, float* dst, cudaStream_t stream) {
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, stream>>>(partials, nblocks, n, dst);
}
