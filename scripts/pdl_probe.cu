// Probe: what does a kernel -> kernel edge cost inside a CUDA graph on this GPU, with and without programmatic dependent
// launch (griddepcontrol.wait at the top of every kernel, programmatic stream serialization on the launch)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_probe scripts/pdl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) work_kernel(float* p, int n, int iters) {
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = p[i];
    for (int k = 0; k < iters; ++k) v = v * 1.0001f + 0.5f;
    p[i] = v;
  }
}

template <bool PDL>
float run(float* p, int n, int grid, int iters, int chain, cudaStream_t st) {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < chain; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, work_kernel<PDL>, p, n, iters);
  }
  cudaStreamEndCapture(st, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(exec, st);
  cudaEventRecord(e0, st);
  const int reps = 10;
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(exec, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  return ms * 1000.f / (reps * chain);
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  const int n = 148 * 8 * 256;
  float* p;
  cudaMalloc(&p, n * sizeof(float));
  cudaMemset(p, 0, n * sizeof(float));
  const int chain = 1000;
  for (int grid : {1, 148, 148 * 8}) {
    for (int iters : {0, 2000, 20000}) {
      const float a = run<false>(p, n, grid, iters, chain, st);
      const float b = run<true>(p, n, grid, iters, chain, st);
      printf("grid %5d iters %6d : plain %.2f us / kernel   pdl %.2f us / kernel\n", grid, iters, a, b);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
