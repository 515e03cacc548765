#define PMV_PDL_FAMILY 4
// gemm_tc_kernel instantiations for BN = 96 (one translation unit per tile width: they compile in parallel)
#include "gemm_tc_kernel.cuh"

int gemm_tc_launch_bn96(int layout, int out_dtype, int kind, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const CUtensorMap& tmD,                           const gemm_tc::TcParams& p, int num_sms, cudaStream_t s) {
  return gemm_tc::launch_bn<96>(layout, out_dtype, kind, tmA, tmB, tmC, tmD, p, num_sms, s);
}
