// internal launchers behind pmv_gemm
#pragma once
#include "epilogue.cuh"

int gemm_simt_launch(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                     int io_dtype, int out_dtype, const EpiDev& e, int split_k, cudaStream_t stream);
int gemm_tc_launch(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                   int out_dtype, const EpiDev& e, int split_k, cudaStream_t stream);
