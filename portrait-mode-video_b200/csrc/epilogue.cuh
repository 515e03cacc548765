// Shared GEMM epilogue (see pmv_gemm in include/pmv_b200.h for the contract).
#pragma once
#include "common.cuh"

struct EpiDev {
  const float* bias;
  int act;
  const void* aux_in;
  void* aux_out;
  int64_t ld_aux;
  const float* row_scale;
  int64_t rows_per_scale;
  const float* residual;
  int64_t ld_residual;
  int accumulate;
  int atomic;  // split-K: atomicAdd partial sums into fp32 out (no other epilogue terms allowed)
  int64_t out_group;
  int64_t out_skip;
  void* out;
  int64_t ldo;
  FastDiv fd_scale;  // rows_per_scale (row indices < 2^31)
  FastDiv fd_group;  // out_group
};

__device__ __forceinline__ int64_t epi_out_row(const EpiDev& e, int64_t row) {
  return e.out_group > 0 ? row + (row / e.out_group + 1) * e.out_skip : row;
}

// one element
template <typename TIO, typename TOut>
__device__ __forceinline__ void epi_store(const EpiDev& e, int64_t row, int64_t col, float acc) {
  if (e.atomic) {
    atomicAdd(reinterpret_cast<float*>(e.out) + row * e.ldo + col, acc);
    return;
  }
  float v = acc;
  if (e.bias) v += e.bias[col];
  if (e.aux_out) reinterpret_cast<TIO*>(e.aux_out)[row * e.ld_aux + col] = from_f32<TIO>(v);
  if (e.act == PMV_ACT_GELU) {
    v = gelu_erf(v);
  } else if (e.act == PMV_ACT_GELU_BWD) {
    v *= gelu_erf_grad(to_f32(reinterpret_cast<const TIO*>(e.aux_in)[row * e.ld_aux + col]));
  }
  if (e.row_scale) v *= e.row_scale[row / e.rows_per_scale];
  if (e.residual) v += e.residual[row * e.ld_residual + col];
  TOut* o = reinterpret_cast<TOut*>(e.out) + epi_out_row(e, row) * e.ldo + col;
  if (e.accumulate) v += to_f32(*o);
  *o = from_f32<TOut>(v);
}

// 4 consecutive columns (col % 4 == 0, all leading dimensions % 4 == 0)
template <typename TIO, typename TOut, bool FAST = false>
__device__ __forceinline__ void epi_store4(const EpiDev& e, int64_t row, int64_t col, const float (&acc)[4]) {
  float v[4] = {acc[0], acc[1], acc[2], acc[3]};
  if (e.atomic) {
    float* o = reinterpret_cast<float*>(e.out) + row * e.ldo + col;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(o + j, v[j]);
    return;
  }
  if (e.bias) {
    float b[4];
    load4(e.bias + col, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += b[j];
  }
  if (e.aux_out) store4(reinterpret_cast<TIO*>(e.aux_out) + row * e.ld_aux + col, v);
  if (e.act == PMV_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = FAST ? gelu_fast(v[j]) : gelu_erf(v[j]);
  } else if (e.act == PMV_ACT_GELU_BWD) {
    float u[4];
    load4(reinterpret_cast<const TIO*>(e.aux_in) + row * e.ld_aux + col, u);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= FAST ? gelu_fast_grad(u[j]) : gelu_erf_grad(u[j]);
  }
  if (e.row_scale) {
    const float s = e.row_scale[row / e.rows_per_scale];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= s;
  }
  if (e.residual) {
    float r[4];
    load4(e.residual + row * e.ld_residual + col, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += r[j];
  }
  TOut* o = reinterpret_cast<TOut*>(e.out) + epi_out_row(e, row) * e.ldo + col;
  if (e.accumulate) {
    float p[4];
    load4(o, p);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += p[j];
  }
  store4(o, v);
}
