// pmv_attention_fwd dispatcher: fp32-math CUDA-core kernel vs tcgen05 kernel.
#include "common.cuh"

int attn_simt_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, void* out,
                  void* out_pre, float* lse, int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, cudaStream_t stream);
int attn_tc_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, void* out,
                void* out_pre, float* lse, int B, int heads, int Nq, int Nk, float scale, int residual, cudaStream_t stream);

extern "C" int pmv_attention_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd,
                                 const void* v, int64_t ld_v, void* out, void* out_pre, float* lse,
                                 int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, int tc, void* stream) {
  PMV_CHECK_ARG(kd % 16 == 0 && kd >= PMV_HEAD_DIM && kd <= 160, "attention: kd=%d must be a multiple of 16 in [96,160]", kd);
  PMV_CHECK_ARG(ld_qk >= kd && ld_qk % 4 == 0 && ld_v % 4 == 0, "attention: bad row strides");
  PMV_CHECK_ARG(B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: bad shape");
  if (tc) {
    PMV_CHECK_ARG(dtype == PMV_BF16, "attention: the tcgen05 kernel takes bf16 operands");
    if (!pmv_has_tcgen05()) {
      pmv_set_error("attention: tcgen05 kernel requested on a device that is not sm_100");
      return PMV_ERR_UNSUPPORTED;
    }
    return attn_tc_fwd(q_aug, k_aug, ld_qk, kd, v, ld_v, out, out_pre, lse, B, heads, Nq, Nk, scale, residual, (cudaStream_t)stream);
  }
  return attn_simt_fwd(q_aug, k_aug, ld_qk, kd, v, ld_v, out, out_pre, lse, B, heads, Nq, Nk, scale, residual, dtype, (cudaStream_t)stream);
}

int attn_simt_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v,
                  const void* out, const void* dout, const float* lse,
                  void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                  int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, void* stream);
int attn_tc_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, const void* o_pre,
                const void* dout, const float* lse, void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                int B, int heads, int Nq, int Nk, float scale, int residual, cudaStream_t stream);

extern "C" int64_t pmv_attention_bwd_workspace_bytes(int B, int heads, int Nq, int Nk) {
  // dK and dV fp32 accumulators + the per-row delta of the tensor-core path
  return ((int64_t)2 * B * heads * Nk * PMV_HEAD_DIM + (int64_t)B * heads * Nq) * (int64_t)sizeof(float);
}

extern "C" int pmv_attention_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v,
                                 const void* out, const void* dout, const float* lse,
                                 void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                                 int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, int tc, void* stream) {
  PMV_CHECK_ARG(kd % 16 == 0 && kd >= PMV_HEAD_DIM && kd <= 160, "attention: kd=%d must be a multiple of 16 in [96,160]", kd);
  PMV_CHECK_ARG(ld_qk % 4 == 0 && ld_v % 4 == 0 && ld_dk % 4 == 0 && ld_dv % 4 == 0, "attention: row strides must be multiples of 4");
  if (tc) {
    PMV_CHECK_ARG(dtype == PMV_BF16, "attention: the tcgen05 kernel takes bf16 operands");
    if (!pmv_has_tcgen05()) {
      pmv_set_error("attention: tcgen05 kernel requested on a device that is not sm_100");
      return PMV_ERR_UNSUPPORTED;
    }
    return attn_tc_bwd(q_aug, k_aug, ld_qk, kd, v, ld_v, out, dout, lse, dq_aug, dk, ld_dk, dv, ld_dv, ws, B, heads, Nq, Nk,
                       scale, residual, (cudaStream_t)stream);
  }
  PMV_CHECK_ARG(dk != nullptr && dv != nullptr, "attention bwd: only the tcgen05 path can leave dk / dv in the workspace");
  return attn_simt_bwd(q_aug, k_aug, ld_qk, kd, v, ld_v, out, dout, lse, dq_aug, dk, ld_dk, dv, ld_dv, ws, B, heads, Nq, Nk, scale,
                       residual, dtype, stream);
}
