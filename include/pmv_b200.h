/* pmv_b200 — C-ABI of the B200-native MViTv2 pooling-attention hot path.
 *
 * The reference (bytedance/Portrait-Mode-Video, MViT/ fork) is 100 % Python: its hot path
 * is a sequence of ATen library calls.  This library replaces those call sites; every
 * entry point below names the reference lines it stands in for (paths relative to
 * MViT/slowfast/models/).  The reference-side binding is a ctypes stub — see
 * INTEGRATION.md.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless named `*_host`.  No allocation happens
 *     inside the library; temporaries are caller-provided workspaces.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Every kernel of a call is ordered on that
 *     stream as seen by the caller.  The pooling entry points may run launches that do not depend on one another on
 *     two library-owned side streams (created lazily per device, cudaStreamNonBlocking) between an event fork and an
 *     event join on `stream` — parallel branches when the call is captured into a CUDA graph; PMV_POOL_STREAMS=0 in the
 *     environment keeps everything on `stream`.  This and the tensor maps / function attributes created on first use
 *     are the only state the library keeps.
 *   - Every function returns 0 on success or a PMV_ERR_* code; pmv_last_error() returns a
 *     thread-local description of the last failure.
 *   - dtype arguments take PMV_F32 or PMV_BF16 ("fp32 mode" / "bf16 mode" of the path:
 *     activations in that type, accumulation and the residual stream always fp32).
 *   - Tokens are channels-last everywhere: [B, N, C] or [B, heads, N, 96].
 *   - head_dim is fixed at PMV_HEAD_DIM = 96 (true for every block of MViTv2-S/B,
 *     SURVEY.md Appendix A); pooling kernels are 3x3x3, pad 1, stride (1, s, s).
 */
#ifndef PMV_B200_H_
#define PMV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMV_HEAD_DIM 96

enum { PMV_F32 = 0, PMV_BF16 = 1 };

enum {
  PMV_OK = 0,
  PMV_ERR_INVALID_ARGUMENT = 1,
  PMV_ERR_CUDA = 2,
  PMV_ERR_UNSUPPORTED = 3,
};

/* epilogue activation of pmv_gemm */
enum {
  PMV_ACT_NONE = 0,
  PMV_ACT_GELU = 1,     /* out = gelu_erf(acc + bias)        (common.py:27-28, nn.GELU exact erf) */
  PMV_ACT_GELU_BWD = 2, /* out = acc * gelu_erf'(aux_in)     (autograd of the above)              */
};

/* operand layouts of pmv_gemm: C[M,N] = op(A) * op(B)
 *   PMV_GEMM_TN : A [M,K] row-major, B [N,K] row-major  (y = x W^T : forward of nn.Linear)
 *   PMV_GEMM_NN : A [M,K] row-major, B [K,N] row-major  (dx = dy W : dgrad)
 *   PMV_GEMM_NT_REDUCE_M : C[N1,N2] = A[M,N1]^T * B[M,N2] (dW = dy^T x : wgrad; reduction over rows) */
enum { PMV_GEMM_TN = 0, PMV_GEMM_NN = 1, PMV_GEMM_NT_REDUCE_M = 2 };

const char* pmv_last_error(void);
int pmv_version(void);
/* 1 if the tcgen05 (tensor-core) kernels are compiled in and the current device is sm_100. */
int pmv_has_tcgen05(void);

/* ---------------------------------------------------------------- LayerNorm ----------
 * norm1 / norm2 / final norm: attention.py:567,578; video_model_builder.py:2163.
 * x fp32 [rows, C] -> y (y_dtype) [rows, C]; mean/rstd [rows] are saved for backward
 * (may be NULL in inference). */
int pmv_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                      float* mean, float* rstd, int64_t rows, int C, float eps, void* stream);
/* dx (fp32) = LN'(dy) [+ dx_base: the gradient arriving over the residual connection, NULL for none, may alias dx];
 * dgamma_dbeta (fp32 [2][C]: dgamma then dbeta) is OVERWRITTEN.
 * ws: fp32 workspace of pmv_layernorm_bwd_workspace_bytes() bytes (one partial vector per CTA, folded by a
 * second small kernel: same-address global atomics from hundreds of CTAs serialise in L2). */
int64_t pmv_layernorm_bwd_workspace_bytes(int64_t rows, int C);
int pmv_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                      const float* mean, const float* rstd, float* dx, const float* dx_base,
                      float* dgamma_dbeta, float* ws, int64_t rows, int C, void* stream);

/* ---------------------------------------------------------------- GEMM family --------
 * qkv / proj / skip-proj / fc1 / fc2 Linear layers and their autograd:
 * attention.py:328,457,570; common.py:27-31.
 *
 *   acc = op(A) op(B)                                   (fp32 accumulate)
 *   v   = acc + bias[col]                               (bias may be NULL)
 *   v   = act(v)                                        (PMV_ACT_*; GELU_BWD multiplies by gelu'(aux_in[row,col]))
 *   if aux_out: aux_out[row,col] = acc + bias           (pre-activation saved for backward, io dtype)
 *   v   = row_scale ? v * row_scale[row / rows_per_scale] : v      (DropPath, common.py:46-59)
 *   v   = residual ? v + residual[row,col] : v          (fp32 residual stream, attention.py:577,585)
 *   out[(row remapped), col] = accumulate ? out + v : v
 *
 * Row remap: out_row = row + (row / out_group + 1) * out_skip when out_group > 0 (used to write
 * PatchEmbed tokens behind the cls slot, video_model_builder.py:2115-2121); identity otherwise.
 * io_dtype is the type of A, B, aux_*; out_dtype the type of out.  tc != 0 selects the tcgen05
 * tensor-core kernel (bf16 operands only); tc == 0 the fp32 FFMA kernel.
 * split_k > 1 (REDUCE_M only) makes CTAs atomically add partial sums into a zeroed fp32 out. */
typedef struct {
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
  int64_t out_group;
  int64_t out_skip;
} pmv_epilogue;

int pmv_gemm(int layout, const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo,
             int64_t M, int64_t N, int64_t K, int io_dtype, int out_dtype, const pmv_epilogue* epi,
             int tc, int split_k, void* stream);

/* column sums: out_sum[c] = sum_r in[r, c] * (row_scale ? row_scale[r / rows_per_scale] : 1); also
 * optionally writes the scaled copy cast to cast_dtype (bias gradients + operand cast of the fp32
 * residual-stream gradient in one pass). out_sum is OVERWRITTEN; it may be NULL; otherwise ws must hold
 * pmv_colsum_workspace_bytes() bytes. */
int64_t pmv_colsum_workspace_bytes(int64_t rows, int64_t cols);
int pmv_colsum_cast(const void* in, int in_dtype, int64_t ld_in, int64_t rows, int64_t cols,
                    const float* row_scale, int64_t rows_per_scale, float* out_sum, float* ws,
                    void* cast_out, int cast_dtype, int64_t ld_cast, void* stream);

/* ---------------------------------------------------------------- pooling ------------
 * attention_pool with depthwise Conv3d(96,96,3^3,stride (1,s,s),pad 1,groups 96,bias=False)
 * + LayerNorm(96, eps) for one of q / k / v: attention.py:14-48, 241-282, 351-371.
 * `in` points at the first channel of this tensor inside the QKV GEMM output
 * [B, 1+T*H*W, 3, heads, 96]; strides are in elements.  The cls token (token 0) bypasses the
 * convolution and is normalised (attention.py:25-26,39-42).
 * w is the reference Conv3d weight [96,1,3,3,3]; out is [B, heads, 1+T*Ho*Wo, out_ld] with the
 * 96 channels in the leading columns of each row (out_ld >= 96). */
int pmv_pool_ln_fwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                    const float* w, const float* gamma, const float* beta, void* out, int64_t out_ld,
                    int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream);
/* Backward.  din is written (not accumulated) with the same strides as `in` (so q/k/v gradients land
 * interleaved in the dQKV buffer); dw_dgamma_dbeta is fp32 [96*27 + 96 + 96] (Conv3d weight gradient in the
 * reference layout, then the LayerNorm weight and bias gradients) and is OVERWRITTEN.
 * ws: fp32 workspace of pmv_pool_ln_bwd_workspace_bytes() bytes (pre-LN gradient + per-CTA partial sums). */
int64_t pmv_pool_ln_bwd_workspace_bytes(int B, int heads, int T, int H, int W, int stride_hw);
int pmv_pool_ln_bwd(const void* in, int64_t in_batch_stride, int64_t in_token_stride, int64_t in_head_stride,
                    const float* w, const float* gamma, const void* dout, int64_t dout_ld,
                    void* din, float* dw_dgamma_dbeta, float* ws,
                    int B, int heads, int T, int H, int W, int stride_hw, float eps, int dtype, void* stream);

/* q, k and v pooled in ONE launch (the three jobs share the QKV buffer and the launch overhead).  Job i reads the
 * tensor that starts `which * which_stride` elements into `qkv` ([B, N, 3, heads, 96]: which_stride = heads*96).
 * Forward fills out / out_ld; backward reads dout / dout_ld, adds into grads (fp32 [96*27 + 96 + 96]) and writes
 * the input gradient into `dqkv` at the same offset.  ws: pmv_pool_ln_qkv_bwd_workspace_bytes() bytes. */
typedef struct {
  const float* w;      /* Conv3d weight [96,1,3,3,3] */
  const float* gamma;  /* LayerNorm weight [96] */
  const float* beta;   /* LayerNorm bias [96] (forward only) */
  void* out;           /* forward: [B, heads, 1+T*Ho*Wo, out_ld] */
  int64_t out_ld;
  const void* dout;    /* backward: gradient of out */
  int64_t dout_ld;
  float* grads;        /* backward: [96*27 + 96 + 96] fp32 (dW, dgamma, dbeta), OVERWRITTEN */
  int stride_hw;
  int which;           /* 0 = q, 1 = k, 2 = v */
  void* xhat;          /* optional [B, heads, 1+T*Ho*Wo, 96] (dtype): forward saves the normalised pre-affine tokens, */
  float* rstd;         /* optional [B, heads, 1+T*Ho*Wo]: ... and 1/sigma; backward then skips the convolution recompute */
  int dout_f32;        /* backward, saved-statistics path only: dout is fp32 whatever the compute dtype (e.g. the fp32
                        * dk / dv accumulators pmv_attention_bwd leaves in its workspace when dk == dv == NULL) */
  int onehot;          /* forward: also fill columns [96, out_ld) of `out` with the one-hot key coordinates of every token
                        * (zeros for the cls row) — what pmv_relpos_augment_k writes into K'; the pooling kernel knows the
                        * (t, h, w) of the token it normalises, so the separate launch disappears */
} pmv_pool_job;
int pmv_pool_ln_qkv_fwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride, int64_t head_stride,
                        const pmv_pool_job* jobs, int njobs, int B, int heads, int T, int H, int W, float eps, int dtype,
                        void* stream);
int64_t pmv_pool_ln_qkv_bwd_workspace_bytes(int B, int heads, int T, int H, int W, const int* strides_hw, int njobs);
int pmv_pool_ln_qkv_bwd(const void* qkv, int64_t batch_stride, int64_t token_stride, int64_t which_stride, int64_t head_stride,
                        const pmv_pool_job* jobs, int njobs, void* dqkv, float* ws, int B, int heads, int T, int H, int W,
                        float eps, int dtype, void* stream);

/* Skip-path MaxPool3d k (1,3,3) s (1,2,2) p (0,1,1) on [B, 1+T*H*W, C] fp32 tokens (cls copied):
 * attention.py:500-502,558-564,571-573.  The forward records the winning window position (0..8, first
 * maximum in window scan order, like ATen) of every output element in `win` (uint8, same shape as y;
 * NULL = not needed); the backward gathers through it and OVERWRITES dx (no atomics, no zero fill). */
int pmv_maxpool_skip_fwd(const float* x, float* y, uint8_t* win, int B, int T, int H, int W, int C, void* stream);
int pmv_maxpool_skip_bwd(const uint8_t* win, const float* dy, float* dx, int B, int T, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------- rel-pos augmentation
 * Decomposed relative position bias, cal_rel_pos_spatial / cal_rel_pos_temporal
 * (attention.py:67-159), folded into the score GEMM:
 *     bias[q,(kt,kh,kw)] = q.Rh[dist_h(qh,kh)] + q.Rw[dist_w(qw,kw)] + q.Rt[dist_t(qt,kt)]
 * is produced by extending the reduction dimension: Q' = [q | rq/scale], K' = [k | onehot(kh), onehot(kw),
 * onehot(kt)] with rq[q, j] the per-query dot products, so that scale*Q'K'^T = scale*q k^T + bias.
 * Rows are `ld` wide (ld = 96 + pad16(kh+kw+kt)); columns [0,96) must already hold q / k
 * (written by pmv_pool_ln_fwd with out_ld = ld).  The cls row gets zeros (attention.py:111,154).
 * idx_h [qh*kh], idx_w [qw*kw], idx_t [qt*kt] are the int32 table rows from the reference's
 * float-ratio / .long() index arithmetic (attention.py:80-99,132-139), computed on the host. */
/* ws: pmv_relpos_fwd_workspace_bytes() bytes (the stacked, pre-scaled tables and the dense
 * [BH*Nq, rows_h+rows_w+rows_t] product Q x Rcat^T, which runs through pmv_gemm: tc selects the tcgen05 kernel). */
int64_t pmv_relpos_fwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw);
int pmv_relpos_augment_q(void* q_aug, int64_t ld, const float* rel_h, const float* rel_w, const float* rel_t,
                         const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t, float* ws,
                         int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                         float inv_scale, int dtype, int tc, void* stream);
int pmv_relpos_augment_k(void* k_aug, int64_t ld, int BH, int kt, int kh, int kw, int dtype, void* stream);
/* Backward of augment_q: given dQ' (same layout), writes the table gradients to d_rel (fp32
 * [rows_h + rows_w + rows_t][96]: the three tables stacked in that order; OVERWRITTEN) and adds the bias path's
 * contribution to dq in place (columns [0,96) of dq_aug, non-cls rows).
 * ws: workspace of pmv_relpos_bwd_workspace_bytes() bytes (stacked tables, their gradient, the dense dRQ;
 * both products run through pmv_gemm). */
int64_t pmv_relpos_bwd_workspace_bytes(int BH, int qt, int qh, int qw, int kt, int kh, int kw);
int pmv_relpos_augment_q_bwd(void* dq_aug, const void* q_aug, int64_t ld, const float* rel_h, const float* rel_w,
                             const float* rel_t, const int32_t* idx_h, const int32_t* idx_w, const int32_t* idx_t,
                             float* d_rel, float* ws,
                             int BH, int qt, int qh, int qw, int kt, int kh, int kw,
                             float inv_scale, int dtype, int tc, void* stream);

/* ---------------------------------------------------------------- attention ----------
 * softmax(scale * Q' K'^T) V  + residual pooling, attention.py:412,446-454, head-merged output
 * (attention.py:456).  Q' [B*heads, Nq, ld_qk], K' [B*heads, Nk, ld_qk], V [B*heads, Nk, 96]
 * (row stride ld_v), out [B, Nq, heads*96], lse [B*heads, Nq] (natural-log sum-exp of the scaled
 * scores, saved for backward; may be NULL).  residual != 0 adds Q'[:, :96] to rows >= 1
 * (un-scaled post-LN q, cls row excluded, attention.py:450-452).  out_pre (may be NULL) additionally receives the
 * output BEFORE that add: the backward pass needs it at its own precision (recovering it as out - q from a
 * bf16-rounded sum loses most of its bits).
 * tc != 0 selects the tcgen05/TMEM/TMA kernel (bf16 only). */
int pmv_attention_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd,
                      const void* v, int64_t ld_v, void* out, void* out_pre, float* lse,
                      int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, int tc, void* stream);
/* Backward: dout [B, Nq, heads*96]; `out` is the PRE-residual output (out_pre of the forward call).  Writes dQ' [B*heads, Nq, ld_qk] (columns [0,kd): the first 96 are dq
 * incl. the residual-pooling path (dout added to rows >= 1), the rest d(rq/scale)), dk [B*heads, Nk, ld_dk]
 * (96 columns; the one-hot columns of K' carry no gradient) and dv [B*heads, Nk, ld_dv].
 * ws: fp32 workspace of pmv_attention_bwd_workspace_bytes() bytes (row deltas + dk/dv accumulators).
 * tc path: dk == dv == NULL skips the down-cast; the gradients are then the fp32 arrays ws[0 : BH*Nk*96) (dk) and
 * ws[BH*Nk*96 : 2*BH*Nk*96) (dv), row stride 96 — the pooling backward reads them directly (pmv_pool_job.dout_f32). */
int64_t pmv_attention_bwd_workspace_bytes(int B, int heads, int Nq, int Nk);
int pmv_attention_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v,
                      const void* out, const void* dout, const float* lse,
                      void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                      int B, int heads, int Nq, int Nk, float scale, int residual, int dtype, int tc, void* stream);

/* ---------------------------------------------------------------- PatchEmbed ---------
 * Conv3d(3 -> 96, k (3,7,7), s (2,4,4), p (1,3,3)) as an implicit GEMM: stem_helper.py:293-325.
 * pmv_patch_im2col gathers the receptive fields of clip [B,3,T,H,W] (fp32) into
 * col [B*To*Ho*Wo, ld_col] (dtype; K = 441 real columns, zero padded to ld_col); the GEMM and its
 * wgrad then run through pmv_gemm. */
int pmv_patch_im2col(const float* clip, void* col, int64_t ld_col, int B, int Cin, int T, int H, int W,
                     int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                     int dtype, void* stream);

/* ---------------------------------------------------------------- head + loss (row f2) ----
 * Tail of the step: final LayerNorm of the cls row (video_model_builder.py:2163-2165), TransformerBasicHead
 * (head_helper.py:561-577: dropout -> Linear -> softmax in eval) and the soft-target cross entropy of
 * losses.py:69-71 (mean over clips of sum_c -y_c log_softmax(logits)_c; integer `labels` = one-hot targets =
 * nn.CrossEntropyLoss).  All fp32.
 *   x: tokens [B, N, C] (row 0 of each clip is read), batch_stride = N*C.  keep_mask: optional uint8 [B, C] dropout
 *   keep mask (kept values are scaled by 1/(1-dropout_p)); the caller owns the random numbers.
 *   logits [B, classes] always; probs (optional) = softmax(logits); loss (optional device scalar, needs exactly one of
 *   labels / soft_targets); xhat [B, C], rstd [B], xd [B, C] (optional): saved for the backward.
 * Backward: dloss = device scalar gradient of the mean loss.  dx [B, N, C] is OVERWRITTEN (row 0 = gradient, other rows
 * zero); dw [classes, C], dbias (optional), dgamma, dbeta are overwritten.  ws: pmv_head_loss_bwd_workspace_bytes(). */
int pmv_head_loss_fwd(const float* x, int64_t batch_stride, const float* gamma, const float* beta, const float* w,
                      const float* bias, const uint8_t* keep_mask, float dropout_p, const int64_t* labels,
                      const float* soft_targets, float* logits, float* probs, float* loss, float* xhat,
                      float* rstd, float* xd, int B, int C, int num_classes, float eps, void* stream);
int64_t pmv_head_loss_bwd_workspace_bytes(int B, int C, int num_classes);
int pmv_head_loss_bwd(const float* dloss, const float* logits, const int64_t* labels, const float* soft_targets,
                      const float* w, const float* gamma, const uint8_t* keep_mask, float dropout_p, const float* xhat,
                      const float* rstd, const float* xd, float* dx, int64_t batch_stride, int N, float* dw, float* dbias,
                      float* dgamma, float* dbeta, float* ws, int B, int C, int num_classes, void* stream);

/* Programmatic dependent launch for the library's kernels: bit mask of kernel families (1 attention fwd, 2 attention
 * bwd, 4 GEMM, 8 column sums, 16 LayerNorm, 32 | 64 pooling, 128 rel-pos, 256 others); -1 = all (default), 0 = off.
 * Also settable with the PMV_PDL environment variable before the first launch. */
void pmv_set_pdl(int family_mask);

/* ---------------------------------------------------------------- optimizer step (row f3) ----
 * Multi-tensor AdamW exactly as torch.optim.AdamW (decoupled weight decay, bias correction, eps outside the
 * square root) — models/optimizer.py:124-131 — with a weight decay PER TENSOR (the reference's zero-WD group for
 * 1-D parameters / biases / no_weight_decay() names, optimizer.py:42-57) and optional global L2-norm clipping of
 * the gradients first (torch.nn.utils.clip_grad_norm_, tools/train_net.py:196-199).  One pass also refreshes the
 * bf16 operand copy (`shadow`) of each weight, so the next forward casts nothing.
 *   lr: DEVICE scalar (fp32) so that a captured CUDA graph follows the schedule; step: DEVICE int32, incremented by
 *   the call; grad_norm_out: optional DEVICE fp32 scalar receiving the pre-clip global gradient norm;
 *   max_grad_norm <= 0: no clipping.  ws: pmv_adamw_workspace_bytes() bytes.  Gradients are not modified. */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  void* shadow;       /* optional bf16 [numel] */
  int64_t numel;
  float weight_decay;
  float lr_scale;     /* multiplies lr (layer-wise decay); 1 otherwise */
} pmv_adamw_tensor;
int64_t pmv_adamw_workspace_bytes(const pmv_adamw_tensor* tensors, int ntensors);
int pmv_adamw_step(const pmv_adamw_tensor* tensors, int ntensors, const float* lr, float beta1, float beta2, float eps,
                   float max_grad_norm, int32_t* step, float* grad_norm_out, float* ws, void* stream);

/* ---------------------------------------------------------------- bring-up probes ----
 * Single-tile tcgen05 probes used by tests/test_tcgen05_probe.py to pin the shared-memory /
 * tensor-memory operand layouts the tensor-core kernels rely on. */
int pmv_probe_umma(const void* smem_image, int smem_bytes, uint64_t desc_a, uint64_t desc_b, uint32_t a_off,
                   uint32_t b_off, uint32_t idesc, int num_k_steps, uint32_t a_step, uint32_t b_step, int a_from_tmem,
                   const uint32_t* tmem_a_image, int tmem_a_cols, float* d_out, int n_cols, void* stream);
int pmv_probe_tma(const void* gsrc, int dtype_bytes, uint64_t dim0, uint64_t dim1, uint64_t stride1_elems,
                  uint32_t box0, uint32_t box1, int swizzle_mode, int c0, int c1, void* smem_dump, int dump_bytes,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMV_B200_H_ */
