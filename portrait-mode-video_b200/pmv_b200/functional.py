"""torch.autograd.Function wrappers: forward AND backward of every op run in the hand-written kernels.

The residual stream stays fp32; activations between kernels are in the compute dtype T (bf16 or fp32).
"""
from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import torch
from torch.autograd import Function

from . import _lib as L
from . import ops


def _cast(w: Optional[torch.Tensor], dtype):
    """Operand copy of a weight in the compute dtype: the copy FusedAdamW keeps current (optim.py), else a cast."""
    if w is None or w.dtype == dtype:
        return w
    sh = getattr(w, "_pmv_lp", None)
    if sh is not None and sh.dtype == dtype and getattr(w, "_pmv_lp_version", -1) == w._version:
        return sh
    return w.to(dtype)


class LayerNormFn(Function):
    """y(T) = LayerNorm(x fp32) — attention.py:567,578."""

    @staticmethod
    def forward(ctx, x, gamma, beta, out_dtype, eps):
        x = x.contiguous()
        need = any(ctx.needs_input_grad)
        y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, out_dtype, save_stats=need, eps=eps)
        if need:
            ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(dy, x, gamma, mean, rstd)
        return dx, dg, db, None, None


class ResidualLayerNormFn(Function):
    """(x, LayerNorm(x)): the residual stream passes through unchanged next to its normalised copy, so that the
    backward kernel adds the gradient arriving over the residual connection on the fly (one pass instead of
    LayerNorm-backward + a separate fp32 add) — attention.py:567-577 / 578-585."""

    @staticmethod
    def forward(ctx, x, gamma, beta, out_dtype, eps):
        x = x.contiguous()
        need = any(ctx.needs_input_grad)
        y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, out_dtype, save_stats=need, eps=eps)
        if need:
            ctx.save_for_backward(x, gamma, mean, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dx_res, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        if dy is None:
            return dx_res, None, None, None, None
        base = None
        if dx_res is not None:
            base = dx_res if (dx_res.dtype == torch.float32 and dx_res.is_contiguous()) else dx_res.float().contiguous()
        dx, dg, db = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dx_base=base)
        return dx, dg, db, None, None


def layer_norm(x, gamma, beta, out_dtype, eps=1e-6):
    return LayerNormFn.apply(x, gamma, beta, out_dtype, float(eps))


def layer_norm_residual(x, gamma, beta, out_dtype, eps=1e-6):
    """Returns (x, LayerNorm(x)); use the returned x for the residual connection."""
    return ResidualLayerNormFn.apply(x, gamma, beta, out_dtype, float(eps))


class LinearFn(Function):
    """y = x W^T + b, optionally  y = residual + row_scale[sample] * (x W^T + b)  (fp32 output).

    Replaces addmm at attention.py:328 (qkv), :457 (proj, fused with the residual add of :577 and the
    DropPath scale of common.py:46-59) and :570 (skip projection)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, row_scale, rows_per_scale, out_fp32):
        T = x.dtype
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w = _cast(weight, T)
        fp32_out = bool(out_fp32) or residual is not None
        res2 = residual.reshape(-1, N) if residual is not None else None
        y = ops.linear_fwd(x2, w, bias, torch.float32 if fp32_out else T, residual=res2, row_scale=row_scale,
                           rows_per_scale=rows_per_scale)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x2, w, row_scale)
        ctx.meta = (x.shape, rows_per_scale, bias is not None, residual is not None, T)
        ctx.arena = getattr(weight, "_pmv_arena", None)
        ctx.param = weight  # the nn.Parameter (not its bf16 copy): where its gradient may be written directly
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, row_scale = ctx.saved_tensors
        xshape, rps, has_bias, has_res, T = ctx.meta
        N = w.shape[0]
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        join = None
        if dy2.dtype != T or row_scale is not None:
            db, dyT = ops.colsum_cast(dy2, T, row_scale, rps, want_sum=has_bias)
        else:
            dyT = dy2
            db = None
            if has_bias:  # sum only: runs beside the two GEMMs below
                db, join = ops.colsum_beside(dy2)
        dx = ops.linear_dgrad(dyT, w, T).view(xshape) if ctx.needs_input_grad[0] else None
        dw = ops.linear_wgrad(dyT, x2, arena=ctx.arena, home=ops.grad_home(ctx.param))
        if join is not None:
            join()
        return dx, dw, db, (dy if has_res else None), None, None, None


def linear(x, weight, bias=None, residual=None, row_scale=None, rows_per_scale=1, out_fp32=False):
    return LinearFn.apply(x, weight, bias, residual, row_scale, rows_per_scale, out_fp32)


class MlpFn(Function):
    """x_out(fp32) = residual + row_scale * fc2(gelu_erf(fc1(x))) — common.py:26-34 fused with attention.py:585.
    Backward fuses GELU' into the fc2 dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, row_scale, rows_per_scale):
        T = x.dtype
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w1c, w2c = _cast(w1, T), _cast(w2, T)
        need = any(ctx.needs_input_grad)
        M, Hd, N = x2.shape[0], w1.shape[0], w2.shape[0]
        u = torch.empty(M, Hd, dtype=T, device=x.device) if need else None
        h = ops.linear_fwd(x2, w1c, b1, T, act=L.ACT_GELU, aux_out=u)
        res2 = residual.reshape(-1, N) if residual is not None else None
        y = ops.linear_fwd(h, w2c, b2, torch.float32, residual=res2, row_scale=row_scale, rows_per_scale=rows_per_scale)
        if need:
            ctx.save_for_backward(x2, w1c, w2c, u, h, row_scale)
        ctx.meta = (x.shape, rows_per_scale, residual is not None, T)
        ctx.arenas = (getattr(w1, "_pmv_arena", None), getattr(w2, "_pmv_arena", None))
        ctx.params = (w1, w2)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w1c, w2c, u, h, row_scale = ctx.saved_tensors
        xshape, rps, has_res, T = ctx.meta
        N = w2c.shape[0]
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if dy2.dtype != T or row_scale is not None:
            db2, dyT = ops.colsum_cast(dy2, T, row_scale, rps)
        else:
            dyT = dy2
            db2 = ops.colsum_cast(dy2, None)[0]
        dw2 = ops.linear_wgrad(dyT, h, arena=ctx.arenas[1], home=ops.grad_home(ctx.params[1]))
        du = ops.linear_dgrad(dyT, w2c, T, act=L.ACT_GELU_BWD, aux_in=u)
        db1, join = ops.colsum_beside(du)  # the fc1 bias gradient, beside the two GEMMs that also read du
        dw1 = ops.linear_wgrad(du, x2, arena=ctx.arenas[0], home=ops.grad_home(ctx.params[0]))
        dx = ops.linear_dgrad(du, w1c, T).view(xshape)
        join()
        return dx, dw1, db1, dw2, db2, (dy if has_res else None), None, None


def mlp(x, w1, b1, w2, b2, residual=None, row_scale=None, rows_per_scale=1):
    return MlpFn.apply(x, w1, b1, w2, b2, residual, row_scale, rows_per_scale)


class MaxPoolSkipFn(Function):
    """Residual-path MaxPool3d (1,3,3)/(1,2,2)/(0,1,1) on fp32 tokens — attention.py:571-573."""

    @staticmethod
    def forward(ctx, x, thw):
        x = x.contiguous()
        ctx.thw = tuple(thw)
        if not any(ctx.needs_input_grad):
            return ops.maxpool_skip_fwd(x, thw)
        y, win = ops.maxpool_skip_fwd(x, thw, want_winner=True)  # 1 byte per output element instead of keeping x alive
        ctx.save_for_backward(win)
        return y

    @staticmethod
    def backward(ctx, dy):
        (win,) = ctx.saved_tensors
        return ops.maxpool_skip_bwd(win, dy, ctx.thw), None


def maxpool_skip(x, thw):
    return MaxPoolSkipFn.apply(x, tuple(thw))


class PoolAttentionFn(Function):
    """q/k/v pooling (+LN) -> rel-pos augmentation -> attention with residual pooling -> head merge.

    Replaces attention.py:351-371 (three attention_pool calls), :412-446 (scores, cal_rel_pos_spatial,
    cal_rel_pos_temporal, softmax), :448 (attn @ v), :450-454 (residual pooling) and :456 (head merge)."""

    @staticmethod
    def forward(ctx, qkv, wq, wk, wv, gq, bq, gk, bk, gv, bv, rel_h, rel_w, rel_t, heads, thw, stride_q, stride_kv,
                scale, residual, use_tc_attn, eps):
        T = qkv.dtype
        B, N, C3 = qkv.shape
        assert C3 == 3 * heads * 96, "head_dim must be 96"
        qkv5 = qkv.contiguous().view(B, N, 3, heads, 96)
        Tn, H, W = thw
        q_shape = (Tn, ops.pooled_hw(H, stride_q), ops.pooled_hw(W, stride_q))
        k_shape = (Tn, ops.pooled_hw(H, stride_kv), ops.pooled_hw(W, stride_kv))
        Nq, Nk = 1 + math.prod(q_shape), 1 + math.prod(k_shape)
        has_rel = rel_h is not None
        ld = ops.aug_ld(k_shape) if has_rel else 96
        dev = qkv.device
        q_aug = torch.empty(B * heads, Nq, ld, dtype=T, device=dev)
        k_aug = torch.empty(B * heads, Nk, ld, dtype=T, device=dev)
        v = torch.empty(B * heads, Nk, 96, dtype=T, device=dev)
        need = any(ctx.needs_input_grad)
        saved = []
        if need:  # normalised pre-affine tokens + 1/sigma: the backward LayerNorm needs no convolution recompute
            for n_tok in (Nq, Nk, Nk):
                saved += [torch.empty(B, heads, n_tok, 96, dtype=T, device=dev),
                          torch.empty(B, heads, n_tok, dtype=torch.float32, device=dev)]
        extra = [tuple(saved[2 * i:2 * i + 2]) for i in range(3)] if need else [(), (), ()]
        ops.pool_ln_qkv_fwd(qkv5, heads, thw, [(0, stride_q, wq, gq, bq, q_aug.view(B, heads, Nq, ld)) + extra[0],
                                               (1, stride_kv, wk, gk, bk, k_aug.view(B, heads, Nk, ld)) + extra[1],
                                               (2, stride_kv, wv, gv, bv, v.view(B, heads, Nk, 96)) + extra[2]], eps,
                            onehot_jobs=(1,) if has_rel else ())  # K' one-hot coordinate columns come with the pooling
        if has_rel:
            ops.relpos_augment_q(q_aug, q_shape, k_shape, rel_h, rel_w, rel_t, 1.0 / scale)
        out, out_pre, lse = ops.attention_fwd(q_aug, k_aug, v, B, heads, ld, scale, residual=residual, want_lse=need,
                                     tc=(1 if (use_tc_attn and T == torch.bfloat16) else 0))
        if need:
            ctx.save_for_backward(qkv5, q_aug, k_aug, v, out_pre, lse, wq, wk, wv, gq, gk, gv, rel_h, rel_w, rel_t, *saved)
        ctx.meta = (heads, tuple(thw), stride_q, stride_kv, q_shape, k_shape, scale, residual, ld, has_rel, eps,
                    bool(use_tc_attn and T == torch.bfloat16 and ld in (128, 160)))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv5, q_aug, k_aug, v, out, lse, wq, wk, wv, gq, gk, gv, rel_h, rel_w, rel_t, *saved = ctx.saved_tensors
        extra = [tuple(saved[2 * i:2 * i + 2]) for i in range(3)]
        heads, thw, sq, skv, q_shape, k_shape, scale, residual, ld, has_rel, eps, tc_bwd = ctx.meta
        B, N = qkv5.shape[0], qkv5.shape[1]
        Nq, Nk = q_aug.shape[1], k_aug.shape[1]
        dq_aug, dk, dv = ops.attention_bwd(q_aug, k_aug, v, out, dout, lse, B, heads, ld, scale, residual=residual,
                                           tc=int(tc_bwd and os.environ.get("PMV_TC_ATTENTION_BWD", "1") == "1"),
                                           fp32_dkv=True)
        drh = drw = drt = None
        if has_rel:
            drh, drw, drt = ops.relpos_augment_q_bwd(dq_aug, q_aug, q_shape, k_shape, rel_h, rel_w, rel_t, 1.0 / scale)
        dqkv = torch.empty_like(qkv5)
        g = torch.empty(3, 96 * 27 + 192, dtype=torch.float32, device=qkv5.device)
        ops.pool_ln_qkv_bwd(qkv5, heads, thw, [(0, sq, wq, gq, dq_aug.view(B, heads, Nq, ld), g[0]) + extra[0],
                                               (1, skv, wk, gk, dk.view(B, heads, Nk, 96), g[1]) + extra[1],
                                               (2, skv, wv, gv, dv.view(B, heads, Nk, 96), g[2]) + extra[2]], dqkv, eps)
        dw = [g[i, :2592].view(96, 1, 3, 3, 3) for i in range(3)]
        dg = [g[i, 2592:2688] for i in range(3)]
        db = [g[i, 2688:] for i in range(3)]
        return (dqkv.view(B, N, 3 * heads * 96), dw[0], dw[1], dw[2], dg[0], db[0], dg[1], db[1], dg[2], db[2],
                drh, drw, drt, None, None, None, None, None, None, None, None)


def pool_attention(qkv, wq, wk, wv, gq, bq, gk, bk, gv, bv, rel_h, rel_w, rel_t, heads, thw, stride_q, stride_kv,
                   scale, residual=True, use_tc_attn=False, eps=1e-6):
    return PoolAttentionFn.apply(qkv, wq, wk, wv, gq, bq, gk, bk, gv, bv, rel_h, rel_w, rel_t, heads, tuple(thw),
                                 int(stride_q), int(stride_kv), float(scale), bool(residual), bool(use_tc_attn), float(eps))


class HeadLossFn(Function):
    """loss, logits = CE(Linear(dropout(LayerNorm(x)[:, 0])), target): video_model_builder.py:2163-2169,
    head_helper.py:561-577, losses.py:69-71 in two launches each way (row f2).  The logits are returned for metrics
    (top-k accuracy in train_net.py) and carry no gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, w, bias, target, keep_mask, dropout_p, eps):
        x = x.contiguous()
        r = ops.head_loss_fwd(x, gamma, beta, w, bias, target=target, keep_mask=keep_mask, dropout_p=dropout_p,
                              save=any(ctx.needs_input_grad), eps=eps)
        ctx.save_for_backward(r["logits"], r["labels"], r["soft"], w, gamma, keep_mask, *r["saved"])
        ctx.meta = (x.shape[1], float(dropout_p), bias is not None)
        ctx.mark_non_differentiable(r["logits"])
        return r["loss"], r["logits"]

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        logits, labels, soft, w, gamma, keep_mask, xhat, rstd, xd = ctx.saved_tensors
        N, p, has_bias = ctx.meta
        dloss = dloss.to(torch.float32).contiguous()
        dx, dw, db, dgamma, dbeta = ops.head_loss_bwd(dloss, logits, labels, soft, w, gamma, keep_mask, p, (xhat, rstd, xd), N,
                                                      has_bias=has_bias)
        return dx, dgamma, dbeta, dw, db, None, None, None, None


def head_loss(x, gamma, beta, w, bias, target, keep_mask=None, dropout_p=0.0, eps=1e-6):
    return HeadLossFn.apply(x, gamma, beta, w, bias, target, keep_mask, float(dropout_p), float(eps))


class PatchEmbedFn(Function):
    """PatchEmbed Conv3d + flatten/transpose + cls-token concat (stem_helper.py:320-325,
    video_model_builder.py:2115-2121) -> fp32 tokens [B, 1+L, 96]."""

    @staticmethod
    def forward(ctx, clip, weight, bias, cls_token, kernel, stride, padding, T):
        B = clip.shape[0]
        Cout = weight.shape[0]
        col, thw, K = ops.patch_im2col(clip.contiguous(), kernel, stride, padding, T)
        wp = torch.zeros(Cout, col.shape[1], dtype=T, device=clip.device)
        wp[:, :K] = weight.reshape(Cout, K).to(T)
        Ltok = thw[0] * thw[1] * thw[2]
        x = torch.empty(B, Ltok + 1, Cout, dtype=torch.float32, device=clip.device)
        x[:, 0] = cls_token.reshape(1, Cout)
        ops.linear_fwd(col, wp, bias, torch.float32, out=x.view(B * (Ltok + 1), Cout), out_group=Ltok, out_skip=1)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(col)
        ctx.meta = (K, weight.shape, thw, T)
        ctx.thw = thw
        return x

    @staticmethod
    def backward(ctx, dx):
        (col,) = ctx.saved_tensors
        K, wshape, thw, T = ctx.meta
        B, Ntok, Cout = dx.shape
        dcls = dx[:, 0].sum(0).reshape(1, 1, Cout)
        dtok = dx[:, 1:].reshape(-1, Cout)
        db, dyT = ops.colsum_cast(dtok.contiguous(), T)
        dw = ops.linear_wgrad(dyT, col)[:, :K].reshape(wshape)
        return None, dw, db, dcls, None, None, None, None


def patch_embed(clip, weight, bias, cls_token, kernel, stride, padding, T):
    x = PatchEmbedFn.apply(clip, weight, bias, cls_token, tuple(kernel), tuple(stride), tuple(padding), T)
    B, Cin, Tn, H, W = clip.shape
    thw = [(Tn + 2 * padding[0] - kernel[0]) // stride[0] + 1, (H + 2 * padding[1] - kernel[1]) // stride[1] + 1,
           (W + 2 * padding[2] - kernel[2]) // stride[2] + 1]
    return x, thw
