"""Drop-in mirror of slowfast/models/attention.py (MultiScaleAttention, MultiScaleBlock) of the
bytedance/Portrait-Mode-Video MViT fork, computing in the hand-written sm_100a kernels of libpmv_b200.so.

Same constructor keyword arguments, same ``forward`` contract and the same ``state_dict`` keys as the
reference (attention.py:162-312, 464-589), so ``video_model_builder.MViT`` can import these symbols instead
of the reference's without any other change.  Configurations outside the MViTv2 path
(pool_first, separate_qkv, avg/max/conv_unshared pooling modes, dropout, layer scale, no cls token,
head_dim != 96, pooling kernels other than 3x3x3 with stride (1,s,s)) raise NotImplementedError instead of
silently diverging.
"""
from __future__ import annotations

import os

import numpy
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.init import trunc_normal_

from . import functional as Fn
from .common import DropPath, Mlp, compute_dtype_of, drop_path_scale

USE_TC_ATTENTION = os.environ.get("PMV_TC_ATTENTION", "1") == "1"  # tcgen05 attention forward in bf16 mode


def _as_list(v):
    return [int(a) for a in v]


def get_rel_pos(rel_pos, d):
    """attention.py:51-64 — identity when lengths match, else 1-D linear interpolation (host-side torch)."""
    ori_d = rel_pos.shape[0]
    if ori_d == d:
        return rel_pos
    new = F.interpolate(rel_pos.reshape(1, ori_d, -1).permute(0, 2, 1), size=d, mode="linear")
    return new.reshape(-1, d).permute(1, 0).contiguous()


class MultiScaleAttention(nn.Module):
    compute_dtype = torch.bfloat16

    def __init__(self, dim, dim_out, input_size, num_heads=8, qkv_bias=False, drop_rate=0.0, kernel_q=(1, 1, 1),
                 kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1), norm_layer=nn.LayerNorm,
                 has_cls_embed=True, mode="conv", pool_first=False, rel_pos_spatial=False, rel_pos_temporal=False,
                 rel_pos_zero_init=False, residual_pooling=False, separate_qkv=False, hw_switch_auto=False):
        super().__init__()
        if mode not in ("conv", "conv_unshared", "avg", "max"):
            raise NotImplementedError(f"Unsupported model {mode}")  # attention.py:284
        if pool_first or separate_qkv or mode != "conv" or drop_rate > 0.0 or not has_cls_embed:
            raise NotImplementedError(
                "pmv_b200.MultiScaleAttention implements the MViTv2 path only: pool_first=False, "
                "separate_qkv=False, mode='conv', drop_rate=0, has_cls_embed=True")
        kernel_q, kernel_kv, stride_q, stride_kv = map(_as_list, (kernel_q, kernel_kv, stride_q, stride_kv))
        if kernel_q != [3, 3, 3] or kernel_kv != [3, 3, 3] or len(stride_q) != 3 or len(stride_kv) != 3 \
                or stride_q[0] != 1 or stride_kv[0] != 1 or stride_q[1] != stride_q[2] or stride_kv[1] != stride_kv[2]:
            raise NotImplementedError(
                "pmv_b200.MultiScaleAttention: pooling kernels must be 3x3x3 with stride (1, s, s) "
                f"(got kernel_q={kernel_q} kernel_kv={kernel_kv} stride_q={stride_q} stride_kv={stride_kv})")
        if dim_out % num_heads != 0 or dim_out // num_heads != 96:
            raise NotImplementedError("pmv_b200.MultiScaleAttention: head_dim must be 96 (all MViTv2-S/B blocks)")
        self.pool_first = pool_first
        self.separate_qkv = separate_qkv
        self.drop_rate = drop_rate
        self.num_heads = num_heads
        self.dim_out = dim_out
        head_dim = dim_out // num_heads
        self.scale = head_dim ** -0.5
        self.has_cls_embed = has_cls_embed
        self.mode = mode
        self.hw_switch_auto = hw_switch_auto
        self.stride_q, self.stride_kv = stride_q, stride_kv

        self.qkv = nn.Linear(dim, dim_out * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim_out, dim_out)
        dim_conv = head_dim
        pad = [1, 1, 1]
        self.pool_q = nn.Conv3d(dim_conv, dim_conv, kernel_q, stride=stride_q, padding=pad, groups=dim_conv, bias=False)
        self.norm_q = norm_layer(dim_conv)
        self.pool_k = nn.Conv3d(dim_conv, dim_conv, kernel_kv, stride=stride_kv, padding=pad, groups=dim_conv, bias=False)
        self.norm_k = norm_layer(dim_conv)
        self.pool_v = nn.Conv3d(dim_conv, dim_conv, kernel_kv, stride=stride_kv, padding=pad, groups=dim_conv, bias=False)
        self.norm_v = norm_layer(dim_conv)
        for n in (self.norm_q, self.norm_k, self.norm_v):
            if not isinstance(n, nn.LayerNorm) or not n.elementwise_affine:
                raise NotImplementedError("pmv_b200: norm_layer must build an affine nn.LayerNorm")

        self.rel_pos_spatial = rel_pos_spatial
        self.rel_pos_temporal = rel_pos_temporal
        if self.rel_pos_spatial:  # attention.py:288-304
            size_h, size_w = input_size[1], input_size[2]
            q_size_h, kv_size_h = size_h // stride_q[1], size_h // stride_kv[1]
            q_size_w, kv_size_w = size_w // stride_q[2], size_w // stride_kv[2]
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * max(q_size_h, kv_size_h) - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * max(q_size_w, kv_size_w) - 1, head_dim))
            if not rel_pos_zero_init:
                trunc_normal_(self.rel_pos_h, std=0.02)
                trunc_normal_(self.rel_pos_w, std=0.02)
        if self.rel_pos_temporal:  # attention.py:305-310
            self.rel_pos_t = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            if not rel_pos_zero_init:
                trunc_normal_(self.rel_pos_t, std=0.02)
        self.residual_pooling = residual_pooling

    # -- tables exactly as cal_rel_pos_spatial / _temporal would see them (attention.py:76-77,96-97,127-129,414-424)
    def _rel_tables(self, thw_shape, q_shape, k_shape):
        if not (self.rel_pos_spatial or self.rel_pos_temporal):
            return None, None, None
        dev = self.qkv.weight.device
        dh = 2 * max(q_shape[1], k_shape[1]) - 1
        dw = 2 * max(q_shape[2], k_shape[2]) - 1
        dt = 2 * max(q_shape[0], k_shape[0]) - 1
        if self.rel_pos_spatial:
            rh, rw = self.rel_pos_h, self.rel_pos_w
            if self.hw_switch_auto and thw_shape[1] > thw_shape[2]:
                rh, rw = rw, rh
            rh, rw = get_rel_pos(rh, dh), get_rel_pos(rw, dw)
        else:
            rh = torch.zeros(dh, 96, device=dev)
            rw = torch.zeros(dw, 96, device=dev)
        rt = get_rel_pos(self.rel_pos_t, dt) if self.rel_pos_temporal else torch.zeros(dt, 96, device=dev)
        return rh.float(), rw.float(), rt.float()

    def _attend(self, x, thw_shape):
        """x: [B, N, dim] in the compute dtype -> head-merged attention output [B, Nq, dim_out] (pre-proj)."""
        T_, H, W = thw_shape
        sq, skv = self.stride_q[1], self.stride_kv[1]
        q_shape = [T_, (H - 1) // sq + 1, (W - 1) // sq + 1]
        k_shape = [T_, (H - 1) // skv + 1, (W - 1) // skv + 1]
        qkv = Fn.linear(x, self.qkv.weight, self.qkv.bias)
        rh, rw, rt = self._rel_tables(thw_shape, q_shape, k_shape)
        o = Fn.pool_attention(qkv, self.pool_q.weight, self.pool_k.weight, self.pool_v.weight,
                              self.norm_q.weight, self.norm_q.bias, self.norm_k.weight, self.norm_k.bias,
                              self.norm_v.weight, self.norm_v.bias, rh, rw, rt, self.num_heads, thw_shape, sq, skv,
                              self.scale, self.residual_pooling, USE_TC_ATTENTION, self.norm_q.eps)
        return o, q_shape

    def forward(self, x, thw_shape, residual=None, row_scale=None):
        """Reference contract (attention.py:314,461): x [B, N, dim] -> (y [B, Nq, dim_out], q_shape).
        ``residual`` / ``row_scale`` are used by MultiScaleBlock to fuse the residual add and DropPath
        into the projection GEMM epilogue."""
        T = compute_dtype_of(self)
        if x.dtype != T:
            x = x.to(T)
        o, q_shape = self._attend(x, list(thw_shape))
        y = Fn.linear(o, self.proj.weight, self.proj.bias, residual=residual, row_scale=row_scale,
                      rows_per_scale=o.shape[1], out_fp32=True)
        return y, q_shape


class MultiScaleBlock(nn.Module):
    compute_dtype = torch.bfloat16

    def __init__(self, dim, dim_out, num_heads, input_size, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0,
                 drop_path=0.0, layer_scale_init_value=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, up_rate=None,
                 kernel_q=(1, 1, 1), kernel_kv=(1, 1, 1), stride_q=(1, 1, 1), stride_kv=(1, 1, 1), mode="conv",
                 has_cls_embed=True, pool_first=False, rel_pos_spatial=False, rel_pos_temporal=False,
                 rel_pos_zero_init=False, residual_pooling=False, dim_mul_in_att=False, separate_qkv=False,
                 hw_switch_auto=False):
        super().__init__()
        if layer_scale_init_value > 0 or (up_rate is not None and up_rate > 1):
            raise NotImplementedError("pmv_b200.MultiScaleBlock: layer scale / up_rate are not on the MViTv2 path")
        if dim != dim_out and not dim_mul_in_att:
            raise NotImplementedError("pmv_b200.MultiScaleBlock: dim change requires dim_mul_in_att=True (MViTv2)")
        self.dim = dim
        self.dim_out = dim_out
        self.norm1 = norm_layer(dim)
        self.dim_mul_in_att = dim_mul_in_att
        stride_q = _as_list(stride_q)
        att_dim = dim_out if dim_mul_in_att else dim
        self.attn = MultiScaleAttention(
            dim, att_dim, num_heads=num_heads, input_size=input_size, qkv_bias=qkv_bias, drop_rate=drop_rate,
            kernel_q=kernel_q, kernel_kv=kernel_kv, stride_q=stride_q, stride_kv=stride_kv, norm_layer=norm_layer,
            has_cls_embed=has_cls_embed, mode=mode, pool_first=pool_first, rel_pos_spatial=rel_pos_spatial,
            rel_pos_temporal=rel_pos_temporal, rel_pos_zero_init=rel_pos_zero_init, residual_pooling=residual_pooling,
            separate_qkv=separate_qkv, hw_switch_auto=hw_switch_auto)
        self.drop_path_prob = float(drop_path)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(att_dim)
        self.has_cls_embed = has_cls_embed
        self.mlp = Mlp(in_features=att_dim, hidden_features=int(att_dim * mlp_ratio), out_features=dim_out,
                       act_layer=act_layer, drop_rate=drop_rate)
        self.gamma_1, self.gamma_2 = None, None
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)
        # attention.py:558-564: MaxPool3d(kernel [s+1 if s>1 else s], stride_q, pad k//2) iff prod(stride_q) > 1
        self.pool_skip = None
        if len(stride_q) > 0 and numpy.prod(stride_q) > 1:
            if stride_q != [1, 2, 2]:
                raise NotImplementedError("pmv_b200.MultiScaleBlock: skip-path max-pool supports stride_q (1,2,2) only")
            self.pool_skip = nn.MaxPool3d([1, 3, 3], stride_q, [0, 1, 1], ceil_mode=False)  # parameter-free marker

    def forward(self, x, thw_shape=None, drop_scales=None):
        """``drop_scales``: optional pre-drawn DropPath factors (attention branch, MLP branch), each [B] fp32 or None —
        MViT draws the factors of all blocks with one set of launches; without it the block draws its own."""
        T = compute_dtype_of(self)
        B = x.shape[0]
        x = x.float() if x.dtype != torch.float32 else x
        thw = list(thw_shape)
        x, x_norm = Fn.layer_norm_residual(x, self.norm1.weight, self.norm1.bias, T, self.norm1.eps)         # :567
        if self.dim_mul_in_att and self.dim != self.dim_out:
            x = Fn.linear(x_norm, self.proj.weight, self.proj.bias, out_fp32=True)           # :569-570
        x_res = Fn.maxpool_skip(x, thw) if self.pool_skip is not None else x                  # :571-573
        if drop_scales is not None:
            ds1, ds2 = drop_scales
        else:
            ds1 = drop_path_scale(B, self.drop_path_prob, self.training, x.device)
        x, thw_new = self.attn(x_norm, thw, residual=x_res, row_scale=ds1)                    # :568,577
        x, x_norm2 = Fn.layer_norm_residual(x, self.norm2.weight, self.norm2.bias, T, self.norm2.eps)        # :578
        if drop_scales is None:
            ds2 = drop_path_scale(B, self.drop_path_prob, self.training, x.device)
        x = self.mlp(x_norm2, residual=x, row_scale=ds2, rows_per_scale=x.shape[1])           # :579,585
        if thw_shape:
            return x, thw_new
        return x


def cache_low_precision_weights(module: nn.Module, dtype: torch.dtype = torch.bfloat16):
    """Inference: keep an operand copy in ``dtype`` of every matrix weight, so that a forward launches no cast kernels
    (68 per MViTv2-S forward, 3 % of the step).  The copies are used while the parameter's ``_version`` is unchanged
    (functional._cast); call this again after the weights were changed by something that does not bump tensor versions
    (a CUDA-graph replay of an optimizer step).  Training with pmv_b200.optim.FusedAdamW maintains the copies itself."""
    n = 0
    for p in module.parameters():
        if p.dim() >= 2 and p.dtype != dtype:
            lp = getattr(p, "_pmv_lp", None)
            if lp is not None and lp.dtype == dtype and lp.shape == p.shape and lp.device == p.device:
                lp.copy_(p.detach())  # in place: FusedAdamW (and a captured graph of its step) keep writing to this tensor
            else:
                p._pmv_lp = p.detach().to(dtype)
            p._pmv_lp_version = p._version
            n += 1
    return n


def set_compute_dtype(module: nn.Module, dtype: torch.dtype):
    """Switch every pmv_b200 module under ``module`` between the bf16 mode and the fp32 mode of the path."""
    assert dtype in (torch.bfloat16, torch.float32)
    for m in module.modules():
        if hasattr(type(m), "compute_dtype"):
            m.compute_dtype = dtype
    return module
