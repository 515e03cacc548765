"""MViTv2 backbone harness: re-states video_model_builder.MViT (video_model_builder.py:1726-2171) for the v2
configurations (cls token, no abs-pos, rel-pos spatial+temporal, residual pooling, dim_mul_in_att) on top of
the pmv_b200 MultiScaleBlock.  The reference MViT itself is kept unchanged by the drop-in;
this class exists so the repository can run and benchmark the full model without the reference's fvcore /
detectron2 / pytorchvideo dependencies.  Parameter names match the reference state_dict one to one."""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from . import functional as Fn
from .attention import MultiScaleBlock, set_compute_dtype
from .common import compute_dtype_of
from .stem import PatchEmbed

MVITV2_S = dict(  # MViT/configs/Kinetics/MVITv2_S_16x4.yaml:8-44
    name="MViTv2-S 16x4", num_frames=16, crop=(224, 224), depth=16, embed_dim=96, num_heads=1, mlp_ratio=4.0,
    patch_kernel=(3, 7, 7), patch_stride=(2, 4, 4), patch_padding=(1, 3, 3),
    dim_mul={1: 2.0, 3: 2.0, 14: 2.0}, head_mul={1: 2.0, 3: 2.0, 14: 2.0},
    pool_q_stride={1: (1, 2, 2), 3: (1, 2, 2), 14: (1, 2, 2)}, kv_stride_adaptive=(1, 8, 8),
    num_classes=400, drop_path_rate=0.2, head_dropout=0.5,
)
MVITV2_B = dict(  # MViT/configs/Kinetics/MVITv2_B_32x3.yaml:8-44
    name="MViTv2-B 32x3", num_frames=32, crop=(224, 224), depth=24, embed_dim=96, num_heads=1, mlp_ratio=4.0,
    patch_kernel=(3, 7, 7), patch_stride=(2, 4, 4), patch_padding=(1, 3, 3),
    dim_mul={2: 2.0, 5: 2.0, 21: 2.0}, head_mul={2: 2.0, 5: 2.0, 21: 2.0},
    pool_q_stride={2: (1, 2, 2), 5: (1, 2, 2), 21: (1, 2, 2)}, kv_stride_adaptive=(1, 8, 8),
    num_classes=400, drop_path_rate=0.3, head_dropout=0.5,
)


def round_width(width, multiplier, min_width=1, divisor=1):
    """models/utils.py:15-31."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def block_schedule(cfg):
    """(dim, dim_out, heads, thw, stride_q, stride_kv) per block — video_model_builder.py:1862-1967."""
    depth = cfg["depth"]
    dim_mul = [cfg["dim_mul"].get(i, 1.0) for i in range(depth + 1)]
    head_mul = [cfg["head_mul"].get(i, 1.0) for i in range(depth + 1)]
    thw = [cfg["num_frames"] // cfg["patch_stride"][0], cfg["crop"][0] // cfg["patch_stride"][1],
           cfg["crop"][1] // cfg["patch_stride"][2]]
    skv = list(cfg["kv_stride_adaptive"])
    embed, heads = cfg["embed_dim"], cfg["num_heads"]
    out = []
    for i in range(depth):
        sq = list(cfg["pool_q_stride"].get(i, (1, 1, 1)))
        skv = [max(skv[d] // sq[d], 1) for d in range(3)]
        heads = round_width(heads, head_mul[i])
        dim_out = round_width(embed, dim_mul[i], divisor=round_width(heads, head_mul[i]))
        out.append(dict(dim=embed, dim_out=dim_out, num_heads=heads, thw=list(thw), stride_q=sq, stride_kv=list(skv)))
        thw = [n // s for n, s in zip(thw, sq)]
        embed = dim_out
    return out


class _Head(nn.Module):
    """TransformerBasicHead (head_helper.py:502-577): dropout (train) -> Linear -> softmax (eval).  Left in
    PyTorch: a [B,768] x [768,400] product is not on the hot path."""

    def __init__(self, dim_in, num_classes, dropout_rate=0.0, act="softmax"):
        super().__init__()
        if dropout_rate > 0.0:
            self.dropout = nn.Dropout(dropout_rate)
        self.projection = nn.Linear(dim_in, num_classes, bias=True)
        self.act = nn.Softmax(dim=1) if act == "softmax" else None

    def forward(self, x):
        if hasattr(self, "dropout"):
            x = self.dropout(x)
        x = self.projection(x)
        if not self.training and self.act is not None:
            x = self.act(x)
        return x.view(x.shape[0], -1)


class MViT(nn.Module):
    compute_dtype = torch.bfloat16

    def __init__(self, cfg=MVITV2_S, compute_dtype=torch.bfloat16):
        super().__init__()
        self.cfg = cfg
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        embed = cfg["embed_dim"]
        self.patch_embed = PatchEmbed(3, embed, cfg["patch_kernel"], cfg["patch_stride"], cfg["patch_padding"])
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed))
        dpr = [x.item() for x in torch.linspace(0, cfg["drop_path_rate"], cfg["depth"])]
        self.blocks = nn.ModuleList()
        sched = block_schedule(cfg)
        self.T, self.H, self.W = sched[0]["thw"]
        for i, b in enumerate(sched):
            self.blocks.append(MultiScaleBlock(
                dim=b["dim"], dim_out=b["dim_out"], num_heads=b["num_heads"], input_size=b["thw"],
                mlp_ratio=cfg["mlp_ratio"], qkv_bias=True, drop_rate=0.0, drop_path=dpr[i], norm_layer=norm_layer,
                kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=b["stride_q"], stride_kv=b["stride_kv"], mode="conv",
                has_cls_embed=True, pool_first=False, rel_pos_spatial=True, rel_pos_temporal=True,
                rel_pos_zero_init=False, residual_pooling=True, dim_mul_in_att=True, separate_qkv=False,
                hw_switch_auto=cfg.get("hw_switch_auto", False)))
        last = sched[-1]["dim_out"]
        self.norm = norm_layer(last)
        self.head = _Head(last, cfg["num_classes"], cfg.get("head_dropout", 0.0))
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)
        set_compute_dtype(self, compute_dtype)

    def no_weight_decay(self):
        """video_model_builder.py:2027-2049 (MVIT.ZERO_DECAY_POS_CLS; False in MVITv2_S_16x4.yaml:21): parameter-name
        fragments the optimizer keeps out of weight decay.  This family has relative positions and a cls token only."""
        names = []
        if self.cfg.get("zero_decay_pos_cls", False):
            names.extend(["rel_pos_h", "rel_pos_w", "rel_pos_hw", "rel_pos_t", "cls_token"])
        return names

    @staticmethod
    def _init_weights(m):  # video_model_builder.py:2018-2025
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv3d)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _drop_path_scales(self, batch, device):
        """DropPath factors mask / keep_prob (common.py:46-59) of every block and both residual branches, drawn with
        one rand / add / floor / div instead of four small launches per branch (120 launches per MViTv2-S step)."""
        if not self.training or all(blk.drop_path_prob == 0.0 for blk in self.blocks):
            return [None] * len(self.blocks)
        if not self.cfg.get("droppath_batched", False):
            # default: every block draws its own factors, branch by branch, in the reference's order and with the
            # reference's call pattern (one torch.rand of B values per DropPath call, common.py:46-59), so a seeded run
            # consumes the generator exactly like the reference model does
            return [None] * len(self.blocks)
        keep = getattr(self, "_dp_keep", None)
        if keep is None or keep.device != device:
            keep = torch.tensor([1.0 - blk.drop_path_prob for blk in self.blocks for _ in range(2)], dtype=torch.float32,
                                device=device).unsqueeze(1)
            self._dp_keep = keep
        scales = (keep + torch.rand(keep.shape[0], batch, device=device)).floor_() / keep
        return [None if blk.drop_path_prob == 0.0 else (scales[2 * i], scales[2 * i + 1]) for i, blk in enumerate(self.blocks)]

    def forward_features(self, clip, thw_expected=None):
        x = self.forward_tokens(clip, thw_expected)
        x = Fn.layer_norm(x, self.norm.weight, self.norm.bias, torch.float32, self.norm.eps)  # :2163
        return x[:, 0]                                                        # :2165

    def forward_tokens(self, clip, thw_expected=None):
        """The token stream before the final norm: [B, 1 + T*H*W / 64, 768] fp32."""
        x, thw = self.patch_embed.forward_tokens(clip, self.cls_token)        # :2100-2121
        assert tuple(thw) == tuple(thw_expected or (self.T, self.H, self.W)), thw  # :2106
        scales = self._drop_path_scales(x.shape[0], x.device)
        ckpt = self.cfg.get("act_checkpoint", False) and torch.is_grad_enabled()
        for blk, ds in zip(self.blocks, scales):                              # :2144-2146
            if ckpt:
                # MODEL.ACT_CHECKPOINT (video_model_builder.py:1958-1959 wraps every block in fairscale's
                # checkpoint_wrapper): the block's activations are dropped after the forward and recomputed in backward
                from torch.utils.checkpoint import checkpoint
                x, thw = checkpoint(blk, x, thw, ds, use_reentrant=False, preserve_rng_state=True)
            else:
                x, thw = blk(x, thw, drop_scales=ds)
        return x

    def _head_no_grad(self, tok):
        """norm -> cls -> head in one launch (row f2 forward without a loss): logits, or softmax probabilities in eval
        mode when the head has its activation (head_helper.py:568-570)."""
        from . import ops
        drop = self.training and hasattr(self.head, "dropout") and self.head.dropout.p > 0.0
        if drop:  # training-mode forward without autograd: keep the module semantics, unfused
            x = Fn.layer_norm(tok, self.norm.weight, self.norm.bias, torch.float32, self.norm.eps)
            return self.head(x[:, 0])
        probs = (not self.training) and self.head.act is not None
        r = ops.head_loss_fwd(tok.contiguous(), self.norm.weight, self.norm.bias, self.head.projection.weight,
                              self.head.projection.bias, want_probs=probs, eps=self.norm.eps)
        return r["probs"] if probs else r["logits"]

    def forward_loss(self, x, target):
        """Training tail fused (row f2): returns (loss, logits) = what ``loss_fun(model(x), target)`` computes in
        tools/train_net.py:172-186 with ``cross_entropy`` (int64 labels) or ``soft_cross_entropy`` (float [B, classes]
        mixup targets, losses.py:69-71).  The head's dropout draws its keep mask from torch's generator."""
        clip = x[0] if isinstance(x, (list, tuple)) else x
        tok = self.forward_tokens(clip)
        p = self.head.dropout.p if (self.training and hasattr(self.head, "dropout")) else 0.0
        keep = None
        if p > 0.0:
            keep = (torch.rand(tok.shape[0], tok.shape[2], device=tok.device) >= p).to(torch.uint8)
        return Fn.head_loss(tok, self.norm.weight, self.norm.bias, self.head.projection.weight, self.head.projection.bias,
                            target, keep_mask=keep, dropout_p=p, eps=self.norm.eps)

    def forward(self, x, pm=None):
        """``x``: clip tensor or the reference's one-element list (:2099).  ``pm``: portrait-mode mask, a bool tensor
        [B] or the reference's list of per-loader-batch tensors (:2076-2077).  Portrait samples arrive transposed
        inside the landscape-shaped batch; they are transposed back and run with H and W swapped (the blocks swap
        their rel-pos tables when built with hw_switch_auto), the landscape samples run as they are, and the rows
        are scattered back into batch order (video_model_builder.py:2075-2096)."""
        clip = x[0] if isinstance(x, (list, tuple)) else x
        if pm is not None and isinstance(pm, (list, tuple)):
            pm = torch.cat(list(pm))
        if pm is None or int(pm.sum()) == 0:
            if not torch.is_grad_enabled():
                return self._head_no_grad(self.forward_tokens(clip))
            return self.head(self.forward_features(clip))
        assert len(pm) == clip.shape[0]
        pm = pm.to(device=clip.device, dtype=torch.bool)
        pm_index = torch.where(pm)[0]
        lm_index = torch.where(~pm)[0]
        pm_x = self.head(self.forward_features(clip[pm_index].transpose(-2, -1), (self.T, self.W, self.H)))
        out = torch.empty((len(pm),) + tuple(pm_x.shape[1:]), device=pm_x.device, dtype=pm_x.dtype)
        out[pm_index] = pm_x
        if len(lm_index) != 0:
            out[lm_index] = self.head(self.forward_features(clip[lm_index]))
        return out
