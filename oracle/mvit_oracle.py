"""CPU oracle for the MViTv2 pooling-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker (or as
the timed CPU baseline), never as the thing shipped.

This is an independent, functional (no ``nn.Module``) restatement of the
reference algorithm in torch-on-CPU arithmetic.  Parameters are passed as a flat
``dict`` keyed by the reference ``state_dict`` names, so the same dict drives the
reference module, this oracle and the CUDA path.  Each function cites the
reference lines it restates (paths relative to ``/root/reference/MViT``).

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself,
generated in the build container by ``oracle/make_golden.py`` (which imports the
reference's own ``slowfast/models/attention.py``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle against
those fixtures on every run.

All maths is done in the dtype of the inputs (fp32 for parity with the
reference, fp64 for a noise-floor estimate).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]
LN_EPS = 1e-6  # video_model_builder.py:1802  partial(nn.LayerNorm, eps=1e-6)


# ---------------------------------------------------------------------------
# small pieces
# ---------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = LN_EPS) -> torch.Tensor:
    """LayerNorm over the last axis (biased variance), attention.py:498,529 / :254-282."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU — ``nn.GELU()`` default, common.py:13,21."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def pooled_size(n: int, k: int, s: int) -> int:
    """Conv/MaxPool output length with padding k//2 (attention.py:199-200,500-502)."""
    p = k // 2
    return (n + 2 * p - k) // s + 1


def conv_pool_tokens(
    x: torch.Tensor,  # [B, nh, 1+T*H*W, C] (or no cls)
    thw: Sequence[int],
    weight: torch.Tensor,  # [C, 1, kt, kh, kw] depthwise, shared over heads
    stride: Sequence[int],
    has_cls: bool,
    ln_w: Optional[torch.Tensor],
    ln_b: Optional[torch.Tensor],
) -> Tuple[torch.Tensor, List[int]]:
    """attention_pool with a depthwise Conv3d + LayerNorm, attention.py:14-48.

    cls token bypasses the convolution (:25-26) and is re-attached before the
    LayerNorm (:39-42), so it is normalised but never convolved.
    """
    B, nh, N, C = x.shape
    T, H, W = thw
    if has_cls:
        cls, tok = x[:, :, :1], x[:, :, 1:]
    else:
        cls, tok = None, x
    kt, kh, kw = weight.shape[2:]
    vol = tok.reshape(B * nh, T, H, W, C).permute(0, 4, 1, 2, 3)
    out = F.conv3d(vol, weight, None, stride=tuple(stride), padding=(kt // 2, kh // 2, kw // 2), groups=C)
    To, Ho, Wo = out.shape[2:]
    out = out.reshape(B, nh, C, To * Ho * Wo).transpose(2, 3)
    if has_cls:
        out = torch.cat([cls, out], dim=2)
    if ln_w is not None:
        out = layer_norm(out, ln_w, ln_b)
    return out, [To, Ho, Wo]


def conv_pool_tokens_taps(x, thw, weight, stride, has_cls, ln_w, ln_b):
    """Same as :func:`conv_pool_tokens` but as an explicit channels-last sum over the
    kernel taps (the formulation the CUDA stencil uses).  Used by the tests to
    cross-check the ATen path on small shapes."""
    B, nh, N, C = x.shape
    T, H, W = thw
    cls, tok = (x[:, :, :1], x[:, :, 1:]) if has_cls else (None, x)
    kt, kh, kw = weight.shape[2:]
    st, sh, sw = stride
    To, Ho, Wo = pooled_size(T, kt, st), pooled_size(H, kh, sh), pooled_size(W, kw, sw)
    vol = tok.reshape(B, nh, T, H, W, C)
    pad = F.pad(vol, (0, 0, kw // 2, kw // 2, kh // 2, kh // 2, kt // 2, kt // 2))
    out = torch.zeros(B, nh, To, Ho, Wo, C, dtype=x.dtype, device=x.device)
    for a in range(kt):
        for b in range(kh):
            for c in range(kw):
                sl = pad[:, :, a : a + (To - 1) * st + 1 : st, b : b + (Ho - 1) * sh + 1 : sh, c : c + (Wo - 1) * sw + 1 : sw]
                out = out + sl * weight[:, 0, a, b, c]
    out = out.reshape(B, nh, To * Ho * Wo, C)
    if has_cls:
        out = torch.cat([cls, out], dim=2)
    if ln_w is not None:
        out = layer_norm(out, ln_w, ln_b)
    return out, [To, Ho, Wo]


def max_pool_tokens(x: torch.Tensor, thw: Sequence[int], kernel, stride, has_cls: bool):
    """Skip-path attention_pool with MaxPool3d and no norm, attention.py:558-564,571-573.
    ``x`` is [B, N, C] (one pseudo-head, :17-23,44-47)."""
    B, N, C = x.shape
    T, H, W = thw
    cls, tok = (x[:, :1], x[:, 1:]) if has_cls else (None, x)
    vol = tok.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
    out = F.max_pool3d(vol, tuple(kernel), tuple(stride), tuple(k // 2 for k in kernel))
    To, Ho, Wo = out.shape[2:]
    out = out.reshape(B, C, To * Ho * Wo).transpose(1, 2)
    if has_cls:
        out = torch.cat([cls, out], dim=1)
    return out, [To, Ho, Wo]


def interp_rel_table(table: torch.Tensor, d: int) -> torch.Tensor:
    """get_rel_pos, attention.py:51-64: identity when the length matches, else 1-D
    linear interpolation (align_corners False) of the [L, C] table to [d, C]."""
    L = table.shape[0]
    if L == d:
        return table
    t = F.interpolate(table.reshape(1, L, -1).permute(0, 2, 1), size=d, mode="linear")
    return t.reshape(-1, d).permute(1, 0)


def rel_index(q_n: int, k_n: int) -> torch.Tensor:
    """Integer table index for every (query, key) coordinate pair on one axis,
    attention.py:80-86,98 (float ratios, ``.long()`` truncation)."""
    q_ratio = max(k_n / q_n, 1.0)
    k_ratio = max(q_n / k_n, 1.0)
    dist = torch.arange(q_n)[:, None] * q_ratio - torch.arange(k_n)[None, :] * k_ratio
    dist = dist + (k_n - 1) * k_ratio
    return dist.long()


def rel_pos_bias_terms(
    q: torch.Tensor,  # [B, nh, Nq, C] un-scaled, post-LN
    has_cls: bool,
    q_shape: Sequence[int],
    k_shape: Sequence[int],
    rel_h: Optional[torch.Tensor],
    rel_w: Optional[torch.Tensor],
    rel_t: Optional[torch.Tensor],
):
    """The three decomposed bias factors of cal_rel_pos_spatial / _temporal
    (attention.py:67-159): returns (bh [B,nh,Lq,k_h], bw [B,nh,Lq,k_w], bt [B,nh,Lq,k_t])
    so that bias[q,(kt,kh,kw)] = bh[q,kh] + bw[q,kw] + bt[q,kt]."""
    s = 1 if has_cls else 0
    qt, qh, qw = q_shape
    kt, kh, kw = k_shape
    B, nh, _, C = q.shape
    r_q = q[:, :, s:].reshape(B, nh, qt, qh, qw, C)
    bh = bw = bt = None
    if rel_h is not None:
        Rh = interp_rel_table(rel_h, 2 * max(qh, kh) - 1)[rel_index(qh, kh).to(q.device)]  # [qh, kh, C]
        Rw = interp_rel_table(rel_w, 2 * max(qw, kw) - 1)[rel_index(qw, kw).to(q.device)]  # [qw, kw, C]
        bh = torch.einsum("bnthwc,hkc->bnthwk", r_q, Rh).reshape(B, nh, -1, kh)
        bw = torch.einsum("bnthwc,wkc->bnthwk", r_q, Rw).reshape(B, nh, -1, kw)
    if rel_t is not None:
        Rt = interp_rel_table(rel_t, 2 * max(qt, kt) - 1)[rel_index(qt, kt).to(q.device)]  # [qt, kt, C]
        bt = torch.einsum("bnthwc,tkc->bnthwk", r_q, Rt).reshape(B, nh, -1, kt)
    return bh, bw, bt


def add_rel_pos_bias(attn, q, has_cls, q_shape, k_shape, rel_h, rel_w, rel_t):
    """attn[:, :, s:, s:] += bias, attention.py:111-115,154-157 (cls row/col get none)."""
    s = 1 if has_cls else 0
    kt, kh, kw = k_shape
    B, nh, Nq, Nk = attn.shape
    bh, bw, bt = rel_pos_bias_terms(q, has_cls, q_shape, k_shape, rel_h, rel_w, rel_t)
    Lq = Nq - s
    bias = torch.zeros(B, nh, Lq, kt, kh, kw, dtype=attn.dtype, device=attn.device)
    if bh is not None:
        bias = bias + bh[:, :, :, None, :, None] + bw[:, :, :, None, None, :]
    if bt is not None:
        bias = bias + bt[:, :, :, :, None, None]
    out = attn.clone()
    out[:, :, s:, s:] = out[:, :, s:, s:] + bias.reshape(B, nh, Lq, kt * kh * kw)
    return out


# ---------------------------------------------------------------------------
# MultiScaleAttention / MultiScaleBlock
# ---------------------------------------------------------------------------
def _maybe(p: Params, key: str):
    return p[key] if key in p else None


def multiscale_attention(
    x: torch.Tensor,  # [B, N, dim]
    thw: Sequence[int],
    p: Params,
    prefix: str,
    num_heads: int,
    stride_q: Sequence[int],
    stride_kv: Sequence[int],
    has_cls: bool = True,
    residual_pooling: bool = True,
    hw_switch_auto: bool = False,
    return_intermediates: bool = False,
):
    """MultiScaleAttention.forward for the v2 settings (pool_first False, mode conv,
    separate_qkv False), attention.py:314-461."""
    B, N, _ = x.shape
    Wqkv, bqkv = p[prefix + "qkv.weight"], _maybe(p, prefix + "qkv.bias")
    C = Wqkv.shape[0] // 3
    hd = C // num_heads
    qkv = F.linear(x, Wqkv, bqkv).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)  # :327-332
    q, k, v = qkv[0], qkv[1], qkv[2]
    q_shape = k_shape = list(thw)
    if prefix + "pool_q.weight" in p:
        q, q_shape = conv_pool_tokens(q, thw, p[prefix + "pool_q.weight"], stride_q, has_cls,
                                      _maybe(p, prefix + "norm_q.weight"), _maybe(p, prefix + "norm_q.bias"))
    if prefix + "pool_k.weight" in p:
        k, k_shape = conv_pool_tokens(k, thw, p[prefix + "pool_k.weight"], stride_kv, has_cls,
                                      _maybe(p, prefix + "norm_k.weight"), _maybe(p, prefix + "norm_k.bias"))
        v, _ = conv_pool_tokens(v, thw, p[prefix + "pool_v.weight"], stride_kv, has_cls,
                                _maybe(p, prefix + "norm_v.weight"), _maybe(p, prefix + "norm_v.bias"))
    scale = hd ** -0.5  # :195
    attn = (q * scale) @ k.transpose(-2, -1)  # :412
    rel_h, rel_w, rel_t = _maybe(p, prefix + "rel_pos_h"), _maybe(p, prefix + "rel_pos_w"), _maybe(p, prefix + "rel_pos_t")
    if rel_h is not None and hw_switch_auto and thw[1] > thw[2]:  # :414-424
        rel_h, rel_w = rel_w, rel_h
    attn = add_rel_pos_bias(attn, q, has_cls, q_shape, k_shape, rel_h, rel_w, rel_t)  # :413-445
    attn = attn.softmax(dim=-1)  # :446
    o = attn @ v  # :448
    if residual_pooling:  # :450-454
        if has_cls:
            o = torch.cat([o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]], dim=2)
        else:
            o = o + q
    o = o.transpose(1, 2).reshape(B, -1, C)  # :456
    y = F.linear(o, p[prefix + "proj.weight"], p[prefix + "proj.bias"])  # :457
    if return_intermediates:
        return y, q_shape, dict(q=q, k=k, v=v, attn=attn, o=o, k_shape=k_shape)
    return y, q_shape


def mlp(x: torch.Tensor, p: Params, prefix: str) -> torch.Tensor:
    """Mlp.forward, common.py:26-34 (drop_rate 0)."""
    h = gelu_erf(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def multiscale_block(
    x: torch.Tensor,
    thw: Sequence[int],
    p: Params,
    prefix: str,
    num_heads: int,
    stride_q: Sequence[int],
    stride_kv: Sequence[int],
    has_cls: bool = True,
    residual_pooling: bool = True,
    dim_mul_in_att: bool = True,
    hw_switch_auto: bool = False,
    drop_scale: Optional[torch.Tensor] = None,  # [2, B]: per-sample DropPath factor mask/keep for the 2 branches
):
    """MultiScaleBlock.forward, attention.py:566-589 (gamma_1/2 None, drop_rate 0)."""
    dim = x.shape[-1]
    xn = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])  # :567
    xb, thw_new = multiscale_attention(xn, thw, p, prefix + "attn.", num_heads, stride_q, stride_kv,
                                       has_cls, residual_pooling, hw_switch_auto)  # :568
    has_proj = prefix + "proj.weight" in p
    if dim_mul_in_att and has_proj:  # :569-570 — residual is proj(LN(x)), not x
        x = F.linear(xn, p[prefix + "proj.weight"], p[prefix + "proj.bias"])
    if len(stride_q) > 0 and math.prod(stride_q) > 1:  # :558-564
        k_skip = [s + 1 if s > 1 else s for s in stride_q]  # :500
        x_res, _ = max_pool_tokens(x, thw, k_skip, stride_q, has_cls)  # :571-573
    else:
        x_res = x
    if drop_scale is not None:
        xb = xb * drop_scale[0][:, None, None]
    x = x_res + xb  # :577
    xn2 = layer_norm(x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])  # :578
    xm = mlp(xn2, p, prefix + "mlp.")  # :579
    if (not dim_mul_in_att) and has_proj:  # :580-581
        x = F.linear(xn2, p[prefix + "proj.weight"], p[prefix + "proj.bias"])
    if drop_scale is not None:
        xm = xm * drop_scale[1][:, None, None]
    return x + xm, thw_new  # :585


# ---------------------------------------------------------------------------
# MViT backbone (harness-level restatement of video_model_builder.py:1726-2171)
# ---------------------------------------------------------------------------
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


MVITV2_S = dict(  # configs/Kinetics/MVITv2_S_16x4.yaml:8-44
    num_frames=16, crop=(224, 224), depth=16, embed_dim=96, num_heads=1, mlp_ratio=4.0,
    patch_kernel=(3, 7, 7), patch_stride=(2, 4, 4), patch_padding=(1, 3, 3),
    dim_mul={1: 2.0, 3: 2.0, 14: 2.0}, head_mul={1: 2.0, 3: 2.0, 14: 2.0},
    pool_q_stride={1: (1, 2, 2), 3: (1, 2, 2), 14: (1, 2, 2)}, kv_stride_adaptive=(1, 8, 8),
    num_classes=400, drop_path_rate=0.2,
)
MVITV2_B = dict(  # configs/Kinetics/MVITv2_B_32x3.yaml:8-44
    num_frames=32, crop=(224, 224), depth=24, embed_dim=96, num_heads=1, mlp_ratio=4.0,
    patch_kernel=(3, 7, 7), patch_stride=(2, 4, 4), patch_padding=(1, 3, 3),
    dim_mul={2: 2.0, 5: 2.0, 21: 2.0}, head_mul={2: 2.0, 5: 2.0, 21: 2.0},
    pool_q_stride={2: (1, 2, 2), 5: (1, 2, 2), 21: (1, 2, 2)}, kv_stride_adaptive=(1, 8, 8),
    num_classes=400, drop_path_rate=0.3,
)


def block_schedule(cfg: dict) -> List[dict]:
    """Per-block (dim, dim_out, heads, input thw, stride_q, stride_kv) schedule,
    video_model_builder.py:1862-1967 with DIM_MUL_IN_ATT and POOL_KV_STRIDE_ADAPTIVE."""
    depth = cfg["depth"]
    dim_mul = [cfg["dim_mul"].get(i, 1.0) for i in range(depth + 1)]
    head_mul = [cfg["head_mul"].get(i, 1.0) for i in range(depth + 1)]
    thw = [cfg["num_frames"] // cfg["patch_stride"][0], cfg["crop"][0] // cfg["patch_stride"][1],
           cfg["crop"][1] // cfg["patch_stride"][2]]
    skv = list(cfg["kv_stride_adaptive"])
    embed, heads = cfg["embed_dim"], cfg["num_heads"]
    out = []
    for i in range(depth):
        sq = list(cfg["pool_q_stride"].get(i, (1, 1, 1)))  # every block is listed in POOL_Q_STRIDE
        skv = [max(skv[d] // sq[d], 1) for d in range(3)]  # :1885-1894
        heads = round_width(heads, head_mul[i])  # :1919
        dim_out = round_width(embed, dim_mul[i], divisor=round_width(heads, head_mul[i]))  # :1920-1925
        out.append(dict(dim=embed, dim_out=dim_out, num_heads=heads, thw=list(thw), stride_q=sq, stride_kv=list(skv)))
        thw = [n // s for n, s in zip(thw, sq)]  # :1961-1965
        embed = dim_out
    return out


def patch_embed(x: torch.Tensor, p: Params, cfg: dict):
    """PatchEmbed.forward, stem_helper.py:320-325."""
    y = F.conv3d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"],
                 stride=tuple(cfg["patch_stride"]), padding=tuple(cfg["patch_padding"]))
    thw = list(y.shape[2:])
    return y.flatten(2).transpose(1, 2), thw


def mvit_forward(clip: torch.Tensor, p: Params, cfg: dict, softmax_head: bool = False,
                 drop_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """MViT.forward_ for the v2 configs (cls token on, no abs-pos, no norm_stem),
    video_model_builder.py:2098-2171; head = TransformerBasicHead (head_helper.py:561-577)
    without dropout (eval) and without the eval softmax unless ``softmax_head``."""
    x, thw = patch_embed(clip, p, cfg)
    B = x.shape[0]
    x = torch.cat([p["cls_token"].expand(B, -1, -1), x], dim=1)  # :2115-2121
    for i, blk in enumerate(block_schedule(cfg)):
        assert thw == blk["thw"], (thw, blk["thw"])
        ds = None if drop_scale is None else drop_scale[i]
        x, thw = multiscale_block(x, thw, p, f"blocks.{i}.", blk["num_heads"], blk["stride_q"], blk["stride_kv"],
                                  drop_scale=ds, hw_switch_auto=cfg.get("hw_switch_auto", False))
    x = layer_norm(x, p["norm.weight"], p["norm.bias"])[:, 0]  # :2163-2165
    y = F.linear(x, p["head.projection.weight"], p["head.projection.bias"])
    return y.softmax(dim=1) if softmax_head else y


def mvit_forward_pm(clip: torch.Tensor, pm: torch.Tensor, p: Params, cfg: dict, softmax_head: bool = False) -> torch.Tensor:
    """MViT.forward with the portrait-mode mask (video_model_builder.py:2075-2096): portrait samples (pm == True)
    arrive transposed inside a landscape-shaped batch; they are transposed back and run with H and W swapped
    (the rel-pos tables swap inside the blocks when hw_switch_auto is set, attention.py:414-435), the landscape
    samples run as they are, and the outputs are scattered back into batch order."""
    if pm is None or int(pm.sum()) == 0:
        return mvit_forward(clip, p, cfg, softmax_head)
    pm_index = torch.where(pm)[0]
    lm_index = torch.where(~pm)[0]
    cfg_p = dict(cfg, crop=(cfg["crop"][1], cfg["crop"][0]))
    pm_x = mvit_forward(clip[pm_index].transpose(-2, -1), p, cfg_p, softmax_head)
    out = torch.empty((len(pm),) + tuple(pm_x.shape[1:]), dtype=pm_x.dtype, device=pm_x.device)
    out[pm_index] = pm_x
    if len(lm_index) != 0:
        out[lm_index] = mvit_forward(clip[lm_index], p, cfg, softmax_head)
    return out


def soft_target_cross_entropy(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """losses.py:69-71 ``soft_cross_entropy`` = pytorchvideo.losses.SoftTargetCrossEntropyLoss(normalize_targets=False).
    pytorchvideo is a third-party dependency that is not vendored under /root/reference (setup.py: pytorchvideo, unpinned;
    the class has been stable since 0.1.3); its published algorithm: integer targets are converted to one-hot rows, then
    ``loss = sum(-y * log_softmax(x, dim=-1), dim=-1)`` with reduction "mean" over the batch.  For one-hot rows this is
    nn.CrossEntropyLoss (losses.py:66), which the test uses as the second anchor."""
    if target.dtype in (torch.int64, torch.int32):
        target = F.one_hot(target.long(), logits.shape[-1]).to(logits.dtype)
    return torch.sum(-target * F.log_softmax(logits, dim=-1), dim=-1).mean()


def head_loss(tokens: torch.Tensor, p: Params, target: torch.Tensor, keep_mask: torch.Tensor = None, dropout_p: float = 0.0):
    """Final norm -> cls row (video_model_builder.py:2163-2165) -> TransformerBasicHead in training mode
    (head_helper.py:561-566: dropout, projection; the dropout keep mask is an input so that the statement is
    deterministic) -> loss.  Returns (loss, logits)."""
    x = layer_norm(tokens, p["norm.weight"], p["norm.bias"])[:, 0]
    if keep_mask is not None:
        x = x * keep_mask.to(x.dtype) / (1.0 - dropout_p)
    logits = F.linear(x, p["head.projection.weight"], p["head.projection.bias"])
    return soft_target_cross_entropy(logits, target), logits


def param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every reference ``state_dict`` entry of ``MViT(cfg)``."""
    sh: Dict[str, Tuple[int, ...]] = {}
    e = cfg["embed_dim"]
    sh["cls_token"] = (1, 1, e)
    sh["patch_embed.proj.weight"] = (e, 3) + tuple(cfg["patch_kernel"])
    sh["patch_embed.proj.bias"] = (e,)
    T0 = cfg["num_frames"] // cfg["patch_stride"][0]
    for i, b in enumerate(block_schedule(cfg)):
        sh.update(block_param_shapes(f"blocks.{i}.", b["dim"], b["dim_out"], b["num_heads"], b["thw"],
                                     b["stride_q"], b["stride_kv"], cfg["mlp_ratio"]))
        last = b["dim_out"]
    sh["norm.weight"] = (last,)
    sh["norm.bias"] = (last,)
    sh["head.projection.weight"] = (cfg["num_classes"], last)
    sh["head.projection.bias"] = (cfg["num_classes"],)
    return sh


def block_param_shapes(prefix, dim, dim_out, nh, thw, stride_q, stride_kv, mlp_ratio=4.0, kernel=(3, 3, 3)):
    """Parameter shapes of one MultiScaleBlock with dim_mul_in_att, attention.py:188-312,495-564."""
    sh = {}
    C = dim_out
    hd = C // nh
    sh[prefix + "norm1.weight"] = (dim,)
    sh[prefix + "norm1.bias"] = (dim,)
    qh, kh = thw[1] // stride_q[1], thw[1] // stride_kv[1]
    qw, kw = thw[2] // stride_q[2], thw[2] // stride_kv[2]
    sh[prefix + "attn.rel_pos_h"] = (2 * max(qh, kh) - 1, hd)
    sh[prefix + "attn.rel_pos_w"] = (2 * max(qw, kw) - 1, hd)
    sh[prefix + "attn.rel_pos_t"] = (2 * thw[0] - 1, hd)
    sh[prefix + "attn.qkv.weight"] = (3 * C, dim)
    sh[prefix + "attn.qkv.bias"] = (3 * C,)
    sh[prefix + "attn.proj.weight"] = (C, C)
    sh[prefix + "attn.proj.bias"] = (C,)
    for n in "qkv":
        sh[prefix + f"attn.pool_{n}.weight"] = (hd, 1) + tuple(kernel)
        sh[prefix + f"attn.norm_{n}.weight"] = (hd,)
        sh[prefix + f"attn.norm_{n}.bias"] = (hd,)
    sh[prefix + "norm2.weight"] = (C,)
    sh[prefix + "norm2.bias"] = (C,)
    hid = int(C * mlp_ratio)
    sh[prefix + "mlp.fc1.weight"] = (hid, C)
    sh[prefix + "mlp.fc1.bias"] = (hid,)
    sh[prefix + "mlp.fc2.weight"] = (dim_out, hid)
    sh[prefix + "mlp.fc2.bias"] = (dim_out,)
    if dim != dim_out:
        sh[prefix + "proj.weight"] = (dim_out, dim)
        sh[prefix + "proj.bias"] = (dim_out,)
    return sh
