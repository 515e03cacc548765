"""Block- and model-level parity of the CUDA path, called through the reference-shaped modules
(pmv_b200.attention.MultiScaleBlock / MultiScaleAttention, pmv_b200.mvit.MViT):
  * against the golden fixtures the UNMODIFIED reference produced (tests/golden/),
  * against the oracle on the real MViTv2-S stage shapes, forward + backward.
Tolerances (normalised max error): fp32 mode 1e-4, bf16 mode 1e-2 on outputs (north star); bf16 gradients
3e-2 (one extra bf16 rounding per backward operand) for every block, see grad_ok."""
import glob
import json
import os
from functools import partial

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BLOCK_FILES = sorted(glob.glob(os.path.join(GOLDEN, "blk_*.npz")))
OUT_TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}
GRAD_TOL = {torch.float32: 1e-4, torch.bfloat16: 3e-2}
DTYPES = [torch.float32, torch.bfloat16]


def nerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_ok(a, b, dtype):
    """Gradient parity, normalised max error: fp32 mode <= 1e-4, bf16 mode <= 3e-2 — for every block.  Downstream of the
    skip-path MaxPool3d the bf16 comparison side is the oracle with its max-pool forced to the winners the kernel chose
    (tests/test_parity2_gpu.py::forced_max_pool): a bf16-rounded activation can move an arg-max to a neighbouring,
    nearly equal window element, which moves single gradient entries by O(1) without being an error of the backward."""
    return nerr(a, b) < GRAD_TOL[dtype]


def make_block(cfg, dtype):
    from pmv_b200.attention import MultiScaleBlock, set_compute_dtype
    blk = MultiScaleBlock(
        dim=cfg["dim"], dim_out=cfg["dim_out"], num_heads=cfg["num_heads"], input_size=cfg["thw"], mlp_ratio=4.0,
        qkv_bias=True, drop_rate=0.0, drop_path=0.0, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
        kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=cfg["stride_q"], stride_kv=cfg["stride_kv"], mode="conv",
        has_cls_embed=True, pool_first=False, rel_pos_spatial=True, rel_pos_temporal=True, rel_pos_zero_init=False,
        residual_pooling=True, dim_mul_in_att=True, separate_qkv=False, hw_switch_auto=cfg.get("hw_switch_auto", False))
    return set_compute_dtype(blk, dtype).cuda()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("path", BLOCK_FILES, ids=[os.path.basename(p)[:-4] for p in BLOCK_FILES])
def test_block_matches_reference_fixture(path, dtype):
    from oracle import detgen, mvit_oracle as orc
    z = np.load(path)
    cfg = json.loads(str(z["cfg"]))
    name = os.path.basename(path)[:-4]
    shapes = orc.block_param_shapes("", cfg["dim"], cfg["dim_out"], cfg["num_heads"], cfg["thw"], cfg["stride_q"], cfg["stride_kv"])
    params = detgen.det_params(shapes, cfg["seed"])
    blk = make_block(cfg, dtype)
    blk.load_state_dict(params, strict=True)  # the reference's own state_dict keys
    N = 1 + int(np.prod(cfg["thw"]))
    x = detgen.det_normal((cfg["B"], N, cfg["dim"]), cfg["seed"], name + ".x").cuda().requires_grad_(True)
    from test_parity2_gpu import capture_maxpool_winners, forced_max_pool
    with capture_maxpool_winners() as cap:
        y, thw = blk(x, cfg["thw"])
    assert list(thw) == cfg["thw_out"]
    assert nerr(y.detach(), z["y"]) < OUT_TOL[dtype]
    dy = detgen.det_normal(tuple(y.shape), cfg["seed"], name + ".dy").cuda()
    y.backward(dy)
    gt = GRAD_TOL[dtype]
    mp = blk.pool_skip is not None and cfg["dim"] != cfg["dim_out"]
    if mp and dtype == torch.bfloat16:
        # gradients against the (reference-pinned) oracle routed through the kernel's own max-pool winners, at the plain
        # bf16 bound; the fixture's gradients belong to the fp32 arg-maxes and are checked in fp32 mode
        po = {k: v.cuda().clone().requires_grad_(True) for k, v in params.items()}
        xo = x.detach().clone().requires_grad_(True)
        saved = orc.max_pool_tokens
        orc.max_pool_tokens = forced_max_pool(cap.wins[0])
        try:
            yo, _ = orc.multiscale_block(xo, cfg["thw"], po, "", cfg["num_heads"], cfg["stride_q"], cfg["stride_kv"],
                                         hw_switch_auto=cfg.get("hw_switch_auto", False))
        finally:
            orc.max_pool_tokens = saved
        yo.backward(dy)
        assert grad_ok(x.grad, xo.grad, dtype)
        for k, p in blk.named_parameters():
            if not k.endswith("norm_k.bias"):
                assert grad_ok(p.grad, po[k].grad, dtype), k
        return
    assert grad_ok(x.grad, z["dx"], dtype)
    for k, p in blk.named_parameters():
        g = p.grad.reshape(-1).double().cpu()
        if f"g::{k}::full" in z.files:
            ref = torch.from_numpy(z[f"g::{k}::full"]).double()
            if k.endswith("norm_k.bias"):  # analytically zero
                assert float(g.abs().max()) < (1e-4 if dtype == torch.float32 else 0.5), k
            else:
                assert nerr(g, ref) < gt, k
        else:
            idx = torch.from_numpy(z[f"g::{k}::idx"])
            rms = float(np.sqrt(z[f"g::{k}::sumsq"] / g.numel()))
            gmax = float(g.abs().max())
            assert float((g[idx] - torch.from_numpy(z[f"g::{k}::val"]).double()).abs().max()) < gt * max(gmax, rms), k
            assert abs(float((g * g).sum()) - float(z[f"g::{k}::sumsq"])) < 4 * gt * float(z[f"g::{k}::sumsq"]), k


STAGES = [  # dim, dim_out, heads, thw, stride_q, stride_kv  (the 7 distinct MViTv2-S rows)
    (96, 96, 1, [8, 56, 56], 1, 8), (96, 192, 2, [8, 56, 56], 2, 4), (192, 192, 2, [8, 28, 28], 1, 4),
    (192, 384, 4, [8, 28, 28], 2, 2), (384, 384, 4, [8, 14, 14], 1, 2), (384, 768, 8, [8, 14, 14], 2, 1),
    (768, 768, 8, [8, 7, 7], 1, 1),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("stage", range(len(STAGES)))
def test_stage_shapes_fwd_bwd_vs_oracle(stage, dtype):
    """BASELINE config 2: one MultiScaleBlock at every MViTv2-S stage shape, fwd + bwd, B = 1, against the
    oracle evaluated on the GPU in fp32."""
    from test_parity2_gpu import _stage_case
    _stage_case(STAGES[stage], dtype, 1, 100 + stage)


def test_droppath_training_mode_matches_oracle():
    """DropPath is folded into the GEMM epilogues; with the same torch RNG state the per-sample mask must
    match the reference's drop_path (common.py:46-59)."""
    from oracle import detgen, mvit_oracle as orc
    cfg = dict(dim=96, dim_out=192, num_heads=2, thw=[2, 8, 8], stride_q=[1, 2, 2], stride_kv=[1, 4, 4], seed=7)
    from pmv_b200.attention import MultiScaleBlock, set_compute_dtype
    blk = MultiScaleBlock(dim=96, dim_out=192, num_heads=2, input_size=cfg["thw"], qkv_bias=True, drop_path=0.5,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3],
                          stride_q=cfg["stride_q"], stride_kv=cfg["stride_kv"], rel_pos_spatial=True, rel_pos_temporal=True,
                          residual_pooling=True, dim_mul_in_att=True)
    set_compute_dtype(blk, torch.float32).cuda().train()
    shapes = orc.block_param_shapes("", 96, 192, 2, cfg["thw"], cfg["stride_q"], cfg["stride_kv"])
    params = {k: v.cuda() for k, v in detgen.det_params(shapes, 7).items()}
    blk.load_state_dict(params, strict=True)
    B = 6
    x = detgen.det_normal((B, 129, 96), 7, "x").cuda()
    torch.manual_seed(123)
    y, _ = blk(x, cfg["thw"])
    torch.manual_seed(123)
    keep = 0.5
    s1 = (keep + torch.rand((B,), device="cuda")).floor_() / keep
    s2 = (keep + torch.rand((B,), device="cuda")).floor_() / keep
    yo, _ = orc.multiscale_block(x, cfg["thw"], params, "", 2, cfg["stride_q"], cfg["stride_kv"], drop_scale=torch.stack([s1, s2]))
    assert nerr(y.detach(), yo) < 1e-4
    assert 0 < int((s1 == 0).sum() + (s2 == 0).sum()) < 2 * B  # the mask actually dropped something


@pytest.mark.parametrize("dtype", DTYPES)
def test_full_model_logits_match_reference_fixture(dtype):
    """BASELINE config 1/3: MViTv2-S 16x4, one deterministic clip, logits vs the reference MViT's own output."""
    from oracle import detgen, mvit_oracle as orc
    from pmv_b200 import mvit
    z = np.load(os.path.join(GOLDEN, "mvitv2_s_logits.npz"))
    model = mvit.MViT(mvit.MVITV2_S, compute_dtype=dtype)
    params = detgen.det_params(orc.param_shapes(orc.MVITV2_S), int(z["seed"]))
    model.load_state_dict(params, strict=True)
    model.cuda().eval()
    model.head.act = None
    clip = detgen.det_normal((1, 3, 16, 224, 224), int(z["seed"]), "clip").cuda()
    with torch.no_grad():
        logits = model([clip])
    assert nerr(logits, z["logits"]) < OUT_TOL[dtype]
    assert int(logits.argmax()) == int(np.argmax(z["logits"]))  # top-1 agrees


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
def test_pm_routing_matches_reference_fixture(dtype):
    """portrait / landscape mixed batch (video_model_builder.py:2075-2096) on a rectangular
    128x96 crop with hw_switch_auto, logits vs the unmodified reference MViT (tests/golden/mvitv2_s_pm_logits.npz)."""
    from oracle import detgen, mvit_oracle as orc
    from pmv_b200 import mvit
    z = np.load(os.path.join(GOLDEN, "mvitv2_s_pm_logits.npz"))
    cfg = dict(mvit.MVITV2_S, crop=(128, 96), hw_switch_auto=True)
    model = mvit.MViT(cfg, compute_dtype=dtype)
    seed = int(z["seed"])
    params = detgen.det_params(orc.param_shapes(dict(orc.MVITV2_S, crop=(128, 96))), seed)
    model.load_state_dict(params, strict=True)
    model.cuda().eval()
    model.head.act = None
    clip = detgen.det_normal((3, 3, 16, 128, 96), seed, "clip").cuda()
    pm = torch.from_numpy(z["pm"]).cuda()
    with torch.no_grad():
        logits = model([clip], pm=[pm])
        plain = model([clip])
    assert nerr(logits, z["logits"]) < OUT_TOL[dtype]
    assert (logits.argmax(1).cpu().numpy() == np.argmax(z["logits"], 1)).all()  # top-1 agrees per clip
    assert nerr(plain[1:2], z["logits"][1:2]) < OUT_TOL[dtype]                   # the landscape clip is unaffected
    assert float((plain[0] - logits[0]).abs().max()) > 1e-3                      # the routing really changes the portrait ones


def test_forward_loss_matches_unfused_tail():
    """Row f2 in the model: MViT.forward_loss (fused final LN / head / cross entropy) gives the loss and the gradients of
    cross_entropy(model(x)) through the unfused tail, for integer labels and for soft (mixup) targets."""
    from pmv_b200 import mvit
    torch.manual_seed(5)
    model = mvit.MViT(mvit.MVITV2_S, compute_dtype=torch.bfloat16).cuda().train()
    model.head.dropout.p = 0.0
    for blk in model.blocks:
        blk.drop_path_prob = 0.0
    clip = torch.randn(2, 3, 16, 224, 224, device="cuda")
    labels = torch.tensor([3, 177], device="cuda")
    soft = torch.zeros(2, 400, device="cuda")
    soft[0, 3], soft[0, 9], soft[1, 177], soft[1, 2] = 0.7, 0.3, 0.4, 0.6
    watch = ["head.projection.weight", "head.projection.bias", "norm.weight", "norm.bias", "blocks.15.mlp.fc2.weight",
             "blocks.0.attn.qkv.weight", "cls_token"]
    sd = dict(model.named_parameters())
    for target in (labels, soft):
        model.zero_grad(set_to_none=True)
        logits = model([clip])
        if target.dtype == torch.int64:
            ref = torch.nn.functional.cross_entropy(logits, target)
        else:
            ref = torch.sum(-target * torch.log_softmax(logits, dim=-1), dim=-1).mean()
        ref.backward()
        g_ref = {n: sd[n].grad.clone() for n in watch}
        model.zero_grad(set_to_none=True)
        loss, logits2 = model.forward_loss([clip], target)
        loss.backward()
        assert nerr(logits2, logits.detach()) < 1e-5
        assert abs(float(loss.detach()) - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
        for n in watch:
            # the tail itself is fp32 on both sides; below it the bf16 backward (and its atomics order) amplifies the
            # last-bit differences of the incoming gradient
            tol = 1e-4 if n.startswith(("head.", "norm.")) else 5e-2
            assert nerr(sd[n].grad, g_ref[n]) < tol, n


def test_cached_low_precision_weights_change_nothing():
    """Inference with cache_low_precision_weights gives bit-identical logits, and a weight update invalidates the copies."""
    from pmv_b200 import mvit
    from pmv_b200.attention import cache_low_precision_weights
    torch.manual_seed(7)
    model = mvit.MViT(mvit.MVITV2_S, compute_dtype=torch.bfloat16).cuda().eval()
    model.head.act = None
    clip = torch.randn(1, 3, 16, 224, 224, device="cuda")
    with torch.no_grad():
        ref = model([clip])
        assert cache_low_precision_weights(model) > 60
        got = model([clip])
        assert torch.equal(ref, got)
        w = model.blocks[5].mlp.fc1.weight
        w.mul_(1.5)  # bumps the version: the stale copy must not be used
        changed = model([clip])
        fresh = mvit.MViT(mvit.MVITV2_S, compute_dtype=torch.bfloat16).cuda().eval()
        fresh.head.act = None
        fresh.load_state_dict(model.state_dict())
        assert torch.equal(changed, fresh([clip]))
        assert not torch.equal(changed, ref)
