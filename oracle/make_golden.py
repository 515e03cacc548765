"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/make_golden.py [--check]

For every case the reference ``MultiScaleBlock`` (MViT/slowfast/models/attention.py:464-589)
is built with the reference constructor, loaded (strict) with deterministic weights from
``oracle/detgen.py`` and run forward + backward in fp32 on CPU; the fixture stores the
configuration, the seeds, the output, the input gradient and the parameter gradients
(full for small tensors, moments + samples for matrices).  A full-model fixture stores the
MViTv2-S logits for one deterministic clip.  With --check nothing is written: the oracle
is compared with the live reference instead.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import detgen, mvit_oracle as orc, ref_loader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

BLOCK_CASES = [
    # name, dim, dim_out, heads, thw, stride_q, stride_kv, B, hw_switch_auto
    ("blk_s0_kv8", 96, 96, 1, [2, 8, 8], [1, 1, 1], [1, 8, 8], 2, False),
    ("blk_s1_q2_kv4", 96, 192, 2, [2, 8, 8], [1, 2, 2], [1, 4, 4], 2, False),
    ("blk_rect_switch", 192, 192, 2, [2, 6, 4], [1, 1, 1], [1, 2, 2], 1, True),
    ("blk_q_lt_k", 96, 96, 1, [3, 4, 4], [1, 2, 2], [1, 1, 1], 2, False),
    ("blk_odd_ratio", 96, 96, 1, [2, 7, 5], [1, 2, 2], [1, 1, 1], 1, False),
    ("blk_odd_interp", 96, 96, 1, [2, 7, 5], [1, 2, 2], [1, 2, 2], 1, False),
]
SAMPLES = 16


def norm_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_ref_block(att, dim, dim_out, nh, thw, sq, skv, hw_switch):
    return att.MultiScaleBlock(
        dim=dim, dim_out=dim_out, num_heads=nh, input_size=thw, mlp_ratio=4.0, qkv_bias=True,
        drop_rate=0.0, drop_path=0.0, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
        kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3], stride_q=sq, stride_kv=skv, mode="conv",
        has_cls_embed=True, pool_first=False, rel_pos_spatial=True, rel_pos_temporal=True,
        rel_pos_zero_init=False, residual_pooling=True, dim_mul_in_att=True, separate_qkv=False,
        hw_switch_auto=hw_switch)


def summarize_grad(name, g: torch.Tensor):
    g = g.detach().reshape(-1).double()
    if g.numel() <= 4096:
        return {"full": g.float().numpy()}
    idx = np.random.Generator(np.random.PCG64(__import__("zlib").crc32(name.encode()))).integers(0, g.numel(), SAMPLES)
    idx = np.sort(idx)
    return {"sum": np.float64(g.sum()), "sumsq": np.float64((g * g).sum()), "idx": idx.astype(np.int64),
            "val": g[idx].float().numpy()}


def run_block_case(att, case, seed=1234):
    name, dim, dim_out, nh, thw, sq, skv, B, hw = case
    shapes = orc.block_param_shapes("", dim, dim_out, nh, thw, sq, skv)
    params = detgen.det_params(shapes, seed)
    blk = build_ref_block(att, dim, dim_out, nh, thw, sq, skv, hw)
    ref_shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    assert ref_shapes == {k: tuple(v) for k, v in shapes.items()}, (name, set(ref_shapes) ^ set(shapes))
    blk.load_state_dict(params, strict=True)
    N = 1 + thw[0] * thw[1] * thw[2]
    x = detgen.det_normal((B, N, dim), seed, name + ".x").requires_grad_(True)
    y, thw_new = blk(x, list(thw))
    dy = detgen.det_normal(tuple(y.shape), seed, name + ".dy")
    y.backward(dy)
    grads = {k: p.grad for k, p in blk.named_parameters()}

    # oracle on the same data
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xo = x.detach().clone().requires_grad_(True)
    yo, thw_o = orc.multiscale_block(xo, thw, po, "", nh, sq, skv, hw_switch_auto=hw)
    yo.backward(dy)
    errs = {"y": norm_err(yo, y), "dx": norm_err(xo.grad, x.grad)}
    for k in grads:
        if k.endswith("norm_k.bias"):
            # analytically zero (a constant added to every key shifts each score row uniformly,
            # SURVEY.md section 4 KAT ii): compare absolutely, both sides are rounding noise
            assert float(grads[k].abs().max()) < 1e-4 and float(po[k].grad.abs().max()) < 1e-4
            continue
        errs["d" + k] = norm_err(po[k].grad, grads[k])
    assert list(thw_o) == list(thw_new)
    return dict(name=name, cfg=dict(dim=dim, dim_out=dim_out, num_heads=nh, thw=thw, stride_q=sq, stride_kv=skv,
                                    B=B, hw_switch_auto=hw, seed=seed, thw_out=list(thw_new)),
                y=y.detach(), dx=x.grad, grads=grads, errs=errs)


def save_block_case(res):
    arrs = {"cfg": np.array(json.dumps(res["cfg"])), "y": res["y"].numpy(), "dx": res["dx"].numpy()}
    for k, g in res["grads"].items():
        for kk, vv in summarize_grad(k, g).items():
            arrs[f"g::{k}::{kk}"] = vv
    np.savez_compressed(os.path.join(GOLDEN, res["name"] + ".npz"), **arrs)


def run_function_cases(att, seed=77):
    """Function-level fixtures: attention_pool, cal_rel_pos_spatial/temporal, get_rel_pos."""
    out = {}
    B, nh, C = 2, 2, 96
    thw, stride = [2, 6, 4], [1, 2, 2]
    N = 1 + thw[0] * thw[1] * thw[2]
    x = detgen.det_normal((B, nh, N, C), seed, "fn.x")
    w = detgen.det_normal((C, 1, 3, 3, 3), seed, "fn.w", 0.2)
    lw = detgen.det_normal((C,), seed, "fn.lw", 0.1, 1.0)
    lb = detgen.det_normal((C,), seed, "fn.lb", 0.1)
    conv = torch.nn.Conv3d(C, C, 3, stride=stride, padding=1, groups=C, bias=False)
    conv.weight.data.copy_(w)
    ln = torch.nn.LayerNorm(C, eps=1e-6)
    ln.weight.data.copy_(lw)
    ln.bias.data.copy_(lb)
    with torch.no_grad():
        y, thw_o = att.attention_pool(x, conv, thw, has_cls_embed=True, norm=ln)
    out["pool_y"] = y.numpy()
    out["pool_thw"] = np.array(thw_o)
    yo, thw_oo = orc.conv_pool_tokens(x, thw, w, stride, True, lw, lb)
    yt, _ = orc.conv_pool_tokens_taps(x, thw, w, stride, True, lw, lb)
    errs = {"pool": norm_err(yo, y), "pool_taps": norm_err(yt, y)}
    # max-pool skip path
    xs = detgen.det_normal((B, N, 192), seed, "fn.xs")
    mp = torch.nn.MaxPool3d([1, 3, 3], [1, 2, 2], [0, 1, 1], ceil_mode=False)
    ys, _ = att.attention_pool(xs, mp, thw, has_cls_embed=True)
    out["maxpool_y"] = ys.numpy()
    errs["maxpool"] = norm_err(orc.max_pool_tokens(xs, thw, [1, 3, 3], [1, 2, 2], True)[0], ys)
    # rel-pos bias: q 2x3x2 vs k 2x6x4 (q<k), plus interpolated tables
    q_shape, k_shape = [2, 3, 2], [2, 6, 4]
    Nq, Nk = 1 + 12, 1 + 48
    q = detgen.det_normal((B, nh, Nq, C), seed, "fn.q")
    k = detgen.det_normal((B, nh, Nk, C), seed, "fn.k")
    attn = detgen.det_normal((B, nh, Nq, Nk), seed, "fn.attn")
    rh = detgen.det_normal((11, C), seed, "fn.rh", 0.3)
    rw = detgen.det_normal((5, C), seed, "fn.rw", 0.3)  # wrong length on purpose -> interpolated to 7
    rt = detgen.det_normal((3, C), seed, "fn.rt", 0.3)
    with torch.no_grad():
        a1 = att.cal_rel_pos_spatial(attn.clone(), q, k, True, q_shape, k_shape, rh, rw)
        a2 = att.cal_rel_pos_temporal(a1.clone(), q, True, q_shape, k_shape, rt)
    out["relpos_attn"] = a2.numpy()
    errs["relpos"] = norm_err(orc.add_rel_pos_bias(attn, q, True, q_shape, k_shape, rh, rw, rt), a2)
    out["interp_5_to_7"] = att.get_rel_pos(rw, 7).numpy()
    errs["interp"] = norm_err(orc.interp_rel_table(rw, 7), att.get_rel_pos(rw, 7))
    return out, errs


def run_full_model(seed=4321):
    model, cfg = ref_loader.load_full_model("configs/Kinetics/MVITv2_S_16x4.yaml")
    shapes = orc.param_shapes(orc.MVITV2_S)
    ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ref_shapes == shapes, set(ref_shapes) ^ set(shapes)
    nparam = sum(int(np.prod(s)) for s in shapes.values())
    assert nparam == 34537744, nparam  # published 34.5 M (projects/mvitv2/README.md:28)
    params = detgen.det_params(shapes, seed)
    model.load_state_dict(params, strict=True)
    model.eval()
    model.head.act = None  # logits instead of eval softmax (head_helper.py:568-570)
    clip = detgen.det_normal((1, 3, 16, 224, 224), seed, "clip")
    with torch.no_grad():
        logits = model([clip])
        lo = orc.mvit_forward(clip, params, orc.MVITV2_S)
    return dict(logits=logits.numpy(), seed=seed, nparam=nparam), {"logits": norm_err(lo, logits)}


PM_CFG = dict(orc.MVITV2_S, crop=(128, 96), hw_switch_auto=True)


def run_pm_model(seed=977):
    """Portrait / landscape routing (video_model_builder.py:2075-2096) on a rectangular crop with
    TRAIN_CROP_SIZE_RECT_SWITCH_AUTO: 3 clips, two of them portrait (stored transposed, as the loader delivers them)."""
    model, cfg = ref_loader.load_full_model("configs/Kinetics/MVITv2_S_16x4.yaml", overrides={
        "DATA.TRAIN_CROP_SIZE_RECT": [128, 96], "DATA.TEST_CROP_SIZE_RECT": [128, 96],
        "DATA.TRAIN_CROP_SIZE_RECT_SWITCH_AUTO": True, "TEST.PROCESS": False})
    shapes = orc.param_shapes(PM_CFG)
    ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ref_shapes == shapes, set(ref_shapes) ^ set(shapes)
    params = detgen.det_params(shapes, seed)
    model.load_state_dict(params, strict=True)
    model.eval()
    model.head.act = None
    clip = detgen.det_normal((3, 3, 16, 128, 96), seed, "clip")
    pm = torch.tensor([True, False, True])
    with torch.no_grad():
        logits = model([clip], pm=[pm])
        lo = orc.mvit_forward_pm(clip, pm, params, PM_CFG)
        # routing really matters: the same clips all treated as landscape give different logits
        plain = model([clip])
    assert float((plain[0] - logits[0]).abs().max()) > 1e-3
    assert float((plain[1] - logits[1]).abs().max()) < 1e-5
    return dict(logits=logits.numpy(), pm=pm.numpy(), seed=seed), {"pm_logits": norm_err(lo, logits)}


def run_checkpoint_case(seed=555):
    """Row f4: the reference's own load_checkpoint (utils/checkpoint.py:191-563) loading a 224-crop MViTv2-S checkpoint
    into a 160-crop model: rel-pos tables of a different length are interpolated.  Stores the resulting tables of three
    blocks and a checksum of every tensor."""
    import tempfile
    model224, _ = ref_loader.load_full_model("configs/Kinetics/MVITv2_S_16x4.yaml")
    params = detgen.det_params(orc.param_shapes(orc.MVITV2_S), seed)
    model224.load_state_dict(params, strict=True)
    path = os.path.join(tempfile.mkdtemp(), "ck.pyth")
    torch.save({"model_state": model224.state_dict(), "epoch": 3}, path)
    model160, _ = ref_loader.load_full_model("configs/Kinetics/MVITv2_S_16x4.yaml",
                                             overrides={"DATA.TRAIN_CROP_SIZE": 160, "DATA.TEST_CROP_SIZE": 160})
    import slowfast.utils.checkpoint as cu

    class _PM:
        exists = staticmethod(os.path.exists)
        open = staticmethod(open)
    cu.pathmgr = _PM
    epoch = cu.load_checkpoint(path, model160, data_parallel=False, optimizer=None, inflation=False,
                               convert_from_caffe2=False, epoch_reset=False, clear_name_pattern=(), image_init=False)
    sd = model160.state_dict()
    out = dict(seed=seed, epoch=int(epoch))
    for blk in (0, 3, 15):
        for n in ("rel_pos_h", "rel_pos_w", "rel_pos_t"):
            out[f"blocks.{blk}.attn.{n}"] = sd[f"blocks.{blk}.attn.{n}"].numpy()
    out["checksum"] = np.array([float(v.double().sum()) for v in sd.values()])
    out["names"] = np.array(list(sd.keys()))
    return out


def run_optimizer_case(seed=8642, steps=3):
    """Row f3 pin: the reference's own construct_optimizer (models/optimizer.py:14-131) on its MViTv2-S, and the
    reference's step sequence (tools/train_net.py:190-199: clip_grad_norm_(CLIP_GRAD_L2NORM) then optimizer.step()) on
    deterministic parameters / gradients.  Stored: group membership by parameter name, the hyper-parameters, the
    gradient norm of every step and per-tensor fp64 (sum, sum of squares) of the parameters after `steps` steps."""
    model, cfg = ref_loader.load_full_model("configs/Kinetics/MVITv2_S_16x4.yaml")
    import slowfast.models.optimizer as ropt
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(detgen.det_params(shapes, seed), strict=True)
    opt = ropt.construct_optimizer(model, cfg)
    names = {id(p): n for n, p in model.named_parameters()}
    out = dict(seed=seed, steps=steps, clip=float(cfg.SOLVER.CLIP_GRAD_L2NORM), lr=float(cfg.SOLVER.BASE_LR),
               betas=np.asarray(cfg.SOLVER.BETAS, np.float64), eps=float(opt.param_groups[0]["eps"]),
               optimizer=type(opt).__name__)
    for gi, g in enumerate(opt.param_groups):
        out[f"group{gi}_weight_decay"] = float(g["weight_decay"])
        out[f"group{gi}_names"] = np.asarray([names[id(p)] for p in g["params"]])
    out["ngroups"] = len(opt.param_groups)
    norms = []
    for it in range(steps):
        for n, p in model.named_parameters():
            # gradients large enough that the clip at 1.0 is active, different every step
            p.grad = detgen.det_normal(p.shape, seed + 1 + it, n, 0.01)
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.SOLVER.CLIP_GRAD_L2NORM)))
        opt.step()
    out["grad_norms"] = np.asarray(norms, np.float64)
    pn = sorted(n for n, _ in model.named_parameters())
    out["param_names"] = np.asarray(pn)
    sd = dict(model.named_parameters())
    out["param_sum"] = np.asarray([float(sd[n].detach().double().sum()) for n in pn], np.float64)
    out["param_sumsq"] = np.asarray([float((sd[n].detach().double() ** 2).sum()) for n in pn], np.float64)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    assert ref_loader.reference_available(), "needs /root/reference (build container)"
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    att = ref_loader.load_attention()
    worst = 0.0
    for case in BLOCK_CASES:
        res = run_block_case(att, case)
        e = max(res["errs"].values())
        worst = max(worst, e)
        print(f"{res['name']:>18s}  oracle-vs-reference max normalised err {e:.2e}  (y {res['errs']['y']:.1e}, dx {res['errs']['dx']:.1e})")
        if not args.check:
            save_block_case(res)
    fn, errs = run_function_cases(att)
    print("function cases:", {k: f"{v:.1e}" for k, v in errs.items()})
    worst = max(worst, max(errs.values()))
    if not args.check:
        np.savez_compressed(os.path.join(GOLDEN, "functions.npz"), **fn)
    if not args.skip_full:
        full, errs = run_full_model()
        print("full MViTv2-S logits:", {k: f"{v:.1e}" for k, v in errs.items()}, "params", full["nparam"])
        worst = max(worst, max(errs.values()))
        if not args.check:
            np.savez_compressed(os.path.join(GOLDEN, "mvitv2_s_logits.npz"), **full)
    if not args.skip_full:
        pmres, errs = run_pm_model()
        print("portrait/landscape routing, rect 128x96:", {k: f"{v:.1e}" for k, v in errs.items()})
        worst = max(worst, max(errs.values()))
        if not args.check:
            np.savez_compressed(os.path.join(GOLDEN, "mvitv2_s_pm_logits.npz"), **pmres)
    if not args.skip_full and not args.check:
        ck = run_checkpoint_case()
        print("checkpoint surgery case: epoch", ck["epoch"], "tables", [k for k in ck if k.startswith("blocks")][:3], "...")
        np.savez_compressed(os.path.join(GOLDEN, "checkpoint_surgery.npz"), **ck)
        og = run_optimizer_case()
        print("optimizer case:", og["optimizer"], "groups", [(og[f"group{i}_weight_decay"], len(og[f"group{i}_names"])) for i in range(og["ngroups"])],
              "grad norms", og["grad_norms"])
        np.savez_compressed(os.path.join(GOLDEN, "optimizer_reference.npz"), **og)
    print(f"worst oracle-vs-reference error {worst:.2e}")
    assert worst < 5e-5, worst


if __name__ == "__main__":
    main()
