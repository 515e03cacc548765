"""The CPU oracle must reproduce the fixtures the UNMODIFIED reference produced
(tests/golden/, written by oracle/make_golden.py in the build container)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import detgen, mvit_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BLOCK_FILES = sorted(glob.glob(os.path.join(GOLDEN, "blk_*.npz")))
TOL = 2e-5  # fp32 vs fp32, different summation orders; observed <= 7e-7


def nerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_oracle_block(cfg):
    shapes = orc.block_param_shapes("", cfg["dim"], cfg["dim_out"], cfg["num_heads"], cfg["thw"],
                                    cfg["stride_q"], cfg["stride_kv"])
    params = {k: v.requires_grad_(True) for k, v in detgen.det_params(shapes, cfg["seed"]).items()}
    N = 1 + int(np.prod(cfg["thw"]))
    return shapes, params, N


@pytest.mark.parametrize("path", BLOCK_FILES, ids=[os.path.basename(p)[:-4] for p in BLOCK_FILES])
def test_block_fwd_bwd_matches_reference(path):
    z = np.load(path)
    cfg = json.loads(str(z["cfg"]))
    name = os.path.basename(path)[:-4]
    shapes, params, N = run_oracle_block(cfg)
    x = detgen.det_normal((cfg["B"], N, cfg["dim"]), cfg["seed"], name + ".x").requires_grad_(True)
    y, thw = orc.multiscale_block(x, cfg["thw"], params, "", cfg["num_heads"], cfg["stride_q"], cfg["stride_kv"],
                                  hw_switch_auto=cfg["hw_switch_auto"])
    assert list(thw) == cfg["thw_out"]
    assert nerr(y.detach(), z["y"]) < TOL
    dy = detgen.det_normal(tuple(y.shape), cfg["seed"], name + ".dy")
    y.backward(dy)
    assert nerr(x.grad, z["dx"]) < TOL
    for k, p in params.items():
        g = p.grad.reshape(-1).double()
        if f"g::{k}::full" in z.files:
            ref = torch.from_numpy(z[f"g::{k}::full"]).double()
            if k.endswith("norm_k.bias"):  # analytically zero (KAT ii)
                assert float(g.abs().max()) < 1e-4 and float(ref.abs().max()) < 1e-4
            else:
                assert nerr(g, ref) < TOL, k
        else:
            idx = torch.from_numpy(z[f"g::{k}::idx"])
            scale = float(np.sqrt(z[f"g::{k}::sumsq"] / g.numel()))
            assert float((g[idx] - torch.from_numpy(z[f"g::{k}::val"]).double()).abs().max()) < 50 * TOL * scale + 1e-7, k
            assert abs(float((g * g).sum()) - float(z[f"g::{k}::sumsq"])) < 1e-4 * float(z[f"g::{k}::sumsq"]), k


def test_function_level_fixtures():
    z = np.load(os.path.join(GOLDEN, "functions.npz"))
    seed, B, nh, C = 77, 2, 2, 96
    thw, stride = [2, 6, 4], [1, 2, 2]
    N = 1 + int(np.prod(thw))
    x = detgen.det_normal((B, nh, N, C), seed, "fn.x")
    w = detgen.det_normal((C, 1, 3, 3, 3), seed, "fn.w", 0.2)
    lw = detgen.det_normal((C,), seed, "fn.lw", 0.1, 1.0)
    lb = detgen.det_normal((C,), seed, "fn.lb", 0.1)
    y, thw_o = orc.conv_pool_tokens(x, thw, w, stride, True, lw, lb)
    assert list(thw_o) == list(z["pool_thw"])
    assert nerr(y, z["pool_y"]) < TOL
    yt, _ = orc.conv_pool_tokens_taps(x, thw, w, stride, True, lw, lb)
    assert nerr(yt, z["pool_y"]) < TOL
    xs = detgen.det_normal((B, N, 192), seed, "fn.xs")
    assert nerr(orc.max_pool_tokens(xs, thw, [1, 3, 3], [1, 2, 2], True)[0], z["maxpool_y"]) == 0.0
    q = detgen.det_normal((B, nh, 13, C), seed, "fn.q")
    attn = detgen.det_normal((B, nh, 13, 49), seed, "fn.attn")
    rh = detgen.det_normal((11, C), seed, "fn.rh", 0.3)
    rw = detgen.det_normal((5, C), seed, "fn.rw", 0.3)
    rt = detgen.det_normal((3, C), seed, "fn.rt", 0.3)
    out = orc.add_rel_pos_bias(attn, q, True, [2, 3, 2], [2, 6, 4], rh, rw, rt)
    assert nerr(out, z["relpos_attn"]) < TOL
    assert nerr(orc.interp_rel_table(rw, 7), z["interp_5_to_7"]) < TOL


def test_full_model_logits_and_param_count():
    z = np.load(os.path.join(GOLDEN, "mvitv2_s_logits.npz"))
    shapes = orc.param_shapes(orc.MVITV2_S)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 34537744 == int(z["nparam"])
    assert sum(int(np.prod(s)) for s in orc.param_shapes(orc.MVITV2_B).values()) == 51230128
    params = detgen.det_params(shapes, int(z["seed"]))
    clip = detgen.det_normal((1, 3, 16, 224, 224), int(z["seed"]), "clip")
    with torch.no_grad():
        logits = orc.mvit_forward(clip, params, orc.MVITV2_S)
        probs = orc.mvit_forward(clip, params, orc.MVITV2_S, softmax_head=True)
    assert nerr(logits, z["logits"]) < TOL
    assert int(logits.argmax()) == int(np.argmax(z["logits"]))
    assert abs(float(probs.sum()) - 1.0) < 1e-5  # eval head sums to 1 (KAT iv)


def test_block_schedule_matches_survey_appendix_a():
    s = orc.block_schedule(orc.MVITV2_S)
    rows = [(b["dim"], b["dim_out"], b["num_heads"], tuple(b["thw"]), tuple(b["stride_q"]), tuple(b["stride_kv"])) for b in s]
    assert rows[0] == (96, 96, 1, (8, 56, 56), (1, 1, 1), (1, 8, 8))
    assert rows[1] == (96, 192, 2, (8, 56, 56), (1, 2, 2), (1, 4, 4))
    assert rows[2] == (192, 192, 2, (8, 28, 28), (1, 1, 1), (1, 4, 4))
    assert rows[3] == (192, 384, 4, (8, 28, 28), (1, 2, 2), (1, 2, 2))
    assert all(r == (384, 384, 4, (8, 14, 14), (1, 1, 1), (1, 2, 2)) for r in rows[4:14])
    assert rows[14] == (384, 768, 8, (8, 14, 14), (1, 2, 2), (1, 1, 1))
    assert rows[15] == (768, 768, 8, (8, 7, 7), (1, 1, 1), (1, 1, 1))
    b = orc.block_schedule(orc.MVITV2_B)
    assert len(b) == 24 and b[2]["dim_out"] == 192 and b[21]["dim_out"] == 768 and b[0]["thw"] == [16, 56, 56]


def test_pm_routing_logits_oracle_vs_reference():
    """Portrait / landscape batch routing (video_model_builder.py:2075-2096) on a rectangular crop with
    hw_switch_auto: oracle restatement against logits of the unmodified reference MViT (oracle/make_golden.py)."""
    from oracle import detgen, mvit_oracle as orc
    z = np.load(os.path.join(GOLDEN, "mvitv2_s_pm_logits.npz"))
    cfg = dict(orc.MVITV2_S, crop=(128, 96), hw_switch_auto=True)
    seed = int(z["seed"])
    params = detgen.det_params(orc.param_shapes(cfg), seed)
    clip = detgen.det_normal((3, 3, 16, 128, 96), seed, "clip")
    pm = torch.from_numpy(z["pm"])
    with torch.no_grad():
        out = orc.mvit_forward_pm(clip, pm, params, cfg)
    assert nerr(out, z["logits"]) < TOL
    # all-landscape mask == plain forward
    with torch.no_grad():
        a = orc.mvit_forward_pm(clip[1:2], torch.tensor([False]), params, cfg)
    assert nerr(a, z["logits"][1:2]) < TOL
