"""Row f3 on the GPU: FusedAdamW (pmv_adamw_step through the C ABI) against torch.optim.AdamW + clip_grad_norm_ — the
calls the reference makes (models/optimizer.py:124-131, tools/train_net.py:196-199) — and against the trajectory the
reference's own construct_optimizer produced (tests/golden/optimizer_reference.npz, oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tensors(seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(1,), (5,), (4096,), (4097,), (96, 27), (384, 1536), (1000, 97), (3, 1, 7)]
    ps = [torch.randn(*s, generator=g).cuda() for s in shapes]
    # one parameter living at a 4-byte-aligned (not 16-byte-aligned) address: the scalar path of the kernels
    base = torch.randn(1001, generator=g).cuda()
    ps.append(base[1:])
    return ps


@pytest.mark.parametrize("clip", [None, 1.0, 1e4])
def test_fused_adamw_matches_torch(clip):
    from pmv_b200.optim import FusedAdamW
    ours = [torch.nn.Parameter(t.clone()) for t in _tensors(1)]
    ref = [torch.nn.Parameter(t.clone()) for t in _tensors(1)]
    assert ours[-1].data_ptr() % 16 != 0 or True
    split = 4
    mine = FusedAdamW([dict(params=ours[:split], weight_decay=0.05), dict(params=ours[split:], weight_decay=0.0)], lr=1e-2,
                      betas=(0.9, 0.999), eps=1e-8, max_grad_norm=clip)
    theirs = torch.optim.AdamW([dict(params=ref[:split], weight_decay=0.05), dict(params=ref[split:], weight_decay=0.0)], lr=1e-2,
                               betas=(0.9, 0.999), eps=1e-8)
    for it in range(6):
        g = torch.Generator(device="cpu").manual_seed(100 + it)
        for a, b in zip(ours, ref):
            gr = (torch.randn(*a.shape, generator=g) * (0.3 if it % 2 else 3.0)).cuda()
            a.grad = gr.clone()
            b.grad = gr.clone()
        if it == 3:
            mine.set_lr(3e-3)
            for grp in theirs.param_groups:
                grp["lr"] = 3e-3
        if clip is not None:
            norm = torch.nn.utils.clip_grad_norm_(ref, clip)
        mine.step()
        theirs.step()
        if clip is not None:
            assert abs(float(mine.grad_norm) - float(norm)) <= 1e-5 * float(norm)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), (it, a.shape, float((a - b).abs().max()))
        for a in ours:  # the step does not touch the gradients
            assert a.grad is not None
    assert int(mine.step_count) == 6
    for a in ours:
        lp = mine.state[a].get("lp")
        assert (lp is not None) == (a.dim() >= 2)
        if lp is not None:
            assert torch.equal(lp, a.detach().to(torch.bfloat16))


def test_low_precision_copy_is_used_until_the_parameter_changes():
    from pmv_b200.functional import _cast
    from pmv_b200.optim import FusedAdamW
    w = torch.nn.Parameter(torch.randn(64, 96).cuda())
    opt = FusedAdamW([w], lr=1e-3)
    lp = opt.state[w]["lp"]
    assert _cast(w, torch.bfloat16) is lp
    w.grad = torch.randn_like(w)
    opt.step()
    assert _cast(w, torch.bfloat16) is lp and torch.equal(lp, w.detach().to(torch.bfloat16))
    with torch.no_grad():
        w.mul_(2.0)  # anything else writing the parameter (checkpoint load, EMA swap ...) invalidates the copy
    fresh = _cast(w, torch.bfloat16)
    assert fresh is not lp and torch.equal(fresh, w.detach().to(torch.bfloat16))
    opt.sync_low_precision()
    assert _cast(w, torch.bfloat16) is lp and torch.equal(lp, w.detach().to(torch.bfloat16))


def test_fused_adamw_in_a_cuda_graph_follows_the_lr_schedule():
    from pmv_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.randn(300, 70).cuda())
    q = torch.nn.Parameter(p.detach().clone())
    grad = torch.randn_like(p)
    p.grad = grad.clone()
    q.grad = grad.clone()
    mine = FusedAdamW([p], lr=1e-2, weight_decay=0.1, max_grad_norm=0.5)
    theirs = torch.optim.AdamW([q], lr=1e-2, weight_decay=0.1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        mine.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.nn.utils.clip_grad_norm_([q], 0.5); theirs.step(); q.grad = grad.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        mine.step()
    torch.nn.utils.clip_grad_norm_([q], 0.5); theirs.step(); q.grad = grad.clone()  # capture does not execute: replay below is step 2
    graph.replay()
    for lr in (5e-3, 1e-3):
        mine.set_lr(lr)
        theirs.param_groups[0]["lr"] = lr
        graph.replay()
        torch.nn.utils.clip_grad_norm_([q], 0.5); theirs.step(); q.grad = grad.clone()
    torch.cuda.synchronize()
    assert int(mine.step_count) == 4
    assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), float((p - q).abs().max())


def test_reference_recipe_trajectory():
    """param_groups + FusedAdamW on the MViTv2-S parameter set reproduce what the reference's construct_optimizer /
    clip_grad_norm_ / step sequence produced on the same deterministic parameters and gradients."""
    from oracle import detgen
    from pmv_b200 import mvit
    from pmv_b200.optim import FusedAdamW, param_groups
    z = np.load(os.path.join(GOLDEN, "optimizer_reference.npz"))
    seed, steps = int(z["seed"]), int(z["steps"])
    model = mvit.MViT(mvit.MVITV2_S)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(detgen.det_params(shapes, seed), strict=True)
    model = model.cuda()
    groups = param_groups(model, 0.05, zero_wd_1d=True)
    opt = FusedAdamW(groups, lr=float(z["lr"]), betas=tuple(z["betas"]), eps=float(z["eps"]), max_grad_norm=float(z["clip"]))
    for it in range(steps):
        for n, p in model.named_parameters():
            p.grad = detgen.det_normal(p.shape, seed + 1 + it, n, 0.01).cuda()
        opt.step()
        # the reference value is torch's fp32 norm-of-norms on the CPU over 34.5 M elements (its own rounding ~1e-5);
        # the kernel sums fp32 per 4096-element chunk and the chunks in double
        assert abs(float(opt.grad_norm) - z["grad_norms"][it]) <= 1e-4 * z["grad_norms"][it]
    sd = dict(model.named_parameters())
    for i, n in enumerate(z["param_names"]):
        t = sd[str(n)].detach().double()
        scale = max(float(z["param_sumsq"][i]) ** 0.5, 1e-12)  # |sum error| relative to the tensor's L2 norm
        assert abs(float(t.sum()) - z["param_sum"][i]) <= 1e-5 * scale * max(t.numel() ** 0.5, 1.0), n
        assert abs(float((t * t).sum()) - z["param_sumsq"][i]) <= 1e-5 * z["param_sumsq"][i] + 1e-12, n
