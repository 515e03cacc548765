"""Kernel-level parity on the GPU: every C-ABI kernel against the oracle / a plain fp32 torch statement of
the same op, on seeded inputs.  Tolerances: fp32 mode 1e-4, bf16 mode 1e-2 (normalised max error
max|a-b| / max|b|, the metric BASELINE.json's north star states)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}
DTYPES = [torch.float32, torch.bfloat16]


def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C", [(1000, 96), (777, 768), (33, 384)])
def test_layernorm_fwd_bwd(dtype, rows, C):
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    x = randn(rows, C, seed=1) * 2 + 0.3
    g, b = randn(C, seed=2) * 0.1 + 1, randn(C, seed=3) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b, dtype)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = orc.layer_norm(xr, gr, br)
    assert nerr(y.float(), yr.detach()) < TOL[dtype]
    dy = randn(rows, C, seed=4).to(dtype)
    yr.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd)
    assert nerr(dx, xr.grad) < TOL[dtype]
    assert nerr(dg, gr.grad) < TOL[dtype]
    assert nerr(db, br.grad) < TOL[dtype]
    acc = torch.ones_like(x)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx_accum=acc)
    assert nerr(acc, xr.grad + 1) < TOL[dtype]


POOL_CASES = [  # B, heads, thw, stride
    (2, 1, (2, 8, 8), 8), (2, 2, (2, 8, 8), 2), (1, 2, (2, 7, 5), 2), (1, 4, (8, 14, 14), 1), (2, 2, (3, 12, 10), 4),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,heads,thw,s", POOL_CASES)
def test_pool_ln_fwd_bwd(dtype, B, heads, thw, s):
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    T, H, W = thw
    N = 1 + T * H * W
    qkv = randn(B, N, 3, heads, 96, seed=5).to(dtype)
    w = randn(96, 1, 3, 3, 3, seed=6) * 0.2
    g, b = randn(96, seed=7) * 0.1 + 1, randn(96, seed=8) * 0.1
    Lo = T * ops.pooled_hw(H, s) * ops.pooled_hw(W, s)
    for which in range(3):
        ld = 128 if which == 0 else 96
        out = torch.zeros(B, heads, 1 + Lo, ld, dtype=dtype, device="cuda")
        ops.pool_ln_fwd(qkv, which, heads, thw, s, w, g, b, out)
        xin = qkv[:, :, which].permute(0, 2, 1, 3).float().clone().requires_grad_(True)
        wr, gr, br = (t.clone().requires_grad_(True) for t in (w, g, b))
        ref, thw_o = orc.conv_pool_tokens(xin, thw, wr, (1, s, s), True, gr, br)
        assert ref.shape[2] == 1 + Lo
        assert nerr(out[..., :96].float(), ref.detach()) < TOL[dtype], which
        assert float(out[..., 96:].abs().max()) == 0.0 if ld > 96 else True
        dout = torch.zeros(B, heads, 1 + Lo, ld, dtype=dtype, device="cuda")
        dout[..., :96] = randn(B, heads, 1 + Lo, 96, seed=9 + which).to(dtype)
        ref.backward(dout[..., :96].float())
        dqkv = torch.full_like(qkv, float("nan"))
        dwg = torch.zeros(96 * 27 + 192, device="cuda")
        ops.pool_ln_bwd(qkv, which, heads, thw, s, w, g, dout, dqkv, dwg)
        got_dx = dqkv[:, :, which].permute(0, 2, 1, 3).float()
        assert nerr(got_dx, xin.grad) < TOL[dtype], which
        assert nerr(dwg[:2592].view(96, 27), wr.grad.view(96, 27)) < TOL[dtype]
        assert nerr(dwg[2592:2688], gr.grad) < TOL[dtype]
        assert nerr(dwg[2688:], br.grad) < TOL[dtype]


@pytest.mark.parametrize("B,thw,C", [(2, (2, 8, 8), 192), (1, (2, 7, 5), 96), (2, (8, 14, 14), 768), (1, (1, 1, 1), 32), (1, (2, 9, 16), 64)])
def test_maxpool_skip(B, thw, C):
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    T, H, W = thw
    x = randn(B, 1 + T * H * W, C, seed=11)
    y, win = ops.maxpool_skip_fwd(x, thw, want_winner=True)
    assert torch.equal(ops.maxpool_skip_fwd(x, thw), y)
    xr = x.clone().requires_grad_(True)
    yr, _ = orc.max_pool_tokens(xr, thw, (1, 3, 3), (1, 2, 2), True)
    assert torch.equal(y, yr.detach())
    dy = randn(*y.shape, seed=12)
    yr.backward(dy)
    dx = ops.maxpool_skip_bwd(win, dy, thw)
    assert nerr(dx, xr.grad) < 1e-6


GEMM_SHAPES = [(300, 288, 96), (1000, 96, 384), (129, 1152, 384), (513, 768, 3072), (64, 400, 768)]


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_layouts(tc, M, N, K):
    """TN / NN / reduce-over-rows against torch fp32 matmul (bf16 operands on both sides)."""
    from pmv_b200 import _lib as L, ops
    dtype = torch.bfloat16 if tc else torch.float32
    x = randn(M, K, seed=20).to(dtype)
    w = (randn(N, K, seed=21) * 0.05).to(dtype)
    dy = randn(M, N, seed=22).to(dtype)
    tol = TOL[dtype]
    y = ops.linear_fwd(x, w, None, dtype, tc=tc)
    assert nerr(y.float(), x.float() @ w.float().t()) < tol
    dx = ops.linear_dgrad(dy, w, torch.float32, tc=tc)
    assert nerr(dx, dy.float() @ w.float()) < tol
    dw = ops.linear_wgrad(dy, x, tc=tc)
    assert nerr(dw, dy.float().t() @ x.float()) < tol


@pytest.mark.parametrize("tc", [0, 1])
def test_gemm_epilogues(tc):
    from pmv_b200 import _lib as L, ops
    dtype = torch.bfloat16 if tc else torch.float32
    tol = TOL[dtype]
    B, Nq, K, N = 3, 101, 192, 384
    M = B * Nq
    x = randn(M, K, seed=30).to(dtype)
    w = (randn(N, K, seed=31) * 0.05).to(dtype)
    bias = randn(N, seed=32) * 0.1
    res = randn(M, N, seed=33)
    scale = torch.tensor([0.0, 1.25, 1.25], device="cuda")
    ref_lin = x.float() @ w.float().t() + bias
    # bias + DropPath scale + residual -> fp32
    y = ops.linear_fwd(x, w, bias, torch.float32, residual=res, row_scale=scale, rows_per_scale=Nq, tc=tc)
    assert nerr(y, res + ref_lin * scale.repeat_interleave(Nq)[:, None]) < tol
    # GELU + saved pre-activation
    u = torch.empty(M, N, dtype=dtype, device="cuda")
    h = ops.linear_fwd(x, w, bias, dtype, act=L.ACT_GELU, aux_out=u, tc=tc)
    assert nerr(u.float(), ref_lin) < tol
    assert nerr(h.float(), torch.nn.functional.gelu(ref_lin)) < tol
    # GELU backward in the dgrad epilogue: du = (dh @ W2) * gelu'(u)
    w2 = (randn(K, N, seed=34) * 0.05).to(dtype)  # fc2 weight [out=K, in=N]
    dyo = randn(M, K, seed=35).to(dtype)
    du = ops.linear_dgrad(dyo, w2, dtype, act=L.ACT_GELU_BWD, aux_in=u, tc=tc)
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(dyo.float() @ w2.float())
    assert nerr(du.float(), uu.grad) < tol
    # accumulate + row remap (PatchEmbed tokens written behind the cls slot)
    out = torch.zeros(B * (Nq + 1), N, device="cuda")
    ops.linear_fwd(x, w, bias, torch.float32, out=out, out_group=Nq, out_skip=1, tc=tc)
    o3 = out.view(B, Nq + 1, N)
    assert float(o3[:, 0].abs().max()) == 0.0
    assert nerr(o3[:, 1:].reshape(M, N), ref_lin) < tol
    acc = torch.ones(M, N, device="cuda")
    ops.linear_fwd(x, w, None, torch.float32, out=acc, accumulate=True, tc=tc)
    assert nerr(acc, ref_lin - bias + 1) < tol


def test_colsum_cast():
    from pmv_b200 import ops
    x = randn(1234, 384, seed=40)
    scale = torch.tensor([1.0, 0.0], device="cuda")
    s, c = ops.colsum_cast(x, torch.bfloat16, row_scale=scale, rows_per_scale=617)
    ref = x * scale.repeat_interleave(617)[:, None]
    assert nerr(s, ref.sum(0)) < 1e-5
    assert nerr(c.float(), ref) < 1e-2


ATTN_CASES = [  # B, heads, q_shape, k_shape
    (2, 2, (2, 4, 4), (2, 2, 2)), (1, 1, (8, 14, 14), (8, 7, 7)), (2, 2, (2, 3, 2), (2, 6, 4)), (1, 2, (2, 7, 5), (2, 4, 3)),
    (1, 2, (8, 7, 7), (8, 14, 14)),
]


def _attn_reference(q, k, v, q_shape, k_shape, rh, rw, rt, scale):
    from oracle import mvit_oracle as orc
    attn = (q * scale) @ k.transpose(-2, -1)
    attn = orc.add_rel_pos_bias(attn, q, True, q_shape, k_shape, rh, rw, rt).softmax(dim=-1)
    o = attn @ v
    o = torch.cat([o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]], dim=2)
    B, nh, Nq, C = q.shape
    return o.transpose(1, 2).reshape(B, Nq, nh * C)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,heads,q_shape,k_shape", ATTN_CASES)
def test_relpos_attention_fwd_bwd(dtype, B, heads, q_shape, k_shape):
    """rel-pos augmentation + attention (CUDA-core kernels) against the oracle's materialised-score formulation."""
    from pmv_b200 import ops
    tol = TOL[dtype]
    Nq, Nk = 1 + math.prod(q_shape), 1 + math.prod(k_shape)
    q = randn(B, heads, Nq, 96, seed=50).to(dtype)
    k = randn(B, heads, Nk, 96, seed=51).to(dtype)
    v = randn(B, heads, Nk, 96, seed=52).to(dtype)
    rh = randn(2 * max(q_shape[1], k_shape[1]) - 1, 96, seed=53) * 0.05
    rw = randn(2 * max(q_shape[2], k_shape[2]) - 1, 96, seed=54) * 0.05
    rt = randn(2 * max(q_shape[0], k_shape[0]) - 1, 96, seed=55) * 0.05
    scale = 96 ** -0.5
    ld = ops.aug_ld(k_shape)
    q_aug = torch.zeros(B * heads, Nq, ld, dtype=dtype, device="cuda")
    k_aug = torch.zeros(B * heads, Nk, ld, dtype=dtype, device="cuda")
    q_aug[..., :96] = q.reshape(B * heads, Nq, 96)
    k_aug[..., :96] = k.reshape(B * heads, Nk, 96)
    ops.relpos_augment_q(q_aug, q_shape, k_shape, rh, rw, rt, 1.0 / scale)
    ops.relpos_augment_k(k_aug, k_shape)
    vv = v.reshape(B * heads, Nk, 96).contiguous()
    out, out_pre, lse = ops.attention_fwd(q_aug, k_aug, vv, B, heads, ld, scale, residual=True, tc=0)
    leaves = [t.float().clone().requires_grad_(True) for t in (q, k, v, rh, rw, rt)]
    ref = _attn_reference(*leaves[:3], q_shape, k_shape, *leaves[3:], scale)
    assert nerr(out.float(), ref.detach()) < tol
    dout = randn(*ref.shape, seed=56).to(dtype)
    ref.backward(dout.float())
    dq_aug, dk, dv = ops.attention_bwd(q_aug, k_aug, vv, out_pre, dout, lse, B, heads, ld, scale, residual=True)
    drh, drw, drt = ops.relpos_augment_q_bwd(dq_aug, q_aug, q_shape, k_shape, rh, rw, rt, 1.0 / scale)
    gtol = tol if dtype == torch.float32 else 3e-2  # bf16 gradients: one more bf16 rounding per operand
    assert nerr(dq_aug[..., :96].float().reshape(B, heads, Nq, 96), leaves[0].grad) < gtol
    assert nerr(dk.float().reshape(B, heads, Nk, 96), leaves[1].grad) < gtol
    assert nerr(dv.float().reshape(B, heads, Nk, 96), leaves[2].grad) < gtol
    assert nerr(drh, leaves[3].grad) < gtol
    assert nerr(drw, leaves[4].grad) < gtol
    assert nerr(drt, leaves[5].grad) < gtol


def test_patch_embed_im2col_gemm():
    from pmv_b200 import ops
    clip = randn(2, 3, 4, 32, 24, seed=60)
    w = randn(96, 3, 3, 7, 7, seed=61) * 0.05
    b = randn(96, seed=62) * 0.1
    ref = torch.nn.functional.conv3d(clip, w, b, stride=(2, 4, 4), padding=(1, 3, 3))
    ref = ref.flatten(2).transpose(1, 2)
    for dtype in DTYPES:
        col, thw, K = ops.patch_im2col(clip, (3, 7, 7), (2, 4, 4), (1, 3, 3), dtype)
        wp = torch.zeros(96, col.shape[1], dtype=dtype, device="cuda")
        wp[:, :K] = w.reshape(96, K).to(dtype)
        L_ = thw[0] * thw[1] * thw[2]
        out = torch.zeros(2 * (L_ + 1), 96, device="cuda")
        ops.linear_fwd(col, wp, b, torch.float32, out=out, out_group=L_, out_skip=1)
        assert nerr(out.view(2, L_ + 1, 96)[:, 1:], ref) < TOL[dtype]


TC_ATTN_CASES = [  # B, heads, q_shape, k_shape — incl. ragged tails and both kd = 128 / 160
    (2, 2, (2, 4, 4), (2, 2, 2)), (1, 2, (8, 14, 14), (8, 7, 7)), (1, 2, (8, 7, 7), (8, 14, 14)),
    (1, 1, (8, 28, 28), (8, 14, 14)), (2, 1, (8, 56, 56), (8, 7, 7)), (1, 2, (2, 7, 5), (2, 4, 3)),
]


@pytest.mark.parametrize("B,heads,q_shape,k_shape", TC_ATTN_CASES)
def test_attention_tcgen05_fwd(B, heads, q_shape, k_shape):
    """tcgen05/TMEM/TMA attention forward (bf16) against the oracle's materialised-score formulation and
    against the CUDA-core kernel (same inputs, same Q'/K' buffers)."""
    from pmv_b200 import ops
    dtype = torch.bfloat16
    Nq, Nk = 1 + math.prod(q_shape), 1 + math.prod(k_shape)
    q = randn(B, heads, Nq, 96, seed=70).to(dtype)
    k = randn(B, heads, Nk, 96, seed=71).to(dtype)
    v = randn(B, heads, Nk, 96, seed=72).to(dtype)
    rh = randn(2 * max(q_shape[1], k_shape[1]) - 1, 96, seed=73) * 0.05
    rw = randn(2 * max(q_shape[2], k_shape[2]) - 1, 96, seed=74) * 0.05
    rt = randn(2 * max(q_shape[0], k_shape[0]) - 1, 96, seed=75) * 0.05
    scale = 96 ** -0.5
    ld = ops.aug_ld(k_shape)
    q_aug = torch.zeros(B * heads, Nq, ld, dtype=dtype, device="cuda")
    k_aug = torch.zeros(B * heads, Nk, ld, dtype=dtype, device="cuda")
    q_aug[..., :96] = q.reshape(B * heads, Nq, 96)
    k_aug[..., :96] = k.reshape(B * heads, Nk, 96)
    ops.relpos_augment_q(q_aug, q_shape, k_shape, rh, rw, rt, 1.0 / scale)
    ops.relpos_augment_k(k_aug, k_shape)
    vv = v.reshape(B * heads, Nk, 96).contiguous()
    out_tc, pre_tc, lse_tc = ops.attention_fwd(q_aug, k_aug, vv, B, heads, ld, scale, residual=True, tc=1)
    out_cc, pre_cc, lse_cc = ops.attention_fwd(q_aug, k_aug, vv, B, heads, ld, scale, residual=True, tc=0)
    torch.cuda.synchronize()
    ref = _attn_reference(q.float(), k.float(), v.float(), q_shape, k_shape, rh, rw, rt, scale)
    assert nerr(out_cc.float(), ref) < 1e-2
    assert nerr(out_tc.float(), ref) < 1e-2
    assert nerr(lse_tc, lse_cc) < 1e-2
    assert nerr(pre_tc.float(), pre_cc.float()) < 1e-2
    # backward: tensor-core kernels against the CUDA-core kernels on the same saved tensors
    dout = randn(*out_tc.shape, seed=76).to(dtype)
    dq_t, dk_t, dv_t = ops.attention_bwd(q_aug, k_aug, vv, pre_cc, dout, lse_cc, B, heads, ld, scale, residual=True, tc=1)
    dq_c, dk_c, dv_c = ops.attention_bwd(q_aug, k_aug, vv, pre_cc, dout, lse_cc, B, heads, ld, scale, residual=True, tc=0)
    torch.cuda.synchronize()
    assert nerr(dq_t.float(), dq_c.float()) < 2e-2
    assert nerr(dk_t.float(), dk_c.float()) < 2e-2
    assert nerr(dv_t.float(), dv_c.float()) < 2e-2
    # no residual, no lse
    out2, _, _ = ops.attention_fwd(q_aug, k_aug, vv, B, heads, ld, scale, residual=False, want_lse=False, tc=1)
    out3, _, _ = ops.attention_fwd(q_aug, k_aug, vv, B, heads, ld, scale, residual=False, want_lse=False, tc=0)
    assert nerr(out2.float(), out3.float()) < 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
def test_pool_ln_qkv_fused_launch(dtype):
    """q / k / v pooled by ONE launch each way (different strides per job) against the oracle."""
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    B, heads, thw, sq, skv = 2, 2, (3, 12, 10), 2, 4
    T, H, W = thw
    N = 1 + T * H * W
    qkv = randn(B, N, 3, heads, 96, seed=80).to(dtype)
    ws = [randn(96, 1, 3, 3, 3, seed=81 + i) * 0.2 for i in range(3)]
    gs = [randn(96, seed=84 + i) * 0.1 + 1 for i in range(3)]
    bs = [randn(96, seed=87 + i) * 0.1 for i in range(3)]
    strides = [sq, skv, skv]
    Ls = [1 + T * ops.pooled_hw(H, s) * ops.pooled_hw(W, s) for s in strides]
    lds = [128, 128, 96]
    outs = [torch.zeros(B, heads, Ls[i], lds[i], dtype=dtype, device="cuda") for i in range(3)]
    xhats = [torch.empty(B, heads, Ls[i], 96, dtype=dtype, device="cuda") for i in range(3)]
    rstds = [torch.empty(B, heads, Ls[i], device="cuda") for i in range(3)]
    ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xhats[i], rstds[i]) for i in range(3)])
    douts = [torch.zeros_like(o) for o in outs]
    grads = torch.zeros(3, 96 * 27 + 192, device="cuda")
    dqkv = torch.full_like(qkv, float("nan"))
    refs = []
    for i in range(3):
        xin = qkv[:, :, i].permute(0, 2, 1, 3).float().clone().requires_grad_(True)
        wr, gr, br = (t.clone().requires_grad_(True) for t in (ws[i], gs[i], bs[i]))
        ref, _ = orc.conv_pool_tokens(xin, thw, wr, (1, strides[i], strides[i]), True, gr, br)
        assert nerr(outs[i][..., :96].float(), ref.detach()) < TOL[dtype], i
        douts[i][..., :96] = randn(B, heads, Ls[i], 96, seed=90 + i).to(dtype)
        ref.backward(douts[i][..., :96].float())
        refs.append((xin, wr, gr, br))
    # once with the convolution recompute, once from the statistics the forward saved
    for saved in (False, True):
        grads.zero_()
        dqkv.fill_(float("nan"))
        extra = [(xhats[i], rstds[i]) if saved else () for i in range(3)]
        ops.pool_ln_qkv_bwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], douts[i], grads[i]) + extra[i] for i in range(3)], dqkv)
        for i, (xin, wr, gr, br) in enumerate(refs):
            assert nerr(dqkv[:, :, i].permute(0, 2, 1, 3).float(), xin.grad) < TOL[dtype], (i, saved)
            assert nerr(grads[i, :2592].view(96, 27), wr.grad.view(96, 27)) < TOL[dtype], (i, saved)
            assert nerr(grads[i, 2592:2688], gr.grad) < TOL[dtype], (i, saved)
            assert nerr(grads[i, 2688:], br.grad) < TOL[dtype], (i, saved)


@pytest.mark.parametrize("B,N,C,ncls,target_kind,p", [(8, 393, 768, 400, "hard", 0.5), (3, 17, 768, 400, "soft", 0.0),
                                                      (5, 2, 96, 10, "soft", 0.25), (2, 1, 100, 7, "hard", 0.0)])
def test_head_loss_fused(B, N, C, ncls, target_kind, p):
    """Row f2: final LN + cls select + head + (soft-target) cross entropy against the oracle, forward and backward."""
    import torch.nn.functional as F
    from oracle import mvit_oracle as orc
    from pmv_b200 import functional as Fn
    tok = randn(B, N, C, seed=31) * 1.5 + 0.2
    prm = {"norm.weight": randn(C, seed=32) * 0.1 + 1, "norm.bias": randn(C, seed=33) * 0.1,
           "head.projection.weight": randn(ncls, C, seed=34) * 0.05, "head.projection.bias": randn(ncls, seed=35) * 0.1}
    g = torch.Generator(device="cpu").manual_seed(36)
    if target_kind == "hard":
        target = torch.randint(0, ncls, (B,), generator=g).cuda()
    else:  # mixup-style rows: two classes share the mass, plus label smoothing
        target = torch.full((B, ncls), 0.1 / ncls)
        for b in range(B):
            i, j = torch.randint(0, ncls, (2,), generator=g).tolist()
            lam = float(torch.rand((), generator=g))
            target[b, i] += 0.9 * lam
            target[b, j] += 0.9 * (1 - lam)
        target = target.cuda()
    keep = (torch.rand(B, C, generator=g) >= p).to(torch.uint8).cuda() if p > 0 else None
    ours = [t.clone().requires_grad_(True) for t in (tok, prm["norm.weight"], prm["norm.bias"], prm["head.projection.weight"],
                                                     prm["head.projection.bias"])]
    loss, logits = Fn.head_loss(ours[0], ours[1], ours[2], ours[3], ours[4], target, keep_mask=keep, dropout_p=p)
    ref_in = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    tok_r = tok.clone().requires_grad_(True)
    loss_r, logits_r = orc.head_loss(tok_r, ref_in, target, keep, p)
    assert nerr(logits, logits_r.detach()) < 1e-5
    loss_v, loss_rv = float(loss.detach()), float(loss_r.detach())
    assert abs(loss_v - loss_rv) < 1e-5 * max(1.0, abs(loss_rv))
    if target_kind == "hard":  # second anchor: nn.CrossEntropyLoss (losses.py:66)
        assert abs(loss_v - float(F.cross_entropy(logits_r.detach(), target))) < 1e-5 * max(1.0, abs(loss_rv))
    (loss * 1.7).backward()
    (loss_r * 1.7).backward()
    refs = [tok_r, ref_in["norm.weight"], ref_in["norm.bias"], ref_in["head.projection.weight"], ref_in["head.projection.bias"]]
    for a, b, name in zip(ours, refs, ["dx", "dgamma", "dbeta", "dW", "dbias"]):
        assert nerr(a.grad, b.grad) < 1e-4, name
    assert float(ours[0].grad[:, 1:].abs().max()) == 0.0 if N > 1 else True


def test_head_eval_probabilities():
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    B, N, C, ncls = 4, 50, 768, 400
    tok = randn(B, N, C, seed=41)
    gm, bt = randn(C, seed=42) * 0.1 + 1, randn(C, seed=43) * 0.1
    w, bias = randn(ncls, C, seed=44) * 0.05, randn(ncls, seed=45) * 0.1
    r = ops.head_loss_fwd(tok, gm, bt, w, bias, want_probs=True)
    x = orc.layer_norm(tok, gm, bt)[:, 0]
    ref = torch.nn.functional.linear(x, w, bias)
    assert nerr(r["logits"], ref) < 1e-5
    assert nerr(r["probs"], ref.softmax(dim=1)) < 1e-5
    assert r["loss"] is None
