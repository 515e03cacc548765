"""Out-of-bounds WRITE detector for the C-ABI kernels (compute-sanitizer is closed on this GPU pool, see
profiles/r02_sanitizer.md).  Every tensor the thin wrappers of pmv_b200/ops.py allocate for a kernel (outputs, saved
statistics, workspaces) is carved out of a larger buffer whose margins are filled with a sentinel byte pattern; after the
kernels of a family have run (ragged sizes: token counts that are no multiple of any tile, odd grids, every pooling
stride) the margins must be untouched.  The inside of the tensors is what the parity tests check."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on each side
SENT = 0xA5


class guarded_allocations:
    """Routes torch.empty / empty_like / zeros inside pmv_b200.ops (and functional) through guarded buffers."""

    def __enter__(self):
        from pmv_b200 import functional, ops
        self.mods = [ops, functional]
        self.saved = [m.torch for m in self.mods]
        self.bufs = []
        real = torch
        outer = self

        class Proxy:
            def __getattr__(self, name):
                return getattr(real, name)

            @staticmethod
            def _carve(shape, dtype, device, zero):
                shape = tuple(int(s) for s in (shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list, real.Size)) else shape))
                n = 1
                for s in shape:
                    n *= s
                esz = real.empty((), dtype=dtype).element_size()
                nbytes = (n * esz + 255) // 256 * 256
                raw = real.full((GUARD + nbytes + GUARD,), SENT, dtype=real.uint8, device=device)
                outer.bufs.append((raw, n * esz))
                t = raw[GUARD:GUARD + n * esz].view(dtype).view(shape)
                if zero:
                    t.zero_()
                return t

            def empty(self, *shape, dtype=None, device=None, **kw):
                return self._carve(shape, dtype or real.float32, device or "cpu", False)

            def zeros(self, *shape, dtype=None, device=None, **kw):
                return self._carve(shape, dtype or real.float32, device or "cpu", True)

            def empty_like(self, t, **kw):
                return self._carve((tuple(t.shape),), kw.get("dtype", t.dtype), t.device, False)

        for m in self.mods:
            m.torch = Proxy()
        return self

    def __exit__(self, *exc):
        for m, s in zip(self.mods, self.saved):
            m.torch = s

    def check(self, what):
        torch.cuda.synchronize()
        for raw, used in self.bufs:
            if raw.device.type != "cuda":
                continue
            lo, hi = raw[:GUARD], raw[GUARD + used:]
            assert bool((lo == SENT).all()), f"{what}: write below a {used}-byte tensor"
            assert bool((hi == SENT).all()), f"{what}: write beyond a {used}-byte tensor"
        n = len(self.bufs)
        self.bufs.clear()
        return n


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("case", [(2, 1, (3, 9, 11), 1, 8), (1, 2, (2, 13, 7), 2, 4), (2, 2, (5, 7, 5), 2, 1), (1, 4, (8, 14, 14), 1, 2),
                                  (1, 1, (1, 5, 5), 1, 1)])
def test_pooling_kernels_write_inside_their_tensors(case, dtype):
    from pmv_b200 import ops
    B, heads, thw, sq, skv = case
    T, H, W = thw
    N = 1 + T * H * W
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3, heads, 96, device="cuda").to(dtype)
    ws = [torch.randn(96, 1, 3, 3, 3, device="cuda") * 0.2 for _ in range(3)]
    gs = [torch.rand(96, device="cuda") + 0.5 for _ in range(3)]
    bs = [torch.randn(96, device="cuda") * 0.1 for _ in range(3)]
    strides = [sq, skv, skv]
    Ls = [1 + T * ops.pooled_hw(H, s) * ops.pooled_hw(W, s) for s in strides]
    with guarded_allocations() as g:
        P = ops.torch
        outs = [P.empty(B, heads, Ls[i], [128, 128, 96][i], dtype=dtype, device="cuda") for i in range(3)]
        xh = [P.empty(B, heads, Ls[i], 96, dtype=dtype, device="cuda") for i in range(3)]
        rs = [P.empty(B, heads, Ls[i], dtype=torch.float32, device="cuda") for i in range(3)]
        ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xh[i], rs[i]) for i in range(3)])
        assert g.check("pool forward") >= 9
        douts = [torch.randn_like(o) for o in outs]
        grads = P.zeros(3, 96 * 27 + 192, dtype=torch.float32, device="cuda")
        dqkv = P.empty_like(qkv)
        ops.pool_ln_qkv_bwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], douts[i], grads[i], xh[i], rs[i]) for i in range(3)], dqkv)
        assert g.check("pool backward") >= 3
        assert torch.isfinite(dqkv.float()).all() and torch.isfinite(grads).all()


@pytest.mark.parametrize("shape", [(1, 2, 131, 77, 128), (2, 1, 393, 393, 128), (1, 4, 50, 1569, 160), (1, 1, 1, 1, 128)])
def test_attention_kernels_write_inside_their_tensors(shape):
    from pmv_b200 import ops
    B, heads, Nq, Nk, ld = shape
    dt = torch.bfloat16
    torch.manual_seed(1)
    q = (torch.randn(B * heads, Nq, ld, device="cuda") * 0.5).to(dt)
    k = (torch.randn(B * heads, Nk, ld, device="cuda") * 0.5).to(dt)
    v = torch.randn(B * heads, Nk, 96, device="cuda").to(dt)
    with guarded_allocations() as g:
        out, out_pre, lse = ops.attention_fwd(q, k, v, B, heads, ld, 96 ** -0.5, residual=True, want_lse=True, tc=1)
        assert g.check("attention forward") >= 3
        dout = torch.randn_like(out)
        dq, dk, dv = ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, 96 ** -0.5, residual=True, tc=1, fp32_dkv=True)
        assert g.check("attention backward") >= 2
        assert torch.isfinite(dq.float()).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("mnk", [(393, 288, 96), (1569, 96, 384), (131, 400, 768), (7, 96, 96), (3144, 1536, 384)])
def test_gemm_layernorm_misc_write_inside_their_tensors(mnk, dtype):
    from pmv_b200 import _lib as L, ops
    M, N, K = mnk
    torch.manual_seed(2)
    x = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(dtype)
    b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    with guarded_allocations() as g:
        P = ops.torch
        y = ops.linear_fwd(x, w, b, dtype)
        u = P.empty(M, N, dtype=dtype, device="cuda")
        h = ops.linear_fwd(x, w, b, dtype, act=L.ACT_GELU, aux_out=u)
        yr = ops.linear_fwd(x, w, b, torch.float32, residual=res)
        dy = torch.randn(M, N, device="cuda").to(dtype)
        dx = ops.linear_dgrad(dy, w, dtype)
        du = ops.linear_dgrad(x, w.t().contiguous(), dtype, act=L.ACT_GELU_BWD, aux_in=u) if K == N else None
        dw = ops.linear_wgrad(dy, x)
        s, c = ops.colsum_cast(res, dtype)
        xf = torch.randn(M, N, device="cuda") if N in (96, 384, 768) else None
        if xf is not None:
            yl, mean, rstd = ops.layernorm_fwd(xf, torch.ones(N, device="cuda"), torch.zeros(N, device="cuda"), dtype)
            ops.layernorm_bwd(yl, xf, torch.ones(N, device="cuda"), mean, rstd)
        assert g.check("gemm / layernorm / colsum") >= 6
    assert torch.isfinite(y.float()).all() and torch.isfinite(dw).all() and torch.isfinite(dx.float()).all()


def test_skip_maxpool_and_relpos_write_inside_their_tensors():
    from pmv_b200 import ops
    torch.manual_seed(3)
    with guarded_allocations() as g:
        x = torch.randn(2, 1 + 3 * 7 * 5, 192, device="cuda")
        y, win = ops.maxpool_skip_fwd(x, (3, 7, 5), want_winner=True)
        ops.maxpool_skip_bwd(win, torch.randn_like(y), (3, 7, 5))
        assert g.check("skip max-pool") >= 3
        q_shape, k_shape = (3, 7, 5), (3, 4, 3)
        ld = ops.aug_ld(k_shape)
        P = ops.torch
        q_aug = P.empty(4, 1 + 105, ld, dtype=torch.bfloat16, device="cuda")
        q_aug.copy_(torch.randn(4, 106, ld, device="cuda"))
        k_aug = P.zeros(4, 1 + 36, ld, dtype=torch.bfloat16, device="cuda")
        rh, rw, rt = (torch.randn(2 * 7 - 1, 96, device="cuda") * .02, torch.randn(2 * 5 - 1, 96, device="cuda") * .02,
                      torch.randn(5, 96, device="cuda") * .02)
        ops.relpos_augment_q(q_aug, q_shape, k_shape, rh, rw, rt, 96 ** 0.5)
        ops.relpos_augment_k(k_aug, k_shape)
        dq = P.empty_like(q_aug)
        dq.copy_(torch.randn(4, 106, ld, device="cuda"))
        ops.relpos_augment_q_bwd(dq, q_aug, q_shape, k_shape, rh, rw, rt, 96 ** 0.5)
        assert g.check("rel-pos augmentation") >= 3
