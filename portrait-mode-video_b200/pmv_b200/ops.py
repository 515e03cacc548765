"""Thin tensor-level wrappers over the C ABI (no autograd here — see functional.py).

Every function allocates its outputs with torch (device memory + current stream are the only things
torch provides) and launches the hand-written kernels through ctypes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

LN_EPS = 1e-6


def _tc_default(dtype) -> int:
    return 1 if dtype == torch.bfloat16 else 0


# ----------------------------------------------------------------------------- launch accounting
LAUNCHES = 0       # kernels of libpmv_b200.so launched so far (bench.py reports the delta as gpu_launches)
_RECORDER = None   # optional list of (kind, meta, start_event, end_event) for per-kernel CUDA-event timing


class record_kernels:
    """Context manager: time every C-ABI launch with CUDA events on the launching stream (bench.py roofline)."""

    def __enter__(self):
        global _RECORDER
        _RECORDER = []
        return _RECORDER

    def __exit__(self, *exc):
        global _RECORDER
        _RECORDER = None


def _ws(nbytes: int, device) -> torch.Tensor:
    """fp32 scratch for the block-partial reductions (stream-ordered: torch's caching allocator recycles it)."""
    return torch.empty(max(int(nbytes) // 4, 1), dtype=torch.float32, device=device)


def _run(name: str, nkernels: int, meta, *args):
    global LAUNCHES
    LAUNCHES += nkernels
    fn = getattr(L.lib(), name)
    if _RECORDER is None:
        L.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(fn(*args), name)
    e1.record()
    _RECORDER.append((name, meta, e0, e1))


# ----------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x: torch.Tensor, gamma, beta, out_dtype, save_stats=True, eps=LN_EPS):
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cdim = x.shape[-1]
    rows = x.numel() // Cdim
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _run("pmv_layernorm_fwd", 1, dict(bytes=rows * Cdim * (4 + y.element_size())), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), L.dt(out_dtype), L.ptr(mean),
                                      L.ptr(rstd), rows, Cdim, eps, L.stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_accum: Optional[torch.Tensor] = None, dx_base: Optional[torch.Tensor] = None):
    """Returns (dx fp32, dgamma, dbeta).  ``dx_base`` (the gradient arriving over the residual connection) is added
    into a fresh dx; ``dx_accum`` adds the input gradient into that tensor in place."""
    Cdim = x.shape[-1]
    rows = x.numel() // Cdim
    dy = dy.contiguous()
    base = dx_accum if dx_accum is not None else dx_base
    if base is not None:
        assert base.dtype == torch.float32 and base.is_contiguous() and base.numel() == x.numel()
    dx = dx_accum if dx_accum is not None else torch.empty_like(x)
    dgb = torch.empty(2, Cdim, dtype=torch.float32, device=x.device)  # overwritten by the kernel
    ws = _ws(L.lib().pmv_layernorm_bwd_workspace_bytes(rows, Cdim), x.device)
    _run("pmv_layernorm_bwd", 2, dict(bytes=rows * Cdim * (8 + dy.element_size() + (4 if base is not None else 0))), L.ptr(dy), L.dt(dy),
         L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(dx), L.ptr(base), L.ptr(dgb), L.ptr(ws), rows, Cdim, L.stream())
    return dx, dgb[0], dgb[1]


# ----------------------------------------------------------------------------- GEMM
def gemm(layout: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, out: torch.Tensor, *, bias=None, act=L.ACT_NONE,
         aux_in=None, aux_out=None, row_scale=None, rows_per_scale=1, residual=None, accumulate=False,
         out_group=0, out_skip=0, tc: Optional[bool] = None, split_k: int = 1, ldo: Optional[int] = None):
    """C = op(A) op(B) with the fused epilogue of pmv_gemm.  A/B/out are 2-D row-major (last stride 1)."""
    assert A.stride(-1) == 1 and B.stride(-1) == 1 and out.stride(-1) == 1
    assert A.dtype == B.dtype
    lda, ldb = A.stride(0), B.stride(0)
    ldo = out.stride(0) if ldo is None else ldo
    use_tc = _tc_default(A.dtype) if tc is None else int(tc)
    need_epi = any(v is not None for v in (bias, aux_in, aux_out, row_scale, residual)) or act or accumulate or out_group
    epi = None
    if need_epi:
        aux = aux_in if aux_in is not None else aux_out
        epi = L.Epilogue(L.ptr(bias), act, L.ptr(aux_in), L.ptr(aux_out), aux.stride(0) if aux is not None else 0,
                         L.ptr(row_scale), rows_per_scale, L.ptr(residual),
                         residual.stride(0) if residual is not None else 0, int(accumulate), out_group, out_skip)
    _run("pmv_gemm", 1, dict(flops=2 * M * N * K, layout=layout, tc=use_tc, shape=(M, N, K)), layout, L.ptr(A), lda, L.ptr(B), ldb, L.ptr(out), ldo, M, N, K, L.dt(A), L.dt(out),
                             C.byref(epi) if epi is not None else None, use_tc, split_k, L.stream())
    return out


def linear_fwd(x2d, weight, bias, out_dtype, **kw):
    """y[M,N] = x[M,K] @ weight[N,K]^T (+ epilogue)."""
    M, K = x2d.shape
    N = weight.shape[0]
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=x2d.device)
    return gemm(L.GEMM_TN, x2d, weight, M, N, K, out, bias=bias, **kw)


def linear_dgrad(dy2d, weight, out_dtype, **kw):
    """dx[M,K] = dy[M,N] @ weight[N,K]."""
    M, N = dy2d.shape
    K = weight.shape[1]
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty(M, K, dtype=out_dtype, device=dy2d.device)
    return gemm(L.GEMM_NN, dy2d, weight, M, K, N, out, **kw)


def _pick_split(rows: int, n1: int, n2: int) -> int:
    tiles = ((n1 + 127) // 128) * ((n2 + 95) // 96)
    want = max(1, (148 * 2) // max(tiles, 1))
    return max(1, min(want, rows // 512 if rows >= 512 else 1))


class ZeroArena:
    """Pre-zeroed fp32 memory for the split-K weight gradients of one step of ONE model.  Owned by that model's
    GradAllReducer / FusedAdamW and attached to its parameters (``p._pmv_arena``, ``attach_arena``): the autograd functions
    take the arena of the weight they differentiate, so two models in one process (teacher / student, a second reducer)
    never share memory.  ``reset()`` (from the owner's ``zero_grad``, i.e. when last step's gradients are dead) clears the
    whole arena with ONE fill and rewinds it; ``take(n)`` hands out the next n zeroed floats, or None once the arena is used
    up — it never hands out memory that was not cleared since the last reset, so forgetting reset() only costs the fallback
    (a torch.zeros per gradient: ~80 fill launches per MViTv2-S step).  ``exhaust()`` (``zero_grad(set_to_none=False)``:
    the gradients stay alive inside the arena) makes take() return None until the next reset.  The arena sizes itself from
    the demand of the previous step; once a reset()/take() ran under CUDA-graph capture the buffer is frozen (a captured
    graph keeps writing to its addresses, so it is never reallocated afterwards)."""

    def __init__(self):
        self.buf = None
        self.used = 0
        self.demand = 0
        self.frozen = False

    @staticmethod
    def _capturing(device) -> bool:
        return device.type == "cuda" and torch.cuda.is_current_stream_capturing()

    def reset(self, device):
        want = self.demand
        if self._capturing(device):
            self.frozen = True
        if want > 0 and not self.frozen and (self.buf is None or self.buf.numel() < want or self.buf.device != device):
            self.buf = torch.empty(int(want * 1.05) + 1024, dtype=torch.float32, device=device)
        if self.buf is not None:
            self.buf.zero_()
        self.used = 0
        self.demand = 0

    def exhaust(self):
        """Nothing more is handed out until the next reset() (the slices given out so far stay valid)."""
        self.used = -1

    def take(self, n, device):
        n_al = (n + 63) // 64 * 64  # 256-byte granules: TMA / vector alignment of every gradient
        self.demand += n_al
        if self.buf is None or self.buf.device != device or self.used < 0 or self.used + n_al > self.buf.numel():
            return None
        if self._capturing(device):
            self.frozen = True
        out = self.buf[self.used:self.used + n]
        self.used += n_al
        return out


def attach_arena(params, arena: Optional["ZeroArena"] = None) -> "ZeroArena":
    """Binds one arena to the parameters of a model (re-uses the one already attached to the first parameter, so a
    reducer and an optimizer over the same model share it)."""
    params = list(params)
    if arena is None:
        arena = next((getattr(p, "_pmv_arena") for p in params if getattr(p, "_pmv_arena", None) is not None), None) or ZeroArena()
    for p in params:
        p._pmv_arena = arena
    return arena


def grad_home(param) -> Optional[torch.Tensor]:
    """The place a weight gradient should be written to directly: the parameter's slot in its all-reduce bucket
    (GradAllReducer sets ``_pmv_grad_home``; zeroed by its zero_grad).  Only while the parameter holds no gradient yet —
    with gradient accumulation autograd adds the new gradient to the old one, which must then be a different tensor."""
    home = getattr(param, "_pmv_grad_home", None)
    if home is None or param.grad is not None:
        return None
    return home


def linear_wgrad(dy2d, x2d, tc=None, arena: Optional[ZeroArena] = None, home: Optional[torch.Tensor] = None):
    """dW[N,K] = dy[M,N]^T @ x[M,K] in fp32 (split over the token rows, fp32 atomics into zeroed memory: the parameter's
    slot in its all-reduce bucket (``home``), else a slice of the owning model's ``arena``, else a fresh torch.zeros).
    (Launching it on the side stream beside the dgrad GEMM of the same layer was measured: 14.11 ms per step against
    14.21 / 13.96 / 14.22 without — two persistent 148-CTA GEMMs do not share SMs, nothing to gain.)"""
    M, N = dy2d.shape
    K = x2d.shape[1]
    split = _pick_split(M, N, K)
    if home is not None and tuple(home.shape) == (N, K) and home.is_contiguous():
        out = home
    elif split > 1:
        flat = arena.take(N * K, dy2d.device) if arena is not None else None
        out = flat.view(N, K) if flat is not None else torch.zeros(N, K, dtype=torch.float32, device=dy2d.device)
    else:
        out = torch.empty(N, K, dtype=torch.float32, device=dy2d.device)
    return gemm(L.GEMM_NT_REDUCE_M, dy2d, x2d, M, N, K, out, split_k=split, tc=tc)


def colsum_cast(x2d, cast_dtype=None, row_scale=None, rows_per_scale=1, want_sum=True):
    rows, cols = x2d.shape
    s = torch.empty(cols, dtype=torch.float32, device=x2d.device) if want_sum else None  # overwritten by the kernel
    c = torch.empty(rows, cols, dtype=cast_dtype, device=x2d.device) if cast_dtype is not None else None
    ws = _ws(L.lib().pmv_colsum_workspace_bytes(rows, cols), x2d.device) if want_sum else None
    _run("pmv_colsum_cast", 2 if want_sum else 1, dict(bytes=rows * cols * (x2d.element_size() + (c.element_size() if c is not None else 0))), L.ptr(x2d), L.dt(x2d), x2d.stride(0), rows, cols, L.ptr(row_scale), rows_per_scale,
                                    L.ptr(s), L.ptr(ws), L.ptr(c), L.dt(cast_dtype) if cast_dtype is not None else 0,
                                    c.stride(0) if c is not None else 0, L.stream())
    return s, c


class _SideStream:
    """One side stream per device for work that only has to finish before the calling autograd node returns: the
    bias-gradient column sums, which are memory-bound 256-thread CTAs without shared memory and fit beside the persistent
    GEMM CTAs of the same backward node.  fork(): the side stream waits for everything issued so far on the current
    stream; join(): the current stream waits for the side stream.  Inside a captured CUDA graph the side work becomes a
    parallel branch.  PMV_SIDE_STREAM=0 disables it (everything on the current stream)."""

    _by_device = {}

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.fork_ev = torch.cuda.Event()
        self.join_ev = torch.cuda.Event()

    @classmethod
    def get(cls, device):
        import os
        if os.environ.get("PMV_SIDE_STREAM", "1") == "0" or device.type != "cuda":
            return None
        key = device.index if device.index is not None else torch.cuda.current_device()
        st = cls._by_device.get(key)
        if st is None:
            st = cls._by_device[key] = cls(device)
        return st

    def fork(self):
        self.fork_ev.record(torch.cuda.current_stream())
        self.stream.wait_event(self.fork_ev)

    def join(self):
        self.join_ev.record(self.stream)
        torch.cuda.current_stream().wait_event(self.join_ev)


def colsum_beside(x2d):
    """Column sums of x2d (a bias gradient) launched on the side stream: returns (sum, join) — call join() before the sum
    is handed to anything else.  The output and the workspace are allocated on the CURRENT stream (the caching allocator
    keys blocks by the allocating stream; the join orders every later use after the side kernels)."""
    side = _SideStream.get(x2d.device)
    if side is None or _RECORDER is not None:  # per-kernel timing (bench.py) measures on the current stream only
        return colsum_cast(x2d, None)[0], (lambda: None)
    global LAUNCHES
    rows, cols = x2d.shape
    s = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    ws = _ws(L.lib().pmv_colsum_workspace_bytes(rows, cols), x2d.device)
    side.fork()
    LAUNCHES += 2
    L.check(L.lib().pmv_colsum_cast(L.ptr(x2d), L.dt(x2d), x2d.stride(0), rows, cols, None, 1, L.ptr(s), L.ptr(ws), None, 0, 0,
                                    C.c_void_p(side.stream.cuda_stream)), "pmv_colsum_cast")

    def join(keep=(ws, x2d)):
        # `ws` must stay allocated until the current stream has been ordered after the side kernels: freed earlier, the
        # caching allocator (which only knows the allocating stream) hands the block to the next allocation of the current
        # stream while the side stream still writes its partial sums into it (seen as a corrupted fc1 weight gradient)
        side.join()

    return s, join


# ----------------------------------------------------------------------------- pooling
def pooled_hw(n: int, s: int) -> int:
    return (n - 1) // s + 1


def pool_ln_fwd(qkv: torch.Tensor, which: int, heads: int, thw: Sequence[int], stride_hw: int, w, gamma, beta,
                out: torch.Tensor, eps=LN_EPS):
    """qkv: [B, N, 3, heads, 96] contiguous; writes out [B, heads, 1+L', ld] columns [0,96)."""
    B, N = qkv.shape[0], qkv.shape[1]
    T, H, W = thw
    view = qkv[:, :, which]
    _run("pmv_pool_ln_fwd", 1, dict(bytes=(B * N * heads * 96 + out.shape[0] * out.shape[1] * out.shape[2] * 96) * qkv.element_size()), view.data_ptr(), qkv.stride(0), qkv.stride(1), qkv.stride(3), L.ptr(w), L.ptr(gamma),
                                    L.ptr(beta), L.ptr(out), out.stride(2), B, heads, T, H, W, stride_hw, eps, L.dt(qkv),
                                    L.stream())
    return out


def pool_ln_bwd(qkv, which, heads, thw, stride_hw, w, gamma, dout, dqkv, grads, eps=LN_EPS):
    """grads: fp32 [96*27 + 96 + 96] (Conv3d weight gradient, LayerNorm weight gradient, bias gradient), overwritten."""
    B = qkv.shape[0]
    T, H, W = thw
    ws = _ws(L.lib().pmv_pool_ln_bwd_workspace_bytes(B, heads, T, H, W, stride_hw), qkv.device)
    _run("pmv_pool_ln_bwd", 3, dict(bytes=(2 * B * qkv.shape[1] * heads * 96 + dout.shape[0] * dout.shape[1] * dout.shape[2] * 96) * qkv.element_size()),
         qkv[:, :, which].data_ptr(), qkv.stride(0), qkv.stride(1), qkv.stride(3), L.ptr(w), L.ptr(gamma), L.ptr(dout),
         dout.stride(2), dqkv[:, :, which].data_ptr(), L.ptr(grads), L.ptr(ws), B, heads, T, H, W, stride_hw, eps,
         L.dt(qkv), L.stream())


def pool_ln_qkv_fwd(qkv, heads, thw, jobs, eps=LN_EPS, onehot_jobs=()):
    """q / k / v pooled in one launch.  qkv: [B, N, 3, heads, 96]; jobs: list of (which, stride_hw, w, gamma, beta, out
    [, xhat, rstd]) with out [B, heads, 1+L', ld]; xhat [B, heads, 1+L', 96] / rstd [B, heads, 1+L'] (optional) receive the
    normalised pre-affine tokens and 1/sigma for the backward pass.  ``onehot_jobs``: indices of jobs whose columns
    [96, ld) are also filled with the one-hot key coordinates (K' of the rel-pos scheme; replaces relpos_augment_k)."""
    B, N = qkv.shape[0], qkv.shape[1]
    T, H, W = thw
    arr = (L.PoolJob * len(jobs))()
    nbytes = 0
    for i, job in enumerate(jobs):
        which, s, w, gamma, beta, out = job[:6]
        xhat, rstd = (job[6], job[7]) if len(job) > 6 else (None, None)
        arr[i] = L.PoolJob(L.ptr(w), L.ptr(gamma), L.ptr(beta), L.ptr(out), out.stride(2), None, 0, None, s, which,
                           L.ptr(xhat), L.ptr(rstd), 0, int(i in onehot_jobs))
        nbytes += (B * N * heads * 96 + out.shape[0] * out.shape[1] * out.shape[2] * 96) * qkv.element_size()
    _run("pmv_pool_ln_qkv_fwd", 1, dict(bytes=nbytes, shape=(B, heads, T, H, W, [j[1] for j in jobs])), L.ptr(qkv), qkv.stride(0),
         qkv.stride(1), qkv.stride(2), qkv.stride(3), arr, len(jobs), B, heads, T, H, W, eps, L.dt(qkv), L.stream())


def pool_ln_qkv_bwd(qkv, heads, thw, jobs, dqkv, eps=LN_EPS):
    """jobs: list of (which, stride_hw, w, gamma, dout, grads [, xhat, rstd]) — grads fp32 [96*27 + 192], overwritten.
    With xhat / rstd (saved by the forward) the LayerNorm backward skips the convolution recompute."""
    B, N = qkv.shape[0], qkv.shape[1]
    T, H, W = thw
    arr = (L.PoolJob * len(jobs))()
    strides = (C.c_int * len(jobs))(*[j[1] for j in jobs])
    nbytes = 0
    for i, job in enumerate(jobs):
        which, s, w, gamma, dout, grads = job[:6]
        xhat, rstd = (job[6], job[7]) if len(job) > 6 else (None, None)
        arr[i] = L.PoolJob(L.ptr(w), L.ptr(gamma), None, None, 0, L.ptr(dout), dout.stride(2), L.ptr(grads), s, which,
                           L.ptr(xhat), L.ptr(rstd), int(dout.dtype == torch.float32 and qkv.dtype != torch.float32))
        nbytes += (2 * B * N * heads * 96 + dout.shape[0] * dout.shape[1] * dout.shape[2] * 96) * qkv.element_size()
    ws = _ws(L.lib().pmv_pool_ln_qkv_bwd_workspace_bytes(B, heads, T, H, W, strides, len(jobs)), qkv.device)
    _run("pmv_pool_ln_qkv_bwd", 3, dict(bytes=nbytes, shape=(B, heads, T, H, W, [j[1] for j in jobs])), L.ptr(qkv), qkv.stride(0),
         qkv.stride(1), qkv.stride(2), qkv.stride(3), arr, len(jobs), L.ptr(dqkv), L.ptr(ws), B, heads, T, H, W, eps, L.dt(qkv),
         L.stream())


def maxpool_skip_fwd(x, thw, want_winner=False):
    """Returns y, or (y, win) with the uint8 winning-window-position map the backward gathers through."""
    B, N, Cdim = x.shape
    T, H, W = thw
    Lo = T * pooled_hw(H, 2) * pooled_hw(W, 2)
    y = torch.empty(B, 1 + Lo, Cdim, dtype=torch.float32, device=x.device)
    win = torch.empty(B, 1 + Lo, Cdim, dtype=torch.uint8, device=x.device) if want_winner else None
    _run("pmv_maxpool_skip_fwd", 1, dict(bytes=(x.numel() + y.numel()) * 4), L.ptr(x), L.ptr(y), L.ptr(win), B, T, H, W, Cdim, L.stream())
    return (y, win) if want_winner else y


def maxpool_skip_bwd(win, dy, thw):
    B, _, Cdim = dy.shape
    T, H, W = thw
    dx = torch.empty(B, 1 + T * H * W, Cdim, dtype=torch.float32, device=dy.device)
    _run("pmv_maxpool_skip_bwd", 1, dict(bytes=dx.numel() * 4 + dy.numel() * 5), L.ptr(win), L.ptr(dy.contiguous()), L.ptr(dx), B, T, H, W, Cdim,
         L.stream())
    return dx


# ----------------------------------------------------------------------------- rel-pos augmentation + attention
_IDX_CACHE = {}


def rel_index_table(q_n: int, k_n: int, device) -> torch.Tensor:
    """int32 [q_n*k_n] rows of the rel-pos table for every (query, key) coordinate pair on one axis.
    Same float32 arithmetic and truncation as the reference (attention.py:80-86,98)."""
    key = (q_n, k_n, str(device))
    t = _IDX_CACHE.get(key)
    if t is None:
        q_ratio = max(k_n / q_n, 1.0)
        k_ratio = max(q_n / k_n, 1.0)
        dist = torch.arange(q_n)[:, None] * q_ratio - torch.arange(k_n)[None, :] * k_ratio
        dist = dist + (k_n - 1) * k_ratio
        t = dist.long().to(torch.int32).reshape(-1).contiguous().to(device)
        _IDX_CACHE[key] = t
    return t


def aug_ld(k_shape) -> int:
    """Row width of Q'/K': 96 channels + the one-hot key-coordinate columns, padded to a multiple of 32."""
    rk = k_shape[0] + k_shape[1] + k_shape[2]
    return 96 + ((rk + 31) // 32) * 32


def relpos_augment_q(q_aug, q_shape, k_shape, rel_h, rel_w, rel_t, inv_scale, tc=None):
    BH, Nq, ld = q_aug.shape
    dev = q_aug.device
    ih, iw, it = (rel_index_table(q_shape[1], k_shape[1], dev), rel_index_table(q_shape[2], k_shape[2], dev),
                  rel_index_table(q_shape[0], k_shape[0], dev))
    ncat = rel_h.shape[0] + rel_w.shape[0] + rel_t.shape[0]
    ws = _ws(L.lib().pmv_relpos_fwd_workspace_bytes(BH, *q_shape, *k_shape), dev)
    use_tc = _tc_default(q_aug.dtype) if tc is None else int(tc)
    e = q_aug.element_size()
    _run("pmv_relpos_augment_q", 3, dict(bytes=BH * Nq * (ld + 2 * ncat) * e, flops=2 * BH * Nq * 96 * ncat), L.ptr(q_aug), ld,
         L.ptr(rel_h), L.ptr(rel_w), L.ptr(rel_t), L.ptr(ih), L.ptr(iw), L.ptr(it), L.ptr(ws), BH, *q_shape, *k_shape, inv_scale,
         L.dt(q_aug), use_tc, L.stream())


def relpos_augment_k(k_aug, k_shape):
    BH, Nk, ld = k_aug.shape
    _run("pmv_relpos_augment_k", 1, dict(bytes=BH * Nk * (ld - 96) * k_aug.element_size()), L.ptr(k_aug), ld, BH, *k_shape, L.dt(k_aug), L.stream())


def relpos_augment_q_bwd(dq_aug, q_aug, q_shape, k_shape, rel_h, rel_w, rel_t, inv_scale, tc=None):
    """In place: dq_aug[:, :, :96] += bias-path gradient.  Returns fp32 (d_rel_h, d_rel_w, d_rel_t)."""
    BH, Nq, ld = q_aug.shape
    dev = q_aug.device
    ih, iw, it = (rel_index_table(q_shape[1], k_shape[1], dev), rel_index_table(q_shape[2], k_shape[2], dev),
                  rel_index_table(q_shape[0], k_shape[0], dev))
    nh, nw, nt = rel_h.shape[0], rel_w.shape[0], rel_t.shape[0]
    ncat = nh + nw + nt
    d_rel = torch.empty(ncat, 96, dtype=torch.float32, device=dev)
    ws = _ws(L.lib().pmv_relpos_bwd_workspace_bytes(BH, *q_shape, *k_shape), dev)
    use_tc = _tc_default(q_aug.dtype) if tc is None else int(tc)
    e = q_aug.element_size()
    _run("pmv_relpos_augment_q_bwd", 5, dict(bytes=BH * Nq * (3 * ld + 3 * ncat) * e, flops=4 * BH * Nq * 96 * ncat), L.ptr(dq_aug),
         L.ptr(q_aug), ld, L.ptr(rel_h), L.ptr(rel_w), L.ptr(rel_t), L.ptr(ih), L.ptr(iw), L.ptr(it), L.ptr(d_rel), L.ptr(ws), BH,
         *q_shape, *k_shape, inv_scale, L.dt(q_aug), use_tc, L.stream())
    return d_rel[:nh], d_rel[nh:nh + nw], d_rel[nh + nw:]


def attention_fwd(q_aug, k_aug, v, B, heads, kd, scale, residual=True, want_lse=True, tc=None):
    """Returns (out, out_pre, lse): out_pre is the pre-residual output kept for backward (None in inference;
    aliases out when there is no residual pooling)."""
    BH, Nq, ld = q_aug.shape
    Nk = k_aug.shape[1]
    out = torch.empty(B, Nq, heads * 96, dtype=q_aug.dtype, device=q_aug.device)
    out_pre = torch.empty_like(out) if (want_lse and residual) else None
    lse = torch.empty(BH, Nq, dtype=torch.float32, device=q_aug.device) if want_lse else None
    use_tc = _tc_default(q_aug.dtype) if tc is None else int(tc)
    _run("pmv_attention_fwd", 1, dict(flops=4 * BH * Nq * Nk * 96, tc=use_tc, shape=(BH, Nq, Nk, kd)), L.ptr(q_aug), L.ptr(k_aug), ld, kd, L.ptr(v), v.stride(1), L.ptr(out), L.ptr(out_pre), L.ptr(lse),
                                      B, heads, Nq, Nk, scale, int(residual), L.dt(q_aug), use_tc, L.stream())
    return out, (out_pre if out_pre is not None else (out if want_lse else None)), lse


def attention_bwd(q_aug, k_aug, v, out, dout, lse, B, heads, kd, scale, residual=True, tc=None, fp32_dkv=False):
    BH, Nq, ld = q_aug.shape
    Nk = k_aug.shape[1]
    use_tc = _tc_default(q_aug.dtype) if tc is None else int(tc)
    dq_aug = torch.empty_like(q_aug)
    ws = torch.empty(L.lib().pmv_attention_bwd_workspace_bytes(B, heads, Nq, Nk) // 4, dtype=torch.float32,
                     device=q_aug.device)
    if use_tc and fp32_dkv:
        # the tcgen05 path accumulates dk / dv in fp32 inside ws; hand those out instead of a bf16 copy
        dk = dv = None
    else:
        dk = torch.empty(BH, Nk, 96, dtype=q_aug.dtype, device=q_aug.device)
        dv = torch.empty(BH, Nk, 96, dtype=q_aug.dtype, device=q_aug.device)
    _run("pmv_attention_bwd", 4 if use_tc else 3, dict(flops=10 * BH * Nq * Nk * 96, tc=use_tc, shape=(BH, Nq, Nk, kd)), L.ptr(q_aug), L.ptr(k_aug), ld, kd, L.ptr(v), v.stride(1), L.ptr(out),
                                      L.ptr(dout.contiguous()), L.ptr(lse), L.ptr(dq_aug), L.ptr(dk), 96, L.ptr(dv), 96,
                                      L.ptr(ws), B, heads, Nq, Nk, scale, int(residual), L.dt(q_aug), use_tc, L.stream())
    if dk is None:
        n = BH * Nk * 96
        dk, dv = ws[:n].view(BH, Nk, 96), ws[n:2 * n].view(BH, Nk, 96)
    return dq_aug, dk, dv


# ----------------------------------------------------------------------------- head + loss (row f2)
def head_loss_fwd(x, gamma, beta, w, bias, target=None, keep_mask=None, dropout_p=0.0, want_probs=False, save=False, eps=LN_EPS):
    """x [B, N, C] fp32 tokens.  target: int64 labels [B], float soft targets [B, classes] or None (no loss).
    Returns dict(logits, probs, loss, saved) — saved = (xhat, rstd, xd) when ``save``."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, N, Cdim = x.shape
    ncls = w.shape[0]
    dev = x.device
    logits = torch.empty(B, ncls, dtype=torch.float32, device=dev)
    probs = torch.empty(B, ncls, dtype=torch.float32, device=dev) if want_probs else None
    labels = soft = loss = None
    if target is not None:
        if target.dtype in (torch.int64, torch.int32):
            labels = target.to(torch.int64).contiguous()
        else:
            soft = target.to(torch.float32).contiguous()
            assert soft.shape == (B, ncls)
        loss = torch.empty((), dtype=torch.float32, device=dev)
    xhat = rstd = xd = None
    if save:
        xhat = torch.empty(B, Cdim, dtype=torch.float32, device=dev)
        xd = torch.empty(B, Cdim, dtype=torch.float32, device=dev)
        rstd = torch.empty(B, dtype=torch.float32, device=dev)
    if keep_mask is not None:
        assert keep_mask.dtype == torch.uint8 and keep_mask.shape == (B, Cdim) and keep_mask.is_contiguous()
    _run("pmv_head_loss_fwd", 2 if (loss is not None or probs is not None) else 1, dict(bytes=(B * ncls * Cdim + B * Cdim) * 4), L.ptr(x), x.stride(0), L.ptr(gamma),
         L.ptr(beta), L.ptr(w), L.ptr(bias), L.ptr(keep_mask), float(dropout_p), L.ptr(labels), L.ptr(soft), L.ptr(logits), L.ptr(probs),
         L.ptr(loss), L.ptr(xhat), L.ptr(rstd), L.ptr(xd), B, Cdim, ncls, eps, L.stream())
    return dict(logits=logits, probs=probs, loss=loss, labels=labels, soft=soft, saved=(xhat, rstd, xd))


def head_loss_bwd(dloss, logits, labels, soft, w, gamma, keep_mask, dropout_p, saved, N, has_bias=True):
    xhat, rstd, xd = saved
    B, ncls = logits.shape
    Cdim = w.shape[1]
    dev = logits.device
    dx = torch.empty(B, N, Cdim, dtype=torch.float32, device=dev)
    dw = torch.empty(ncls, Cdim, dtype=torch.float32, device=dev)
    db = torch.empty(ncls, dtype=torch.float32, device=dev) if has_bias else None
    dgamma = torch.empty(Cdim, dtype=torch.float32, device=dev)
    dbeta = torch.empty(Cdim, dtype=torch.float32, device=dev)
    ws = _ws(L.lib().pmv_head_loss_bwd_workspace_bytes(B, Cdim, ncls), dev)
    _run("pmv_head_loss_bwd", 4, dict(bytes=(2 * B * ncls * Cdim + B * N * Cdim) * 4), L.ptr(dloss), L.ptr(logits), L.ptr(labels), L.ptr(soft),
         L.ptr(w), L.ptr(gamma), L.ptr(keep_mask), float(dropout_p), L.ptr(xhat), L.ptr(rstd), L.ptr(xd), L.ptr(dx), dx.stride(0), N,
         L.ptr(dw), L.ptr(db), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), B, Cdim, ncls, L.stream())
    return dx, dw, db, dgamma, dbeta


# ----------------------------------------------------------------------------- PatchEmbed
def patch_im2col(clip, kernel, stride, padding, dtype):
    B, Cin, T, H, W = clip.shape
    To = (T + 2 * padding[0] - kernel[0]) // stride[0] + 1
    Ho = (H + 2 * padding[1] - kernel[1]) // stride[1] + 1
    Wo = (W + 2 * padding[2] - kernel[2]) // stride[2] + 1
    K = Cin * kernel[0] * kernel[1] * kernel[2]
    ld = (K + 63) // 64 * 64
    col = torch.empty(B * To * Ho * Wo, ld, dtype=dtype, device=clip.device)
    _run("pmv_patch_im2col", 1, dict(bytes=clip.numel() * 4 + col.numel() * col.element_size()), L.ptr(clip), L.ptr(col), ld, B, Cin, T, H, W, *kernel, *stride, *padding, L.dt(dtype),
                                     L.stream())
    return col, (To, Ho, Wo), K
