"""Per-kernel micro-benchmarks at the MViTv2-S stage shapes (B = 8 clips), CUDA-event timed with an L2 flush
between iterations.  Usage: python scripts/bench_kernels.py [pool] [gemm] [relpos] [attn] [ln]
Prints one line per (kernel, shape): median microseconds and achieved GB/s or TFLOP/s (algorithmic work)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch  # noqa: E402

from pmv_b200 import _lib as L  # noqa: E402
from pmv_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
dt = torch.bfloat16
FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B = int(os.environ.get("PMV_BENCH_B", "8"))

# (dim, att_dim, heads, thw, stride_q, stride_kv)
STAGES = [(96, 96, 1, (8, 56, 56), 1, 8), (96, 192, 2, (8, 56, 56), 2, 4), (192, 192, 2, (8, 28, 28), 1, 4),
          (192, 384, 4, (8, 28, 28), 2, 2), (384, 384, 4, (8, 14, 14), 1, 2), (384, 768, 8, (8, 14, 14), 2, 1),
          (768, 768, 8, (8, 7, 7), 1, 1)]


def timeit(fn, iters=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, shape, us, nbytes=None, flops=None):
    extra = ""
    if nbytes is not None:
        extra += f"  {nbytes / us / 1e3:8.1f} GB/s ({nbytes / 1e6:.1f} MB)"
    if flops is not None:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    print(f"{name:28s} {str(shape):44s} {us:9.1f} us{extra}", flush=True)


def bench_pool():
    for dim, att, heads, thw, sq, skv in STAGES:
        T, H, W = thw
        N = 1 + T * H * W
        qkv = torch.randn(B, N, 3, heads, 96, device=dev).to(dt)
        ws = [torch.randn(96, 1, 3, 3, 3, device=dev) * 0.2 for _ in range(3)]
        gs = [torch.ones(96, device=dev) for _ in range(3)]
        bs = [torch.zeros(96, device=dev) for _ in range(3)]
        strides = [sq, skv, skv]
        Ls = [1 + T * ops.pooled_hw(H, s) * ops.pooled_hw(W, s) for s in strides]
        lds = [128, 128, 96]
        outs = [torch.zeros(B, heads, Ls[i], lds[i], dtype=dt, device=dev) for i in range(3)]
        douts = [torch.randn_like(o) for o in outs]
        grads = torch.zeros(3, 96 * 27 + 192, device=dev)
        dqkv = torch.empty_like(qkv)
        e = qkv.element_size()
        fwd_bytes = (3 * (N - 1) * heads * 96 + sum(Ls) * heads * 96) * B * e   # read 3(N-1)C, write (Lq+2Lk+3)C
        xh = [torch.empty(B, heads, Ls[i], 96, dtype=dt, device=dev) for i in range(3)]
        rs = [torch.empty(B, heads, Ls[i], device=dev) for i in range(3)]
        us = timeit(lambda: ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i]) for i in range(3)]))
        report("pool_ln_qkv_fwd (infer)", (B, heads, thw, strides), us, nbytes=fwd_bytes)
        us = timeit(lambda: ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xh[i], rs[i]) for i in range(3)]))
        report("pool_ln_qkv_fwd", (B, heads, thw, strides), us, nbytes=fwd_bytes)
        bwd_bytes = (2 * 3 * (N - 1) * heads * 96 + sum(Ls) * heads * 96) * B * e
        us = timeit(lambda: ops.pool_ln_qkv_bwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], douts[i], grads[i], xh[i], rs[i]) for i in range(3)], dqkv))
        report("pool_ln_qkv_bwd", (B, heads, thw, strides), us, nbytes=bwd_bytes)


def bench_relpos():
    for dim, att, heads, thw, sq, skv in STAGES:
        T, H, W = thw
        q_shape = (T, ops.pooled_hw(H, sq), ops.pooled_hw(W, sq))
        k_shape = (T, ops.pooled_hw(H, skv), ops.pooled_hw(W, skv))
        Nq, Nk = 1 + q_shape[0] * q_shape[1] * q_shape[2], 1 + k_shape[0] * k_shape[1] * k_shape[2]
        ld = ops.aug_ld(k_shape)
        nh = 2 * max(q_shape[1], k_shape[1]) - 1
        rh, rw, rt = (torch.randn(nh, 96, device=dev) * .02, torch.randn(nh, 96, device=dev) * .02,
                      torch.randn(2 * T - 1, 96, device=dev) * .02)
        q_aug = torch.randn(B * heads, Nq, ld, device=dev).to(dt)
        k_aug = torch.randn(B * heads, Nk, ld, device=dev).to(dt)
        dq = torch.randn_like(q_aug)
        ncol = k_shape[0] + k_shape[1] + k_shape[2]
        nb = B * heads * Nq * (96 + ncol) * 2
        us = timeit(lambda: ops.relpos_augment_q(q_aug, q_shape, k_shape, rh, rw, rt, 96 ** 0.5))
        report("relpos_augment_q", (B * heads, Nq, ld, k_shape), us, nbytes=nb, flops=2.0 * B * heads * Nq * 96 * ncol)
        us = timeit(lambda: ops.relpos_augment_k(k_aug, k_shape))
        report("relpos_augment_k", (B * heads, Nk, ld), us, nbytes=B * heads * Nk * (ld - 96) * 2)
        us = timeit(lambda: ops.relpos_augment_q_bwd(dq, q_aug, q_shape, k_shape, rh, rw, rt, 96 ** 0.5))
        report("relpos_augment_q_bwd", (B * heads, Nq, ld, k_shape), us, nbytes=B * heads * Nq * (2 * 96 + ncol + 96) * 2,
               flops=4.0 * B * heads * Nq * 96 * ncol)


def bench_gemm():
    shapes = set()
    for dim, att, heads, thw, sq, skv in STAGES:
        T, H, W = thw
        M = B * (1 + T * H * W)
        Mq = B * (1 + T * ops.pooled_hw(H, sq) * ops.pooled_hw(W, sq))
        shapes.add(("qkv", M, 3 * att, dim))
        shapes.add(("proj", Mq, att, att))
        shapes.add(("fc1", Mq, 4 * att, att))
        shapes.add(("fc2", Mq, att, 4 * att))
    for name, M, N, K in sorted(shapes, key=lambda s: (s[0], -s[1])):
        x = torch.randn(M, K, device=dev).to(dt)
        w = torch.randn(N, K, device=dev).to(dt) * 0.05
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        dy = torch.randn(M, N, device=dev).to(dt)
        u = torch.randn(M, N, device=dev).to(dt)
        fl = 2.0 * M * N * K
        if name == "fc1":
            aux = torch.empty(M, N, dtype=dt, device=dev)
            us = timeit(lambda: ops.linear_fwd(x, w, bias, dt, act=L.ACT_GELU, aux_out=aux))
            report("gemm fwd fc1+gelu(+aux)", (M, N, K), us, flops=fl, nbytes=(M * K + N * K + 2 * M * N) * 2)
            us = timeit(lambda: ops.linear_fwd(x, w, bias, dt, act=L.ACT_GELU))
            report("gemm fwd fc1+gelu", (M, N, K), us, flops=fl, nbytes=(M * K + N * K + M * N) * 2)
            us = timeit(lambda: ops.linear_dgrad(dy, w, dt))
            report("gemm dgrad", (M, K, N), us, flops=fl, nbytes=(M * K + N * K + M * N) * 2)
        elif name in ("fc2", "proj"):
            us = timeit(lambda: ops.linear_fwd(x, w, bias, torch.float32, residual=res))
            report(f"gemm fwd {name}+res(f32)", (M, N, K), us, flops=fl, nbytes=(M * K + N * K) * 2 + 8 * M * N)
            if name == "fc2":
                uu = torch.randn(M, K, device=dev).to(dt)
                us = timeit(lambda: ops.linear_dgrad(dy, w, dt, act=L.ACT_GELU_BWD, aux_in=uu))
                report("gemm dgrad+gelu'", (M, K, N), us, flops=fl, nbytes=(M * N + N * K + 2 * M * K) * 2)
            else:
                us = timeit(lambda: ops.linear_dgrad(dy, w, dt))
                report("gemm dgrad", (M, K, N), us, flops=fl, nbytes=(M * K + N * K + M * N) * 2)
        else:
            us = timeit(lambda: ops.linear_fwd(x, w, bias, dt))
            report("gemm fwd qkv", (M, N, K), us, flops=fl, nbytes=(M * K + N * K + M * N) * 2)
            us = timeit(lambda: ops.linear_dgrad(dy, w, dt))
            report("gemm dgrad", (M, K, N), us, flops=fl, nbytes=(M * K + N * K + M * N) * 2)
        us = timeit(lambda: ops.linear_wgrad(dy, x))
        report("gemm wgrad", (N, K, M), us, flops=fl, nbytes=(M * K + M * N) * 2 + N * K * 4)


def bench_attn():
    for dim, att, heads, thw, sq, skv in STAGES:
        T, H, W = thw
        q_shape = (T, ops.pooled_hw(H, sq), ops.pooled_hw(W, sq))
        k_shape = (T, ops.pooled_hw(H, skv), ops.pooled_hw(W, skv))
        Nq, Nk = 1 + q_shape[0] * q_shape[1] * q_shape[2], 1 + k_shape[0] * k_shape[1] * k_shape[2]
        ld = ops.aug_ld(k_shape)
        q_aug = (torch.randn(B * heads, Nq, ld, device=dev) * 0.5).to(dt)
        k_aug = (torch.randn(B * heads, Nk, ld, device=dev) * 0.5).to(dt)
        v = torch.randn(B * heads, Nk, 96, device=dev).to(dt)
        scale = 96 ** -0.5
        fl = 4.0 * B * heads * Nq * Nk * 96
        out, out_pre, lse = ops.attention_fwd(q_aug, k_aug, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
        us = timeit(lambda: ops.attention_fwd(q_aug, k_aug, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1))
        report("attention_fwd", (B * heads, Nq, Nk, ld), us, flops=fl)
        dout = torch.randn_like(out)
        us = timeit(lambda: ops.attention_bwd(q_aug, k_aug, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1))
        report("attention_bwd", (B * heads, Nq, Nk, ld), us, flops=2.5 * fl)


def bench_ln():
    for dim, att, heads, thw, sq, skv in STAGES:
        T, H, W = thw
        M = B * (1 + T * H * W)
        x = torch.randn(M, dim, device=dev)
        g, b = torch.ones(dim, device=dev), torch.zeros(dim, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, b, dt)
        us = timeit(lambda: ops.layernorm_fwd(x, g, b, dt))
        report("layernorm_fwd", (M, dim), us, nbytes=M * dim * 6)
        dy = torch.randn(M, dim, device=dev).to(dt)
        acc = torch.zeros_like(x)
        us = timeit(lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dx_accum=acc))
        report("layernorm_bwd(+acc)", (M, dim), us, nbytes=M * dim * 14)
        us = timeit(lambda: ops.colsum_cast(x, dt))
        report("colsum_cast f32->bf16", (M, dim), us, nbytes=M * dim * 6)


if __name__ == "__main__":
    what = sys.argv[1:] or ["pool", "relpos", "gemm", "attn", "ln"]
    print(f"device {torch.cuda.get_device_name(0)}  B={B}")
    for w in what:
        {"pool": bench_pool, "relpos": bench_relpos, "gemm": bench_gemm, "attn": bench_attn, "ln": bench_ln}[w]()
