"""Phase timeline inside the pooling t-march CTAs (debug build with -DPMV_ATTN_TRACE: PTRACE() points in
csrc/pool_tma.cu).  python scripts/build_trace_lib.py, then
    PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/pool_trace.py
Per mode (0 forward, 2 dW, 3 stride-1 input gradient) and stage shape: the per-step durations seen by thread 0 of the
CTAs of job 0 (q) — wait for the TMA plane, conv FFMA2 section, barrier, LayerNorm phase — and the CTA lifetimes."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import numpy as np
import torch

from pmv_b200 import _lib as L, ops

torch.manual_seed(0)
dev, dt = "cuda", torch.bfloat16
B = 8
CTAS, SLOTS = 1024, 64
handle = ctypes.CDLL(L.LIB_PATH)
buf = (ctypes.c_longlong * (4 * CTAS * SLOTS))()
GHZ = 1.965


def dump(mode, title):
    assert handle.pmv_debug_pool_trace(buf) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(4, CTAS, SLOTS)[mode].astype(np.float64)
    live = t[:, 2] > t[:, 2].max() - 3e5  # the device buffer is never cleared: keep the CTAs of the latest launch (ns)
    t = t[live]
    tag = t[:, 4].astype(int)
    print(f"--- {title}: {len(t)} CTAs traced; job tags {sorted(set(tag.tolist()))}")
    g0 = t[:, 2].min()
    print(f"    kernel span (globaltimer) {(t[:, 3].max() - g0) / 1e3:.2f} us; CTA lifetime median {np.median(t[:, 3] - t[:, 2]) / 1e3:.2f} us, "
          f"max {np.max(t[:, 3] - t[:, 2]) / 1e3:.2f}; CTA start p50 {np.median(t[:, 2] - g0) / 1e3:.2f} p90 {np.percentile(t[:, 2] - g0, 90) / 1e3:.2f} us")
    for job in sorted(set(tag.tolist())):
        if job % 1000 == 999:
            continue
        tj = t[tag == job]
        print(f"    stride {job // 10000} job {job // 1000 % 10}: {len(tj)} CTAs; prologue (entry -> march) {np.median(tj[:, 5] - tj[:, 0]) / GHZ / 1e3:.2f} us; "
              f"whole CTA {np.median(tj[:, 6] - tj[:, 0]) / GHZ / 1e3:.2f} us")
        print(f"      {'step':>4s} {'wait':>7s} {'conv':>7s} {'barrier':>7s} {'LN':>7s} {'total':>7s}   (us, median over CTAs, first item)")
        for tin in range(9):
            b = 8 + 6 * tin
            s0, s1, s2, s3, s4 = (tj[:, b + k] for k in range(5))
            nxt = tj[:, b + 6] if tin < 8 else tj[:, 6]
            f = lambda a, bb: np.median((bb - a)[(a > 0) & (bb > 0)]) / GHZ / 1e3 if np.any((a > 0) & (bb > 0)) else float("nan")
            print(f"      {tin:4d} {f(s0, s1):7.2f} {f(s1, s2):7.2f} {f(s2, s3):7.2f} {f(s3, s4):7.2f} {f(s0, nxt):7.2f}")
    # clear
    ctypes.memset(buf, 0, ctypes.sizeof(buf))


STAGES = [(1, (8, 56, 56), 1, 8), (4, (8, 14, 14), 1, 2), (2, (8, 56, 56), 2, 4), (8, (8, 7, 7), 1, 1)]
for heads, thw, sq, skv in STAGES:
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
    xh = [torch.empty(B, heads, Ls[i], 96, dtype=dt, device=dev) for i in range(3)]
    rs = [torch.empty(B, heads, Ls[i], device=dev) for i in range(3)]
    for _ in range(3):
        ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xh[i], rs[i]) for i in range(3)])
        ops.pool_ln_qkv_bwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], douts[i], grads[i], xh[i], rs[i]) for i in range(3)], dqkv)
    torch.cuda.synchronize()
    # NOTE: a launch per job class overwrites the buffer of its mode: the dense class (stride 1, 2) runs first, the
    # tap-tile class (stride >= 3) second, so for block 0 the forward trace shows the LAST class launched.
    dump(0, f"forward  heads {heads} thw {thw} strides {strides}")
    dump(2, f"dW       heads {heads} thw {thw} strides {strides}")
    dump(3, f"din s=1  heads {heads} thw {thw} strides {strides}")
