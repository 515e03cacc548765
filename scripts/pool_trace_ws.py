"""Phase timeline of the warp-specialised pooling forward (pool_ws_fwd_kernel; debug build, WTRACE points).
    python scripts/build_trace_lib.py && PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/pool_trace_ws.py
First item of every CTA.  Conv warp 0, per step: wait(full) / FFMA2 section / park (cfree wait + STS).  LayerNorm warp 0,
per frame: wait (empty + TMA request + parked) / LayerNorm + stores."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import numpy as np
import torch
from pmv_b200 import _lib as L, ops
torch.manual_seed(0)
dev, dt = "cuda", torch.bfloat16
B = 8
CTAS, SLOTS, GHZ = 1024, 64, 1.965
handle = ctypes.CDLL(L.LIB_PATH)
buf = (ctypes.c_longlong * (4 * CTAS * SLOTS))()


def med(a, b):
    m = (a > 0) & (b > 0)
    return np.median((b - a)[m]) / GHZ / 1e3 if np.any(m) else float("nan")


def dump(title):
    assert handle.pmv_debug_pool_trace(buf) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(4, CTAS, SLOTS)[1].astype(np.float64)
    t = t[t[:, 2] > t[:, 2].max() - 3e5]
    tag = t[:, 4].astype(int)
    g0 = t[:, 2].min()
    print(f"--- {title}: {len(t)} CTAs; kernel span {(t[:, 3].max() - g0) / 1e3:.2f} us; CTA lifetime median {np.median(t[:, 3] - t[:, 2]) / 1e3:.2f} max {np.max(t[:, 3] - t[:, 2]) / 1e3:.2f} us")
    for job in sorted(set(tag.tolist())):
        if job % 1000 == 999:
            continue
        tj = t[tag == job]
        print(f"    stride {job // 10000} job {job // 1000 % 10}: {len(tj)} CTAs, conv warp 0 alive {np.median(tj[:, 6] - tj[:, 0]) / GHZ / 1e3:.2f} us")
        print(f"      {'step':>4s} {'c.wait':>7s} {'c.fma':>7s} {'c.park':>7s} {'c.total':>7s} | {'ln.wait':>7s} {'lds+sum':>7s} {'shfl1':>7s} {'ctr+sh2':>7s} {'stores':>7s}")
        for st in range(8):
            c = 8 + 4 * st
            l = 40 + 6 * st
            nxt = tj[:, c + 4] if st < 7 else tj[:, 6]
            ln = " ".join(f"{med(tj[:, l + q], tj[:, l + q + 1]):7.2f}" for q in range(5)) if st < 4 else ""
            print(f"      {st:4d} {med(tj[:, c], tj[:, c + 1]):7.2f} {med(tj[:, c + 1], tj[:, c + 2]):7.2f} {med(tj[:, c + 2], tj[:, c + 3]):7.2f} {med(tj[:, c], nxt):7.2f} | {ln}")

for heads, thw, sq, skv in [(4, (8, 14, 14), 1, 2), (1, (8, 56, 56), 1, 8), (2, (8, 56, 56), 2, 4)]:
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
    xh = [torch.empty(B, heads, Ls[i], 96, dtype=dt, device=dev) for i in range(3)]
    rs = [torch.empty(B, heads, Ls[i], device=dev) for i in range(3)]
    for _ in range(3):
        ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xh[i], rs[i]) for i in range(3)])
    torch.cuda.synchronize()
    dump(f"forward (ws) heads {heads} thw {thw} strides {strides}")
