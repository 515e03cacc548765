"""Micro-driver for ncu: pooling forward / backward at the MViTv2-S mid-stage shape (blocks 4-13) and at block 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops
torch.manual_seed(0)
dt = torch.bfloat16
def case(B, heads, thw, sq, skv):
    T, H, W = thw
    N = 1 + T * H * W
    qkv = torch.randn(B, N, 3, heads, 96, device="cuda").to(dt)
    ws = [torch.randn(96, 1, 3, 3, 3, device="cuda") * 0.2 for _ in range(3)]
    gs = [torch.ones(96, device="cuda") for _ in range(3)]
    bs = [torch.zeros(96, device="cuda") for _ in range(3)]
    strides = [sq, skv, skv]
    Ls = [1 + T * ops.pooled_hw(H, s) * ops.pooled_hw(W, s) for s in strides]
    lds = [128, 128, 96]
    outs = [torch.zeros(B, heads, Ls[i], lds[i], dtype=dt, device="cuda") for i in range(3)]
    xh = [torch.empty(B, heads, Ls[i], 96, dtype=dt, device="cuda") for i in range(3)]
    rs = [torch.empty(B, heads, Ls[i], device="cuda") for i in range(3)]
    douts = [torch.randn_like(o) for o in outs]
    grads = torch.zeros(3, 96 * 27 + 192, device="cuda")
    dqkv = torch.empty_like(qkv)
    def run():
        ops.pool_ln_qkv_fwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], bs[i], outs[i], xh[i], rs[i]) for i in range(3)])
        ops.pool_ln_qkv_bwd(qkv, heads, thw, [(i, strides[i], ws[i], gs[i], douts[i], grads[i], xh[i], rs[i]) for i in range(3)], dqkv)
    return run
runs = [case(8, 4, (8, 14, 14), 1, 2), case(8, 1, (8, 56, 56), 1, 8)]
for it in range(3):
    for r in runs:
        r()
torch.cuda.synchronize()
print("ok")
