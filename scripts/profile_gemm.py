"""Micro-driver for ncu: three GEMM cases that bracket the epilogue cost (small-K qkv, GELU fwd, GELU' dgrad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
def mk(M, N, K):
    return (torch.randn(M, K, device="cuda").to(dt), (torch.randn(N, K, device="cuda") * .05).to(dt), torch.randn(N, device="cuda"))
x1, w1, b1 = mk(12552, 1152, 384)
x2, w2, b2 = mk(12552, 1536, 384)
dy3 = torch.randn(12552, 384, device="cuda").to(dt); w3 = (torch.randn(384, 1536, device="cuda") * .05).to(dt); u3 = torch.randn(12552, 1536, device="cuda").to(dt)
for it in range(2):
    ops.linear_fwd(x1, w1, b1, dt)
    ops.linear_fwd(x2, w2, b2, dt, act=L.ACT_GELU)
    ops.linear_dgrad(dy3, w3, dt, act=L.ACT_GELU_BWD, aux_in=u3)
torch.cuda.synchronize()
print("ok")
