"""Micro-driver for ncu: attention backward at the MViTv2-S mid-stage shape (blocks 4-13, B = 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops
torch.manual_seed(0)
dt = torch.bfloat16
B, heads, Nq, Nk, ld = 8, 4, 1569, 393, 128
q = (torch.randn(B * heads, Nq, ld, device="cuda") * .5).to(dt); k = (torch.randn(B * heads, Nk, ld, device="cuda") * .5).to(dt)
v = torch.randn(B * heads, Nk, 96, device="cuda").to(dt)
scale = 96 ** -0.5
out, out_pre, lse = ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
dout = torch.randn_like(out)
for it in range(3):
    ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
    ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1)
torch.cuda.synchronize()
print("ok")
