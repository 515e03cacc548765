"""One tcgen05 GEMM shape per kind, a few launches each: the target of `ncu --set full --import-source on` captures
(profiles/r02_gemm_ncu_*).  python scripts/gemm_one.py [plain|gelu|gelu_bwd|res|dgrad|wgrad]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
kind = sys.argv[1] if len(sys.argv) > 1 else "plain"
M = 12552


def lin(M, N, K):
    return (torch.randn(M, K, device="cuda") * .5).to(dt), (torch.randn(N, K, device="cuda") * .05).to(dt), torch.randn(N, device="cuda")


if kind == "plain":
    x, w, b = lin(M, 1152, 384)
    fn = lambda: ops.linear_fwd(x, w, b, dt)
elif kind == "gelu":
    x, w, b = lin(M, 1536, 384)
    u = torch.empty(M, 1536, dtype=dt, device="cuda")
    fn = lambda: ops.linear_fwd(x, w, b, dt, act=L.ACT_GELU, aux_out=u)
elif kind == "gelu_bwd":
    x, w, b = lin(M, 1536, 384)  # dy [M, 384] x W2 [384, 1536] -> [M, 1536] * gelu'(u)
    dy = (torch.randn(M, 384, device="cuda") * .5).to(dt)
    w2 = (torch.randn(384, 1536, device="cuda") * .05).to(dt)
    u = (torch.randn(M, 1536, device="cuda")).to(dt)
    fn = lambda: ops.linear_dgrad(dy, w2, dt, act=L.ACT_GELU_BWD, aux_in=u)
elif kind == "res":
    x, w, b = lin(M, 384, 384)
    res = torch.randn(M, 384, device="cuda")
    fn = lambda: ops.linear_fwd(x, w, b, torch.float32, residual=res)
elif kind == "dgrad":
    dy = (torch.randn(M, 1536, device="cuda") * .5).to(dt)
    w = (torch.randn(1536, 384, device="cuda") * .05).to(dt)
    fn = lambda: ops.linear_dgrad(dy, w, dt)
else:
    dy = (torch.randn(M, 1536, device="cuda") * .5).to(dt)
    x = (torch.randn(M, 384, device="cuda") * .5).to(dt)
    fn = lambda: ops.linear_wgrad(dy, x)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok", kind)
