"""Attention backward time against the dK/dV chunk count (PMV_ATTN_DKV_CHUNKS; 0 = the model of csrc/attn_tc_bwd.cu).
One process per setting (the variable is read once).  python scripts/dkv_chunk_sweep.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops
torch.manual_seed(0)
dt = torch.bfloat16
for (B, heads, Nq, Nk, ld) in [(8, 1, 25089, 393, 128), (8, 2, 6273, 1569, 160), (8, 2, 6273, 393, 128), (8, 4, 1569, 1569, 160), (8, 4, 1569, 393, 128), (8, 8, 393, 1569, 160), (8, 8, 393, 393, 128)]:
    q = (torch.randn(B * heads, Nq, ld, device="cuda") * .5).to(dt); k = (torch.randn(B * heads, Nk, ld, device="cuda") * .5).to(dt)
    v = torch.randn(B * heads, Nk, 96, device="cuda").to(dt)
    scale = 96 ** -0.5
    out, out_pre, lse = ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
    dout = torch.randn_like(out)
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    for _ in range(3):
        ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1, fp32_dkv=True)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1, fp32_dkv=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"  {B*heads:3d} x {Nq:5d} x {Nk:4d} kd {ld}: {ts[len(ts)//2]:7.1f} us", flush=True)
''' % (ROOT, ROOT)
for c in sys.argv[1:] or ["0", "1", "2", "3", "4", "6", "9"]:
    print(f"PMV_ATTN_DKV_CHUNKS={c}", flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, PMV_ATTN_DKV_CHUNKS=c), check=False)
