"""Phase timeline inside the attention backward dQ CTAs (debug build with -DPMV_ATTN_TRACE; BTRACE() points in
csrc/attn_tc_bwd.cu).  python scripts/build_trace_lib.py, then PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/attn_trace_bwd.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import numpy as np
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
B, heads, Nq, Nk, ld = 8, 4, 1569, 393, 128
q = (torch.randn(B * heads, Nq, ld, device="cuda") * .5).to(dt); k = (torch.randn(B * heads, Nk, ld, device="cuda") * .5).to(dt)
v = torch.randn(B * heads, Nk, 96, device="cuda").to(dt)
scale = 96 ** -0.5
out, out_pre, lse = ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
dout = torch.randn_like(out)
for _ in range(3):
    ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1)
torch.cuda.synchronize()
CTAS, SLOTS = 1024, 24
buf = (ctypes.c_longlong * (CTAS * SLOTS))()
handle = ctypes.CDLL(L.LIB_PATH)
assert handle.pmv_debug_attn_bwd_trace(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)[:416].astype(np.float64)
ghz = 1.965
names = {1: "prologue done"}
order = [1]
for hh in range(7):
    names[2 + 2 * hh] = f"softmax: S/dP of half {hh} ready"; names[3 + 2 * hh] = f"softmax: dS of half {hh} written"
    order += [2 + 2 * hh, 3 + 2 * hh]
names.update({21: "MMA thread: starts issuing S/dP of half 2", 22: "MMA thread: S/dP of half 2 issued + committed", 23: "MMA thread: dQ += dS K of half 1 issued", 18: "softmax: dQ final", 19: "epilogue stores issued", 20: "TMEM freed (CTA end)"})
order += [21, 22, 23, 18, 19, 20]
print(f"{'phase':40s} {'median us':>10s} {'p10':>8s} {'p90':>8s}")
for slot in order:
    d = (t[:, slot] - t[:, 0]) / ghz / 1e3
    print(f"{names[slot]:40s} {np.median(d):10.2f} {np.percentile(d, 10):8.2f} {np.percentile(d, 90):8.2f}")
