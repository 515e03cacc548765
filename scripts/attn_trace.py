"""Phase timeline inside the attention forward CTAs (debug build with -DPMV_ATTN_TRACE, see the TRACE() points in
csrc/attn_tc.cu).  Build the traced library first (python scripts/build_trace_lib.py), then:  PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/attn_trace.py
Prints, per phase, the median / p90 time since CTA start over all CTAs, and the CTA start offsets per SM."""
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
for _ in range(3):
    ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
torch.cuda.synchronize()
CTAS, SLOTS = 1024, 24
buf = (ctypes.c_longlong * (CTAS * SLOTS))()
lib = L.lib()
handle = ctypes.CDLL(L.LIB_PATH)
assert handle.pmv_debug_attn_trace(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)[:416].astype(np.float64)
ghz = 1.965
names = {1: "prologue done (barriers, TMEM alloc, sync)", 2: "TMA: Q + first K/V requested", 3: "MMA: Q landed", 4: "MMA: K0/V0 landed",
         8: "softmax: S0 ready", 12: "softmax: P0 written", 5: "epilogue: O in registers", 9: "softmax: S1 ready", 13: "softmax: P1 written",
         6: "epilogue: tiles staged in smem", 10: "softmax: S2 ready", 14: "softmax: P2 written", 7: "epilogue: bulk stores issued", 11: "softmax: S3 ready",
         15: "softmax: P3 written", 22: "softmax: S1 in registers", 23: "softmax: exp loop 1 issued", 16: "softmax: O final", 17: "epilogue stores done", 18: "TMEM freed (CTA end)"}
print(f"{'phase':48s} {'median us':>10s} {'p10':>8s} {'p90':>8s}")
for slot in [1, 2, 3, 4, 8, 12, 9, 22, 23, 13, 10, 14, 11, 15, 16, 5, 6, 7, 17, 18]:
    d = (t[:, slot] - t[:, 0]) / ghz / 1e3
    print(f"{names[slot]:48s} {np.median(d):10.2f} {np.percentile(d, 10):8.2f} {np.percentile(d, 90):8.2f}")
g0 = t[:, 19].min()
start = (t[:, 19] - g0) / 1e3
end = (t[:, 20] - g0) / 1e3
print(f"kernel span by globaltimer: {end.max():.2f} us; CTA lifetime median {np.median(end - start):.2f} us")
sm = t[:, 21].astype(int)
per_sm = {}
for i in range(len(sm)):
    per_sm.setdefault(sm[i], []).append((start[i], end[i]))
gaps = []
for s_, lst in per_sm.items():
    lst.sort()
    for a, b in zip(lst, lst[1:]):
        gaps.append(b[0] - a[1])
print(f"CTAs per SM: {np.mean([len(v) for v in per_sm.values()]):.2f}; gap between consecutive CTAs on an SM: median {np.median(gaps):.2f} us, p90 {np.percentile(gaps, 90):.2f}")
print(f"first-wave start spread: p90 {np.percentile(sorted(start)[:148], 90):.2f} us")
