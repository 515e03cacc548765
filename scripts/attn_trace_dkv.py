"""Phase timeline inside the attention backward dK/dV CTAs (debug build with -DPMV_ATTN_TRACE; KTRACE() points in
csrc/attn_tc_bwd.cu) plus the CTA schedule (SM id, global timer at CTA start / end).
python scripts/build_trace_lib.py, then PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/attn_trace_dkv.py [B heads Nq Nk ld]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import numpy as np
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
args = [int(a) for a in sys.argv[1:6]]
B, heads, Nq, Nk, ld = args if len(args) == 5 else (8, 4, 1569, 393, 128)
q = (torch.randn(B * heads, Nq, ld, device="cuda") * .5).to(dt); k = (torch.randn(B * heads, Nk, ld, device="cuda") * .5).to(dt)
v = torch.randn(B * heads, Nk, 96, device="cuda").to(dt)
scale = 96 ** -0.5
out, out_pre, lse = ops.attention_fwd(q, k, v, B, heads, ld, scale, residual=True, want_lse=True, tc=1)
dout = torch.randn_like(out)
for _ in range(3):
    ops.attention_bwd(q, k, v, out_pre, dout, lse, B, heads, ld, scale, residual=True, tc=1)
torch.cuda.synchronize()
CTAS, SLOTS = 4096, 32
buf = (ctypes.c_longlong * (CTAS * SLOTS))()
handle = ctypes.CDLL(L.LIB_PATH)
assert handle.pmv_debug_attn_dkv_trace(buf) == 0
k_tiles = (Nk + 127) // 128
q_tiles = (Nq + 127) // 128
BH = B * heads
def dkv_chunks():  # csrc/attn_tc_bwd.cu: dkv_chunks()
    forced = int(os.environ.get("PMV_ATTN_DKV_CHUNKS", "0"))
    if forced > 0:
        return min(forced, q_tiles)
    best, best_c = 1e30, 1
    for c in range(1, min(q_tiles, 64) + 1):
        rounds = (k_tiles * BH * c + 147) // 148
        t = rounds * ((4.5 if c == 1 else 5.8) + ((q_tiles + c - 1) // c) * (2.05 if ld == 128 else 2.3))
        if t < best - 1e-9:
            best, best_c = t, c
    return best_c
chunks = dkv_chunks()
per = (q_tiles + chunks - 1) // chunks
n = min(CTAS, k_tiles * BH * chunks)
t = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)[:n].astype(np.float64)
ghz = 1.965
print(f"shape B={B} heads={heads} Nq={Nq} Nk={Nk} ld={ld}: grid {k_tiles} x {BH} x {chunks} = {k_tiles * BH * chunks} CTAs, <= {per} query tiles per chunk")
gt0 = t[:, 30].min()
start = (t[:, 30] - gt0) / 1e3
end = (t[:, 31] - gt0) / 1e3
print(f"kernel makespan (global timer, first CTA start -> last CTA end): {end.max():.1f} us; CTA duration median {np.median(end - start):.2f} us, "
      f"p10 {np.percentile(end - start, 10):.2f}, p90 {np.percentile(end - start, 90):.2f}")
sm = t[:, 29].astype(int)
per_sm = np.bincount(sm, minlength=148)
busy = np.zeros(148)
for s_, a, b in zip(sm, start, end):
    busy[s_] += b - a
print(f"CTAs per SM: min {per_sm.min()} max {per_sm.max()}; SM busy time: median {np.median(busy):.1f} us, min {busy.min():.1f}, max {busy.max():.1f}")
cta = np.arange(n)
kx = cta % k_tiles
z = cta // (k_tiles * BH)
for name, mask in [("full key tiles", kx < k_tiles - 1), ("tail key tile", kx == k_tiles - 1)]:
    if mask.any():
        print(f"  {name:32s}: {int(mask.sum()):4d} CTAs, duration median {np.median((end - start)[mask]):.2f} us")
names = {1: "prologue done (barriers, TMEM)", 2: "MMA thread: K', V landed", 3: "MMA thread: Q'/dO tile 0 landed", 4: "MMA thread: Q'/dO tile 1 landed",
         5: "MMA thread: Q'/dO tile 2 landed", 6: "MMA thread: starts issuing S^T/dP^T of half 2", 7: "MMA thread: S^T/dP^T of half 2 committed",
         8: "MMA thread: dV/dK products of half 1 issued", 25: "softmax: loop done", 26: "softmax: accumulators final", 27: "softmax: reductions issued",
         28: "TMEM freed (CTA end)"}
order = [1, 2, 3, 4, 5]
for hh in range(8):
    names[9 + 2 * hh] = f"softmax: S^T/dP^T of half {hh} ready"; names[10 + 2 * hh] = f"softmax: P^T/dS^T of half {hh} written"
    order += [9 + 2 * hh, 10 + 2 * hh]
order += [6, 7, 8, 25, 26, 27, 28]
sel = (kx < k_tiles - 1) if per >= 4 else np.ones(n, bool)
print(f"{'phase (full key tiles of full chunks)':48s} {'median us':>10s} {'p10':>8s} {'p90':>8s}")
for slot in order:
    d = (t[sel, slot] - t[sel, 0]) / ghz / 1e3
    d = d[t[sel, slot] > 0]
    if d.size:
        print(f"{names[slot]:48s} {np.median(d):10.2f} {np.percentile(d, 10):8.2f} {np.percentile(d, 90):8.2f}")
