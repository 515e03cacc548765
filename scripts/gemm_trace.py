"""Per-tile timeline inside the persistent tcgen05 GEMM CTAs (debug build with -DPMV_ATTN_TRACE; GTRACE() points in
csrc/gemm_tc_kernel.cuh): when the MMA thread gets an accumulator stage, when its first operands have landed, when its last
MMA is committed; when the TMA thread starts / finishes a tile's loads; when the epilogue sees the accumulator, has it in
registers, has stored it.
python scripts/build_trace_lib.py, then PMV_B200_LIB=scripts/bin/libpmv_b200_trace.so python scripts/gemm_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import numpy as np
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
CTAS, ITEMS, SLOTS = 148, 12, 8
handle = ctypes.CDLL(L.LIB_PATH)
handle.pmv_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
ghz = 1.965


def trace(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    assert handle.pmv_debug_gemm_trace(None, 1) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (CTAS * ITEMS * SLOTS))()
    assert handle.pmv_debug_gemm_trace(buf, 0) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, ITEMS, SLOTS).astype(np.float64)
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us (event time of the call)")
    t0 = t[:, 0, 3:4]  # first TMA issue of the CTA
    names = ["MMA: accumulator stage free", "MMA: first operands landed", "MMA: last MMA committed", "TMA: first load of the tile issued",
             "TMA: last load issued", "epilogue: accumulator complete", "epilogue: accumulator in registers (stage released)", "epilogue: tile stored"]
    n_items = int((t[:, :, 0] > 0).sum(1).max())
    print(f"   work items per CTA: up to {n_items}; microseconds since the CTA's first TMA issue, median over CTAs")
    print("   item " + " ".join(f"{n.split(':')[0][:3]}{i}".rjust(7) for i, n in enumerate(names)))
    for it in range(min(n_items, ITEMS)):
        row = []
        for sl in range(SLOTS):
            v = t[:, it, sl]
            ok = v > 0
            row.append(np.median((v[ok] - t0[ok, 0]) / ghz / 1e3) if ok.any() else float("nan"))
        print(f"   {it:4d} " + " ".join(f"{x:7.2f}" for x in row))
    for i, n in enumerate(names):
        print(f"      {n.split(':')[0][:3]}{i} = {n}")


def lin(M, N, K, **kw):
    x = (torch.randn(M, K, device="cuda") * .5).to(dt)
    w = (torch.randn(N, K, device="cuda") * .05).to(dt)
    b = torch.randn(N, device="cuda")
    return x, w, b


M = 12552
x, w, b = lin(M, 1152, 384)
trace("qkv fwd PLAIN (12552, 1152, 384)", lambda: ops.linear_fwd(x, w, b, dt))
x2, w2, b2 = lin(M, 1536, 384)
u = torch.empty(M, 1536, dtype=dt, device="cuda")
trace("fc1 + GELU (+aux) (12552, 1536, 384)", lambda: ops.linear_fwd(x2, w2, b2, dt, act=L.ACT_GELU, aux_out=u))
x3, w3, b3 = lin(M, 384, 1536)
res = torch.randn(M, 384, device="cuda")
trace("fc2 + residual fp32 (12552, 384, 1536)", lambda: ops.linear_fwd(x3, w3, b3, torch.float32, residual=res))
x4, w4, b4 = lin(M, 384, 384)
trace("proj + residual fp32 (12552, 384, 384)", lambda: ops.linear_fwd(x4, w4, b4, torch.float32, residual=res))
dy = (torch.randn(M, 1536, device="cuda") * .5).to(dt)
trace("dgrad (12552, 384, 1536)", lambda: ops.linear_dgrad(dy, w2, dt))
trace("wgrad (1536, 384, 12552)", lambda: ops.linear_wgrad(dy, x2))
