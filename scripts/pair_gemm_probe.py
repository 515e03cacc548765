"""CTA-pair GEMM bring-up: correctness against torch and timing, with PMV_GEMM_PAIR from the environment."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from pmv_b200 import ops, _lib as L
torch.manual_seed(0)
dt = torch.bfloat16
FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=7):
    fn(); fn()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
print("PMV_GEMM_PAIR =", os.environ.get("PMV_GEMM_PAIR"))
for (M, N, K) in [(12552, 1152, 384), (12552, 1536, 384), (12552, 384, 1536), (12552, 384, 384), (50184, 768, 192), (200712, 384, 96), (3144, 3072, 768), (9999, 256, 200)]:
    x = torch.randn(M, K, device="cuda").to(dt); w = (torch.randn(N, K, device="cuda") * .05).to(dt); b = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    ref = x.float() @ w.float().t() + b
    y = ops.linear_fwd(x, w, b, dt)
    torch.cuda.synchronize()
    e1 = float((y.float() - ref).abs().max() / ref.abs().max())
    aux = torch.empty(M, N, dtype=dt, device="cuda")
    h = ops.linear_fwd(x, w, b, dt, act=L.ACT_GELU, aux_out=aux)
    e2 = float((h.float() - torch.nn.functional.gelu(ref)).abs().max() / ref.abs().max())
    e2b = float((aux.float() - ref).abs().max() / ref.abs().max())
    r = ops.linear_fwd(x, w, b, torch.float32, residual=res)
    e3 = float((r - (ref + res)).abs().max() / ref.abs().max())
    t1 = timeit(lambda: ops.linear_fwd(x, w, b, dt))
    t2 = timeit(lambda: ops.linear_fwd(x, w, b, dt, act=L.ACT_GELU, aux_out=aux))
    t3 = timeit(lambda: ops.linear_fwd(x, w, b, torch.float32, residual=res))
    print(f"{(M, N, K)}: err plain {e1:.1e} gelu {e2:.1e} aux {e2b:.1e} res {e3:.1e} | us plain {t1:.1f} gelu {t2:.1f} res {t3:.1f}  ({2*M*N*K/t1/1e6:.0f} TF/s)", flush=True)
print("ok")
