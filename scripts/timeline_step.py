"""Kernel timeline of ONE replay of the graphed MViTv2-S training (or inference) step.

Writes gpurun_out/timeline_<tag>.json: [[kernel name, start_us, dur_us], ...] in start order, and prints the busy time,
the idle time between kernels and the largest gaps.  A timeline under the profiler is diagnostic only (never a bench value).
usage: python scripts/timeline_step.py [train|infer] [tag]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from pmv_b200 import mvit
from pmv_b200.ddp import GradAllReducer
from pmv_b200.graphs import GraphedStep

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
tag = sys.argv[2] if len(sys.argv) > 2 else mode
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mvit.MViT(mvit.MVITV2_S, compute_dtype=torch.bfloat16).to(dev)
B = 8
clips = torch.randn(B, 3, 16, 224, 224, device=dev)
labels = torch.randint(0, 400, (B,), device=dev)
if mode == "train":
    model.train()
    from pmv_b200.optim import FusedAdamW, param_groups
    reducer = GradAllReducer(model, bucket_mb=25.0)
    opt = FusedAdamW(param_groups(model, 0.05, zero_wd_1d=True), lr=1e-4, max_grad_norm=1.0)  # as in bench.py

    def step(c, l):
        reducer.zero_grad()
        loss, _ = model.forward_loss([c], l)
        loss.backward()
        reducer.finish()
        opt.step()
        return loss
else:
    model.eval(); model.head.act = None
    from pmv_b200.attention import cache_low_precision_weights
    cache_low_precision_weights(model)  # as in bench.py

    def step(c, l):
        with torch.no_grad():
            return model([c])
g = GraphedStep(step, [clips, labels])
for _ in range(3):
    g(clips, labels)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(clips, labels)
    torch.cuda.synchronize()
ev = [(e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in prof.events()
      if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda r: r[1])
t0 = ev[0][1]
rows = [[n, round(s - t0, 3), round(d, 3)] for n, s, d in ev]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.json"), "w") as f:
    json.dump(rows, f)
busy = sum(r[2] for r in rows)
span = rows[-1][1] + rows[-1][2]
print(f"{tag}: {len(rows)} device activities, span {span/1e3:.3f} ms, busy {busy/1e3:.3f} ms, idle {(span-busy)/1e3:.3f} ms")
