"""One MViTv2-S training step under torch.profiler: every CUDA kernel (ours and torch's eager ones) by total time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from pmv_b200 import mvit
from pmv_b200.ddp import GradAllReducer
mode = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = mvit.MViT(mvit.MVITV2_S, compute_dtype=torch.bfloat16).to(dev)
B = 8
clips = torch.randn(B, 3, 16, 224, 224, device=dev)
labels = torch.randint(0, 400, (B,), device=dev)
if mode == "train":
    model.train()
    reducer = GradAllReducer(model, bucket_mb=25.0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05, fused=True)
    def step():
        reducer.zero_grad()
        loss = torch.nn.functional.cross_entropy(model([clips]), labels)
        loss.backward()
        reducer.finish()
        opt.step()
else:
    model.eval(); model.head.act = None
    def step():
        with torch.no_grad():
            model([clips])
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/5:.2f} ms/step, wall {1e3*(t2-t0)/5:.2f} ms/step")
