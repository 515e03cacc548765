"""Kernel timeline of ONE replay of the graphed MViTv2-S training step under torchrun (N ranks, NCCL), rank 0's view:
where the collectives sit, what runs beside them, and what is exposed after the last gradient.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/timeline_ddp.py [tag]
env: PMV_BUCKET_MB (default 25), PMV_LAST_BUCKET_MB (size of the bucket holding the FIRST layers = the last one reduced).
Diagnostic only (a run under the profiler is never a bench value).  Writes gpurun_out/timeline_ddp_<tag>.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from pmv_b200 import mvit
from pmv_b200.ddp import GradAllReducer
from pmv_b200.graphs import GraphedStep
from pmv_b200.optim import FusedAdamW, param_groups

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
tag = sys.argv[1] if len(sys.argv) > 1 else f"n{world}"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = mvit.MViT(dict(mvit.MVITV2_S, droppath_batched=True), compute_dtype=torch.bfloat16).to(dev).train()
B = 8
clips = torch.randn(B, 3, 16, 224, 224, device=dev)
labels = torch.randint(0, 400, (B,), device=dev)
kw = {"last_bucket_mb": float(os.environ.get("PMV_LAST_BUCKET_MB", "2"))}
reducer = GradAllReducer(model, bucket_mb=float(os.environ.get("PMV_BUCKET_MB", "25")), **kw)
opt = FusedAdamW(param_groups(model, 0.05, zero_wd_1d=True), lr=1e-4, max_grad_norm=1.0)


def step(c, l):
    reducer.zero_grad()
    loss, _ = model.forward_loss([c], l)
    loss.backward()
    reducer.finish()
    opt.step()
    return loss


g = GraphedStep(step, [clips, labels])
for _ in range(5):
    g(clips, labels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
# device-timed replays without the profiler (context for the timeline below)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g(clips, labels)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(clips, labels)
    torch.cuda.synchronize()
if rank == 0:
    ev = [(e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in prof.events()
          if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda r: r[1])
    t0 = ev[0][1]
    rows = [[n, round(s - t0, 3), round(d, 3)] for n, s, d in ev]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"timeline_ddp_{tag}.json"), "w"))
    nccl = [r for r in rows if "nccl" in r[0].lower()]
    comp = [r for r in rows if "nccl" not in r[0].lower()]
    span = max(r[1] + r[2] for r in rows)
    adam = [r for r in comp if "adamw" in r[0] or "grad_sumsq" in r[0]]
    first_opt = min(r[1] for r in adam) if adam else span
    last_bwd = max(r[1] + r[2] for r in comp if r[1] < first_opt)
    print(f"[{tag}] world {world}: {ms:.3f} ms per replay (CUDA events, no profiler); profiled span {span / 1e3:.3f} ms; "
          f"compute kernels busy {sum(r[2] for r in comp) / 1e3:.3f} ms; {len(nccl)} NCCL kernels busy {sum(r[2] for r in nccl) / 1e3:.3f} ms; "
          f"buckets {reducer.num_buckets}")
    for r in nccl:
        beside = sum(min(c[1] + c[2], r[1] + r[2]) - max(c[1], r[1]) for c in comp if c[1] < r[1] + r[2] and c[1] + c[2] > r[1])
        print(f"    NCCL {r[0][:48]:48s} start {r[1] / 1e3:8.3f} ms  dur {r[2]:8.1f} us  compute beside it {beside:8.1f} us")
    print(f"    last backward kernel ends {last_bwd / 1e3:.3f} ms, optimizer starts {first_opt / 1e3:.3f} ms -> exposed {first_opt - last_bwd:.1f} us")
    fam = {}
    for n, s, d in comp:
        k = n.split("<")[0].split("(")[0][-40:]
        fam[k] = fam.get(k, 0.0) + d
    top = sorted(fam.items(), key=lambda kv: -kv[1])[:12]
    print("    compute by kernel (us): " + ", ".join(f"{k}={v:.0f}" for k, v in top))
if world > 1:
    del g
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
