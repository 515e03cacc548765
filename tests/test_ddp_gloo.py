"""world_size-2 gloo test of the bucketed gradient all-reduce (host logic of the N>1 path, CPU only)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "portrait-mode-video_b200"))
    from pmv_b200.ddp import GradAllReducer
    torch.manual_seed(rank)  # rank-dependent initialisation: the reducer must broadcast rank 0's weights like DDP does
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.LayerNorm(8))
    red = GradAllReducer(net, bucket_mb=0.001)  # force several buckets
    assert red.num_buckets > 1
    w0 = [p.detach().clone() for p in net.parameters()]
    dist.broadcast_object_list(chk := [[float(t.sum()) for t in w0]], src=0)
    assert all(abs(a - float(t.sum())) < 1e-6 for a, t in zip(chk[0], w0)), "parameters differ across ranks after construction"
    g = torch.Generator().manual_seed(1)
    xs = torch.randn(world * 4, 16, generator=g)
    for _ in range(2):  # two steps: buckets must re-arm
        red.zero_grad()
        x = xs[rank * 4:(rank + 1) * 4]
        net(x).square().mean().backward()
        red.finish()
    got = [p.grad.clone() for p in net.parameters()]
    # single-process gradient on the concatenated batch (mean loss over the global batch)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8), torch.nn.LayerNorm(8))
    ref.load_state_dict(net.state_dict())
    ref(xs).square().mean().backward()
    err = max(float((a - b.grad).abs().max()) for a, b in zip(got, ref.parameters()))
    # gradient accumulation: two micro-batches per rank, the first under no_sync(); same global gradient
    red.zero_grad()
    x = xs[rank * 4:(rank + 1) * 4]
    with red.no_sync():
        (net(x[:2]).square().mean() * 0.5).backward()
    (net(x[2:]).square().mean() * 0.5).backward()
    red.finish()
    err2 = max(float((p.grad - b.grad).abs().max()) for p, b in zip(net.parameters(), ref.parameters()))
    q.put((rank, max(err, err2)))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_process_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(err < 1e-6 for _, err in res), res
