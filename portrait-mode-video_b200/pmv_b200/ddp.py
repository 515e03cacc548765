"""Batch-sharded data parallelism: bucketed gradient all-reduce overlapped with backward.

Mirrors what the reference gets from ``torch.nn.parallel.DistributedDataParallel`` in
models/build.py:71-79 (bucketed all-reduce launched as gradients become ready, mean over ranks,
find_unused_parameters=False) and the optional fp16-compressed hook of build.py:80-83 (here: bf16).

Design: parameters are grouped, in reverse registration order (the order backward produces gradients), into
buckets with a flat fp32 staging buffer.  A post-accumulate-grad hook counts down the bucket; when its last gradient
has been produced the bucket is packed (one multi-tensor copy) and all-reduced on a dedicated communication stream
(NCCL over NVLink / NVSwitch) while backward keeps running on the compute stream.  ``finish()`` makes the compute stream wait for the reductions.
Inference needs no collective (pure batch partitioning, tools/test_net.py:131-132 gathers logits only).
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.distributed as dist


class _Bucket:
    ALIGN = 64  # floats: every slot starts on a 256-byte boundary (the wgrad kernels write their slot with 16-byte vector
                # reductions / bulk tensor stores when it is the parameter's gradient home)

    def __init__(self, params, device, compress_dtype, need_flat):
        self.params = params
        A = self.ALIGN
        n = sum((p.numel() + A - 1) // A * A for p in params)
        # flat fp32 staging buffer of the bucket (only needed when there is something to reduce)
        self.flat = torch.zeros(n, dtype=torch.float32, device=device) if need_flat else None
        self.views = []
        if need_flat:
            off = 0
            for p in params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += (p.numel() + A - 1) // A * A
        self.pending = len(params)
        self.work = None
        self.avg_done = False
        self.event = torch.cuda.Event() if device.type == "cuda" else None
        self.compressed = torch.empty(n, dtype=compress_dtype, device=device) if (compress_dtype is not None and need_flat) else None


class GradAllReducer:
    """Gradients are NOT pre-allocated views: ``zero_grad()`` sets ``p.grad = None`` so autograd adopts each freshly
    computed gradient tensor as is (no ``grad += dW`` kernel and no zero-fill per parameter: 397 + 200 launches per
    MViTv2-S step otherwise).  When the last gradient of a bucket has arrived the bucket is packed into its flat
    buffer with one multi-tensor copy on the communication stream and all-reduced there; ``finish()`` re-points
    ``p.grad`` at the reduced flat views.  With one rank nothing is copied at all."""

    def __init__(self, module: torch.nn.Module, bucket_mb: float = 25.0, compress_dtype=None, process_group=None,
                 broadcast_init: bool = True, last_bucket_mb: float = None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, "no trainable parameters"
        self.params = params
        self.device = params[0].device
        if self.world > 1 and broadcast_init:
            # DistributedDataParallel broadcasts rank 0's parameters and buffers at construction (build.py:71-79 relies on
            # it: only rank 0 loads the checkpoint / draws the initial weights that count)
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        self._sync = True
        if self.device.type == "cuda":
            from . import ops
            self.arena = ops.attach_arena(params)
        else:
            self.arena = None
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.buckets: List[_Bucket] = []
        # Buckets in backward order.  ``last_bucket_mb``: the gradients of the FIRST layers arrive last and nothing is left
        # to hide their all-reduce behind, so they get a small bucket of their own (its reduction is what the step
        # waits for after backward; profiles/r02_ddp_timeline.md)
        order = list(reversed(params))
        tail = []
        if last_bucket_mb is not None and last_bucket_mb > 0:
            tb = 0
            while order and tb + order[-1].numel() * 4 <= int(last_bucket_mb * 2 ** 20):
                tb += order[-1].numel() * 4
                tail.insert(0, order.pop())
        cur, cur_bytes, limit = [], 0, int(bucket_mb * 2 ** 20)
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= limit:
                self.buckets.append(_Bucket(cur, self.device, compress_dtype, self.world > 1))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, self.device, compress_dtype, self.world > 1))
        if tail:
            self.buckets.append(_Bucket(tail, self.device, compress_dtype, self.world > 1))
        self._owner = {}
        for b in self.buckets:
            for i, p in enumerate(b.params):
                self._owner[p] = b
                p.register_post_accumulate_grad_hook(self._hook)
                if b.views and self.device.type == "cuda" and p.dim() == 2:
                    # matrix weights (99 % of the bytes): the split-K wgrad kernels write straight into the bucket, so the
                    # pack copy below only moves the small tensors (the timeline of round 2 showed 0.23 ms of pack kernels
                    # per step, profiles/r02_ddp_timeline.md)
                    p._pmv_grad_home = b.views[i]

    @property
    def num_buckets(self):
        return len(self.buckets)

    def zero_grad(self):
        if self.arena is not None:
            self.arena.reset(self.device)  # one fill for all split-K weight gradients of the coming step
        for b in self.buckets:
            for p in b.params:
                p.grad = None
            if b.flat is not None and self.device.type == "cuda":
                b.flat.zero_()  # the gradient homes accumulate split-K partial products
        self._rearm()

    def _rearm(self):
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation (DistributedDataParallel.no_sync): backward passes inside the context only accumulate
        into ``p.grad``; the first backward outside it reduces the accumulated gradients."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _hook(self, p):
        if not self._sync:
            return
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0 and self.world > 1:
            self._launch(b)

    def _launch(self, b: _Bucket):
        if self.stream is not None:
            b.event.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(b.event)
                self._reduce(b)
        else:
            self._reduce(b)

    def _reduce(self, b: _Bucket):
        # pack: one multi-tensor copy of the gradients that were not written into the bucket by their producer
        dst, src = [], []
        for v, p in zip(b.views, b.params):
            if p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, src)
        buf = b.flat
        if b.compressed is not None:
            b.compressed.copy_(b.flat)
            buf = b.compressed
        avg = dist.ReduceOp.AVG if self.device.type == "cuda" else dist.ReduceOp.SUM
        b.work = dist.all_reduce(buf, op=avg, group=self.group, async_op=True)
        b.avg_done = self.device.type == "cuda"

    def finish(self):
        """Call after backward(): waits for every bucket reduction; gradients are then the mean over ranks."""
        if self.world == 1:
            self._rearm()
            return
        for b in self.buckets:
            assert b.pending == 0, "a parameter received no gradient (find_unused_parameters=False semantics)"
            if self.stream is not None:
                with torch.cuda.stream(self.stream):
                    b.work.wait()
                    if b.compressed is not None:
                        b.flat.copy_(b.compressed)
            else:
                b.work.wait()
                if b.compressed is not None:
                    b.flat.copy_(b.compressed)
                if not b.avg_done:
                    b.flat.div_(self.world)
            for p, v in zip(b.params, b.views):
                p.grad = v
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self._rearm()  # a further backward before zero_grad() (accumulation) counts the buckets down again
