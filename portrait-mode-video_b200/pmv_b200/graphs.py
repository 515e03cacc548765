"""CUDA-graph capture of a whole step (forward, or forward + backward + gradient all-reduce + optimizer).

The path issues ~750 kernel launches per MViTv2-S training step through ctypes; at ~35 us of host work each the CPU,
not the GPU, sets the step time once the kernels are fast (measured: 29.6 ms of host enqueue for 21.8 ms of device
work).  Capturing the step removes that: every launch of libpmv_b200.so goes to ``torch.cuda.current_stream()``, the
library allocates nothing and never synchronises, tensor maps travel as kernel parameters, and the workspaces come
from torch's graph-private memory pool, so the captured addresses stay valid across replays.

Usage:
    step = GraphedStep(fn, example_inputs)      # fn(*static_inputs) -> tensor (loss / logits)
    out = step(*inputs)                         # copies the inputs into the static buffers, replays, returns the
                                                # static output tensor (valid until the next call)
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.fn = fn
        self.static_inputs = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up off the default stream: lazy initialisation (func attributes, caches)
            for _ in range(warmup):
                fn(*self.static_inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_output = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        for s, t in zip(self.static_inputs, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_output
