"""CUDA-graph replay of a tokenizer call (serving path).

A tokenizer step is ~10 kernel launches through Python custom-op wrappers; at 1.3 ms of GPU work per step the
host side (dispatch, ctypes, tensor allocation) is a comparable cost and makes end-to-end throughput depend on
host load.  `GraphedTokenizer` captures one call of a p3tok module at fixed shapes into a CUDA graph - every
kernel of libp3tok.so is capturable: launches go to the capturing stream, tensor maps are by-value kernel
parameters, nothing synchronises or allocates outside torch's graph-private pool - and replays it with a
single launch.  Inputs are written into static device buffers (`.inputs`), the result is read from `.output`.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedTokenizer:
    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 stream: "torch.cuda.Stream | None" = None):
        """fn(*inputs) -> tensor; example_inputs fix shapes/dtypes/device (their values seed the static buffers).
        stream: run `__call__` / `replay_on_stream` on this stream - two instances on two streams keep two steps in flight
        (the index kernels of one step overlap the embedding of the other), what bench.py's device loop does."""
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedTokenizer: CUDA example inputs required")
        self.device = example_inputs[0].device
        self.stream = stream
        self.inputs = [t.clone() for t in example_inputs]
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):            # kernel attributes / lazy init happen outside the capture
                fn(*self.inputs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.output = fn(*self.inputs)

    def replay(self) -> torch.Tensor:
        """Run on the current contents of `.inputs`; returns the static `.output` tensor (overwritten next replay)."""
        self.graph.replay()
        return self.output

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                for dst, src in zip(self.inputs, inputs):
                    dst.copy_(src, non_blocking=True)
                return self.replay()
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        return self.replay()


class GraphedHostTokenizer:
    """The whole serving step as ONE CUDA graph on its own stream: pinned host inputs -> device (memcpy nodes) -> the
    captured module call -> pinned host output.  A step then costs the host one graph launch; two instances replayed
    alternately on their two streams overlap the copies of one step with the kernels of the other (what bench.py's `e2e`
    measures).  `host_inputs` are read at every replay (refill them in place); `host_output` is overwritten by the next
    replay of the same instance - call `synchronize()` (or wait on `done`) before reading it."""

    def __init__(self, fn: Callable[..., torch.Tensor], host_inputs: Sequence[torch.Tensor], device, warmup: int = 3):
        self.device = torch.device(device)
        self.host_inputs = [t if t.is_pinned() else t.pin_memory() for t in host_inputs]
        self.inputs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in self.host_inputs]
        self.stream = torch.cuda.Stream(device=self.device)
        self.graph = torch.cuda.CUDAGraph()
        self.done = torch.cuda.Event()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(max(1, warmup)):            # lazy initialisation and the output shape, outside the capture
                for d, h in zip(self.inputs, self.host_inputs):
                    d.copy_(h, non_blocking=True)
                out = fn(*self.inputs)
        self.stream.synchronize()
        self.host_output = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=self.stream):
            for d, h in zip(self.inputs, self.host_inputs):
                d.copy_(h, non_blocking=True)
            self.output = fn(*self.inputs)
            self.host_output.copy_(self.output, non_blocking=True)

    def replay(self) -> None:
        """Enqueue one step on this instance's stream (asynchronous)."""
        with torch.cuda.stream(self.stream):
            self.graph.replay()
            self.done.record()

    def synchronize(self) -> torch.Tensor:
        self.stream.synchronize()
        return self.host_output
