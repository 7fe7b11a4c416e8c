"""Whole-step CUDA graph: forward + loss + backward (+ gradient all-reduce) + optimizer step captured once and
replayed, so the ~330 kernel launches of a ViT-B/16 training step cost one host call.

Replaces the reference's host loop (Lightning automatic optimisation around training_step,
model/CE/createViTmodel.py:68-77), whose per-module Python overhead is comparable to a B200 step.
Every libvitseg entry point is capture-safe (no allocation, no host sync, stream-ordered memsets only)."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    """step_fn(batch, idx) -> loss tensor must run the full optimisation step (e.g. DataParallel.step or a closure
    doing training_step / backward / optimizer.step / zero_grad) using only the tensors of `batch`.

    The optimizer must be capture-safe (torch.optim.Adam/AdamW(..., fused=True, capturable=True))."""

    def __init__(self, step_fn: Callable, example_batch: Sequence[torch.Tensor], warmup: int = 3, engines=()):
        self.step_fn = step_fn
        self.engines = tuple(engines)  # their bf16 weight shadows are one optimizer step behind after a replay
        self.static_batch = tuple(t.clone() for t in example_batch)
        dev = self.static_batch[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(warmup):
                self.step_fn(self.static_batch, i)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.static_loss = self.step_fn(self.static_batch, 0)
        self.steps = 0

    def load(self, batch: Sequence[torch.Tensor], non_blocking: bool = True) -> None:
        """copies a (host or device) batch into the graph's static input tensors on the current stream."""
        for s, t in zip(self.static_batch, batch):
            s.copy_(t, non_blocking=non_blocking)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        self.steps += 1
        for eng in self.engines:
            eng._dirty = True  # force a re-cast of the weight shadows on the next eager forward
        return self.static_loss

    def __call__(self, batch: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        return self.replay()


class GraphedInference:
    """fn(x) -> tensor (e.g. a LightningViTModel in eval mode, or ViTSegmentationModel.predict_mask) captured as one
    CUDA graph for a fixed input shape: the ~100 launches of a ViT-B/16 forward cost ~3 ms of Python/ctypes enqueue
    time eagerly, about as much as the 3.5 ms the GPU needs for a batch of 64 — the Celery-worker path replays the
    graph instead.  The weights are read at replay time (bf16 shadows must be current: call after load_state_dict /
    the last optimizer step and a first eager forward, which this constructor performs)."""

    def __init__(self, fn: Callable, example: torch.Tensor, warmup: int = 2):
        self.fn = fn
        self.static_in = example.clone()
        dev = example.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.static_out = self.fn(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """returns the graph's static output tensor (overwritten by the next call)."""
        if x.shape != self.static_in.shape:
            raise ValueError(f"GraphedInference was captured for {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
