"""Data-parallel training / sharded inference for the ViT-segmentation path: one process per GPU
(torchrun), full replica per rank, batch sharded by image.

The reference has no distributed code at all (devices=1 everywhere, model/CE/createViTmodel.py:72-73); this is the
net-new exchange step: ONE collective per optimizer step — a bucketed all-reduce of the flat fp32 gradient arena over
NCCL (NVLink 5 / NVSwitch), issued bucket by bucket from inside backward so it overlaps the remaining backward work —
plus a 6-scalar all-reduce inside the PAEDTrainer loss for exact global-batch Dice / |PAED| semantics.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class BucketReducer:
    """All-reduces contiguous slices ("buckets") of a flat gradient tensor as soon as each becomes final.

    Device-agnostic host logic (tested with gloo on CPU): `ready(name)` launches the asynchronous all-reduce of that
    bucket, `finish()` waits for all of them and applies the 1/world scaling when `average` is set."""

    def __init__(self, flat: torch.Tensor, ranges: Sequence[Tuple[str, int, int]], group=None, average: bool = True):
        self.flat = flat
        self.ranges = {n: (s, e) for n, s, e in ranges}
        self.order = [n for n, _, _ in ranges]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        self._works: List = []
        self._done: List[str] = []
        backend = dist.get_backend(group) if dist.is_initialized() else ""
        self._native_avg = backend == "nccl"

    def ready(self, name: str):
        if self.world == 1:
            return
        s, e = self.ranges[name]
        if e <= s:
            return
        view = self.flat[s:e]
        op = dist.ReduceOp.AVG if (self.average and self._native_avg) else dist.ReduceOp.SUM
        self._works.append(dist.all_reduce(view, op=op, group=self.group, async_op=True))
        self._done.append(name)

    def finish(self):
        if self.world == 1:
            return
        for w in self._works:
            w.wait()
        if self.average and not self._native_avg:
            for n in self._done:
                s, e = self.ranges[n]
                self.flat[s:e].mul_(1.0 / self.world)
        missing = [n for n in self.order if n not in self._done and self.ranges[n][1] > self.ranges[n][0]]
        self._works, self._done = [], []
        if missing:
            raise RuntimeError(f"gradient buckets never reduced: {missing}")


class MultimemReducer:
    """Same interface as BucketReducer, but the reduction happens in the NVSwitch (NVLS) through the library's own
    kernel (csrc/collective.cu) instead of NCCL: the gradient arena lives in symmetric memory; when a bucket is
    final the communication stream waits for it, passes a cross-GPU barrier, runs vs_multimem_allreduce_f32 on the
    bucket (one small CTA per SM, resident NEXT TO the backward GEMM CTAs rather than displacing them) and passes a
    second barrier.  Requires NVLink multicast support (NVSwitch systems); DataParallel falls back to NCCL otherwise."""

    def __init__(self, engine, group=None, average: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.average = average
        dev = engine.grads.device
        n = engine.grads.numel()
        self.buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.mc_ptr = int(self.hdl.multicast_ptr)
        if self.mc_ptr == 0:
            raise RuntimeError("NVLink multicast (NVLS) is not available: the symmetric-memory rendezvous returned no "
                               "multicast address")
        engine.set_grad_arena(self.buf)
        self.flat = self.buf
        self.ranges = {name: (s, e) for name, s, e in engine.bucket_ranges()}
        self.order = [name for name, _, _ in engine.bucket_ranges()]
        self.stream = torch.cuda.Stream(device=dev)
        self._done: List[str] = []

    def ready(self, name: str):
        s, e = self.ranges[name]
        if e <= s:
            return
        from . import kernels as K
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)          # this rank's bucket is final
            self.hdl.barrier(channel=0)         # ... and so is every other rank's
            K.multimem_allreduce(self.mc_ptr + 4 * s, e - s, self.rank, self.world,
                                 1.0 / self.world if self.average else 1.0)
            self.hdl.barrier(channel=0)         # every slice of the bucket has been written back everywhere
        self._done.append(name)

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.stream)
        missing = [n for n in self.order if n not in self._done and self.ranges[n][1] > self.ranges[n][0]]
        self._done = []
        if missing:
            raise RuntimeError(f"gradient buckets never reduced: {missing}")


def shard_batch(batch, rank: int, world: int):
    """Even split of every tensor of a batch along dim 0 (images are independent: no BatchNorm, per-token LN)."""
    out = []
    for t in batch:
        n = t.shape[0]
        if n % world != 0:
            raise ValueError(f"global batch {n} is not divisible by world size {world}")
        per = n // world
        out.append(t[rank * per:(rank + 1) * per])
    return tuple(out)


class DataParallel:
    """Wraps a LightningViTModel / PAEDTrainer-shaped module (anything with `.model` = ViTSegmentationModel and a
    `training_step(batch, idx)` returning the loss).

    step(batch): forward, backward with overlapped bucketed gradient all-reduce, optimizer step.
      * losses that are means over the local shard (CE, multi-class PAED): gradients are averaged over ranks;
      * PAEDTrainer: its loss already is the global-batch loss (cross-rank sums inside), so gradients are summed."""

    def __init__(self, module, optimizer: Optional[torch.optim.Optimizer] = None, group=None):
        self.module = module
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.optimizer = optimizer
        self._global_loss = hasattr(module, "dp_world_size")
        if self._global_loss:
            module.dp_group = group
            module.dp_world_size = self.world
        self._reducer = None
        self._synced = False
        # gradient exchange: "multimem" = in-switch reduction by the library's own co-resident kernel (NVLS), "nccl" =
        # bucketed dist.all_reduce; "auto" tries multimem on NCCL/CUDA process groups and falls back to NCCL
        self.reduce_mode = os.environ.get("VS_DP_REDUCE", "auto")
        self.reduce_backend = None
        self.reduce_fallback_reason = None

    def _engine(self):
        return self.module.model.engine

    def broadcast_parameters(self, src: int = 0):
        """Makes every replica start from rank `src`'s weights (one flat broadcast of the master arena)."""
        eng = self._engine()
        dev = next(self.module.parameters()).device
        eng.ensure_packed(dev)
        if self.world > 1:
            dist.broadcast(eng.master, src=src, group=self.group)
        eng._dirty = True
        self._synced = True

    def _make_reducer(self, eng):
        backend = dist.get_backend(self.group) if dist.is_initialized() else ""
        if self.world > 1 and self.reduce_mode in ("auto", "multimem") and backend == "nccl" and eng.grads.is_cuda:
            try:
                r = MultimemReducer(eng, self.group, average=not self._global_loss)
                self.reduce_backend = "multimem"
                return r
            except Exception as e:  # noqa: BLE001  (no NVSwitch multicast, old driver, ...)
                if self.reduce_mode == "multimem":
                    raise
                self.reduce_fallback_reason = f"{type(e).__name__}: {e}"
        self.reduce_backend = "nccl" if backend == "nccl" else backend
        return BucketReducer(eng.grads, eng.bucket_ranges(), self.group, average=not self._global_loss)

    def _attach(self):
        eng = self._engine()
        if self._reducer is None or self._reducer.flat.data_ptr() != eng.grads.data_ptr():
            self._reducer = self._make_reducer(eng)
        eng.grad_ready_hook = self._reducer.ready if self.world > 1 else None

    def step(self, batch, batch_idx: int = 0, sync_grads: bool = True):
        """One optimisation step on this rank's shard.  sync_grads=False skips the all-reduce (gradient
        accumulation micro-batches before the last one, model/CE/createViTmodel.py:74)."""
        if not self._synced:
            self.broadcast_parameters()
        loss = self.module.training_step(batch, batch_idx)
        eng = self._engine()
        if sync_grads:
            self._attach()
        else:
            eng.grad_ready_hook = None
        loss.backward()
        if sync_grads and self.world > 1:
            self._reducer.finish()
        eng.finish_foreign_grads()
        eng.grad_ready_hook = None
        if sync_grads and self.optimizer is not None:
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
        return loss.detach()

    @torch.no_grad()
    def predict_masks(self, images: torch.Tensor, gather: bool = False):
        """Sharded inference: each rank runs predict_mask on its slice; replicas only, no data-path collective.
        gather=True all-gathers the uint8 masks (256*512*512 B = 67 MB for BASELINE config 5)."""
        (local,) = shard_batch((images,), self.rank, self.world)
        masks = self.module.model.predict_mask(local)
        if not gather or self.world == 1:
            return masks
        out = [torch.empty_like(masks) for _ in range(self.world)]
        dist.all_gather(out, masks, group=self.group)
        return torch.cat(out, 0)
