"""FusedAdam — torch.optim.Adam / AdamW semantics executed as ONE kernel over the engine's flat arenas
(vs_adam_step): parameter update + bf16 weight shadow + gradient zeroing in a single HBM pass.

Replaces `torch.optim.Adam(self.parameters(), lr=1e-5)` (model/CE/classes.py:296-297),
`Adam(lr=1e-4)` (model/PAED/classes.py:486-487) and `AdamW(lr=1e-4)` (model/PAED/classes.py:536-548) for users
who opt in; the stock torch optimizers keep working (the engine then re-casts the shadow itself)."""
from __future__ import annotations

import torch

from . import kernels as K


def _engine_of(module):
    m = module.model if hasattr(module, "model") and hasattr(module.model, "engine") else module
    return m.engine


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        self.engine = _engine_of(module)
        params = [p for p in module.parameters() if p.requires_grad]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled_weight_decay)
        super().__init__(params, defaults)
        self._m = self._v = None
        self._step = None
        self._lr_dev = None
        self._lr_host = torch.zeros(1, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None

    def _ensure_state(self):
        eng = self.engine
        dev = next(eng.module.parameters()).device
        eng.ensure_packed(dev)
        if self._m is None or self._m.data_ptr() == 0 or self._m.numel() != eng.master.numel() or self._m.device != dev:
            self._m = torch.zeros_like(eng.master)
            self._v = torch.zeros_like(eng.master)
            self._step = torch.zeros(1, device=dev, dtype=torch.int32)
            self._lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_state()
        eng = self.engine
        g = self.param_groups[0]
        # lr through pinned host -> device copy: a captured graph re-reads the host value at every replay
        self._lr_host[0] = float(g["lr"])
        self._lr_dev.copy_(self._lr_host, non_blocking=True)
        self._step.add_(1)
        s0 = eng.slots["backbone.pooler.dense.weight"].offset
        sb = eng.slots["backbone.pooler.dense.bias"]
        s1 = sb.offset + (sb.numel + 63) // 64 * 64
        K.adam_step(eng.master, eng.grads, self._m, self._v, eng.shadow, self._lr_dev, self._step, g["betas"][0],
                    g["betas"][1], g["eps"], g["weight_decay"], g["decoupled"], 1.0, True, s0, s1)
        K.pack_conv3x3(eng.w32("seg_head.0.weight"), eng.head_w_packed)
        eng.note_optimizer_step(grads_zeroed=True)
        return loss

    def set_lr(self, lr: float):
        """Changes the learning rate; also effective for an already captured CUDA graph (the replayed host->device
        copy re-reads this pinned value)."""
        for group in self.param_groups:
            group["lr"] = float(lr)
        if self._lr_host is not None:
            self._lr_host[0] = float(lr)

    def zero_grad(self, set_to_none: bool = True):
        # the arena was zeroed by the update kernel; dropping the views makes the next backward re-attach them
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None


class FusedAdamW(FusedAdam):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(module, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled_weight_decay=True)
