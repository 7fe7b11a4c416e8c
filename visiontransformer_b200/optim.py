"""FusedAdam — torch.optim.Adam / AdamW semantics executed as ONE kernel over the engine's flat arenas
(vs_adam_step): parameter update + bf16 weight shadow + gradient zeroing in a single HBM pass.

Replaces `torch.optim.Adam(self.parameters(), lr=1e-5)` (model/CE/classes.py:296-297),
`Adam(lr=1e-4)` (model/PAED/classes.py:486-487) and `AdamW(lr=1e-4)` (model/PAED/classes.py:536-548) for users
who opt in; the stock torch optimizers keep working (the engine then re-casts the shadow itself).

Checkpointing: `state_dict()` / `load_state_dict()` use torch.optim.Adam's own layout (per-parameter `step`,
`exp_avg`, `exp_avg_sq`), so a Lightning checkpoint written with FusedAdam resumes under torch Adam and vice versa.
The per-parameter tensors in `self.state` are VIEWS into the flat moment arenas the kernel updates."""
from __future__ import annotations

import torch

from . import kernels as K


def _engine_of(module):
    m = module.model if hasattr(module, "model") and hasattr(module.model, "engine") else module
    return m.engine


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        self.engine = _engine_of(module)
        self._root = self.engine.module
        params = [p for p in self._root.parameters() if p.requires_grad]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled_weight_decay)
        super().__init__(params, defaults)
        self._m = self._v = None
        self._step = None
        self._lr_dev = None
        self._ranges = None
        self._lr_host = torch.zeros(1, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None

    # ------------------------------------------------------------------------------------------ state arenas
    def _slot_of(self):
        """id(parameter) -> arena slot, for the parameters this optimizer owns."""
        eng = self.engine
        by_id = {id(p): n for n, p in self._root.named_parameters()}
        return {id(p): eng.slots[by_id[id(p)]] for g in self.param_groups for p in g["params"]}

    def _trainable_ranges(self):
        """contiguous [begin, end) element ranges of the arenas that this optimizer updates: every owned parameter that
        can receive a gradient (the pooler never does: dead compute in the reference forward, TF:456), merged when
        adjacent.  Frozen parameters (requires_grad=False) fall outside every range: no update, no weight decay."""
        eng = self.engine
        owned = {id(p) for g in self.param_groups for p in g["params"] if p.requires_grad}
        spans = []
        for n, p in self._root.named_parameters():
            if id(p) not in owned or n.startswith("backbone.pooler."):
                continue
            s = eng.slots[n]
            spans.append((s.offset, s.offset + (s.numel + 63) // 64 * 64))
        spans.sort()
        merged = []
        for b, e in spans:
            if merged and merged[-1][1] == b:
                merged[-1][1] = e
            else:
                merged.append([b, e])
        return [(b, e) for b, e in merged]

    def _ensure_state(self):
        eng = self.engine
        dev = next(self._root.parameters()).device
        eng.ensure_packed(dev)
        n = eng.master.numel()
        if self._m is None or self._m.numel() != n:
            self._m = torch.zeros(n, device=dev, dtype=torch.float32)
            self._v = torch.zeros(n, device=dev, dtype=torch.float32)
            self._step = torch.zeros(1, device=dev, dtype=torch.int32)
            self._lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)
            self._alias_state()
        elif self._m.device != dev:
            # the module moved: the moments move with it (they are optimizer state, not scratch)
            self._m, self._v = self._m.to(dev), self._v.to(dev)
            self._step, self._lr_dev = self._step.to(dev), self._lr_dev.to(dev)
            self._alias_state()
        self._ranges = self._trainable_ranges()

    def _alias_state(self):
        """self.state[p] = {step, exp_avg, exp_avg_sq} as views of the arenas (torch.optim.Adam's layout)."""
        slots = self._slot_of()
        names = {id(p): n for n, p in self._root.named_parameters()}
        for g in self.param_groups:
            for p in g["params"]:
                if names[id(p)].startswith("backbone.pooler.") or not p.requires_grad:
                    continue   # torch.optim.Adam creates no state for parameters that never receive a gradient
                s = slots[id(p)]
                self.state[p] = {
                    "step": self._step.view(()),
                    "exp_avg": self._m[s.offset:s.offset + s.numel].view(s.shape),
                    "exp_avg_sq": self._v[s.offset:s.offset + s.numel].view(s.shape),
                }

    def state_dict(self):
        sd = super().state_dict()
        # torch.optim.Adam stores `step` as a float tensor; emit that so its load_state_dict accepts the checkpoint
        for st in sd["state"].values():
            if "step" in st:
                st["step"] = st["step"].detach().to(device="cpu", dtype=torch.float32)
                st["exp_avg"] = st["exp_avg"].detach().clone()
                st["exp_avg_sq"] = st["exp_avg_sq"].detach().clone()
        return sd

    def load_state_dict(self, state_dict):
        """accepts FusedAdam's and torch.optim.Adam / AdamW's state dicts (same parameter order): moments are copied
        into the flat arenas, the common step count into the device counter."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:   # a torch.optim.Adam checkpoint has no "decoupled" key (and extra ones)
            for k, v in self.defaults.items():
                group.setdefault(k, v)
        loaded = {p: dict(st) for p, st in self.state.items()}
        self.state.clear()
        self._m = None
        self._ensure_state()
        steps = set()
        with torch.no_grad():
            for p, st in loaded.items():
                if p not in self.state:
                    continue
                mine = self.state[p]
                mine["exp_avg"].copy_(st["exp_avg"])
                mine["exp_avg_sq"].copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError(f"FusedAdam keeps ONE step counter for all parameters; checkpoint has {sorted(steps)}")
            if steps:
                self._step.fill_(steps.pop())

    # ------------------------------------------------------------------------------------------ update
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
        n = eng.master.numel()
        args = (g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], g["decoupled"], 1.0, True)
        s0 = eng.slots["backbone.pooler.dense.weight"].offset
        sb = eng.slots["backbone.pooler.dense.bias"]
        s1 = sb.offset + (sb.numel + 63) // 64 * 64
        if self._ranges == [(0, s0), (s1, n)] or self._ranges == [(0, n)]:
            # the usual case (everything trains): one launch over the whole arena with the pooler hole
            K.adam_step(eng.master, eng.grads, self._m, self._v, eng.shadow, self._lr_dev, self._step, *args, s0, s1)
        else:
            # frozen parameters: one launch per contiguous trainable range (their weights, moments and bf16 shadow
            # stay untouched; the engine clears the whole gradient arena itself when anything is frozen)
            for b, e in self._ranges:
                K.adam_step(eng.master[b:e], eng.grads[b:e], self._m[b:e], self._v[b:e], eng.shadow[b:e], self._lr_dev,
                            self._step, *args, 0, 0)
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
