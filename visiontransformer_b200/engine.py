"""Host-side execution engine of the ViT-segmentation hot path: owns the packed weights, the activation
workspace and the sequence of libvitseg kernel launches for forward and backward.

Reference call stack being replaced: ViTSegmentationModel.forward (model/CE/classes.py:246-262)
-> transformers ViTModel.forward (TF:428-458) -> ViTEmbeddings / ViTLayer x L / LayerNorm -> seg_head -> upsample.

Data layout in HBM
  * fp32 master parameters live in ONE flat arena; every nn.Parameter of the module is a view into it, laid out so
    that query/key/value weights (and biases) of a layer are adjacent: the fused [3D, D] QKV GEMM reads them in place.
  * a bf16 shadow arena with the same offsets feeds the tensor cores; it is refreshed by one cast kernel whenever a
    parameter version changes (optimizer step / load_state_dict).  The 3x3 head conv gets an extra packed
    [(ky,kx,c)]-ordered bf16 copy.
  * gradients accumulate into a flat fp32 arena with the same offsets (param.grad are views), which is what the
    data-parallel all-reduce buckets slice.
  * residual stream fp32 [B*(T+1), D]; GEMM operands bf16; per-layer activations saved for backward, no recompute.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed

from . import kernels as K

BF16, F32 = torch.bfloat16, torch.float32
LN_EPS = 1e-12  # ViTConfig.layer_norm_eps default used by the reference (TF:325-326)
HEAD_CH = 256   # model/CE/classes.py:241


class _Slot:
    __slots__ = ("name", "offset", "numel", "shape")

    def __init__(self, name, offset, numel, shape):
        self.name, self.offset, self.numel, self.shape = name, offset, numel, shape


def param_order(num_layers: int) -> List[str]:
    """Arena order: embeddings, then per layer (ln1, q, k, v, out-proj, ln2, fc1, fc2), final LN, pooler, head.
    Weights of q/k/v are adjacent, then their biases adjacent, so both fuse without copies."""
    names = [
        "backbone.embeddings.cls_token",
        "backbone.embeddings.position_embeddings",
        "backbone.embeddings.patch_embeddings.projection.weight",
        "backbone.embeddings.patch_embeddings.projection.bias",
    ]
    for i in range(num_layers):
        p = f"backbone.encoder.layer.{i}."
        names += [
            p + "layernorm_before.weight", p + "layernorm_before.bias",
            p + "attention.attention.query.weight", p + "attention.attention.key.weight",
            p + "attention.attention.value.weight",
            p + "attention.attention.query.bias", p + "attention.attention.key.bias",
            p + "attention.attention.value.bias",
            p + "attention.output.dense.weight", p + "attention.output.dense.bias",
            p + "layernorm_after.weight", p + "layernorm_after.bias",
            p + "intermediate.dense.weight", p + "intermediate.dense.bias",
            p + "output.dense.weight", p + "output.dense.bias",
        ]
    names += [
        "backbone.layernorm.weight", "backbone.layernorm.bias",
        "backbone.pooler.dense.weight", "backbone.pooler.dense.bias",
        "seg_head.0.weight", "seg_head.0.bias", "seg_head.2.weight", "seg_head.2.bias",
    ]
    return names


class Engine:
    def __init__(self, module, cfg):
        self.module = module
        self.cfg = cfg
        self.device: Optional[torch.device] = None
        self.slots: Dict[str, _Slot] = {}
        self.master: Optional[torch.Tensor] = None
        self.shadow: Optional[torch.Tensor] = None
        self.grads: Optional[torch.Tensor] = None
        self.head_w_packed: Optional[torch.Tensor] = None
        self._versions = None
        self._dirty = True
        self._shadow_fresh = False   # set by FusedAdam: shadow already matches the master arena
        self._grads_zeroed = False   # set by FusedAdam: gradient arena already zero
        self._ws: Dict = {}
        self._saved = None
        self._foreign = []
        self.grad_ready_hook = None  # callable(bucket_name) used by the data-parallel wrapper
        self.rng_step = None         # device uint32 counter: advances once per dropout-enabled forward (graph-safe)
        self._drop = (0.0, 0.0)      # (hidden p, attention p) used by the last training forward
        self.launches = 0            # kernel launches issued by the last forward/backward (for bench bookkeeping)
        self._two_streams = os.environ.get("VS_BWD_STREAMS", "1") == "2"   # opt-in: measured neutral (11.03-11.10 vs 11.05 ms/step, r02)
        self._side = None
        self._fuse_colsum = os.environ.get("VS_FUSE_COLSUM", "1") != "0"   # A/B switch: 0 = separate column-sum pass
        self._attn_bwd_short = os.environ.get("VS_ATTN_BWD", "short") != "blocks"   # A/B switch: key-block kernel always

    # ------------------------------------------------------------------------------------------ parameters
    def _named(self):
        return dict(self.module.named_parameters())

    def ensure_packed(self, device: torch.device):
        """(Re)build the flat arenas when the module moved device or its parameters were re-allocated."""
        params = self._named()
        names = param_order(self.cfg.num_hidden_layers)
        assert set(names) == set(params.keys()), "parameter tree does not match the expected ViT layout"
        ok = self.master is not None and self.device == device
        if ok:
            base = self.master.data_ptr()
            for n in names:
                s = self.slots[n]
                if params[n].data_ptr() != base + s.offset * 4:
                    ok = False
                    break
        if ok:
            return
        if device.type != "cuda":
            raise RuntimeError("visiontransformer_b200: parameters must live on a CUDA device (no CPU fallback)")
        self.device = device
        off = 0
        self.slots = {}
        for n in names:
            p = params[n]
            self.slots[n] = _Slot(n, off, p.numel(), tuple(p.shape))
            off += (p.numel() + 63) // 64 * 64  # 256-byte aligned slots (TMA needs 16 B; cast kernel 16 B)
        total = off
        master = torch.zeros(total, device=device, dtype=F32)
        grads = torch.zeros(total, device=device, dtype=F32)
        with torch.no_grad():
            for n in names:
                s = self.slots[n]
                p = params[n]
                view = master[s.offset:s.offset + s.numel].view(s.shape)
                view.copy_(p.data.to(device=device, dtype=F32))
                old_grad = p.grad
                p.data = view
                if old_grad is not None:
                    gview = grads[s.offset:s.offset + s.numel].view(s.shape)
                    gview.copy_(old_grad.to(device=device, dtype=F32))
                    p.grad = gview
        self.master, self.grads = master, grads
        self.shadow = torch.empty(total, device=device, dtype=BF16)
        D = self.cfg.hidden_size
        self.head_w_packed = torch.empty(HEAD_CH, 9 * D, device=device, dtype=BF16)
        self.head_wgrad_packed = torch.zeros(HEAD_CH, 9 * D, device=device, dtype=F32)
        self._versions = None
        self._ws = {}

    def _side_stream(self):
        if self._side is None or self._side.device != self.device:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def set_grad_arena(self, buf: torch.Tensor) -> None:
        """Re-homes the flat gradient arena into `buf` (same size, fp32, same device) — used by the data-parallel
        wrapper to place it in NVLink symmetric memory.  param.grad views are dropped and re-attached by the next
        backward."""
        assert buf.dtype == F32 and buf.numel() == self.grads.numel() and buf.device == self.grads.device
        buf.copy_(self.grads)
        self.grads = buf
        for p in self.module.parameters():
            p.grad = None
        self._grads_zeroed = False

    def refresh_shadow(self, train: bool = False):
        """fp32 master arena -> bf16 shadow (+ packed head conv weight).

        Parameter versions catch load_state_dict / in-place edits, but fused optimizers (torch.optim.Adam(fused=True),
        multi-tensor kernels) update parameters WITHOUT bumping Tensor._version, so: every training-mode forward
        re-casts (the weights changed since the previous step), and an inference forward re-casts when the versions
        changed or a training forward happened since the last cast.  One pass over the arena: ~90 us for ViT-B/16."""
        params = self._named()
        versions = tuple(p._version for p in params.values())
        if self._shadow_fresh and versions == self._versions:
            # FusedAdam wrote the shadow (and re-packed the head weight) in its own pass
            self._shadow_fresh = False
            self._dirty = False
            return
        self._shadow_fresh = False
        if not train and not self._dirty and versions == self._versions:
            return
        self._dirty = train   # an optimizer step is expected to follow a training forward
        K.cast_bf16(self.master, self.shadow)
        K.pack_conv3x3(self.w32("seg_head.0.weight"), self.head_w_packed)
        self.launches += 2
        self._versions = versions

    def note_optimizer_step(self, grads_zeroed: bool):
        """called by visiontransformer_b200.optim.FusedAdam after its fused update pass."""
        self._shadow_fresh = True
        self._grads_zeroed = grads_zeroed
        self._versions = tuple(p._version for p in self._named().values())

    def w32(self, name):
        s = self.slots[name]
        return self.master[s.offset:s.offset + s.numel].view(s.shape)

    def w16(self, name, rows=None, cols=None):
        s = self.slots[name]
        t = self.shadow[s.offset:s.offset + s.numel]
        if rows is not None:
            return t.view(rows, cols)
        return t.view(s.shape)

    def g32(self, name, shape=None):
        s = self.slots[name]
        return self.grads[s.offset:s.offset + s.numel].view(shape if shape is not None else s.shape)

    def fused_qkv(self, i, arena="shadow"):
        """[3D, D] weight and [3D] bias of layer i as single views (q,k,v slots are adjacent and unpadded)."""
        D = self.cfg.hidden_size
        p = f"backbone.encoder.layer.{i}.attention.attention."
        sw, sb = self.slots[p + "query.weight"], self.slots[p + "query.bias"]
        if arena == "shadow":
            return self.shadow[sw.offset:sw.offset + 3 * D * D].view(3 * D, D), self.master[sb.offset:sb.offset + 3 * D]
        return self.grads[sw.offset:sw.offset + 3 * D * D].view(3 * D, D), self.grads[sb.offset:sb.offset + 3 * D]

    def bucket_ranges(self):
        """Contiguous [start, end) element ranges of the gradient arena, in the order backward completes them:
        head (+final LN, pooler), layers L-1..0, embeddings."""
        L = self.cfg.num_hidden_layers
        names = param_order(L)
        first = {n: self.slots[n].offset for n in names}
        end_total = self.grads.numel()
        starts = [first[f"backbone.encoder.layer.{i}.layernorm_before.weight"] for i in range(L)]
        tail_start = first["backbone.layernorm.weight"]
        ranges = [("head", tail_start, end_total)]
        for i in reversed(range(L)):
            e = starts[i + 1] if i + 1 < L else tail_start
            ranges.append((f"layer{i}", starts[i], e))
        ranges.append(("embed", 0, starts[0] if L > 0 else tail_start))
        return ranges

    # ------------------------------------------------------------------------------------------ workspace
    def workspace(self, B: int, S: int, train: bool):
        key = (B, S, train)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        cfg, dev = self.cfg, self.device
        D, I, L, P = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.patch_size
        g = S // P
        T1 = g * g + 1
        M = B * T1
        H = cfg.num_attention_heads
        e = lambda *shape, dtype=BF16: torch.empty(*shape, device=dev, dtype=dtype)  # noqa: E731
        ws = {"g": g, "T1": T1, "M": M}
        ws["patches"] = torch.zeros(M, 3 * P * P, device=dev, dtype=BF16)  # CLS rows stay zero
        nsave = L if train else 1
        ws["x_in"] = [e(M, D, dtype=F32) for _ in range(nsave + 1 if train else 1)]
        ws["x_mid"] = [e(M, D, dtype=F32) for _ in range(nsave)]
        ws["ln1"] = [e(M, D) for _ in range(nsave)]
        ws["qkv"] = [e(M, 3 * D) for _ in range(nsave)]
        ws["ctx"] = [e(M, D) for _ in range(nsave)]
        ws["ln2"] = [e(M, D) for _ in range(nsave)]
        ws["h_act"] = [e(M, I) for _ in range(nsave)]
        if train:
            ws["h_pre"] = [e(M, I) for _ in range(L)]
            ws["lse"] = [e(B, H, T1, dtype=F32) for _ in range(L)]
            ws["stats"] = [e(4, M, dtype=F32) for _ in range(L)]  # mean1, rstd1, mean2, rstd2
            ws["fstats"] = e(2, M, dtype=F32)
            ws["dx_a"] = e(M, D, dtype=F32)
            ws["dx_b"] = e(M, D, dtype=F32)
            ws["dx16"] = e(M, D)
            ws["dh"] = e(M, I)
            ws["d_ln"] = e(M, D)
            ws["dctx"] = e(M, D)
            ws["dqkv"] = e(M, 3 * D)
            ws["dq_acc"] = e(M, D, dtype=F32)
            ws["delta"] = e(B, H, T1, dtype=F32)
            ws["dfeat"] = e(B * g * g, HEAD_CH)
            ws["dcol"] = e(B * g * g, 9 * D)
            ws["dtok"] = e(M, D, dtype=F32)
        ws["tok"] = e(M, D)
        ws["col"] = e(B * g * g, 9 * D)
        ws["feat"] = e(B * g * g, HEAD_CH)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------ forward
    def seed_dropout(self, seed: int) -> None:
        """sets the per-step dropout counter (tests / reproducibility)."""
        if self.device is None:
            self.ensure_packed(next(self.module.parameters()).device)
        self.rng_step = torch.tensor([seed & 0x7FFFFFFF], device=self.device, dtype=torch.int32)

    def _site(self, p: float, site: int):
        return (p, self.rng_step, site) if p > 0.0 else None

    def forward_lowres(self, x: torch.Tensor, train: bool, dropout: bool = False) -> torch.Tensor:
        """image fp32 [B,3,S,S] -> low-resolution logits fp32 [B,C,g,g] (seg_head output before the upsample).
        dropout=True applies hidden / attention dropout with the config's probabilities (module.training)."""
        cfg = self.cfg
        K.require_cuda(x, "ViTSegmentationModel.forward")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected [B,3,S,S] input, got {tuple(x.shape)}")
        B, _, S, _ = x.shape
        if S != cfg.image_size:
            # same failure mode as ViTPatchEmbeddings.forward (TF:160-165)
            raise ValueError(f"Input image size ({S}*{S}) doesn't match model ({cfg.image_size}*{cfg.image_size}).")
        self.ensure_packed(x.device)
        self.launches = 0
        self.refresh_shadow(train)
        x = x.contiguous().to(F32)
        D, I, L, P, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.patch_size, cfg.num_attention_heads
        Cn = cfg.num_classes
        ws = self.workspace(B, S, train)
        g, T1, M = ws["g"], ws["T1"], ws["M"]
        scale = 1.0 / math.sqrt(D // H)
        n = 0
        # dropout follows module.training as in the reference, also on a no-grad forward (train=False only means that
        # no activations are saved for a backward pass)
        p_hid = float(cfg.hidden_dropout_prob) if dropout else 0.0
        p_att = float(cfg.attention_probs_dropout_prob) if dropout else 0.0
        if p_hid > 0.0 or p_att > 0.0:
            if self.rng_step is None or self.rng_step.device != x.device:
                seed = int(torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64))
                # data parallel: ranks that seeded the host generator identically must still draw independent masks
                if torch.distributed.is_available() and torch.distributed.is_initialized():
                    seed = (seed + torch.distributed.get_rank() * 0x3C6EF35F) & 0x7FFFFFFF
                self.rng_step = torch.tensor([seed], dtype=torch.int32).to(x.device)
            self.rng_step.add_(1)   # device-side: a captured graph advances it on every replay
        if train:
            self._drop = (p_hid, p_att)

        # --- embeddings: patch projection (+bias +pos) and CLS rows (TF:100-128)
        K.patchify(x, ws["patches"], P)
        xcur = ws["x_in"][0]
        K.gemm(ws["patches"], self.w16("backbone.embeddings.patch_embeddings.projection.weight", D, 3 * P * P), xcur,
               bias=self.w32("backbone.embeddings.patch_embeddings.projection.bias"),
               residual=self.w32("backbone.embeddings.position_embeddings").view(T1, D), row_tokens=T1)
        K.cls_rows(self.w32("backbone.embeddings.cls_token"), self.w32("backbone.embeddings.position_embeddings"),
                   xcur, B, T1, D)
        n += 3
        if p_hid > 0.0:   # embedding dropout (TF:126), site 0
            K.dropout_rows(xcur, None, self._site(p_hid, 0))
            n += 1
        # --- encoder layers (TF:328-346)
        for i in range(L):
            j = i if train else 0
            p = f"backbone.encoder.layer.{i}."
            st = ws["stats"][i] if train else None
            x_in = ws["x_in"][i] if train else ws["x_in"][0]
            x_mid = ws["x_mid"][j]
            x_out = ws["x_in"][i + 1] if train else ws["x_in"][0]
            K.layernorm_fwd(x_in, self.w32(p + "layernorm_before.weight"), self.w32(p + "layernorm_before.bias"),
                            LN_EPS, y_bf16=ws["ln1"][j], mean=st[0] if train else None, rstd=st[1] if train else None)
            wqkv, bqkv = self.fused_qkv(i)
            K.gemm(ws["ln1"][j], wqkv, ws["qkv"][j], bias=bqkv)
            K.attention_fwd(ws["qkv"][j], ws["ctx"][j], ws["lse"][i] if train else None, B, T1, H, scale,
                            dropout=self._site(p_att, 1000 + i))
            K.gemm(ws["ctx"][j], self.w16(p + "attention.output.dense.weight"), x_mid,
                   bias=self.w32(p + "attention.output.dense.bias"), residual=x_in,
                   dropout=self._site(p_hid, 1 + 2 * i))
            K.layernorm_fwd(x_mid, self.w32(p + "layernorm_after.weight"), self.w32(p + "layernorm_after.bias"),
                            LN_EPS, y_bf16=ws["ln2"][j], mean=st[2] if train else None, rstd=st[3] if train else None)
            K.gemm(ws["ln2"][j], self.w16(p + "intermediate.dense.weight"), ws["h_act"][j],
                   bias=self.w32(p + "intermediate.dense.bias"), act=K.ACT_GELU,
                   out2=ws["h_pre"][i] if train else None)
            K.gemm(ws["h_act"][j], self.w16(p + "output.dense.weight"), x_out,
                   bias=self.w32(p + "output.dense.bias"), residual=x_mid, dropout=self._site(p_hid, 2 + 2 * i))
            n += 7
        x_fin = ws["x_in"][L] if train else ws["x_in"][0]
        fs = ws["fstats"] if train else None
        K.layernorm_fwd(x_fin, self.w32("backbone.layernorm.weight"), self.w32("backbone.layernorm.bias"), LN_EPS,
                        y_bf16=ws["tok"], mean=fs[0] if train else None, rstd=fs[1] if train else None)
        # --- segmentation head (model/CE/classes.py:240-257): conv3x3+ReLU as im2col GEMM, conv1x1
        K.head_im2col(ws["tok"], ws["col"], B, g, D)
        K.gemm(ws["col"], self.head_w_packed, ws["feat"], bias=self.w32("seg_head.0.bias"), act=K.ACT_RELU)
        low = torch.empty(B, Cn, g, g, device=x.device, dtype=F32)
        K.conv1x1_fwd(ws["feat"], self.w32("seg_head.2.weight").view(Cn, HEAD_CH), self.w32("seg_head.2.bias"), low, B, g,
                      HEAD_CH, Cn)
        n += 4
        self.launches += n
        if train:
            self._saved = (B, S)
        return low

    # ------------------------------------------------------------------------------------------ backward
    def _prepare_grads(self):
        """param.grad must alias the gradient arena.  Slots whose .grad was reset (zero_grad(set_to_none=True)) are
        zeroed and re-attached; existing aliased grads are accumulated into (gradient accumulation)."""
        params = self._named()
        base = self.grads.data_ptr()
        fresh = all(p.grad is None for p in params.values())
        frozen = any(not p.requires_grad for p in params.values())
        # frozen parameters: the backward kernels still write their arena slots (the GEMM sequence is fixed), but the
        # slots are never exposed as .grad and never reach an optimizer; they are cleared with the rest every step
        if fresh and (frozen or not self._grads_zeroed):
            self.grads.zero_()
            self.launches += 1
        self._grads_zeroed = False
        foreign = []
        for n, p in params.items():
            if n.startswith("backbone.pooler."):
                continue  # dead compute in the reference forward (TF:456 result unused): no gradient
            if not p.requires_grad:
                continue  # requires_grad=False: no .grad, as autograd would leave it
            s = self.slots[n]
            gview = self.grads[s.offset:s.offset + s.numel].view(s.shape)
            if p.grad is None:
                if not fresh:
                    gview.zero_()
                p.grad = gview
            elif p.grad.data_ptr() != base + s.offset * 4:
                foreign.append((p, gview, p.grad))
                gview.zero_()
        return foreign

    def backward_lowres(self, dlow: torch.Tensor):
        """Backpropagates d(loss)/d(low-res logits) through head, encoder and embeddings, accumulating every
        parameter gradient into the flat arena."""
        assert self._saved is not None, "backward without a training-mode forward"
        B, S = self._saved
        cfg = self.cfg
        D, I, L, P, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.patch_size, cfg.num_attention_heads
        Cn = cfg.num_classes
        ws = self.workspace(B, S, True)
        g, T1, M = ws["g"], ws["T1"], ws["M"]
        scale = 1.0 / math.sqrt(D // H)
        foreign = self._prepare_grads()
        dlow = dlow.contiguous().to(F32)
        n = 0
        hook = self.grad_ready_hook
        p_hid, p_att = self._drop

        # --- head
        K.conv1x1_bwd(dlow, ws["feat"], self.w32("seg_head.2.weight").view(Cn, HEAD_CH), ws["dfeat"],
                      self.g32("seg_head.2.weight", (Cn, HEAD_CH)), self.g32("seg_head.2.bias"), B, g, HEAD_CH, Cn)
        K.colsum(ws["dfeat"], self.g32("seg_head.0.bias"), accumulate=True)
        self.head_wgrad_packed.zero_()
        K.gemm(ws["dfeat"], ws["col"], self.head_wgrad_packed, a_mn=True, b_mn=True, accumulate=True)
        K.unpack_conv3x3_grad(self.head_wgrad_packed, self.g32("seg_head.0.weight"))
        K.gemm(ws["dfeat"], self.head_w_packed, ws["dcol"], b_mn=True)
        K.head_col2im(ws["dcol"], ws["dtok"], B, g, D)
        # final LayerNorm
        fs = ws["fstats"]
        dx, dx_other = ws["dx_a"], ws["dx_b"]
        # the bf16 copy of dx feeds the backward GEMMs of the last layer's fc2, whose output was dropped out (site 2L)
        K.layernorm_bwd(ws["dtok"], ws["x_in"][L], self.w32("backbone.layernorm.weight"), fs[0], fs[1], None, dx,
                        ws["dx16"], self.g32("backbone.layernorm.weight"), self.g32("backbone.layernorm.bias"),
                        dropout=self._site(p_hid, 2 * L) if L > 0 else None,
                        dbias=self.g32(f"backbone.encoder.layer.{L - 1}.output.dense.bias") if L > 0 else None)
        n += 8
        if hook:
            hook("head")
        # --- encoder layers in reverse
        # Weight-gradient GEMMs run on a second stream, concurrently with the data-gradient chain: dW = dY^T X and
        # dX = dY W are independent, and at batch 64 every GEMM ends in a partly filled wave (12608 rows = 49.25 tiles
        # of 256) — the other kernel's CTAs fill the SMs that a kernel's tail wave leaves idle.  Forks / joins are
        # events, so the pattern is captured into the step's CUDA graph as parallel branches (VS_BWD_STREAMS=2).  VS_BWD_STREAMS=1 keeps
        # everything on one stream (the default: the two-stream form measured neutral on one GPU, r02).
        side = self._side_stream() if self._two_streams else None
        cur = torch.cuda.current_stream()

        def fork():   # side stream continues from the current point of the main stream
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)

        def join():   # main stream waits for everything queued on the side stream
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(side)
                cur.wait_event(ev)

        def wgrad(dy, xsaved, gw):
            if side is None:
                K.gemm(dy, xsaved, gw, a_mn=True, b_mn=True, accumulate=True)
            else:
                with torch.cuda.stream(side):
                    K.gemm(dy, xsaved, gw, a_mn=True, b_mn=True, accumulate=True)

        for i in reversed(range(L)):
            p = f"backbone.encoder.layer.{i}."
            st = ws["stats"][i]
            # fc2: x_out = x_mid + h_act W2^T + b2   (its bias gradient = column sums of dx16: fused into the
            # LayerNorm backward that produced dx16)
            fork()
            wgrad(ws["dx16"], ws["h_act"][i], self.g32(p + "output.dense.weight"))
            # the fc1 bias gradient (column sums of dh) comes out of this GEMM's epilogue (out_colsum, column-persistent
            # tile order: sums carried in registers, one reduction per warp and kernel).  r02 session 13, same box:
            # 11.46 -> 11.24 ms per step, 12 launches and a 77 MB re-read per layer gone; VS_FUSE_COLSUM=0 restores the
            # separate column-sum pass
            K.gemm(ws["dx16"], self.w16(p + "output.dense.weight"), ws["dh"], b_mn=True, aux=ws["h_pre"][i],
                   aux_mode=K.AUX_GELU_GRAD,
                   colsum=self.g32(p + "intermediate.dense.bias") if self._fuse_colsum else None)
            # fc1
            fork()   # dh is final
            wgrad(ws["dh"], ws["ln2"][i], self.g32(p + "intermediate.dense.weight"))
            if not self._fuse_colsum:
                K.colsum(ws["dh"], self.g32(p + "intermediate.dense.bias"), accumulate=True)
            K.gemm(ws["dh"], self.w16(p + "intermediate.dense.weight"), ws["d_ln"], b_mn=True)
            # LN2 + skip (overwrites dx16, which the fc2 weight gradient reads: join first)
            join()
            K.layernorm_bwd(ws["d_ln"], ws["x_mid"][i], self.w32(p + "layernorm_after.weight"), st[2], st[3], dx,
                            dx_other, ws["dx16"], self.g32(p + "layernorm_after.weight"),
                            self.g32(p + "layernorm_after.bias"), dropout=self._site(p_hid, 1 + 2 * i),
                            dbias=self.g32(p + "attention.output.dense.bias"))
            dx, dx_other = dx_other, dx
            # attention output projection
            fork()
            wgrad(ws["dx16"], ws["ctx"][i], self.g32(p + "attention.output.dense.weight"))
            K.gemm(ws["dx16"], self.w16(p + "attention.output.dense.weight"), ws["dctx"], b_mn=True)
            # attention core
            # T1 <= 256: whole (batch, head) items per CTA, dQ written as bf16 (no fp32 accumulator, no cast pass);
            # longer sequences: key-block kernel + fp32 dQ accumulator, cast inside the bias-gradient pass below
            short = self._attn_bwd_short and T1 <= 256
            K.attention_bwd(ws["qkv"][i], ws["ctx"][i], ws["dctx"], ws["lse"][i], ws["dqkv"],
                            None if short else ws["dq_acc"], ws["delta"],
                            B, T1, H, scale, dropout=self._site(p_att, 1000 + i))
            # fused QKV projection; its bias gradient = column sums of dQKV
            gw, gb = self.fused_qkv(i, arena="grads")
            wqkv, _ = self.fused_qkv(i)
            if short:
                K.colsum(ws["dqkv"], gb, accumulate=True)
            else:
                K.colsum_cast(ws["dqkv"], ws["dq_acc"], gb, accumulate=True)
            fork()   # dqkv is final
            wgrad(ws["dqkv"], ws["ln1"][i], gw)
            K.gemm(ws["dqkv"], wqkv, ws["d_ln"], b_mn=True)
            # LN1 + skip
            # next consumer of dx16: fc2 of layer i-1 (site 2i); for i == 0 the embedding dropout is applied below
            join()
            K.layernorm_bwd(ws["d_ln"], ws["x_in"][i], self.w32(p + "layernorm_before.weight"), st[0], st[1], dx,
                            dx_other, ws["dx16"], self.g32(p + "layernorm_before.weight"),
                            self.g32(p + "layernorm_before.bias"),
                            dropout=self._site(p_hid, 2 * i) if i > 0 else None,
                            dbias=self.g32(f"backbone.encoder.layer.{i - 1}.output.dense.bias") if i > 0 else None)
            dx, dx_other = dx_other, dx
            n += 15 if self._fuse_colsum else 16
            if hook:
                hook(f"layer{i}")
        # --- embeddings
        if p_hid > 0.0:   # same mask as the forward embedding dropout, on the gradient (fp32 in place + bf16 copy)
            K.dropout_rows(dx, ws["dx16"], self._site(p_hid, 0))
            n += 1
        K.embed_bwd(dx, self.g32("backbone.embeddings.cls_token").view(D),
                    self.g32("backbone.embeddings.position_embeddings").view(T1 * D),
                    self.g32("backbone.embeddings.patch_embeddings.projection.bias"), B, T1, D)
        K.gemm(ws["dx16"], ws["patches"], self.g32("backbone.embeddings.patch_embeddings.projection.weight",
                                                   (D, 3 * P * P)), a_mn=True, b_mn=True, accumulate=True)
        n += 2
        if hook:
            hook("embed")
        self._foreign = foreign
        if not hook:
            self.finish_foreign_grads()
        self.launches += n
        self._fire_param_hooks()

    def finish_foreign_grads(self):
        """.grad tensors that do not alias the arena (set by the user / another optimizer) receive this backward's
        gradient by addition.  Under data parallel this runs after the bucket all-reduces have completed
        (DataParallel.step), never while a collective is still in flight on the arena slice."""
        for p_, gview, old in getattr(self, "_foreign", ()):
            old.add_(gview)
        self._foreign = []

    def _fire_param_hooks(self):
        """Parameter gradients are produced by the engine, not by autograd's AccumulateGrad nodes, so hooks registered
        with Tensor.register_post_accumulate_grad_hook are invoked here (once per backward, after all gradients are
        final).  Hooks on the AccumulateGrad node itself (what torch DistributedDataParallel installs) cannot be
        reached: use visiontransformer_b200.dp.DataParallel for multi-GPU training."""
        for p in self.module.parameters():
            hooks = getattr(p, "_post_accumulate_grad_hooks", None)
            if hooks and p.grad is not None:
                for h in list(hooks.values()):
                    h(p)
