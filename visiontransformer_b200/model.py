"""ViTSegmentationModel — same nn.Module surface and state_dict layout as the reference class
(model/CE/classes.py:221-262, identical copy at model/PAED/classes.py:372-413), executed by libvitseg kernels.

The parameter tree reproduces transformers.ViTModel's names (204 tensors for ViT-B/16,
incl. the unused backbone.pooler.dense.*), so `load_state_dict(torch.load(ckpt)['state_dict'])` works unchanged.
The sub-modules below are parameter holders only: their own forward() is never used on the hot path."""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
from torch import nn

from . import kernels as K
from .engine import HEAD_CH, Engine

F32 = torch.float32


def _trunc_normal_(t: torch.Tensor, std: float):
    # HF ViT init (TF:385-398): trunc_normal_(mean=0, std=initializer_range) in fp32
    return nn.init.trunc_normal_(t, mean=0.0, std=std)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guards against accidental eager use
        raise RuntimeError("parameter holder: the ViT forward runs inside the libvitseg engine")


class _Linear(_Holder):
    def __init__(self, fin, fout, std):
        super().__init__()
        self.weight = nn.Parameter(_trunc_normal_(torch.empty(fout, fin), std))
        self.bias = nn.Parameter(torch.zeros(fout))


class _LayerNorm(_Holder):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _SelfAttention(_Holder):
    def __init__(self, d, std):
        super().__init__()
        self.query, self.key, self.value = _Linear(d, d, std), _Linear(d, d, std), _Linear(d, d, std)


class _SelfOutput(_Holder):
    def __init__(self, fin, fout, std):
        super().__init__()
        self.dense = _Linear(fin, fout, std)


class _Attention(_Holder):
    def __init__(self, d, std):
        super().__init__()
        self.attention = _SelfAttention(d, std)
        self.output = _SelfOutput(d, d, std)


class _Layer(_Holder):
    def __init__(self, d, inter, std):
        super().__init__()
        self.attention = _Attention(d, std)
        self.intermediate = _SelfOutput(d, inter, std)
        self.output = _SelfOutput(inter, d, std)
        self.layernorm_before = _LayerNorm(d)
        self.layernorm_after = _LayerNorm(d)


class _Encoder(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList(
            [_Layer(cfg.hidden_size, cfg.intermediate_size, cfg.initializer_range) for _ in range(cfg.num_hidden_layers)])


class _PatchEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.projection = _Holder()
        P = cfg.patch_size
        self.projection.weight = nn.Parameter(
            _trunc_normal_(torch.empty(cfg.hidden_size, cfg.num_channels, P, P), cfg.initializer_range))
        self.projection.bias = nn.Parameter(torch.zeros(cfg.hidden_size))


class _Embeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        T = (cfg.image_size // cfg.patch_size) ** 2
        self.cls_token = nn.Parameter(_trunc_normal_(torch.empty(1, 1, cfg.hidden_size), cfg.initializer_range))
        self.position_embeddings = nn.Parameter(
            _trunc_normal_(torch.empty(1, T + 1, cfg.hidden_size), cfg.initializer_range))
        self.patch_embeddings = _PatchEmbeddings(cfg)


class _Backbone(_Holder):
    """Mirror of transformers.ViTModel's module tree (embeddings / encoder / layernorm / pooler)."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.layernorm = _LayerNorm(cfg.hidden_size)
        self.pooler = _SelfOutput(cfg.hidden_size, cfg.hidden_size, cfg.initializer_range)  # .dense; dead compute


class _LowresFn(torch.autograd.Function):
    """image -> low-res logits; backward fills parameter gradients inside the engine (flat arena)."""

    @staticmethod
    def forward(ctx, x, anchor, engine, dropout):
        low = engine.forward_lowres(x, train=True, dropout=dropout)
        engine._generation = getattr(engine, "_generation", 0) + 1
        ctx.engine = engine
        ctx.generation = engine._generation
        return low

    @staticmethod
    def backward(ctx, dlow):
        eng = ctx.engine
        if ctx.generation != eng._generation:
            raise RuntimeError("ViTSegmentationModel: activations of this forward were overwritten by a later forward; "
                               "call backward() before the next training-mode forward")
        eng.backward_lowres(dlow)
        return None, None, None, None


class _UpsampleFn(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=False) (model/CE/classes.py:260) and its adjoint."""

    @staticmethod
    def forward(ctx, low, size):
        B, C, g, _ = low.shape
        full = torch.empty(B, C, size, size, device=low.device, dtype=F32)
        K.upsample_fwd(low.contiguous(), full)
        ctx.shape = (B, C, g)
        return full

    @staticmethod
    def backward(ctx, dfull):
        B, C, g = ctx.shape
        dlow = torch.empty(B, C, g, g, device=dfull.device, dtype=F32)
        K.upsample_bwd(dfull.contiguous().to(F32), dlow)
        return dlow, None


def upsample_bilinear(low: torch.Tensor, size: int) -> torch.Tensor:
    return _UpsampleFn.apply(low, size)


class ViTSegmentationModel(nn.Module):
    """Drop-in for the reference ViTSegmentationModel(num_classes, patch_size, hidden_size, num_hidden_layers,
    num_attention_heads).  Extra keyword arguments expose what the reference hard-codes."""

    def __init__(self, num_classes, patch_size, hidden_size, num_hidden_layers, num_attention_heads, *,
                 image_size=224, intermediate_size=3072, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1):
        super().__init__()
        if hidden_size % num_attention_heads != 0 or hidden_size // num_attention_heads != 64:
            raise ValueError("visiontransformer_b200 supports head_dim 64 (all reference configs: 768/12, 512/8, 1024/16); "
                             f"got hidden_size={hidden_size}, heads={num_attention_heads}")
        if image_size % patch_size != 0:
            raise ValueError("image_size must be a multiple of patch_size")
        cfg = SimpleNamespace(
            image_size=image_size, patch_size=patch_size, num_channels=3, hidden_size=hidden_size,
            num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
            intermediate_size=intermediate_size, qkv_bias=True, hidden_dropout_prob=hidden_dropout_prob,
            attention_probs_dropout_prob=attention_probs_dropout_prob, initializer_range=0.02,
            layer_norm_eps=1e-12, hidden_act="gelu", num_classes=num_classes)
        self.backbone = _Backbone(cfg)
        # torch's own Conv2d default init, as in the reference (model/CE/classes.py:240-244)
        self.seg_head = nn.Sequential(
            nn.Conv2d(hidden_size, HEAD_CH, kernel_size=3, padding=1),
            nn.ReLU(),
            nn.Conv2d(HEAD_CH, num_classes, kernel_size=1),
        )
        self._engine = Engine(self, cfg)

    # the engine must not be pickled / deep-copied with stale device pointers
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engine"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._engine = Engine(self, self.backbone.config)

    @property
    def engine(self) -> Engine:
        return self._engine

    def _dropout_active(self) -> bool:
        cfg = self.backbone.config
        return self.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)

    def forward_lowres(self, x: torch.Tensor) -> torch.Tensor:
        """Low-resolution logits [B,C,S/P,S/P] = seg_head output before the bilinear upsample."""
        K.require_cuda(x, "ViTSegmentationModel")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            anchor = self.seg_head[2].bias
            return _LowresFn.apply(x, anchor, self._engine, self._dropout_active())
        return self._engine.forward_lowres(x, train=False, dropout=self._dropout_active())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,S,S] fp32 -> logits [B,C,S,S] fp32 (model/CE/classes.py:246-262)."""
        low = self.forward_lowres(x)
        return upsample_bilinear(low, x.shape[-1])

    @torch.no_grad()
    def predict_mask(self, x: torch.Tensor) -> torch.Tensor:
        """Fused upsample+argmax: uint8 class map [B,S,S] == model(x).sigmoid().argmax(1)
        (testViTModel.py:121-126); for num_classes == 1 it is sigmoid(logit) > 0.5."""
        low = self._engine.forward_lowres(x, train=False)
        mask = torch.empty(x.shape[0], x.shape[-1], x.shape[-1], device=x.device, dtype=torch.uint8)
        return K.upsample_argmax(low, mask)


    @torch.no_grad()
    def predict_colored(self, x: torch.Tensor, palette: torch.Tensor) -> torch.Tensor:
        """uint8 RGB mask image [B,S,S,3] = palette[predict_mask(x)] (model/CE/testViTModel.py:139-143: the image the
        inference worker posts back); palette uint8 [num_classes (2 for a binary head), 3] on the same device."""
        return K.colorize_mask(self.predict_mask(x), palette.contiguous())


def flops_per_image(cfg, train: bool) -> float:
    """Algorithmic FLOPs (2*MAC; GEMM/conv/attention only) per image — BASELINE.md §4."""
    P, D, L, I, C = cfg.patch_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size, cfg.num_classes
    T = (cfg.image_size // P) ** 2
    N = T + 1
    fwd = 2 * T * D * 3 * P * P + L * (2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 4 * N * D * I) \
        + 2 * T * HEAD_CH * 9 * D + 2 * T * C * HEAD_CH
    return float(fwd * (3 if train else 1))
