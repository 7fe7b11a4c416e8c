"""CPU ORACLE — test infrastructure, NOT product code.

A plain fp32 PyTorch-on-CPU restatement of the reference's ViT-segmentation hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may import this module; the product
package (visiontransformer_b200/) never does, and has no CPU path of its own.

Parity pin: the reference ships no tests or golden vectors for this path, so the pin was
minted here: oracle/make_golden.py imports the reference classes from /root/reference (with import stubs for its
missing optional dependencies), loads the same seeded weights into the reference and into this restatement, and
commits the reference's outputs under tests/golden/.  tests/test_oracle_golden.py checks this file against them.

Every function cites the reference lines it follows.
  CE   = /root/reference/model/CE/classes.py
  PAED = /root/reference/model/PAED/classes.py
  SEG  = /root/reference/model/PAED/segmentation.py
  TF   = transformers/models/vit/modeling_vit.py (transformers 5.5.0; un-vendored dependency, not pinned by the
         reference's requirements.txt — its ViT math is unchanged since the 4.x series the reference was written on)
  SDPA = transformers/integrations/sdpa_attention.py
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    num_classes: int
    patch_size: int
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    image_size: int = 224           # CE:225 hard-codes 224; SURVEY D3 adds the knob
    intermediate_size: int = 3072   # CE:231
    layer_norm_eps: float = 1e-12   # ViTConfig default (TF:325-326)


def param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """state_dict layout of ViTSegmentationModel."""
    D, P, I, C = cfg.hidden_size, cfg.patch_size, cfg.intermediate_size, cfg.num_classes
    T = (cfg.image_size // P) ** 2
    s = {
        "backbone.embeddings.cls_token": (1, 1, D),
        "backbone.embeddings.position_embeddings": (1, T + 1, D),
        "backbone.embeddings.patch_embeddings.projection.weight": (D, 3, P, P),
        "backbone.embeddings.patch_embeddings.projection.bias": (D,),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"backbone.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (D, D)
            s[p + f"attention.attention.{n}.bias"] = (D,)
        s[p + "attention.output.dense.weight"] = (D, D)
        s[p + "attention.output.dense.bias"] = (D,)
        s[p + "intermediate.dense.weight"] = (I, D)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (D, I)
        s[p + "output.dense.bias"] = (D,)
        s[p + "layernorm_before.weight"] = (D,)
        s[p + "layernorm_before.bias"] = (D,)
        s[p + "layernorm_after.weight"] = (D,)
        s[p + "layernorm_after.bias"] = (D,)
    s["backbone.layernorm.weight"] = (D,)
    s["backbone.layernorm.bias"] = (D,)
    s["backbone.pooler.dense.weight"] = (D, D)
    s["backbone.pooler.dense.bias"] = (D,)
    s["seg_head.0.weight"] = (256, D, 3, 3)
    s["seg_head.0.bias"] = (256,)
    s["seg_head.2.weight"] = (C, 256, 1, 1)
    s["seg_head.2.bias"] = (C,)
    return s


def seeded_state_dict(cfg: OracleConfig, seed: int, bf16_representable: bool = True, head_gain: float = 1.0):
    """Deterministic synthetic weights (CPU generator => identical on every machine).  Distributions follow the
    reference init (TF:385-398 trunc-normal std 0.02 for Linear/Conv/pos/cls, LayerNorm (1,0); torch's default
    uniform init for seg_head), with small random biases / LN affine so that every parameter influences the output.
    Values are rounded to bf16-representable numbers by default (SURVEY.md §7.2-1 parity protocol)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if name.startswith("seg_head") and name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            bound = head_gain / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        elif name.startswith("seg_head"):
            t = (torch.rand(shape, generator=gen) * 2 - 1) * 0.05
        elif "layernorm" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=gen)
        else:
            t = (0.02 * torch.randn(shape, generator=gen)).clamp_(-0.04, 0.04)
        if bf16_representable:
            t = t.to(torch.bfloat16).to(torch.float32)
        sd[name] = t
    return sd


# ------------------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------------------
def _drop(t, p):
    """nn.Dropout in train mode (TF:126,267,310; SDPA:96 dropout_p); p = 0 is the identity used by every parity test."""
    return F.dropout(t, p, training=True) if p > 0.0 else t


def embeddings(sd, x, cfg, p_hidden=0.0):
    """TF:153-167 (conv patch projection, flatten, transpose) + TF:100-128 (CLS, position embeddings, dropout)."""
    P = cfg.patch_size
    h = F.conv2d(x, sd["backbone.embeddings.patch_embeddings.projection.weight"],
                 sd["backbone.embeddings.patch_embeddings.projection.bias"], stride=P)
    h = h.flatten(2).transpose(1, 2)
    cls = sd["backbone.embeddings.cls_token"].expand(x.shape[0], -1, -1)
    h = torch.cat((cls, h), dim=1)
    return _drop(h + sd["backbone.embeddings.position_embeddings"], p_hidden)


def encoder_layer(sd, h, i, cfg, p_hidden=0.0, p_attn=0.0):
    """TF:328-346 pre-LN block; attention TF:220-251 with SDPA:92-101 (scale = head_dim**-0.5, non-causal);
    MLP TF:290-312 with exact-erf GELU."""
    p = f"backbone.encoder.layer.{i}."
    D, H = cfg.hidden_size, cfg.num_attention_heads
    dh = D // H
    B, N, _ = h.shape
    y = F.layer_norm(h, (D,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], cfg.layer_norm_eps)

    def proj(n):
        return F.linear(y, sd[p + f"attention.attention.{n}.weight"], sd[p + f"attention.attention.{n}.bias"]) \
            .view(B, N, H, dh).transpose(1, 2)

    q, k, v = proj("query"), proj("key"), proj("value")
    att = _drop(torch.softmax((q @ k.transpose(-1, -2)) * (dh ** -0.5), dim=-1), p_attn)
    ctx = (att @ v).transpose(1, 2).reshape(B, N, D)
    h = h + _drop(F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]),
                  p_hidden)
    y = F.layer_norm(h, (D,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], cfg.layer_norm_eps)
    y = F.gelu(F.linear(y, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return h + _drop(F.linear(y, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]), p_hidden)


def forward_lowres(sd, x, cfg, dropout=(0.0, 0.0)):
    """CE:246-257: backbone (TF:428-458, pooler result unused), drop CLS, NCHW view, seg_head.
    dropout = (hidden_dropout_prob, attention_probs_dropout_prob) of the train-mode forward (CE:233-234: 0.1 / 0.1)."""
    h = embeddings(sd, x, cfg, dropout[0])
    for i in range(cfg.num_hidden_layers):
        h = encoder_layer(sd, h, i, cfg, dropout[0], dropout[1])
    D = cfg.hidden_size
    h = F.layer_norm(h, (D,), sd["backbone.layernorm.weight"], sd["backbone.layernorm.bias"], cfg.layer_norm_eps)
    h = h[:, 1:, :]
    B, T, _ = h.shape
    g = int(T ** 0.5)
    feat = h.transpose(1, 2).reshape(B, D, g, g)
    out = F.relu(F.conv2d(feat, sd["seg_head.0.weight"], sd["seg_head.0.bias"], padding=1))
    return F.conv2d(out, sd["seg_head.2.weight"], sd["seg_head.2.bias"])


def upsample(low, size):
    """CE:260."""
    return F.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)


def forward(sd, x, cfg, dropout=(0.0, 0.0)):
    """CE:246-262: ViTSegmentationModel.forward."""
    return upsample(forward_lowres(sd, x, cfg, dropout), x.shape[-1])


# ------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------
def resize_target(y, size):
    """CE:273-274 / PAED:496-506: legacy 'nearest' resize of the label map."""
    if y.dim() == 3:
        y = y.unsqueeze(1)
    return F.interpolate(y.float(), size=(size, size), mode="nearest").squeeze(1).long()


def ce_loss(logits, y):
    """CE:268,280: nn.CrossEntropyLoss() — mean over B*H*W, ignore_index -100."""
    return F.cross_entropy(logits, y)


def paed_loss_multiclass_soft(msk, pred_mask, sigma=3, class_penalty=True):
    """PAED:336-369 (19x19 Gaussian depthwise blur of mask and probabilities, |diff|, class-mismatch penalty)."""
    B, C, H, W = msk.shape
    size = int(6 * sigma + 1)
    x = torch.arange(size).float() - size // 2
    gauss = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    k2 = gauss[:, None] * gauss[None, :]
    k2 = (k2 / k2.sum()).unsqueeze(0).unsqueeze(0).repeat(C, 1, 1, 1)
    msk_s = F.conv2d(msk, k2, padding=size // 2, groups=C)
    pred_s = F.conv2d(pred_mask, k2, padding=size // 2, groups=C)
    base = torch.abs(msk_s - pred_s)
    if class_penalty:
        dist = (msk * (1 - pred_mask) * base * 2).mean(dim=[2, 3])
    else:
        dist = base.mean(dim=[2, 3])
    return dist.mean(dim=1).mean()


def paed_multiclass_step_loss(logits, y, num_classes=17):
    """PAED:448-467 training_step of the multi-class LightningViTModel."""
    probs = torch.softmax(logits, dim=1)
    onehot = F.one_hot(y.long(), num_classes).permute(0, 3, 1, 2).float()
    return paed_loss_multiclass_soft(onehot, probs)


def dice_loss(preds, targets, smooth=1e-6):
    """PAED:608-620."""
    p, t = preds.float().reshape(-1), targets.float().reshape(-1)
    return 1 - (2.0 * (p * t).sum() + smooth) / (p.sum() + t.sum() + smooth)


def paed_loss_soft(sdf_ext, sdf_int, preds):
    """PAED:623-661 (sdf_* are [B,1,H,W])."""
    B, _, H, W = preds.shape
    sdf_ext = F.interpolate(sdf_ext, size=(H, W), mode="bilinear", align_corners=False)
    sdf_int = F.interpolate(sdf_int, size=(H, W), mode="bilinear", align_corners=False)
    sx = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], dtype=torch.float32).view(1, 1, 3, 3)
    gx = F.conv2d(preds, sx, padding=1)
    gy = F.conv2d(preds, sx.transpose(2, 3), padding=1)
    edge = torch.sqrt(gx ** 2 + gy ** 2 + 1e-6)
    mx = edge.view(B, -1).max(dim=1)[0].view(B, 1, 1, 1) + 1e-6
    edge = edge / mx
    ext = (sdf_ext * edge).mean()
    inn = (sdf_int * preds).mean()
    return 1 * ext - 0.5 * inn


def paed_binary_step_loss(logits, masks, sdf_ext, sdf_int):
    """PAED:664-681 (_forward_step_paed): loss = bce + 0.1*dice + 5*|paed|.  masks [B,H,W] in {0,1}."""
    preds = torch.sigmoid(logits)
    paed = paed_loss_soft(sdf_ext.unsqueeze(1), sdf_int.unsqueeze(1), preds)
    m = masks.unsqueeze(1).float()
    bce = F.binary_cross_entropy(preds, m)
    return bce + 0.1 * dice_loss(preds, m) + 5.0 * torch.abs(paed)


def compute_sdf(mask: np.ndarray):
    """SEG:6-34."""
    from scipy.ndimage import distance_transform_edt
    mask = mask.astype(bool)
    sdf_ext = distance_transform_edt(~mask).astype(np.float32)
    sdf_int = distance_transform_edt(mask).astype(np.float32)
    if sdf_ext.max() > 0:
        sdf_ext /= sdf_ext.max()
    if sdf_int.max() > 0:
        sdf_int /= sdf_int.max()
    return sdf_ext, sdf_int


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def synthetic_images(B, S, seed=1234, bf16_representable=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, S, S, generator=g)
    return x.to(torch.bfloat16).to(torch.float32) if bf16_representable else x


def synthetic_labels(B, num_classes, seed=1235, size=256):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, num_classes, (B, size, size), generator=g)


def learnable_labels(x, num_classes, patch=16):
    """labels that are a function of the image (block-mean brightness), so the loss can fall during training."""
    m = F.avg_pool2d(x.mean(1, keepdim=True), patch)
    lab = (m * num_classes * 2 - num_classes / 2).floor().clamp(0, num_classes - 1)
    return F.interpolate(lab, size=x.shape[-2:], mode="nearest").squeeze(1).long()


def curve_task(name: str = "regions", B: int = 2, S: int = 224, num_classes: int = 17, block: int = 32,
               noise: float = 0.5, contrast: float = 0.8, seed: int = 77, palette_seed: int = 77):
    """fixed batch for the 200-step loss-curve parity run (tests/golden/vitb16_curve.pt).
      'regions': block x block pixel cells, each of one class, coloured from a fixed palette (scaled by `contrast`) plus
                 uniform pixel noise; a fraction `noise` of the label pixels is replaced by uniformly random classes, so
                 the loss falls smoothly from ln(17) towards an irreducible floor instead of collapsing to 0 (where a
                 relative gap is meaningless);
      'brightness': the learnable_labels task of the 60-step test."""
    g = torch.Generator().manual_seed(seed)
    if name == "regions":
        pal = torch.rand(num_classes, 3, generator=torch.Generator().manual_seed(palette_seed))
        if seed == palette_seed:
            torch.rand(num_classes, 3, generator=g)   # keep the stream of the original single-generator version
        z = torch.randint(0, num_classes, (B, S // block, S // block), generator=g)
        z = z.repeat_interleave(block, 1).repeat_interleave(block, 2)
        x = pal[z].permute(0, 3, 1, 2) * contrast + (1.0 - contrast) * torch.rand(B, 3, S, S, generator=g)
        x = x.to(torch.bfloat16).to(torch.float32)
        flip = torch.rand(B, S, S, generator=g) < noise
        y = torch.where(flip, torch.randint(0, num_classes, (B, S, S), generator=g), z)
        return x, y
    x = synthetic_images(B, S, seed=32)
    return x, learnable_labels(x, num_classes)


def synthetic_binary_targets(B, S, seed=1236):
    """1-3 random discs per image -> mask {0,1} [B,S,S] + SDF maps via compute_sdf (SEG:22-32)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:S, 0:S]
    masks, se, si = [], [], []
    for _ in range(B):
        m = np.zeros((S, S), dtype=bool)
        for _ in range(rng.randint(1, 4)):
            cy, cx, r = rng.randint(20, S - 20), rng.randint(20, S - 20), rng.randint(8, 40)
            m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        e, i = compute_sdf(m.astype(np.uint8))
        masks.append(m.astype(np.float32)); se.append(e); si.append(i)
    t = lambda a: torch.from_numpy(np.stack(a))  # noqa: E731
    return t(masks), t(se), t(si)


# gradient pins of the headline-size training-step goldens (tests/golden/vitb16_train.pt, p8w1024_train.pt): (parameter, index expression applied to its .grad) spanning head / last layer / first layer /
# embeddings; big matrices are strided so the fixture stays small
def headline_grad_pins(L: int):
    last = L - 1
    return [
        ("model.seg_head.2.weight", (slice(None),) * 4),
        ("model.seg_head.0.bias", (slice(None),)),
        ("model.seg_head.0.weight", (slice(None, None, 16), slice(None, None, 24))),
        ("model.backbone.layernorm.weight", (slice(None),)),
        (f"model.backbone.encoder.layer.{last}.output.dense.bias", (slice(None),)),
        (f"model.backbone.encoder.layer.{last}.intermediate.dense.weight", (slice(None, None, 48), slice(None, None, 12))),
        (f"model.backbone.encoder.layer.{last}.attention.attention.value.weight", (slice(None, None, 12), slice(None, None, 12))),
        (f"model.backbone.encoder.layer.{L // 2}.attention.output.dense.weight", (slice(None, None, 12), slice(None, None, 12))),
        ("model.backbone.encoder.layer.0.attention.attention.query.weight", (slice(None, None, 12), slice(None, None, 12))),
        ("model.backbone.encoder.layer.0.attention.attention.query.bias", (slice(None),)),
        ("model.backbone.encoder.layer.0.layernorm_before.weight", (slice(None),)),
        ("model.backbone.encoder.layer.0.output.dense.weight", (slice(None, None, 12), slice(None, None, 48))),
        ("model.backbone.embeddings.cls_token", (slice(None),) * 3),
        ("model.backbone.embeddings.position_embeddings", (slice(None), slice(None, None, 7), slice(None, None, 8))),
        ("model.backbone.embeddings.patch_embeddings.projection.weight", (slice(None, None, 8),)),
        ("model.backbone.embeddings.patch_embeddings.projection.bias", (slice(None),)),
    ]


def to_module_state_dict(sd, prefix=""):
    return {prefix + k: v.clone() for k, v in sd.items()}
