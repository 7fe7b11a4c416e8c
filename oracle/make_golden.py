"""Generates tests/golden/*.pt by running the UNMODIFIED reference classes from /root/reference.

Runs only in the build container (where /root/reference is mounted); the GPU box never executes this.  The
reference's optional imports that are absent here (lightning, segmentation_models_pytorch, torchmetrics, skimage,
matplotlib) are stubbed in sys.modules — none of them is on the arithmetic path.  No
reference source is copied: the classes are imported from where they lie.

usage: python oracle/make_golden.py            (everything)
       python oracle/make_golden.py --headline (only the headline-size training-step goldens)
       python oracle/make_golden.py --curve    (only the 200-step reference loss curves, ~10 min)
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import vitseg_oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def _install_stubs():
    import torchvision  # noqa: F401  (real module; must be imported before any stub exists)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, n):
            return _Any()

    class LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    for name in ("matplotlib", "matplotlib.pyplot", "segmentation_models_pytorch", "skimage", "skimage.morphology",
                 "torchmetrics", "torchmetrics.functional", "torchmetrics.functional.segmentation",
                 "torchmetrics.functional.classification", "torchmetrics.segmentation", "torchmetrics.classification",
                 ):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001
                m = mod(name)
                m.__getattr__ = lambda n: _Any()  # type: ignore
    sys.modules["skimage.morphology"].skeletonize = lambda x: x
    mod("lightning", LightningModule=LightningModule, Trainer=_Any, seed_everything=lambda *a, **k: None)
    mod("lightning.pytorch")
    mod("lightning.pytorch.callbacks", EarlyStopping=_Any, ModelCheckpoint=_Any)
    mod("lightning.pytorch.loggers", CSVLogger=_Any, TensorBoardLogger=_Any)


def load_reference(which: str):
    """imports /root/reference/model/<which>/classes.py as a module object."""
    _install_stubs()
    path = os.path.join(REF, "model", which)
    for m in ("classes", "functions", "segmentation"):
        sys.modules.pop(m, None)
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module("classes")
    finally:
        sys.path.remove(path)
    sys.modules.pop("classes", None)
    return mod


def _set_image_size(ref_model, cfg: O.OracleConfig):
    """the reference hard-codes image_size=224; for other sizes rebuild its backbone with the same
    ViTConfig and the one-field delta."""
    if cfg.image_size == 224:
        return
    from transformers import ViTConfig, ViTModel
    c = ref_model.backbone.config.to_dict()
    c["image_size"] = cfg.image_size
    ref_model.backbone = ViTModel(ViTConfig(**c))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ce = load_reference("CE")
    paed = load_reference("PAED")
    import builtins
    real_print = builtins.print

    # ------------------------------------------------------------------ tiny model: every output + gradients
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, seed=7, head_gain=4.0)
    x = O.synthetic_images(2, 224, seed=11)
    y256 = O.synthetic_labels(2, 17, seed=12)

    ref = ce.LightningViTModel(17, 16, 128, 2, 2)
    ref.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    ref.eval()  # dropout off (parity protocol); eval() does not change any other arithmetic
    captured = {}
    ref.model.seg_head.register_forward_hook(lambda m, i, o: captured.__setitem__("low", o.detach().clone()))
    out = {"cfg": cfg.__dict__, "weights_seed": 7, "head_gain": 4.0, "image_seed": 11, "label_seed": 12}
    with torch.no_grad():
        logits = ref(x)
    out["low"] = captured["low"]
    out["logits_sub"] = logits[:, :, ::7, ::7].clone()
    out["logits_sum"] = logits.double().sum().item()
    out["logits_abs_sum"] = logits.double().abs().sum().item()
    out["resized_labels_sub"] = ref._resize_target(y256, size=(224, 224))[:, ::5, ::5].clone()
    grad_keys = ["model.seg_head.2.weight", "model.seg_head.0.bias", "model.backbone.embeddings.cls_token",
                 "model.backbone.encoder.layer.0.attention.attention.query.bias",
                 "model.backbone.encoder.layer.1.output.dense.bias", "model.backbone.layernorm.weight",
                 "model.backbone.embeddings.patch_embeddings.projection.bias"]

    def grads_of(module, loss):
        module.zero_grad()
        loss.backward()
        named = dict(module.named_parameters())
        g = {k: named[k].grad.detach().clone() for k in grad_keys}
        g["__total_sq__"] = sum(float((p.grad.double() ** 2).sum()) for p in module.parameters() if p.grad is not None)
        g["__pooler_has_grad__"] = named["model.backbone.pooler.dense.weight"].grad is not None
        return g

    loss = ref.training_step((x, y256), 0)
    out["ce_loss"] = loss.item()
    out["ce_grads"] = grads_of(ref, loss)

    # PAED multi-class flavour (PAED/classes.py:415-487)
    builtins.print = lambda *a, **k: None  # the reference prints label stats every step (PAED:454)
    try:
        refp = paed.LightningViTModel(17, 16, 128, 2, 2)
        refp.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
        refp.eval()
        loss = refp.training_step((x, y256), 0)
    finally:
        builtins.print = real_print
    out["paed_multi_loss"] = loss.item()
    out["paed_multi_grads"] = grads_of(refp, loss)

    # free function on dense tensors
    gen = torch.Generator().manual_seed(5)
    pm = torch.softmax(torch.randn(2, 5, 64, 64, generator=gen), 1)
    mk = torch.nn.functional.one_hot(torch.randint(0, 5, (2, 64, 64), generator=gen), 5).permute(0, 3, 1, 2).float()
    out["dense_seed"] = 5
    out["dense_loss_cp"] = paed.paed_loss_multiclass_soft(mk, pm, num_classes=5).item()
    out["dense_loss_nocp"] = paed.paed_loss_multiclass_soft(mk, pm, num_classes=5, class_penalty=False).item()

    # PAEDTrainer (binary), C = 1
    cfg1 = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd1 = O.seeded_state_dict(cfg1, seed=8, head_gain=8.0)
    masks, sdf_e, sdf_i = O.synthetic_binary_targets(2, 224, seed=13)
    reft = paed.PAEDTrainer(1, 16, 128, 2, 2)
    reft.load_state_dict(O.to_module_state_dict(sd1, "model."), strict=True)
    reft.eval()
    # logging-only metrics need torchmetrics (absent): neutralise them, the loss does not depend on them
    zero = lambda *a, **k: torch.zeros(1)  # noqa: E731
    paed.segmentation_metrics = types.SimpleNamespace(mean_iou=zero)
    paed.classification_metrics = types.SimpleNamespace(precision=zero, recall=zero)
    loss = reft.training_step((x, masks, sdf_e, sdf_i), 0)
    out["paed_bin"] = {"weights_seed": 8, "head_gain": 8.0, "target_seed": 13, "loss": loss.item(),
                       "grads": grads_of(reft, loss)}
    # compute_sdf pin
    seg = sys.modules.get("segmentation") or importlib.import_module("segmentation")
    m0 = masks[0].numpy().astype(np.uint8)
    e0, i0 = seg.compute_sdf(m0)
    out["sdf_pin"] = {"ext_sub": torch.from_numpy(e0[::9, ::9].copy()), "int_sub": torch.from_numpy(i0[::9, ::9].copy())}
    torch.save(out, os.path.join(GOLD, "tiny_p16h128.pt"))
    print("tiny golden written: ce_loss", out["ce_loss"], "paed_multi", out["paed_multi_loss"], "paed_bin",
          out["paed_bin"]["loss"])

    # ------------------------------------------------------------------ ViT-B/16 (headline config), forward only
    cfgb = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
    sdb = O.seeded_state_dict(cfgb, seed=0, head_gain=4.0)
    xb = O.synthetic_images(2, 224, seed=1234)
    refb = ce.LightningViTModel(17, 16, 768, 12, 12)
    refb.load_state_dict(O.to_module_state_dict(sdb, "model."), strict=True)
    refb.eval()
    cap = {}
    refb.model.seg_head.register_forward_hook(lambda m, i, o: cap.__setitem__("low", o.detach().clone()))
    with torch.no_grad():
        lg = refb(xb)
    yb = O.synthetic_labels(2, 17, seed=1235)
    with torch.no_grad():
        lb = refb.loss_fn(lg, refb._resize_target(yb, size=(224, 224))).item()
    torch.save({"cfg": cfgb.__dict__, "weights_seed": 0, "head_gain": 4.0, "image_seed": 1234, "label_seed": 1235,
                "low": cap["low"], "logits_sub": lg[:, :, ::7, ::7].clone(), "ce_loss": lb,
                "argmax_sub": lg.argmax(1)[:, ::3, ::3].to(torch.uint8)}, os.path.join(GOLD, "vitb16.pt"))
    print("vitb16 golden written: ce", lb)

    # ------------------------------------------------------------------ non-224 image size, tiny width
    cfgs = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                          image_size=384)
    sds = O.seeded_state_dict(cfgs, seed=9, head_gain=4.0)
    xs = O.synthetic_images(1, 384, seed=21)
    refs = ce.ViTSegmentationModel(17, 16, 128, 1, 2)
    _set_image_size(refs, cfgs)
    refs.load_state_dict(O.to_module_state_dict(sds), strict=True)
    refs.eval()
    with torch.no_grad():
        ls = refs(xs)
    torch.save({"cfg": cfgs.__dict__, "weights_seed": 9, "head_gain": 4.0, "image_seed": 21,
                "logits_sub": ls[:, :, ::11, ::11].clone()}, os.path.join(GOLD, "tiny_s384.pt"))
    print("s384 golden written")


# ---------------------------------------------------------------------------------------------------------------------
# headline-size TRAINING-STEP goldens (round 2): ViT-B/16 (768/12/12) for the three training wrappers, and the
# reference's own PAED configuration, PAEDTrainer(patch 8, hidden 1024, 16 layers, 16 heads) (PAED/ViTscript.py:66)
# ---------------------------------------------------------------------------------------------------------------------
def _pinned_grads(module, loss, L):
    module.zero_grad()
    loss.backward()
    named = dict(module.named_parameters())
    g = {}
    for k, idx in O.headline_grad_pins(L):
        g[k] = named[k].grad.detach()[idx].clone()
    g["__total_sq__"] = sum(float((p.grad.double() ** 2).sum()) for p in module.parameters() if p.grad is not None)
    g["__pooler_has_grad__"] = named["model.backbone.pooler.dense.weight"].grad is not None
    return g


def headline():
    import builtins
    import time
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    ce = load_reference("CE")
    paed = load_reference("PAED")
    real_print = builtins.print
    zero = lambda *a, **k: torch.zeros(1)  # noqa: E731
    paed.segmentation_metrics = types.SimpleNamespace(mean_iou=zero)
    paed.classification_metrics = types.SimpleNamespace(precision=zero, recall=zero)

    # ---- ViT-B/16, B = 2: CE, multi-class PAED (C = 17), PAEDTrainer (C = 1)
    t0 = time.time()
    cfgb = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
    sdb = O.seeded_state_dict(cfgb, seed=0, head_gain=4.0)
    xb = O.synthetic_images(2, 224, seed=1234)
    yb = O.synthetic_labels(2, 17, seed=1235)
    out = {"cfg": cfgb.__dict__, "weights_seed": 0, "head_gain": 4.0, "image_seed": 1234, "label_seed": 1235}
    ref = ce.LightningViTModel(17, 16, 768, 12, 12)
    ref.load_state_dict(O.to_module_state_dict(sdb, "model."), strict=True)
    ref.eval()   # dropout off (parity protocol)
    loss = ref.training_step((xb, yb), 0)
    out["ce_loss"] = loss.item()
    out["ce_grads"] = _pinned_grads(ref, loss, 12)
    del ref
    builtins.print = lambda *a, **k: None
    try:
        refp = paed.LightningViTModel(17, 16, 768, 12, 12)
        refp.load_state_dict(O.to_module_state_dict(sdb, "model."), strict=True)
        refp.eval()
        loss = refp.training_step((xb, yb), 0)
    finally:
        builtins.print = real_print
    out["paed_multi_loss"] = loss.item()
    out["paed_multi_grads"] = _pinned_grads(refp, loss, 12)
    del refp
    cfg1 = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
    sd1 = O.seeded_state_dict(cfg1, seed=3, head_gain=8.0)
    masks, sdf_e, sdf_i = O.synthetic_binary_targets(2, 224, seed=1236)
    reft = paed.PAEDTrainer(1, 16, 768, 12, 12)
    reft.load_state_dict(O.to_module_state_dict(sd1, "model."), strict=True)
    reft.eval()
    loss = reft.training_step((xb, masks, sdf_e, sdf_i), 0)
    out["paed_bin"] = {"weights_seed": 3, "head_gain": 8.0, "target_seed": 1236, "loss": loss.item(),
                       "grads": _pinned_grads(reft, loss, 12)}
    del reft
    torch.save(out, os.path.join(GOLD, "vitb16_train.pt"))
    print(f"vitb16_train golden written ({time.time() - t0:.0f} s): ce {out['ce_loss']:.6f} paed_multi "
          f"{out['paed_multi_loss']:.6f} paed_bin {out['paed_bin']['loss']:.6f}")

    # ---- the reference's PAED model: PAEDTrainer(num_classes=1, patch 8, hidden 1024, 16 layers, 16 heads), B = 1
    t0 = time.time()
    cfg8 = O.OracleConfig(num_classes=1, patch_size=8, hidden_size=1024, num_hidden_layers=16, num_attention_heads=16)
    sd8 = O.seeded_state_dict(cfg8, seed=5, head_gain=8.0)
    x8 = O.synthetic_images(1, 224, seed=1237)
    m8, e8, i8 = O.synthetic_binary_targets(1, 224, seed=1238)
    ref8 = paed.PAEDTrainer(num_classes=1, patch_size=8, hidden_size=1024, num_hidden_layers=16, num_attention_heads=16)
    ref8.load_state_dict(O.to_module_state_dict(sd8, "model."), strict=True)
    ref8.eval()
    cap = {}
    ref8.model.seg_head.register_forward_hook(lambda m, i, o: cap.__setitem__("low", o.detach().clone()))
    loss = ref8.training_step((x8, m8, e8, i8), 0)
    o8 = {"cfg": cfg8.__dict__, "weights_seed": 5, "head_gain": 8.0, "image_seed": 1237, "target_seed": 1238,
          "low": cap["low"], "loss": loss.item(), "grads": _pinned_grads(ref8, loss, 16)}
    torch.save(o8, os.path.join(GOLD, "p8w1024_train.pt"))
    print(f"p8w1024_train golden written ({time.time() - t0:.0f} s): paed_bin {o8['loss']:.6f}")



def curve(steps: int = 200):
    """200 optimisation steps of the UNMODIFIED reference CE module (model/CE/classes.py:276-297) on a fixed batch,
    dropout off: the loss curves the CUDA path has to follow (north_star: within 1 %).
      'adam_lr1e-5': the reference's own optimizer, ref.configure_optimizers() = Adam(lr=1e-5);
      'adam_lr3e-6': the same optimizer object with lr = 3e-6 (a regime without Adam's sign-descent loss spikes, where a
                     per-step 1 % bound is meaningful: see tests/test_headline_gpu.py).
    Weights: fp32 random init NOT snapped to bf16 (north_star: 'random-init weights').  Snapping the initial weights to
    the bf16 grid, as the single-step parity fixtures do, makes any bf16-operand implementation lag the fp32 run by
    several steps at these learning rates (every weight must first travel half a bf16 ulp before its rounded copy
    moves), which is an artefact of the fixture, not of the implementation."""
    import time
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    ce = load_reference("CE")
    cfgb = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
    task = dict(block=32, noise=0.5, contrast=0.8)
    out = {"cfg": cfgb.__dict__, "weights_seed": 31, "head_gain": 1.0, "bf16_representable": False, "steps": steps,
           "task": task, "curves": {}}
    for name, lr in (("adam_lr1e-5", None), ("adam_lr3e-6", 3e-6)):
        t0 = time.time()
        sd = O.seeded_state_dict(cfgb, seed=31, head_gain=1.0, bf16_representable=False)
        x, y = O.curve_task("regions", **task)
        ref = ce.LightningViTModel(17, 16, 768, 12, 12)
        ref.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
        ref.eval()   # dropout off; training_step / backward / optimizer are unaffected by eval()
        opt = ref.configure_optimizers()
        if lr is not None:
            for g in opt.param_groups:
                g["lr"] = lr
        losses = []
        for i in range(steps):
            loss = ref.training_step((x, y), i)     # labels already 224x224: _resize_target is the identity
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(loss.item())
            if i % 20 == 0:
                print(name, i, losses[-1], f"{time.time() - t0:.0f}s", flush=True)
        out["curves"][name] = {"lr": opt.param_groups[0]["lr"], "loss": torch.tensor(losses, dtype=torch.float64)}
        del ref, opt
    torch.save(out, os.path.join(GOLD, "vitb16_curve.pt"))
    print("vitb16_curve golden written")


if __name__ == "__main__":
    if "--curve" in sys.argv:
        curve()
    elif "--headline" in sys.argv:
        headline()
    else:
        main()
        headline()
        curve()
