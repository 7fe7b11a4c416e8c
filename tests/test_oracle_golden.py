"""CPU: the oracle restatement (oracle/vitseg_oracle.py) against the golden vectors minted from the live reference
(oracle/make_golden.py, run where /root/reference is mounted).  Tolerances: fp32 CPU vs fp32 CPU, different op
order only -> 2e-5 relative on logits/losses, 1e-4 on gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import vitseg_oracle as O

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return torch.load(os.path.join(golden_dir, "tiny_p16h128.pt"), weights_only=False)


def _grads(sd, loss_fn, keys):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = loss_fn(leaves)
    loss.backward()
    total = sum(float((v.grad.double() ** 2).sum()) for v in leaves.values() if v.grad is not None)
    return loss.item(), {k: leaves[k[len("model."):]].grad for k in keys}, total, leaves


def test_forward_matches_reference(tiny):
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"])
    with torch.no_grad():
        low = O.forward_lowres(sd, x, cfg)
        full = O.upsample(low, 224)
    assert _rel(low, tiny["low"]) < 2e-5
    assert _rel(full[:, :, ::7, ::7], tiny["logits_sub"]) < 2e-5
    assert abs(full.double().sum().item() - tiny["logits_sum"]) < 1e-4 * tiny["logits_abs_sum"]


def test_resize_target_matches_reference(tiny):
    y = O.synthetic_labels(2, 17, seed=tiny["label_seed"])
    assert torch.equal(O.resize_target(y, 224)[:, ::5, ::5], tiny["resized_labels_sub"])


def test_ce_loss_and_grads(tiny):
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"])
    y = O.resize_target(O.synthetic_labels(2, 17, seed=tiny["label_seed"]), 224)
    g = tiny["ce_grads"]
    keys = [k for k in g if not k.startswith("__")]
    loss, grads, total, leaves = _grads(sd, lambda p: O.ce_loss(O.forward(p, x, cfg), y), keys)
    assert abs(loss - tiny["ce_loss"]) < 2e-5 * abs(tiny["ce_loss"])
    for k in keys:
        assert _rel(grads[k], g[k]) < 1e-4, k
    assert abs(total - g["__total_sq__"]) < 1e-3 * g["__total_sq__"]
    assert not g["__pooler_has_grad__"]
    assert leaves["backbone.pooler.dense.weight"].grad is None


def test_paed_multiclass_loss_and_grads(tiny):
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"])
    y = O.resize_target(O.synthetic_labels(2, 17, seed=tiny["label_seed"]), 224)
    g = tiny["paed_multi_grads"]
    keys = [k for k in g if not k.startswith("__")]
    loss, grads, total, _ = _grads(sd, lambda p: O.paed_multiclass_step_loss(O.forward(p, x, cfg), y), keys)
    assert abs(loss - tiny["paed_multi_loss"]) < 2e-5 * abs(tiny["paed_multi_loss"])
    for k in keys:
        assert _rel(grads[k], g[k]) < 2e-4, k


def test_paed_dense_function(tiny):
    gen = torch.Generator().manual_seed(tiny["dense_seed"])
    pm = torch.softmax(torch.randn(2, 5, 64, 64, generator=gen), 1)
    mk = torch.nn.functional.one_hot(torch.randint(0, 5, (2, 64, 64), generator=gen), 5).permute(0, 3, 1, 2).float()
    assert abs(O.paed_loss_multiclass_soft(mk, pm).item() - tiny["dense_loss_cp"]) < 1e-6
    assert abs(O.paed_loss_multiclass_soft(mk, pm, class_penalty=False).item() - tiny["dense_loss_nocp"]) < 1e-6


def test_paed_binary_loss_and_grads(tiny):
    pb = tiny["paed_bin"]
    cfg = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, pb["weights_seed"], head_gain=pb["head_gain"])
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"])
    masks, se, si = O.synthetic_binary_targets(2, 224, seed=pb["target_seed"])
    g = pb["grads"]
    keys = [k for k in g if not k.startswith("__")]
    loss, grads, total, _ = _grads(
        sd, lambda p: O.paed_binary_step_loss(O.forward(p, x, cfg), O.resize_target(masks, 224), se, si), keys)
    assert abs(loss - pb["loss"]) < 2e-5 * abs(pb["loss"])
    for k in keys:
        assert _rel(grads[k], g[k]) < 2e-4, k


def test_compute_sdf_matches_reference(tiny):
    masks, _, _ = O.synthetic_binary_targets(2, 224, seed=tiny["paed_bin"]["target_seed"])
    e, i = O.compute_sdf(masks[0].numpy().astype(np.uint8))
    assert np.array_equal(e[::9, ::9], tiny["sdf_pin"]["ext_sub"].numpy())
    assert np.array_equal(i[::9, ::9], tiny["sdf_pin"]["int_sub"].numpy())


def test_vitb16_forward_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "vitb16.pt"), weights_only=False)
    cfg = O.OracleConfig(**gold["cfg"])
    sd = O.seeded_state_dict(cfg, gold["weights_seed"], head_gain=gold["head_gain"])
    x = O.synthetic_images(2, 224, seed=gold["image_seed"])
    with torch.no_grad():
        low = O.forward_lowres(sd, x, cfg)
        full = O.upsample(low, 224)
        y = O.resize_target(O.synthetic_labels(2, 17, seed=gold["label_seed"]), 224)
        loss = O.ce_loss(full, y).item()
    assert _rel(low, gold["low"]) < 5e-5
    assert _rel(full[:, :, ::7, ::7], gold["logits_sub"]) < 5e-5
    assert abs(loss - gold["ce_loss"]) < 1e-5 * gold["ce_loss"]
    agree = (full.argmax(1)[:, ::3, ::3].to(torch.uint8) == gold["argmax_sub"]).float().mean().item()
    assert agree > 0.9995


def test_image_size_384_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "tiny_s384.pt"), weights_only=False)
    cfg = O.OracleConfig(**gold["cfg"])
    sd = O.seeded_state_dict(cfg, gold["weights_seed"], head_gain=gold["head_gain"])
    x = O.synthetic_images(1, 384, seed=gold["image_seed"])
    with torch.no_grad():
        full = O.forward(sd, x, cfg)
    assert _rel(full[:, :, ::11, ::11], gold["logits_sub"]) < 2e-5


# ---------------------------------------------------------------------------------------------------------------------
# headline-size training-step goldens (round 2): the reference's loss and pinned gradient slices at ViT-B/16 and at the
# reference's own PAED configuration (patch 8, hidden 1024, 16 layers, 16 heads: PAED/ViTscript.py:66)
# ---------------------------------------------------------------------------------------------------------------------
def _pinned(sd, loss_fn, L, gold):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = loss_fn(leaves)
    loss.backward()
    for k, idx in O.headline_grad_pins(L):
        got = leaves[k[len("model."):]].grad[idx]
        assert got.shape == gold[k].shape, k
        assert _rel(got, gold[k]) < 3e-4, (k, _rel(got, gold[k]))
    total = sum(float((v.grad.double() ** 2).sum()) for v in leaves.values() if v.grad is not None)
    assert abs(total - gold["__total_sq__"]) < 1e-3 * gold["__total_sq__"]
    assert not gold["__pooler_has_grad__"] and leaves["backbone.pooler.dense.weight"].grad is None
    return loss.item()


@pytest.fixture(scope="module")
def vitb_train(golden_dir):
    return torch.load(os.path.join(golden_dir, "vitb16_train.pt"), weights_only=False)


def test_vitb16_ce_training_step_matches_reference(vitb_train):
    g = vitb_train
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.seeded_state_dict(cfg, g["weights_seed"], head_gain=g["head_gain"])
    x = O.synthetic_images(2, 224, seed=g["image_seed"])
    y = O.resize_target(O.synthetic_labels(2, 17, seed=g["label_seed"]), 224)
    loss = _pinned(sd, lambda p: O.ce_loss(O.forward(p, x, cfg), y), 12, g["ce_grads"])
    assert abs(loss - g["ce_loss"]) < 2e-5 * abs(g["ce_loss"])


def test_vitb16_paed_multiclass_training_step_matches_reference(vitb_train):
    g = vitb_train
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.seeded_state_dict(cfg, g["weights_seed"], head_gain=g["head_gain"])
    x = O.synthetic_images(2, 224, seed=g["image_seed"])
    y = O.resize_target(O.synthetic_labels(2, 17, seed=g["label_seed"]), 224)
    loss = _pinned(sd, lambda p: O.paed_multiclass_step_loss(O.forward(p, x, cfg), y), 12, g["paed_multi_grads"])
    assert abs(loss - g["paed_multi_loss"]) < 5e-5 * abs(g["paed_multi_loss"])


def test_vitb16_paed_trainer_step_matches_reference(vitb_train):
    g, pb = vitb_train, vitb_train["paed_bin"]
    cfg = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
    sd = O.seeded_state_dict(cfg, pb["weights_seed"], head_gain=pb["head_gain"])
    x = O.synthetic_images(2, 224, seed=g["image_seed"])
    masks, se, si = O.synthetic_binary_targets(2, 224, seed=pb["target_seed"])
    loss = _pinned(sd, lambda p: O.paed_binary_step_loss(O.forward(p, x, cfg), O.resize_target(masks, 224), se, si),
                   12, pb["grads"])
    assert abs(loss - pb["loss"]) < 2e-5 * abs(pb["loss"])


def test_p8_w1024_paed_trainer_step_matches_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "p8w1024_train.pt"), weights_only=False)
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.seeded_state_dict(cfg, g["weights_seed"], head_gain=g["head_gain"])
    x = O.synthetic_images(1, 224, seed=g["image_seed"])
    masks, se, si = O.synthetic_binary_targets(1, 224, seed=g["target_seed"])
    with torch.no_grad():
        assert _rel(O.forward_lowres(sd, x, cfg), g["low"]) < 5e-5
    loss = _pinned(sd, lambda p: O.paed_binary_step_loss(O.forward(p, x, cfg), O.resize_target(masks, 224), se, si),
                   16, g["grads"])
    assert abs(loss - g["loss"]) < 2e-5 * abs(g["loss"])
