"""GPU (-m gpu): surface behaviour a user of the reference relies on beyond one training step — optimizer
checkpoints, frozen parameters, parameter hooks, the PAEDTrainer helper methods (model/PAED/classes.py:524,608,623)."""
import copy

import pytest
import torch

from oracle import vitseg_oracle as O

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _build(cls, cfg, sd, dev):
    m = cls(cfg.num_classes, cfg.patch_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    return m.to(dev).train()


CFG = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)


def _batch(dev, seed=5):
    return O.synthetic_images(2, 224, seed=seed).to(dev), O.synthetic_labels(2, 17, seed=seed + 1, size=224).to(dev)


def _steps(m, opt, batch, n):
    for _ in range(n):
        m._loss(*batch).backward()
        opt.step()
        opt.zero_grad()


def test_fused_adam_checkpoint_resume_and_torch_adam_interchange():
    """save -> load -> continue gives the trajectory of an uninterrupted run (bit-identical: same kernels, same
    state), and the checkpoint is in torch.optim.Adam's own format: torch Adam loads it and FusedAdam loads torch's."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.optim import FusedAdam
    dev = _dev()
    sd = O.seeded_state_dict(CFG, 101, head_gain=2.0)
    batch = _batch(dev)
    a = _build(LightningViTModel, CFG, sd, dev)
    oa = FusedAdam(a, lr=1e-3)
    assert oa.state_dict()["state"] == {}            # lazily created, like torch
    _steps(a, oa, batch, 3)
    ck_model = copy.deepcopy(a.state_dict())
    ck_opt = copy.deepcopy(oa.state_dict())
    st = ck_opt["state"]
    n_trainable = sum(1 for k, _ in a.named_parameters() if "pooler" not in k)
    assert len(st) == n_trainable
    some = next(iter(st.values()))
    assert set(some) == {"step", "exp_avg", "exp_avg_sq"} and float(some["step"]) == 3.0
    assert any(float(v["exp_avg"].abs().max()) > 0 for v in st.values())
    _steps(a, oa, batch, 2)                          # uninterrupted: steps 4, 5
    # resumed replica
    b = _build(LightningViTModel, CFG, sd, dev)
    b.load_state_dict(ck_model, strict=True)
    ob = FusedAdam(b, lr=1e-3)
    ob.load_state_dict(ck_opt)
    _steps(b, ob, batch, 2)
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        # gradients differ run to run at the last bit (atomics order): Adam turns that into <= ~lr-sized deviations only
        # for near-zero gradients; a resume that lost the moments would be off by O(lr * sqrt(steps)) everywhere
        assert torch.allclose(pa, pb, rtol=0, atol=2e-4), (k, (pa - pb).abs().max().item())
    lost = _build(LightningViTModel, CFG, sd, dev)
    lost.load_state_dict(ck_model, strict=True)
    ol = FusedAdam(lost, lr=1e-3)                    # no optimizer state: must NOT match (the check has teeth)
    _steps(lost, ol, batch, 2)
    dev_lost = max((pa - pl).abs().max().item() for (_, pa), (_, pl) in zip(a.named_parameters(), lost.named_parameters()))
    dev_ok = max((pa - pb).abs().max().item() for (_, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()))
    assert dev_lost > 5 * max(dev_ok, 1e-6), (dev_lost, dev_ok)
    # interchange with torch.optim.Adam (same parameter order; torch holds no state for the gradient-less pooler)
    c = _build(LightningViTModel, CFG, sd, dev)
    c.load_state_dict(ck_model, strict=True)
    oc = torch.optim.Adam(c.parameters(), lr=1e-3)
    ck_for_torch = copy.deepcopy(ck_opt)
    # torch's parameter list also contains the two pooler tensors: re-index our state onto its numbering
    names_t = [k for k, _ in c.named_parameters()]
    names_f = [k for k, p in b.named_parameters() if p.requires_grad]
    assert names_t == names_f                        # FusedAdam owns the pooler too (it just never updates it)
    for g in ck_for_torch["param_groups"]:
        g.pop("decoupled", None)
        for k, v in oc.state_dict()["param_groups"][0].items():
            g.setdefault(k, v)
    oc.load_state_dict(ck_for_torch)
    _steps(c, oc, batch, 2)
    for (k, pa), (_, pc) in zip(a.named_parameters(), c.named_parameters()):
        assert torch.allclose(pa, pc, rtol=0, atol=2e-4), (k, (pa - pc).abs().max().item())
    d = _build(LightningViTModel, CFG, sd, dev)
    d.load_state_dict(c.state_dict(), strict=True)
    od = FusedAdam(d, lr=1e-3)
    od.load_state_dict(oc.state_dict())              # torch -> fused
    assert int(od._step.item()) == 5
    _steps(d, od, batch, 1)
    _steps(a, oa, batch, 1)
    for (k, pa), (_, pd) in zip(a.named_parameters(), d.named_parameters()):
        assert torch.allclose(pa, pd, rtol=0, atol=3e-4), (k, (pa - pd).abs().max().item())


def test_frozen_parameters_get_no_grad_and_no_update():
    """requires_grad=False (fine-tuning only the head): no .grad appears on frozen parameters, FusedAdam with weight
    decay leaves them bit-identical, trainable ones still move and match a run of torch.optim.AdamW."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.optim import FusedAdamW
    dev = _dev()
    sd = O.seeded_state_dict(CFG, 111, head_gain=2.0)
    batch = _batch(dev, 7)

    def make():
        m = _build(LightningViTModel, CFG, sd, dev)
        for k, p in m.named_parameters():
            if "seg_head" not in k and "encoder.layer.1." not in k:
                p.requires_grad_(False)
        return m

    a, b = make(), make()
    oa = FusedAdamW(a, lr=1e-3, weight_decay=0.1)
    ob = torch.optim.AdamW([p for p in b.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.1)
    before = {k: p.detach().clone() for k, p in a.named_parameters()}
    for _ in range(3):
        a._loss(*batch).backward()
        for k, p in a.named_parameters():
            assert (p.grad is None) == (not p.requires_grad or "pooler" in k), k
        b.zero_grad(set_to_none=True)
        for (_, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            pb.grad = None if pa.grad is None else pa.grad.clone()
        oa.step()
        oa.zero_grad()
        ob.step()
    moved = 0
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        if pa.requires_grad:
            assert torch.allclose(pa, pb, rtol=2e-5, atol=3e-6), (k, (pa - pb).abs().max().item())
            moved += int(not torch.equal(pa, before[k]))
        else:
            assert torch.equal(pa, before[k]), k
    assert moved > 10
    # the frozen weights' bf16 shadows are intact: inference still matches the oracle on the current weights
    a.eval()
    now = {k[len("model."):]: v.detach().cpu().clone() for k, v in a.state_dict().items()}
    x = batch[0]
    with torch.no_grad():
        ours, ref = a(x).cpu(), O.forward(now, x.cpu(), CFG)
    assert ((ours - ref).abs().max() / ref.abs().max()).item() < 1e-2


def test_post_accumulate_grad_hooks_fire_once_per_backward():
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    m = _build(LightningViTModel, CFG, O.seeded_state_dict(CFG, 121), dev)
    seen = []
    p = dict(m.named_parameters())["model.backbone.encoder.layer.0.output.dense.weight"]
    p.register_post_accumulate_grad_hook(lambda t: seen.append(float(t.grad.abs().sum())))
    m._loss(*_batch(dev, 9)).backward()
    assert len(seen) == 1 and seen[0] > 0


def test_no_grad_forward_in_train_mode_applies_dropout():
    """the reference's modules drop out whenever .training is set, with or without autograd."""
    from visiontransformer_b200.ce.classes import ViTSegmentationModel
    dev = _dev()
    m = ViTSegmentationModel(17, 16, 128, 2, 2).to(dev)
    x = _batch(dev, 11)[0]
    with torch.no_grad():
        m.eval()
        e1, e2 = m(x), m(x)
        m.train()
        t1, t2 = m(x), m(x)
    assert torch.equal(e1, e2)
    assert not torch.equal(t1, t2) and not torch.equal(t1, e1)


def test_paed_trainer_helper_methods_and_test_step():
    from visiontransformer_b200.paed.classes import PAEDTrainer
    dev = _dev()
    cfg = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    m = _build(PAEDTrainer, cfg, O.seeded_state_dict(cfg, 131, head_gain=8.0), dev)
    x = O.synthetic_images(2, 224, seed=132).to(dev)
    masks, se, si = O.synthetic_binary_targets(2, 224, seed=133)
    g = torch.Generator().manual_seed(134)
    preds = torch.rand(2, 1, 224, 224, generator=g)
    # dense helper methods against the oracle restatement of the reference lines, values and gradients
    pa = preds.to(dev).requires_grad_(True)
    pr = preds.clone().requires_grad_(True)
    la = m.dice_loss(pa, masks.to(dev)) + m.paed_loss_soft(se.unsqueeze(1).to(dev), si.unsqueeze(1).to(dev), pa)
    lr_ = O.dice_loss(pr, masks.unsqueeze(1)) + O.paed_loss_soft(se.unsqueeze(1), si.unsqueeze(1), pr)
    assert abs(la.item() - lr_.item()) < 1e-5 * abs(lr_.item()) + 1e-7
    la.backward()
    lr_.backward()
    assert ((pa.grad.cpu() - pr.grad).abs().max() / pr.grad.abs().max()).item() < 5e-4   # torch CUDA vs torch CPU, fp32
    # a subclass written like the reference's own step (dense tensors + the helper methods) reproduces the fused loss
    logits = m(x)
    p = torch.sigmoid(logits)
    mk = m._resize_target(masks.to(dev), size=(224, 224)).unsqueeze(1).float()
    dense = torch.nn.functional.binary_cross_entropy(p, mk) + 0.1 * m.dice_loss(p, mk) + 5.0 * torch.abs(
        m.paed_loss_soft(se.unsqueeze(1).to(dev), si.unsqueeze(1).to(dev), p))
    fused = m.training_step((x, masks.to(dev), se.to(dev), si.to(dev)), 0)
    assert abs(dense.item() - fused.item()) < 1e-4 * abs(fused.item())
    out = m.test_step((x, masks.to(dev), se.to(dev), si.to(dev)), 0)
    assert set(out) == {"test_acc", "test_IoU", "test_recall", "test_dice", "test_precision"}
    assert all(0.0 <= float(v) <= 1.0 for v in out.values())
