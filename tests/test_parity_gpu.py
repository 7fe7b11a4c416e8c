"""GPU (-m gpu): module-level parity of the CUDA path against the CPU oracle and the committed golden vectors.

Protocol: oracle and candidate share bf16-representable weights and inputs; the oracle computes in
fp32 on the CPU; the candidate uses bf16 tensor-core operands with fp32 accumulation, fp32 residual stream, LayerNorm,
softmax and loss math.  Tolerances from BASELINE.json north_star: logits max|d|/max|ref| <= 1e-2; argmax agreement
>= 99.9 % (on pixels whose fp32 top-2 margin exceeds 2x the measured max|d|, plus the raw figure); loss within 1 %."""
import os

import pytest
import torch

from oracle import vitseg_oracle as O

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-2
GRAD_TOL = 3e-2


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _relmax(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _build(cls, cfg, sd, dev, prefix="model.", **kw):
    m = cls(cfg.num_classes, cfg.patch_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
            image_size=cfg.image_size, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)
    m.load_state_dict(O.to_module_state_dict(sd, prefix), strict=True)
    return m.to(dev)


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return torch.load(os.path.join(golden_dir, "tiny_p16h128.pt"), weights_only=False)


def test_tiny_forward_vs_golden_and_oracle(tiny):
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).eval()
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"])
    with torch.no_grad():
        low = m.model.forward_lowres(x.to(dev))
        full = m(x.to(dev))
    assert full.shape == (2, 17, 224, 224) and full.dtype == torch.float32
    assert _relmax(low, tiny["low"]) < LOGIT_TOL
    assert _relmax(full[:, :, ::7, ::7], tiny["logits_sub"]) < LOGIT_TOL
    with torch.no_grad():
        ref = O.forward(sd, x, cfg)
    assert _relmax(full, ref) < LOGIT_TOL


def _check_grads(module, gold_grads, oracle_leaves=None):
    named = dict(module.named_parameters())
    for k, g in gold_grads.items():
        if k.startswith("__"):
            continue
        assert named[k].grad is not None, k
        assert _relmax(named[k].grad, g) < GRAD_TOL, (k, _relmax(named[k].grad, g))
    total = sum(float((p.grad.double() ** 2).sum()) for p in module.parameters() if p.grad is not None)
    assert abs(total - gold_grads["__total_sq__"]) < 3e-2 * gold_grads["__total_sq__"]
    assert named["model.backbone.pooler.dense.weight"].grad is None
    assert named["model.backbone.pooler.dense.bias"].grad is None


def test_tiny_ce_training_step_vs_golden(tiny):
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).train()
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"]).to(dev)
    y = O.synthetic_labels(2, 17, seed=tiny["label_seed"]).to(dev)
    loss = m.training_step((x, y), 0)
    assert abs(loss.item() - tiny["ce_loss"]) < 1e-2 * abs(tiny["ce_loss"])
    loss.backward()
    _check_grads(m, tiny["ce_grads"])


def test_generic_loss_on_full_logits_matches_fused_path(tiny):
    """model(x) -> nn.CrossEntropyLoss (the reference's literal call sequence) gives the same loss / gradients as the
    fused low-res path."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).train()
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"]).to(dev)
    y = m._resize_target(O.synthetic_labels(2, 17, seed=tiny["label_seed"]).to(dev), (224, 224))
    loss = m.loss_fn(m(x), y)
    assert abs(loss.item() - tiny["ce_loss"]) < 1e-2 * abs(tiny["ce_loss"])
    loss.backward()
    _check_grads(m, tiny["ce_grads"])


def test_tiny_paed_multiclass_step_vs_golden(tiny):
    from visiontransformer_b200.paed.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).train()
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"]).to(dev)
    y = O.synthetic_labels(2, 17, seed=tiny["label_seed"]).to(dev)
    loss = m.training_step((x, y), 0)
    assert abs(loss.item() - tiny["paed_multi_loss"]) < 1e-2 * abs(tiny["paed_multi_loss"])
    loss.backward()
    _check_grads(m, tiny["paed_multi_grads"])


def test_paed_dense_function_vs_golden(tiny):
    from visiontransformer_b200.paed.classes import paed_loss_multiclass_soft
    dev = _dev()
    gen = torch.Generator().manual_seed(tiny["dense_seed"])
    pm = torch.softmax(torch.randn(2, 5, 64, 64, generator=gen), 1)
    mk = torch.nn.functional.one_hot(torch.randint(0, 5, (2, 64, 64), generator=gen), 5).permute(0, 3, 1, 2).float()
    p = pm.to(dev).requires_grad_(True)
    l1 = paed_loss_multiclass_soft(mk.to(dev), p, num_classes=5)
    assert abs(l1.item() - tiny["dense_loss_cp"]) < 1e-5 * abs(tiny["dense_loss_cp"]) + 1e-8
    l1.backward()
    pr = pm.clone().requires_grad_(True)
    O.paed_loss_multiclass_soft(mk, pr).backward()
    assert _relmax(p.grad, pr.grad) < 1e-4
    l2 = paed_loss_multiclass_soft(mk.to(dev), pm.to(dev), num_classes=5, class_penalty=False)
    assert abs(l2.item() - tiny["dense_loss_nocp"]) < 1e-5 * abs(tiny["dense_loss_nocp"])


def test_tiny_paed_binary_step_vs_golden(tiny):
    from visiontransformer_b200.paed.classes import PAEDTrainer
    dev = _dev()
    pb = tiny["paed_bin"]
    cfg = O.OracleConfig(num_classes=1, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, pb["weights_seed"], head_gain=pb["head_gain"])
    m = _build(PAEDTrainer, cfg, sd, dev).train()
    x = O.synthetic_images(2, 224, seed=tiny["image_seed"]).to(dev)
    masks, se, si = [t.to(dev) for t in O.synthetic_binary_targets(2, 224, seed=pb["target_seed"])]
    loss = m.training_step((x, masks, se, si), 0)
    assert abs(loss.item() - pb["loss"]) < 1e-2 * abs(pb["loss"])
    loss.backward()
    _check_grads(m, pb["grads"])


def test_vitb16_logits_and_argmax_vs_golden(golden_dir):
    """BASELINE config 1/2 model: bf16 logits within 1e-2 max relative error; argmax-mask agreement >= 99.9 %."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "vitb16.pt"), weights_only=False)
    cfg = O.OracleConfig(**gold["cfg"])
    sd = O.seeded_state_dict(cfg, gold["weights_seed"], head_gain=gold["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).eval()
    x = O.synthetic_images(2, 224, seed=gold["image_seed"])
    with torch.no_grad():
        low = m.model.forward_lowres(x.to(dev))
        full = m(x.to(dev))
        mask = m.model.predict_mask(x.to(dev))
    err = _relmax(low, gold["low"])
    assert err < LOGIT_TOL, err
    assert _relmax(full[:, :, ::7, ::7], gold["logits_sub"]) < LOGIT_TOL
    # full-tensor comparison against the live oracle on this box's CPU
    with torch.no_grad():
        ref = O.forward(sd, x, cfg)
    d = (full.cpu() - ref).abs().max().item()
    assert d / ref.abs().max().item() < LOGIT_TOL
    ref_arg = ref.argmax(1)
    raw = (full.argmax(1).cpu() == ref_arg).float().mean().item()
    top2 = ref.topk(2, dim=1).values
    confident = (top2[:, 0] - top2[:, 1]) > 2 * d
    conf_agree = (full.argmax(1).cpu() == ref_arg)[confident].float().mean().item()
    print(f"vitb16: logits err {err:.2e}, raw argmax agreement {raw:.5f}, margin-filtered {conf_agree:.5f} "
          f"({confident.float().mean().item():.3f} of pixels)")
    # random-init logits are nearly flat (top-2 margins down to 1e-4), so raw agreement is bounded by near-ties
    # (even fp32-vs-bf16-weights alone gives 99.4-99.9 %); the 99.9 % bar is asserted on pixels whose
    # fp32 margin exceeds twice the measured logit error, and on trained weights in test_ce_loss_curve_tracks_oracle.
    assert conf_agree >= 0.999
    assert raw >= 0.995
    # fused upsample+argmax equals argmax of the materialised logits
    assert (mask.cpu().long() == full.argmax(1).cpu()).float().mean().item() > 0.9999
    y = m._resize_target(O.synthetic_labels(2, 17, seed=gold["label_seed"]).to(dev), (224, 224))
    with torch.no_grad():
        loss = m.loss_fn(full, y).item()
    assert abs(loss - gold["ce_loss"]) < 1e-2 * gold["ce_loss"]


def test_image_size_384_vs_golden(golden_dir):
    from visiontransformer_b200.ce.classes import ViTSegmentationModel
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "tiny_s384.pt"), weights_only=False)
    cfg = O.OracleConfig(**gold["cfg"])
    sd = O.seeded_state_dict(cfg, gold["weights_seed"], head_gain=gold["head_gain"])
    m = _build(ViTSegmentationModel, cfg, sd, dev, prefix="").eval()
    x = O.synthetic_images(1, 384, seed=gold["image_seed"]).to(dev)
    with torch.no_grad():
        full = m(x)
    assert _relmax(full[:, :, ::11, ::11], gold["logits_sub"]) < LOGIT_TOL
    with pytest.raises(ValueError, match="doesn't match model"):
        m(torch.rand(1, 3, 224, 224, device=dev))


def test_wrong_size_and_small_patch_variants():
    """P8 variant (785 tokens) of the reference sweep (datasetTestViTmodel.py:97-107) runs and matches the oracle."""
    from visiontransformer_b200.ce.classes import ViTSegmentationModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=8, hidden_size=128, num_hidden_layers=1, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 21, head_gain=4.0)
    m = _build(ViTSegmentationModel, cfg, sd, dev, prefix="").eval()
    x = O.synthetic_images(1, 224, seed=22)
    with torch.no_grad():
        full = m(x.to(dev))
        ref = O.forward(sd, x, cfg)
    assert _relmax(full, ref) < LOGIT_TOL


def test_ce_loss_curve_tracks_oracle():
    """Loop-level parity: Adam steps on a learnable synthetic task, dropout off; per-step loss within 1 % of the fp32
    oracle trained from the same weights (north_star: loss curve within 1 %).  This is the quick 60-step check on a
    2-layer model; the 200-step ViT-B/16 run is tests/test_headline_gpu.py::test_vitb16_200_step_loss_curve_and_trained_argmax."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 31, head_gain=1.0)
    x = O.synthetic_images(8, 224, seed=32)
    y = O.learnable_labels(x, 17)
    m = _build(LightningViTModel, cfg, sd, dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    used = [v for k, v in leaves.items() if not k.startswith("backbone.pooler")]
    opt_ref = torch.optim.Adam(used, lr=1e-3)
    xg, yg = x.to(dev), y.to(dev)
    worst, first, last = 0.0, None, None
    for step in range(60):
        loss = m._loss(xg, yg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        lr_ = O.ce_loss(O.forward(leaves, x, cfg), y)
        opt_ref.zero_grad(set_to_none=True)
        lr_.backward()
        opt_ref.step()
        a, b = loss.item(), lr_.item()
        worst = max(worst, abs(a - b) / abs(b))
        first = first if first is not None else b
        last = b
    print(f"loss curve: oracle {first:.4f} -> {last:.4f}; worst relative gap {worst:.3e}")
    assert last < 0.9 * first, "the synthetic task should be learnable"
    assert worst < 1e-2
    # argmax-mask agreement on TRAINED weights (margins O(1)): the candidate's weights evaluated by both paths
    m.eval()
    trained = {k[len("model."):]: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    xe = O.synthetic_images(4, 224, seed=33)
    with torch.no_grad():
        ours = m(xe.to(dev)).argmax(1).cpu()
        mask = m.model.predict_mask(xe.to(dev)).cpu().long()
        ref = O.forward(trained, xe, cfg).argmax(1)
    agree = (ours == ref).float().mean().item()
    print(f"trained-weights argmax agreement {agree:.5f}")
    assert agree >= 0.999
    assert (mask == ours).float().mean().item() >= 0.9999


def test_gradient_accumulation_matches_single_batch():
    """accumulate_grad_batches semantics (createViTmodel.py:74): two micro-batches accumulate into the same .grad."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=1, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 41, head_gain=2.0)
    m = _build(LightningViTModel, cfg, sd, dev).train()
    x = O.synthetic_images(4, 224, seed=42).to(dev)
    y = O.synthetic_labels(4, 17, seed=43, size=224).to(dev)
    m._loss(x, y).backward()
    full = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    (m._loss(x[:2], y[:2]) / 2).backward()
    (m._loss(x[2:], y[2:]) / 2).backward()
    for k, p in m.named_parameters():
        if p.grad is not None:
            # absolute floor: key biases have an exactly-zero true gradient (softmax shift invariance)
            err = (p.grad - full[k]).abs().max().item()
            assert err <= 2e-2 * full[k].abs().max().item() + 2e-5, (k, err)


def test_fused_optimizer_updates_reach_the_kernels():
    """torch.optim.Adam(fused=True) updates parameters without bumping Tensor._version: the bf16 weight shadows must
    still follow (regression test: the first engine version keyed the re-cast on versions only and trained on stale
    bf16 weights under fused optimizers / CUDA graphs)."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=1, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 51, head_gain=2.0)
    m = _build(LightningViTModel, cfg, sd, dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-3, fused=True)
    x = O.synthetic_images(2, 224, seed=52)
    y = O.synthetic_labels(2, 17, seed=53, size=224)
    for _ in range(3):
        loss = m._loss(x.to(dev), y.to(dev))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    m.eval()
    now = {k[len("model."):]: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    assert max((now[k] - sd[k]).abs().max().item() for k in sd if "pooler" not in k) > 1e-3   # weights did move
    with torch.no_grad():
        ours = m(x.to(dev)).cpu()
        ref = O.forward(now, x, cfg)
        stale = O.forward(sd, x, cfg)
    assert _relmax(ours, ref) < LOGIT_TOL
    assert _relmax(stale, ref) > 5 * LOGIT_TOL    # the check would catch stale weights


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (False, 0.01), (True, 0.01)])
def test_fused_adam_matches_torch_adam(decoupled, wd):
    """visiontransformer_b200.optim.FusedAdam (one pass over the flat arenas) against torch.optim.Adam / AdamW fed
    the same gradients; the pooler (never receives a gradient) must stay untouched even with weight decay."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.optim import FusedAdam
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=1, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 61, head_gain=2.0)
    a = _build(LightningViTModel, cfg, sd, dev).train()
    b = _build(LightningViTModel, cfg, sd, dev).train()
    opt_a = FusedAdam(a, lr=1e-3, weight_decay=wd, decoupled_weight_decay=decoupled)
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    opt_b = cls(b.parameters(), lr=1e-3, weight_decay=wd)
    x = O.synthetic_images(2, 224, seed=62).to(dev)
    y = O.synthetic_labels(2, 17, seed=63, size=224).to(dev)
    for step in range(4):
        la = a._loss(x, y)
        la.backward()
        # feed b exactly a's gradients (isolates the optimizer arithmetic from kernel-level fp noise)
        b.zero_grad(set_to_none=True)
        for (_, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            pb.grad = None if pa.grad is None else pa.grad.clone()
        opt_a.step()
        opt_a.zero_grad()
        opt_b.step()
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            # atol = 0.3 % of one update (lr 1e-3): with coupled weight decay an element whose gradient cancels
            # wd * p (g' = g + wd p ~ 1e-8 from two 2.6e-5 terms) amplifies the fma-vs-mul/add rounding of g' to
            # ~1e-6 in p — observed in 1 of 10 runs, position_embeddings idx 22664
            assert torch.allclose(pa, pb, rtol=2e-5, atol=3e-6), (step, k, (pa - pb).abs().max().item())
    assert torch.equal(dict(a.named_parameters())["model.backbone.pooler.dense.weight"].cpu(),
                       sd["backbone.pooler.dense.weight"])
    # the bf16 shadow written by the fused pass is what the next forward uses
    a.eval()
    now = {k[len("model."):]: v.detach().cpu().clone() for k, v in a.state_dict().items()}
    with torch.no_grad():
        assert _relmax(a(x).cpu(), O.forward(now, x.cpu(), cfg)) < LOGIT_TOL


def test_graphed_train_step_matches_eager_and_honours_set_lr():
    """GraphedTrainStep (whole step in one CUDA graph) follows the eager trajectory, and FusedAdam.set_lr reaches an
    already captured graph (the lr lives in pinned host memory that the replayed copy node re-reads)."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.graph import GraphedTrainStep
    from visiontransformer_b200.optim import FusedAdam
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 71, head_gain=2.0)
    x = O.synthetic_images(2, 224, seed=72).to(dev)
    y = O.synthetic_labels(2, 17, seed=73, size=224).to(dev)

    def make():
        m = _build(LightningViTModel, cfg, sd, dev).train()
        opt = FusedAdam(m, lr=1e-3)

        def step(batch, i):
            loss = m._loss(*batch)
            loss.backward()
            opt.step()
            opt.zero_grad()
            return loss.detach()
        return m, opt, step

    a, _, step_a = make()
    eager = [step_a((x, y), i).item() for i in range(4)]
    b, opt_b, step_b = make()
    g = GraphedTrainStep(step_b, (x, y), warmup=1, engines=[b.model.engine])     # eager step 0, then capture
    replayed = [g((x, y)).item() for _ in range(3)]                              # steps 1..3
    for le, lr_ in zip(eager[1:], replayed):
        assert abs(le - lr_) <= 2e-3 * abs(le), (eager, replayed)
    # (no element-wise weight comparison: Adam turns the atomics-order noise of near-zero gradients into +-lr steps)
    opt_b.set_lr(0.0)
    before = {k: p.detach().clone() for k, p in b.named_parameters()}
    g((x, y))
    torch.cuda.synchronize()
    for k, p in b.named_parameters():
        assert torch.equal(p, before[k]), k


def test_graphed_inference_equals_eager(tiny):
    """GraphedInference (one CUDA graph per input shape: the worker path) returns exactly the eager results and sees
    later weight updates only after a fresh capture — here: same weights, two different inputs."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.graph import GraphedInference
    dev = _dev()
    cfg = O.OracleConfig(**tiny["cfg"])
    sd = O.seeded_state_dict(cfg, tiny["weights_seed"], head_gain=tiny["head_gain"])
    m = _build(LightningViTModel, cfg, sd, dev).eval()
    x1 = O.synthetic_images(2, 224, seed=81).to(dev)
    x2 = O.synthetic_images(2, 224, seed=82).to(dev)
    with torch.no_grad():
        g_logits = GraphedInference(m, x1)
        g_mask = GraphedInference(m.model.predict_mask, x1)
        for x in (x1, x2, x1):
            assert torch.equal(g_logits(x), m(x))
            assert torch.equal(g_mask(x), m.model.predict_mask(x))
    with pytest.raises(ValueError):
        g_logits(x1[:1])


def test_patch4_variant_forward_and_training_step_vs_oracle():
    """P4 variant (3137 tokens, 56x56 grid; IDs 2/5/8 of the reference sweep, datasetTestViTmodel.py:97-107): logits,
    CE loss and head gradients against the fp32 oracle.  Exercises the K = N = 48 patch-embedding GEMMs, the streaming
    attention kernels at N = 3137 and the row-range staging of the upsample kernels (C * g * g floats = 213 KB)."""
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=4, hidden_size=128, num_hidden_layers=1, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 91, head_gain=4.0)
    m = _build(LightningViTModel, cfg, sd, dev)
    x = O.synthetic_images(1, 224, seed=92)
    y = O.synthetic_labels(1, 17, seed=93)
    m.eval()
    with torch.no_grad():
        full = m(x.to(dev))
        mask = m.model.predict_mask(x.to(dev))
        ref = O.forward(sd, x, cfg)
    assert _relmax(full, ref) < LOGIT_TOL
    assert (mask.long().cpu() == full.cpu().argmax(1)).float().mean().item() > 0.9999
    m.train()
    loss = m.training_step((x.to(dev), y.to(dev)), 0)
    loss.backward()
    keys = ("seg_head.2.weight", "seg_head.0.bias", "backbone.embeddings.patch_embeddings.projection.weight")
    sdr = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    ref_loss = O.ce_loss(O.forward(sdr, x, cfg), O.resize_target(y, 224))     # model/CE/classes.py:276-285
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    for key in keys:
        g = dict(m.model.named_parameters())[key].grad.cpu()
        r = sdr[key].grad
        assert (g - r).abs().max().item() <= 3e-2 * r.abs().max().item() + 1e-7, key
