"""Full-size (BASELINE.json configs[1]: ViT-B/16, batch 64, 224x224, C=17) checks through size-independent
properties — the fp32 oracle cannot run this size in seconds, so parity at the full size is established as
  * batch independence: every image of a 64-image batch gets exactly the logits it gets in an 8-image batch (each output
    row of every GEMM, LayerNorm, attention tile and upsample is computed independently of M), and the 8-image case is
    pinned to the reference by tests/test_parity_gpu.py::test_vitb16_logits_and_argmax_vs_golden;
  * the fused uint8 mask equals sigmoid().argmax() of the full logits (testViTModel.py:121-126);
  * linearity of the gradient in the batch: the mean-loss gradient of 64 images equals the mean of the gradients of its
    two halves (this is exactly what the data-parallel all-reduce relies on);
  * the loss of random weights is finite and not better than the uniform prediction ln(17)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def vitb():
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    torch.manual_seed(1234)
    m = LightningViTModel(17, 16, 768, 12, 12, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0).to(dev)
    with torch.no_grad():   # a head with real dynamic range (the default init gives near-constant logits)
        m.model.seg_head[2].weight.mul_(8.0)
    x = torch.rand(64, 3, 224, 224, device=dev)
    y = torch.randint(0, 17, (64, 256, 256), device=dev)
    return m, x, y


def test_full_batch_logits_are_batch_independent(vitb):
    m, x, _ = vitb
    m.eval()
    with torch.no_grad():
        full = m(x)
        assert full.shape == (64, 17, 224, 224) and full.dtype == torch.float32
        for i in range(0, 64, 8):
            part = m(x[i:i + 8])
            assert torch.equal(part, full[i:i + 8]), f"images {i}..{i + 7} depend on their batch"
        perm = torch.randperm(64, device=x.device)
        assert torch.equal(m(x[perm]), full[perm])
    assert torch.isfinite(full).all()


def test_full_batch_fused_mask_equals_argmax_of_logits(vitb):
    m, x, _ = vitb
    m.eval()
    with torch.no_grad():
        logits = m(x)
        mask = m.model.predict_mask(x)
    assert mask.shape == (64, 224, 224) and mask.dtype == torch.uint8
    sig = logits.sigmoid()
    want = sig.argmax(1)
    diff = mask.long() != want
    # sigmoid is not injective in fp32: two close logits can round to the same probability, where torch's argmax
    # returns the first index and the fused kernel (which orders by the logits themselves) the larger logit.  Any
    # disagreement must be such a tie (equal probabilities up to 1 ulp) and rare.
    assert diff.float().mean().item() < 1e-5
    if diff.any():
        p_ours = sig.gather(1, mask.long().unsqueeze(1)).squeeze(1)[diff]
        p_ref = sig.gather(1, want.unsqueeze(1)).squeeze(1)[diff]
        assert (p_ours - p_ref).abs().max().item() <= 1.2e-7


def test_full_batch_gradient_is_the_mean_of_half_batch_gradients(vitb):
    m, x, y = vitb
    m.train()

    def grads(xs, ys):
        m.zero_grad(set_to_none=True)
        loss = m._loss(xs, ys)
        loss.backward()
        return loss.item(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    l_full, g_full = grads(x, y)
    l_a, g_a = grads(x[:32], y[:32])
    l_b, g_b = grads(x[32:], y[32:])
    assert abs(l_full - 0.5 * (l_a + l_b)) < 1e-5 * abs(l_full)
    assert math.log(17) - 0.5 < l_full < 20.0        # random weights: no better than the uniform prediction
    gmax = max(g.abs().max().item() for g in g_full.values())
    worst = 0.0
    for k, g in g_full.items():
        ref = 0.5 * (g_a[k] + g_b[k])
        scale = ref.abs().max().item()
        if scale < 1e-5 * gmax:   # key biases: softmax is shift-invariant, their true gradient is zero (pure noise)
            assert k.endswith("attention.key.bias"), k
            continue
        worst = max(worst, (g - ref).abs().max().item() / scale)
    # the 1/64 vs 1/32 loss scaling is a power of two, so every bf16 rounding of the backward tensors is identical in
    # both computations; what remains is fp32 summation order (split-K, atomics): measured 7.5e-4, the same as the
    # run-to-run noise of one computation
    assert worst < 5e-3, worst
    m.zero_grad(set_to_none=True)
