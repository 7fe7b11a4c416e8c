"""GPU (-m gpu): dropout (hidden_dropout_prob / attention_probs_dropout_prob, model/CE/classes.py:233-234).

torch's RNG stream cannot be bit-matched, so the kernels are checked EXACTLY against PyTorch fp32
math that uses the kernels' own masks (exported by vs_dropout_mask), plus keep-rate statistics and module-level
train/eval behaviour."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vitseg_oracle as O

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _mask(K, n, scheme, drop, row_len=0):
    if scheme == 1 and row_len == 0:
        row_len = 256
    return K.dropout_mask(torch.empty(n, device=drop[1].device, dtype=torch.uint8), scheme, drop, row_len).float()


def test_mask_statistics_and_determinism():
    from visiontransformer_b200 import kernels as K
    dev = _dev()
    seed = torch.tensor([1234], device=dev, dtype=torch.int32)
    n = 1 << 22
    for p in (0.1, 0.5):
        for scheme in (0, 1):
            m = _mask(K, n, scheme, (p, seed, 7))
            assert abs(m.mean().item() - (1 - p)) < 2e-3
            assert torch.equal(m, _mask(K, n, scheme, (p, seed, 7)))
            m2 = _mask(K, n, scheme, (p, seed, 8))          # other site -> independent mask
            assert abs((m * m2).mean().item() - (1 - p) ** 2) < 3e-3
    seed2 = seed + 1                                          # next step -> independent mask
    a, b = _mask(K, n, 0, (0.1, seed, 3)), _mask(K, n, 0, (0.1, seed2, 3))
    assert abs((a * b).mean().item() - 0.81) < 3e-3


def test_gemm_epilogue_dropout_and_layernorm_bwd_mask():
    from visiontransformer_b200 import kernels as K
    dev = _dev()
    torch.manual_seed(0)
    M, N, Kd, p = 1000, 768, 512, 0.1
    seed = torch.tensor([99], device=dev, dtype=torch.int32)
    drop = (p, seed, 5)
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = (torch.randn(N, Kd, device=dev) * 0.1).bfloat16()
    bias, res = torch.randn(N, device=dev), torch.randn(M, N, device=dev)
    mask = _mask(K, M * N, 0, drop).view(M, N)
    thresh = round(p * 65536)
    scale = 1.0 / (1.0 - thresh / 65536.0)
    ref = (a.float() @ w.float().t() + bias) * mask * scale + res
    # every tile configuration: 1, 2, 4 use the smem-transposed fp32 epilogue, 3 and 5 (128-column tiles) the TMA
    # residual-prefetch epilogue — the element -> mask map must be the same in all of them
    for cfg in (0, 1, 2, 3, 4, 5):
        out = torch.empty(M, N, device=dev)
        K.gemm(a, w, out, bias=bias, residual=res, dropout=drop, tile_cfg=cfg)
        assert _rel(out, ref) < 1e-5, cfg
    # LayerNorm backward: fp32 dx unmasked, bf16 copy masked with the same site
    D = N
    x = torch.randn(M, D, device=dev)
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    K.layernorm_fwd(x, g, b, 1e-12, y, None, mean, rstd)
    dy = torch.randn(M, D, device=dev)
    dx, dx16 = torch.empty(M, D, device=dev), torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    K.layernorm_bwd(dy, x, g, mean, rstd, None, dx, dx16, dg, db, dropout=drop)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), g, b, 1e-12).backward(dy)
    assert _rel(dx, xr.grad) < 1e-4
    assert _rel(dx16, xr.grad * mask * scale) < 1e-2
    assert ((dx16 == 0) == (mask == 0)).float().mean().item() > 0.999
    # dropout_rows: in place + bf16 copy
    z = torch.randn(M, D, device=dev)
    z0 = z.clone()
    z16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    K.dropout_rows(z, z16, drop)
    assert _rel(z, z0 * mask * scale) < 1e-6
    assert _rel(z16, z) < 1e-2


@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (1, 300, 2), (1, 256, 2), (2, 40, 1)])
def test_attention_dropout_forward_backward(B, N, H):
    from visiontransformer_b200 import kernels as K
    dev = _dev()
    torch.manual_seed(1)
    p = 0.1
    seed = torch.tensor([4242], device=dev, dtype=torch.int32)
    drop = (p, seed, 1003)
    D = H * 64
    qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
    ctx = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    K.attention_fwd(qkv, ctx, lse, B, N, H, 0.125, dropout=drop)
    mask = _mask(K, B * H * N * N, 1, drop, N).view(B, H, N, N)
    scale = 1.0 / (1.0 - round(p * 65536) / 65536.0)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).float().requires_grad_(True) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = (torch.softmax(s, -1) * mask * scale) @ v
    assert _rel(ctx.permute(0, 2, 1, 3), ref) < 1e-2
    assert _rel(lse, torch.logsumexp(s, -1)) < 1e-3
    dctx = torch.randn(B, N, H, 64, device=dev).bfloat16()
    ref.backward(dctx.permute(0, 2, 1, 3).float())
    dqkv = torch.zeros(B, N, 3, H, 64, device=dev, dtype=torch.bfloat16)
    dq_acc, delta = torch.empty(B, N, D, device=dev), torch.empty(B, H, N, device=dev)
    K.attention_bwd(qkv, ctx, dctx, lse, dqkv, dq_acc, delta, B, N, H, 0.125, dropout=drop)
    assert _rel(dqkv[:, :, 1].permute(0, 2, 1, 3), k.grad) < 2e-2
    assert _rel(dqkv[:, :, 2].permute(0, 2, 1, 3), v.grad) < 2e-2
    assert _rel(dq_acc.view(B, N, H, 64).permute(0, 2, 1, 3), q.grad) < 2e-2
    if N <= 256:   # short-sequence kernel (dq_accum = None): same masks, all three gradients as bf16
        dqkv2 = torch.full_like(dqkv, float("nan"))
        K.attention_bwd(qkv, ctx, dctx, lse, dqkv2, None, delta, B, N, H, 0.125, dropout=drop)
        for i, g in enumerate((q.grad, k.grad, v.grad)):
            assert _rel(dqkv2[:, :, i].permute(0, 2, 1, 3), g) < 2e-2


def test_module_train_eval_dropout_semantics():
    """train(): dropout on (different masks per step, loss finite, gradients flow); eval(): deterministic and equal to
    the dropout-free model; same seed -> same result."""
    from oracle import vitseg_oracle as O
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 7, head_gain=4.0)
    m = LightningViTModel(17, 16, 128, 2, 2)     # reference defaults: p = 0.1 / 0.1
    m.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    m = m.to(dev)
    x = O.synthetic_images(2, 224, seed=11).to(dev)
    y = O.synthetic_labels(2, 17, seed=12).to(dev)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(x), m(x)
        ref = O.forward(sd, x.cpu(), cfg)
    assert torch.equal(e1, e2)
    assert _rel(e1.cpu(), ref) < 1e-2
    m.train()
    eng = m.model.engine
    eng.seed_dropout(5)
    l1 = m.training_step((x, y), 0)
    l1.backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    l2 = m.training_step((x, y), 1)              # counter advanced -> different masks
    eng.seed_dropout(5)
    l3 = m.training_step((x, y), 2)              # same counter -> identical result
    l3.backward()
    assert torch.isfinite(l1) and l1.item() != l2.item()
    assert abs(l1.item() - l3.item()) < 1e-4 * abs(l1.item())   # float atomics in the loss reduction
    for k, p in m.named_parameters():
        if p.grad is not None:
            # atomics reorder sums: not bitwise; key biases have a zero true gradient, hence the absolute floor
            # (dQ is accumulated with fp32 atomics and then rounded to bf16: a reordered sum flips single roundings,
            #  2^-9 relative on those elements, which the earlier layers inherit)
            assert (p.grad - g1[k]).abs().max().item() <= 5e-3 * g1[k].abs().max().item() + 1e-5, k
    with torch.no_grad():
        m.eval()
        lo = m._loss(x, y).item()
    assert abs(l1.item() - lo) / lo < 0.2          # dropout perturbs, it does not destroy


def _masked_reference_loss(sd, x, y, cfg, masks, scale_h, scale_a):
    """fp32 PyTorch restatement of the training-mode forward (oracle/vitseg_oracle.py structure, TF:100-128,220-346)
    with the dropout masks supplied explicitly."""
    D, H = cfg.hidden_size, cfg.num_attention_heads
    dh = D // H
    h = O.embeddings(sd, x, cfg)
    B, N, _ = h.shape
    h = h * masks[0].view(B, N, D) * scale_h
    for i in range(cfg.num_hidden_layers):
        p = f"backbone.encoder.layer.{i}."
        yv = F.layer_norm(h, (D,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], 1e-12)
        q, k, v = [F.linear(yv, sd[p + f"attention.attention.{n}.weight"], sd[p + f"attention.attention.{n}.bias"])
                   .view(B, N, H, dh).transpose(1, 2) for n in ("query", "key", "value")]
        att = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, -1) * masks[1000 + i].view(B, H, N, N) * scale_a
        ctx = (att @ v).transpose(1, 2).reshape(B, N, D)
        o = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        h = h + o * masks[1 + 2 * i].view(B, N, D) * scale_h
        yv = F.layer_norm(h, (D,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], 1e-12)
        yv = F.gelu(F.linear(yv, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        o = F.linear(yv, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
        h = h + o * masks[2 + 2 * i].view(B, N, D) * scale_h
    h = F.layer_norm(h, (D,), sd["backbone.layernorm.weight"], sd["backbone.layernorm.bias"], 1e-12)[:, 1:]
    g = int((N - 1) ** 0.5)
    feat = h.transpose(1, 2).reshape(B, D, g, g)
    out = F.conv2d(F.relu(F.conv2d(feat, sd["seg_head.0.weight"], sd["seg_head.0.bias"], padding=1)),
                   sd["seg_head.2.weight"], sd["seg_head.2.bias"])
    return O.ce_loss(O.upsample(out, x.shape[-1]), y)


def test_training_step_with_dropout_matches_masked_reference():
    """whole training step with dropout ON against the fp32 reference that uses the kernels' own masks: loss and
    gradients (every site's mask must be regenerated consistently in backward)."""
    from visiontransformer_b200 import kernels as K
    from visiontransformer_b200.ce.classes import LightningViTModel
    dev = _dev()
    cfg = O.OracleConfig(num_classes=17, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2)
    sd = O.seeded_state_dict(cfg, 9, head_gain=4.0)
    ph, pa = 0.1, 0.2
    m = LightningViTModel(17, 16, 128, 2, 2, hidden_dropout_prob=ph, attention_probs_dropout_prob=pa)
    m.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    m = m.to(dev).train()
    x = O.synthetic_images(2, 224, seed=3)
    y = O.resize_target(O.synthetic_labels(2, 17, seed=4), 224)
    eng = m.model.engine
    eng.seed_dropout(77)
    loss = m._loss(x.to(dev), O.synthetic_labels(2, 17, seed=4).to(dev))
    loss.backward()
    # export the masks of this step (the counter was advanced once by the forward)
    B, N, D, H = 2, 197, 128, 2
    masks = {}
    for site in (0, 1, 2, 3, 4):
        masks[site] = _mask(K, B * N * D, 0, (ph, eng.rng_step, site)).cpu()
    for i in range(2):
        masks[1000 + i] = _mask(K, B * H * N * N, 1, (pa, eng.rng_step, 1000 + i), N).cpu()
    sc = lambda p: 1.0 / (1.0 - round(p * 65536) / 65536.0)  # noqa: E731
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = _masked_reference_loss(leaves, x, y, cfg, masks, sc(ph), sc(pa))
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-2 * abs(ref.item())
    named = dict(m.named_parameters())
    for k, v in leaves.items():
        if v.grad is None:
            continue
        g = named["model." + k].grad
        assert g is not None, k
        # key biases have an exactly-zero true gradient (softmax shift invariance): absolute floor for such tensors
        err = (g.cpu() - v.grad).abs().max().item()
        assert err <= 4e-2 * v.grad.abs().max().item() + 2e-5, (k, err, v.grad.abs().max().item())


def test_attention_dropout_index_space_is_independent_of_the_batch():
    """Patch-4 scale (N = 3137) with batch*heads = 456: B*H*N*(N+1) exceeds 2^32, which the mask index used to
    overflow.  The mask is seeded per (batch, head), so a batch entry gets the same mask wherever it sits... in its own
    (batch, head) slot: slot 0 of a 1-image call equals slot 0 of the big call."""
    from visiontransformer_b200 import kernels as K
    dev = _dev()
    B, N, H = 38, 3137, 12
    torch.manual_seed(0)
    qkv = (torch.randn(B, N, 3, H, 64, device=dev) * 0.5).to(torch.bfloat16)
    ctx = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    seed = torch.tensor([11], device=dev, dtype=torch.int32)
    K.attention_fwd(qkv, ctx, lse, B, N, H, 0.125, dropout=(0.1, seed, 1003))
    assert torch.isfinite(ctx.float()).all() and torch.isfinite(lse).all()
    ctx1 = torch.empty(1, N, H, 64, device=dev, dtype=torch.bfloat16)
    lse1 = torch.empty(1, H, N, device=dev)
    K.attention_fwd(qkv[:1].contiguous(), ctx1, lse1, 1, N, H, 0.125, dropout=(0.1, seed, 1003))
    assert torch.equal(ctx1[0], ctx[0])          # (b, h) = (0, h) slots coincide in both calls
    nodrop = torch.empty_like(ctx1)
    K.attention_fwd(qkv[:1].contiguous(), nodrop, lse1, 1, N, H, 0.125)
    assert not torch.equal(nodrop, ctx1)


def test_attention_mask_matches_the_numpy_restatement_of_the_generator():
    """The attention-probability mask (four decisions per hash, drop_keep4 in csrc/common.cuh) bit for bit against
    tools/dropout_quality.py's numpy restatement — the offline quality numbers quoted in DESIGN.md are about THIS generator."""
    import os
    import sys

    import numpy as np

    from visiontransformer_b200 import kernels as K
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from dropout_quality import quad_fields

    def h3(x, seed):   # drop_hash
        x = (int(x) ^ int(seed)) & 0xFFFFFFFF
        x = (x * 0x9E3779B1) & 0xFFFFFFFF; x ^= x >> 16
        x = (x * 0x85EBCA6B) & 0xFFFFFFFF; x ^= x >> 13
        x = (x * 0xC2B2AE35) & 0xFFFFFFFF; x ^= x >> 16
        return x

    dev = _dev()
    step, site, p, N, BH = 4321, 1003, 0.1, 37, 3
    seed = torch.tensor([step], device=dev, dtype=torch.int32)
    got = K.dropout_mask(torch.empty(BH * N * N, device=dev, dtype=torch.uint8), 1, (p, seed, site), N).cpu().numpy()
    thresh = int(p * 65536.0 + 0.5)
    sd = h3((site + 0x27D4EB2F) & 0xFFFFFFFF, h3(step, 0x9E3779B9))            # drop_seed(step, site)
    exp = np.empty((BH, N, N), dtype=np.uint8)
    nquad = (N + 3) >> 2
    for bh in range(BH):
        quads = (np.arange(N, dtype=np.uint64)[:, None] * np.uint64(nquad) + np.arange(nquad, dtype=np.uint64)[None, :])
        f = quad_fields(quads.reshape(-1), h3(bh, sd)).reshape(N, nquad * 4)   # [query, key] uniforms
        exp[bh] = (f[:, :N] >= thresh)
    assert np.array_equal(got.reshape(BH, N, N), exp)
