"""vs_upsample_argmax_stats: exact integer parity of the fused argmax + per-class pixel counts with counts formed from
the (already parity-tested) class map and the labels, and the wrappers' logged metrics against the reference formulas."""
import pytest
import torch

from _metric_refs import (binary_iou_ref, counts_from_maps, dice_score_ref, iou_score_ref, pixel_accuracy_ref)

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,C,g,S", [(3, 17, 14, 224), (2, 1, 14, 224), (2, 4, 24, 384), (1, 17, 32, 512), (5, 2, 7, 56)])
def test_counts_are_exact(B, C, g, S):
    from visiontransformer_b200 import kernels as K
    dev = _dev()
    torch.manual_seed(B * 100 + C)
    low = torch.randn(B, C, g, g, device=dev) * 3
    nc = 2 if C == 1 else C
    labels = torch.randint(0, nc, (B, S, S), device=dev)
    labels[0, :5, :7] = -100                                  # ignore_index pixels: predicted, never a target
    mask = torch.empty(B, S, S, device=dev, dtype=torch.uint8)
    counts = K.upsample_argmax_stats(low, labels, S, mask)
    want_mask = K.upsample_argmax(low, torch.empty_like(mask))
    assert torch.equal(mask, want_mask)
    ref = counts_from_maps(want_mask.long().cpu(), labels.cpu(), nc)
    assert torch.equal(counts.cpu(), ref)
    assert int(counts[..., 1].sum()) == B * S * S             # every pixel is predicted exactly once
    no_mask = K.upsample_argmax_stats(low, labels, S)          # mask output is optional
    assert torch.equal(no_mask, counts)


def test_wrapper_metrics_follow_the_reference_formulas():
    from oracle import vitseg_oracle as O
    from visiontransformer_b200 import metrics as M
    from visiontransformer_b200.paed.classes import LightningViTModel, PAEDTrainer
    dev = _dev()
    torch.manual_seed(5)
    x = torch.rand(2, 3, 224, 224, device=dev)
    y = torch.randint(0, 17, (2, 256, 256), device=dev)
    m = LightningViTModel(17, 16, 128, 1, 2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0).to(dev).train()
    with torch.no_grad():
        m.model.seg_head[2].weight.mul_(20.0)
    loss, iou = m._step((x, y))
    with torch.no_grad():
        pred = m.model.predict_mask(x).long().cpu()
        tgt = m._resize_target(y, size=(224, 224)).cpu()
    assert abs(iou.item() - iou_score_ref(pred, tgt, 17).item()) < 1e-6
    masks, se, si = [t.to(dev) for t in O.synthetic_binary_targets(2, 224, seed=3)]
    t = PAEDTrainer(1, 16, 128, 1, 2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0).to(dev).train()
    _, acc, biou, dice, prec, rec = t._forward_step((x, masks, se, si), 0)
    with torch.no_grad():
        bp = t.model.predict_mask(x).int().cpu()
        mk = masks.int().cpu()
    assert abs(acc.item() - pixel_accuracy_ref(mk, bp).item()) < 1e-6
    assert abs(biou.item() - binary_iou_ref(mk, bp).item()) < 1e-6
    assert abs(dice.item() - dice_score_ref(mk, bp).item()) < 1e-6
    tp = (bp & mk).sum().float()
    assert abs(prec.item() - (tp / bp.sum().clamp_min(1)).item()) < 1e-6
    assert abs(rec.item() - (tp / mk.sum().clamp_min(1)).item()) < 1e-6
    c = M.segmentation_counts(t.model.forward_lowres(x).detach(), masks.long(), 224)
    assert torch.equal(M.all_reduce_sum_counts(c)[0], c.sum(0).to(torch.int64))


@pytest.mark.parametrize("S", [224, 96, 384])
def test_sdf_targets_are_bit_identical_to_compute_sdf(S):
    """vs_sdf_targets (exact EDT on the device) against compute_sdf = scipy.ndimage.distance_transform_edt, the
    reference's own call (model/PAED/segmentation.py:22-32): random discs, a single pixel, an empty and a full mask."""
    import numpy as np
    from oracle import vitseg_oracle as O
    from visiontransformer_b200.paed.segmentation import compute_sdf_batch
    dev = _dev()
    masks, se, si = O.synthetic_binary_targets(4, S, seed=S)
    extra = torch.zeros(4, S, S)
    extra[0, S // 3, S // 2] = 1.0                      # one object pixel
    extra[1] = 1.0                                      # full mask: EDT(~mask) = 0, EDT(mask) hits SciPy's virtual zero
    extra[2, :, : S // 2] = 1.0                         # half plane (whole rows / columns without a zero)
    # extra[3] stays empty
    masks = torch.cat([masks, extra])
    ref = [O.compute_sdf(m.numpy().astype(np.uint8)) for m in masks]
    ext, inn = compute_sdf_batch(masks.to(dev))
    for b in range(masks.shape[0]):
        assert torch.equal(ext[b].cpu(), torch.from_numpy(ref[b][0])), ("ext", b)
        assert torch.equal(inn[b].cpu(), torch.from_numpy(ref[b][1])), ("int", b)


def test_colorize_mask_is_the_palette_lookup():
    """vs_colorize_mask == index_to_color[pred_labels] (model/CE/testViTModel.py:139-143), byte-exact, odd sizes."""
    from visiontransformer_b200 import kernels as K
    from visiontransformer_b200.ce.classes import ViTSegmentationModel
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    palette = torch.randint(0, 256, (17, 3), generator=g, dtype=torch.uint8)
    for shape in ((2, 224, 224), (1, 7, 9), (3, 5)):
        mask = torch.randint(0, 17, shape, generator=g, dtype=torch.uint8)
        rgb = K.colorize_mask(mask.to(dev), palette.to(dev))
        assert torch.equal(rgb.cpu(), palette[mask.long()])
    m = ViTSegmentationModel(17, 16, 128, 1, 2).to(dev).eval()
    x = torch.rand(2, 3, 224, 224, device=dev)
    img = m.predict_colored(x, palette.to(dev))
    assert img.shape == (2, 224, 224, 3) and torch.equal(img.cpu(), palette[m.predict_mask(x).long().cpu()])
