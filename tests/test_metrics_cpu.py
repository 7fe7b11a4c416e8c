"""metrics.py (count-based segmentation metrics) against the reference definitions, on CPU:
the count algebra is device-agnostic; the counts themselves come from the CUDA kernel (tests/test_metrics_gpu.py)."""
import numpy as np
import torch

from _metric_refs import (binary_iou_ref, counts_from_maps, dice_score_ref, iou_score_ref, per_image_eval_ref,
                          pixel_accuracy_ref)
from visiontransformer_b200 import metrics as M


def _maps(B, S, nc, seed, missing=()):
    g = torch.Generator().manual_seed(seed)
    pred = torch.randint(0, nc, (B, S, S), generator=g)
    tgt = torch.randint(0, nc, (B, S, S), generator=g)
    for c in missing:          # classes absent from both maps exercise the NaN / eps branches
        pred[pred == c] = 0
        tgt[tgt == c] = 0
    return pred, tgt


def test_multiclass_iou_score_and_accuracy():
    pred, tgt = _maps(3, 40, 17, 0, missing=(5, 16))
    c = counts_from_maps(pred, tgt, 17)
    assert torch.allclose(M.iou_score(c), iou_score_ref(pred, tgt, 17), rtol=1e-6)
    assert torch.allclose(M.pixel_accuracy(c), pixel_accuracy_ref(tgt, pred), rtol=1e-6)


def test_binary_metrics():
    pred, tgt = _maps(4, 32, 2, 1)
    c = counts_from_maps(pred, tgt, 2)
    assert torch.allclose(M.intersection_over_union(c), binary_iou_ref(tgt, pred), rtol=1e-6)
    assert torch.allclose(M.dice_score(c), dice_score_ref(tgt, pred), rtol=1e-6)
    assert torch.allclose(M.pixel_accuracy(c), pixel_accuracy_ref(tgt, pred), rtol=1e-6)
    tp = ((pred == 1) & (tgt == 1)).sum().float()
    prec, rec = M.binary_precision_recall(c)
    assert torch.allclose(prec, tp / (pred == 1).sum()) and torch.allclose(rec, tp / (tgt == 1).sum())
    empty = counts_from_maps(torch.zeros(1, 8, 8, dtype=torch.long), torch.zeros(1, 8, 8, dtype=torch.long), 2)
    p0, r0 = M.binary_precision_recall(empty)           # no positives anywhere: defined as 0, not NaN
    assert p0.item() == 0.0 and r0.item() == 0.0
    assert abs(M.intersection_over_union(empty).item() - 1.0) < 1e-6   # (0 + eps) / (0 + eps), as the reference


def test_per_image_eval_matches_the_eval_script():
    pred, tgt = _maps(3, 24, 6, 2, missing=(4,))
    out = M.per_image_eval(counts_from_maps(pred, tgt, 6))
    for b in range(3):
        acc, ious, dices = per_image_eval_ref(tgt[b].numpy(), pred[b].numpy(), 6)
        assert abs(out["accuracy"][b].item() - acc) < 1e-4
        assert np.allclose(out["iou"][b].numpy(), ious, rtol=1e-6, equal_nan=True)
        assert np.allclose(out["dice"][b].numpy(), dices, rtol=1e-6, equal_nan=True)
        assert abs(out["mean_iou"][b].item() - np.nanmean(ious)) < 1e-6
        assert abs(out["mean_dice"][b].item() - np.nanmean(dices)) < 1e-6


def test_all_reduce_sum_counts_without_process_group():
    pred, tgt = _maps(2, 16, 3, 3)
    c = counts_from_maps(pred, tgt, 3)
    tot = M.all_reduce_sum_counts(c)
    assert tot.shape == (1, 3, 3) and torch.equal(tot[0], c.sum(0).to(torch.int64))
    assert torch.allclose(M.pixel_accuracy(tot), M.pixel_accuracy(c))
