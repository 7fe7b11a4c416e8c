"""On-device segmentation metrics from the exact per-image / per-class pixel counts produced by the fused
upsample + argmax + statistics kernel (vs_upsample_argmax_stats) — SURVEY.md §8f rank 2.

counts: int32 [B, NC, 3] = {intersection, predicted, target}.  Every function below is O(B * NC) tensor algebra on the
device (no host sync) and reproduces a reference definition:

  iou_score            model/PAED/classes.py:430-447   mean over classes of the batch mean of the per-image IoU
  pixel_accuracy       model/PAED/segmentation.py:38-50
  intersection_over_union / dice_score (binary, positives, global over the batch)   segmentation.py:53-86
  binary_precision_recall   torchmetrics precision / recall, task='binary', multidim_average='global'
                            (model/PAED/classes.py:686-689)
  per_image_eval       model/CE/datasetTestViTmodel.py:188-217   accuracy %, per-class IoU / Dice with NaN for absent
                            classes, nan-means

Data parallel: counts of different ranks concatenate along the batch axis (per-image metrics) or add (global ones);
`all_reduce_sum_counts` does the latter."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import kernels as K


def segmentation_counts(low: torch.Tensor, labels: torch.Tensor, size: int, mask: torch.Tensor | None = None):
    """low: fp32 [B,C,g,g] low-resolution logits (ViTSegmentationModel.forward_lowres); labels int64 [B,size,size]."""
    return K.upsample_argmax_stats(low.detach().contiguous(), labels.contiguous(), size, mask)


def _parts(counts):
    c = counts.to(torch.float32)
    inter, pred, tgt = c[..., 0], c[..., 1], c[..., 2]
    return inter, pred, tgt, pred + tgt - inter


def iou_score(counts: torch.Tensor) -> torch.Tensor:
    inter, _, _, union = _parts(counts)
    return ((inter + 1e-6) / (union + 1e-6)).mean(0).mean()


def pixel_accuracy(counts: torch.Tensor) -> torch.Tensor:
    inter, pred, _, _ = _parts(counts)
    return inter.sum() / pred.sum()          # every pixel is predicted as exactly one class


def intersection_over_union(counts: torch.Tensor, positive: int = 1, eps: float = 1e-6) -> torch.Tensor:
    inter, _, _, union = _parts(counts)
    return (inter[:, positive].sum() + eps) / (union[:, positive].sum() + eps)


def dice_score(counts: torch.Tensor, positive: int = 1, eps: float = 1e-6) -> torch.Tensor:
    inter, pred, tgt, _ = _parts(counts)
    return (2 * inter[:, positive].sum() + eps) / (tgt[:, positive].sum() + pred[:, positive].sum() + eps)


def binary_precision_recall(counts: torch.Tensor, positive: int = 1):
    inter, pred, tgt, _ = _parts(counts)
    tp, pp, ap = inter[:, positive].sum(), pred[:, positive].sum(), tgt[:, positive].sum()
    zero = torch.zeros((), device=counts.device)
    return torch.where(pp > 0, tp / pp.clamp_min(1.0), zero), torch.where(ap > 0, tp / ap.clamp_min(1.0), zero)


def per_image_eval(counts: torch.Tensor):
    """-> dict of [B] tensors: accuracy (percent), mean_iou, mean_dice (nan-means over the classes present in either
    map), plus the [B, NC] per-class iou / dice with NaN where the class is absent from both maps."""
    inter, pred, tgt, union = _parts(counts)
    nan = torch.full_like(inter, float("nan"))
    iou = torch.where(union > 0, inter / union.clamp_min(1.0), nan)
    den = pred + tgt
    dice = torch.where(den > 0, 2 * inter / den.clamp_min(1.0), nan)
    acc = 100.0 * inter.sum(1) / pred.sum(1)
    return {"accuracy": acc, "iou": iou, "dice": dice, "mean_iou": torch.nanmean(iou, dim=1),
            "mean_dice": torch.nanmean(dice, dim=1)}


def all_reduce_sum_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Global (batch-summed) counts [1, NC, 3] over all ranks — for the metrics that are global over the batch."""
    tot = counts.sum(0, keepdim=True).to(torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, group=group)
    return tot
