"""Plain-PyTorch restatements of the reference's metric definitions (test infrastructure), used by the CPU and GPU
metric tests.  Each function cites the reference lines it follows."""
import numpy as np
import torch


def counts_from_maps(pred: torch.Tensor, tgt: torch.Tensor, nc: int) -> torch.Tensor:
    """[B,H,W] integer class maps -> int32 [B, nc, 3] = {intersection, predicted, target}."""
    B = pred.shape[0]
    out = torch.zeros(B, nc, 3, dtype=torch.int32)
    for c in range(nc):
        p, t = pred == c, tgt == c
        out[:, c, 0] = (p & t).flatten(1).sum(1)
        out[:, c, 1] = p.flatten(1).sum(1)
        out[:, c, 2] = t.flatten(1).sum(1)
    return out


def iou_score_ref(preds, targets, num_classes):
    """model/PAED/classes.py:430-447."""
    po = torch.nn.functional.one_hot(preds, num_classes).permute(0, 3, 1, 2).float()
    to = torch.nn.functional.one_hot(targets, num_classes).permute(0, 3, 1, 2).float()
    per_class = []
    for c in range(num_classes):
        inter = (po[:, c] * to[:, c]).sum((1, 2))
        union = (po[:, c] + to[:, c]).clamp(0, 1).sum((1, 2))
        per_class.append(((inter + 1e-6) / (union + 1e-6)).mean())
    return torch.tensor(per_class).mean()


def pixel_accuracy_ref(gt, pred):
    """model/PAED/segmentation.py:38-50."""
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (gt == pred).float().sum() / torch.numel(gt)


def binary_iou_ref(gt, pred, eps=1e-6):
    """model/PAED/segmentation.py:53-68."""
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return ((gt & pred).float().sum() + eps) / ((gt | pred).float().sum() + eps)


def dice_score_ref(gt, pred, eps=1e-6):
    """model/PAED/segmentation.py:71-86."""
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (2 * (gt & pred).float().sum() + eps) / (gt.float().sum() + pred.float().sum() + eps)


def per_image_eval_ref(gt: np.ndarray, pred: np.ndarray, num_classes: int):
    """model/CE/datasetTestViTmodel.py:188-217 for one image: accuracy %, per-class IoU / Dice (NaN when absent)."""
    acc = 100 * (1 - (gt != pred).astype(float).sum() / gt.size)
    ious, dices = [], []
    for c in range(num_classes):
        g, p = gt == c, pred == c
        inter, union = np.logical_and(g, p).sum(), np.logical_or(g, p).sum()
        ious.append(float("nan") if union == 0 else inter / union)
        den = g.sum() + p.sum()
        dices.append(float("nan") if den == 0 else 2 * inter / den)
    return acc, np.array(ious), np.array(dices)
