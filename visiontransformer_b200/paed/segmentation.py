"""Host-side helpers of model/PAED/segmentation.py: SDF target generation (dataset side, SciPy EDT on the CPU —
an input producer, not part of the device hot path) and the binary monitoring metrics."""
import numpy as np
import torch


def compute_sdf(mask: np.ndarray):
    """model/PAED/segmentation.py:6-34: Euclidean distance transforms of the background and of the object, each
    divided by its own maximum."""
    from scipy.ndimage import distance_transform_edt
    mask = mask.astype(bool)
    sdf_ext = distance_transform_edt(~mask).astype(np.float32)
    sdf_int = distance_transform_edt(mask).astype(np.float32)
    if sdf_ext.max() > 0:
        sdf_ext /= sdf_ext.max()
    if sdf_int.max() > 0:
        sdf_int /= sdf_int.max()
    return sdf_ext, sdf_int


def pixel_accuracy(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (gt == pred).float().sum() / torch.numel(gt)


def intersection_over_union(gt: torch.Tensor, pred: torch.Tensor, eps=1e-6) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return ((gt & pred).float().sum() + eps) / ((gt | pred).float().sum() + eps)


def dice_score(gt: torch.Tensor, pred: torch.Tensor, eps=1e-6) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (2 * (gt & pred).float().sum() + eps) / (gt.float().sum() + pred.float().sum() + eps)
