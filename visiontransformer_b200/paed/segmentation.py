"""Helpers of model/PAED/segmentation.py: SDF target generation — compute_sdf as in the reference (dataset side, SciPy
EDT on the CPU) and compute_sdf_batch, the same transform for a whole batch on the device — and the binary monitoring
metrics."""
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


def compute_sdf_batch(masks: torch.Tensor):
    """Device version of compute_sdf for a batch: masks [B,S,S] or [B,1,S,S] (any dtype, object = non-zero) on a CUDA
    device -> (sdf_ext, sdf_int) fp32 [B,S,S], bit-identical to compute_sdf per image (vs_sdf_targets: exact EDT)."""
    from .. import kernels as K
    if masks.dim() == 4:
        masks = masks.squeeze(1)
    return K.sdf_targets((masks != 0).to(torch.float32).contiguous())


def pixel_accuracy(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (gt == pred).float().sum() / torch.numel(gt)


def intersection_over_union(gt: torch.Tensor, pred: torch.Tensor, eps=1e-6) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return ((gt & pred).float().sum() + eps) / ((gt | pred).float().sum() + eps)


def dice_score(gt: torch.Tensor, pred: torch.Tensor, eps=1e-6) -> torch.Tensor:
    gt, pred = gt.squeeze().int(), pred.squeeze().int()
    return (2 * (gt & pred).float().sum() + eps) / (gt.float().sum() + pred.float().sum() + eps)
