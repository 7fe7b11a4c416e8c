"""Fused segmentation losses on the low-resolution logits (the [B,C,S,S] tensor is never materialised).

  upsample_cross_entropy      == nn.CrossEntropyLoss()(F.interpolate(low, S, 'bilinear'), y)   model/CE/classes.py:276-285
  paed_binary_loss            == PAEDTrainer._forward_step_paed loss                           model/PAED/classes.py:608-701
  paed_multiclass_soft_fused  == paed_loss_multiclass_soft(one_hot(y), softmax(up(low)))       model/PAED/classes.py:336-369,448-467
  paed_loss_multiclass_soft   == the free function on dense (mask, probability) tensors        model/PAED/classes.py:336-369
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import kernels as K

F32 = torch.float32


class _UpsampleCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, low, labels, size):
        low = low.contiguous()
        labels = labels.contiguous()
        if labels.dim() == 4 and labels.shape[1] == 1:
            labels = labels[:, 0]
        if labels.dim() != 3 or labels.shape[0] != low.shape[0]:
            raise ValueError(f"labels must be [B,H,W], got {tuple(labels.shape)}")
        if labels.dtype not in (torch.int64, torch.uint8):
            labels = labels.long()
        loss_sum = torch.zeros(2, device=low.device, dtype=F32)
        need = ctx.needs_input_grad[0]
        dlow = torch.zeros_like(low) if need else None
        K.upsample_ce(low, labels, loss_sum, dlow, size)
        ctx.save_for_backward(dlow if need else torch.empty(0), loss_sum)
        return loss_sum[0] / loss_sum[1]

    @staticmethod
    def backward(ctx, gout):
        dlow, loss_sum = ctx.saved_tensors
        return dlow * (gout / loss_sum[1]), None, None


def upsample_cross_entropy(low: torch.Tensor, labels: torch.Tensor, size: int) -> torch.Tensor:
    """mean over non-ignored pixels of -log_softmax(bilinear_up(low))[label].  labels int64 (or uint8) [B,H,W]; when
    H x W differs from size x size the kernel reads them through the legacy-'nearest' index map of
    F.interpolate(y.float(), size, mode='nearest') — LightningViTModel._resize_target without the three extra passes."""
    K.require_cuda(low, "upsample_cross_entropy")
    return _UpsampleCE.apply(low, labels, size)


# --------------------------------------------------------------------------------------------------------
class _PaedBinaryStats(torch.autograd.Function):
    """per-image sums {bce, p*t, p, t, sdf_int*p, sdf_ext*edge} and the per-image max edge (7 columns)."""

    @staticmethod
    def forward(ctx, low, mask, sdf_ext, sdf_int):
        B = low.shape[0]
        stats = torch.zeros(B, 8, device=low.device, dtype=F32)
        keys = torch.zeros(B, device=low.device, dtype=torch.int64)
        K.paed_binary_stats(low, mask, sdf_ext, sdf_int, stats, keys)
        stats[:, 6] = (keys >> 32).to(torch.int32).view(F32)
        ctx.save_for_backward(low, mask, sdf_ext, sdf_int, keys)
        return stats[:, :7]

    @staticmethod
    def backward(ctx, gstats):
        low, mask, sdf_ext, sdf_int, keys = ctx.saved_tensors
        coef = torch.zeros(low.shape[0], 8, device=low.device, dtype=F32)
        coef[:, :7] = gstats
        dlow = torch.zeros_like(low)
        K.paed_binary_bwd(low, mask, sdf_ext, sdf_int, coef, keys, dlow)
        return dlow, None, None, None


def _all_reduce_sum(t: torch.Tensor, group) -> torch.Tensor:
    """Differentiable cross-rank sum (gradient of a sum w.r.t. each addend is the identity)."""

    class _AR(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            y = x.clone()
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
            return y

        @staticmethod
        def backward(ctx, g):
            return g

    return _AR.apply(t)


def paed_binary_loss(low, masks, sdf_ext, sdf_int, size, *, group=None, world_size=1):
    """loss = BCE + 0.1*Dice + 5*|ext - 0.5*int| of PAEDTrainer (model/PAED/classes.py:664-681), from the low-res
    single-channel logits.  masks/sdf_* fp32 [B,S,S].

    With world_size > 1 the five batch-wide sums are all-reduced so that the loss equals the single-process
    global-batch value (Dice and |paed| are not means of per-shard losses); the returned loss is
    the GLOBAL loss and its gradient w.r.t. local logits is exact (so gradients are summed, not averaged, over ranks)."""
    K.require_cuda(low, "paed_binary_loss")
    B = low.shape[0]
    low = low.contiguous()
    masks = masks.reshape(B, size, size).contiguous().to(F32)
    sdf_ext = sdf_ext.reshape(B, size, size).contiguous().to(F32)
    sdf_int = sdf_int.reshape(B, size, size).contiguous().to(F32)
    st = _PaedBinaryStats.apply(low, masks, sdf_ext, sdf_int)
    ext_img = st[:, 5] / (st[:, 6] + 1e-6)
    sums = torch.cat([st[:, :5].sum(0), ext_img.sum(0, keepdim=True)])
    npix = float(B * size * size)
    if world_size > 1:
        sums = _all_reduce_sum(sums, group)
        npix *= world_size
    bce = sums[0] / npix
    dice = 1.0 - (2.0 * sums[1] + 1e-6) / (sums[2] + sums[3] + 1e-6)
    paed = 1.0 * (sums[5] / npix) - 0.5 * (sums[4] / npix)
    return bce + 0.1 * dice + 5.0 * torch.abs(paed)


# --------------------------------------------------------------------------------------------------------
class _PaedMulticlass(torch.autograd.Function):
    @staticmethod
    def forward(ctx, low, labels, size):
        low = low.contiguous()
        labels = labels.contiguous()
        B, C, g, _ = low.shape
        need = ctx.needs_input_grad[0]
        bufs = _scratch(low.device, B * C * size * size, 3 if need else 2)
        loss_sum = torch.zeros(1, device=low.device, dtype=F32)
        dlow = torch.zeros_like(low) if need else None
        K.paed_multiclass(low, labels, bufs[0], bufs[1], bufs[2] if need else None, loss_sum, dlow)
        ctx.save_for_backward(dlow if need else torch.empty(0))
        ctx.count = float(B * C * size * size)
        return loss_sum[0] / ctx.count

    @staticmethod
    def backward(ctx, gout):
        (dlow,) = ctx.saved_tensors
        return dlow * (gout / ctx.count), None, None


_SCRATCH = {}
_RETIRED = []   # outgrown scratch buffers stay alive: a captured CUDA graph may have their addresses baked in


def _scratch(device, numel, n):
    """n fp32 scratch planes of >= numel elements, keyed by (device, stream, n): the buffers are stream-ordered
    temporaries, so two streams never share one, and growing never frees what an earlier capture recorded."""
    key = (device, torch.cuda.current_stream(device).cuda_stream, n)
    cur = _SCRATCH.get(key)
    if cur is None or cur[0].numel() < numel:
        if cur is not None:
            _RETIRED.append(cur)
        cur = [torch.empty(numel, device=device, dtype=F32) for _ in range(n)]
        _SCRATCH[key] = cur
    return cur


def paed_multiclass_soft_fused(low, labels, size):
    """paed_loss_multiclass_soft(one_hot(labels), softmax(bilinear_up(low))) with sigma=3, class_penalty=True."""
    K.require_cuda(low, "paed_multiclass_soft_fused")
    return _PaedMulticlass.apply(low, labels, size)


class _PaedMulticlassDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, msk, prob, class_penalty):
        msk = msk.contiguous().to(F32)
        prob = prob.contiguous()
        need = ctx.needs_input_grad[1]
        bufs = _scratch(prob.device, prob.numel(), 3 if need else 2)
        loss_sum = torch.zeros(1, device=prob.device, dtype=F32)
        dprob = torch.empty_like(prob) if need else None
        K.paed_multiclass_dense(msk, prob, bufs[0], bufs[1], bufs[2] if need else None, loss_sum, dprob, class_penalty)
        ctx.save_for_backward(dprob if need else torch.empty(0))
        ctx.count = float(prob.numel())
        return loss_sum[0] / ctx.count

    @staticmethod
    def backward(ctx, gout):
        (dprob,) = ctx.saved_tensors
        return None, dprob * (gout / ctx.count), None


def paed_multiclass_dense(msk, pred_mask, sigma=3, class_penalty=True):
    """model/PAED/classes.py:336-369 on dense [B,C,H,W] tensors (H == W); only sigma == 3 (the reference default and
    only call-site value) is implemented by the 19-tap kernel."""
    K.require_cuda(pred_mask, "paed_loss_multiclass_soft")
    if sigma != 3:
        raise ValueError("paed_loss_multiclass_soft: only sigma=3 is supported by the CUDA blur kernel")
    if pred_mask.dtype != F32:
        raise RuntimeError("paed_loss_multiclass_soft expects float32 probabilities (the reference builds its kernel "
                           "in fp32, SURVEY.md Appendix D8)")
    return _PaedMulticlassDense.apply(msk, pred_mask, class_penalty)
