"""CPU, world_size 2, gloo: the host logic of the data-parallel path — bucketed all-reduce of the flat gradient
arena (visiontransformer_b200/dp.py) and the cross-rank sums that make the PAEDTrainer loss a global-batch loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vitseg_oracle as O
from visiontransformer_b200.dp import BucketReducer, shard_batch
from visiontransformer_b200.engine import Engine, param_order
from visiontransformer_b200.losses import _all_reduce_sum


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _reduce_job(rank, world):
    torch.manual_seed(100 + rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    ranges = [("head", 700, 1000), ("layer1", 400, 700), ("layer0", 100, 400), ("embed", 0, 100)]
    red = BucketReducer(flat, ranges, average=True)
    for name, _, _ in ranges:      # backward completion order
        red.ready(name)
    red.finish()
    return mine, flat


def test_bucket_reducer_averages_across_ranks():
    out = _run(_reduce_job)
    expect = (out[0][0] + out[1][0]) / 2
    for r in (0, 1):
        assert torch.allclose(out[r][1], expect, atol=1e-6)


def _missing_bucket_job(rank, world):
    flat = torch.zeros(10)
    red = BucketReducer(flat, [("a", 0, 5), ("b", 5, 10)], average=False)
    red.ready("a")
    try:
        red.finish()
    except RuntimeError as e:
        return str(e)
    return ""


def test_bucket_reducer_detects_unreduced_bucket():
    out = _run(_missing_bucket_job)
    assert "never reduced" in out[0]


def _paed_global_job(rank, world):
    # global-batch PAED-binary loss from per-rank sums == single-process loss on the concatenated batch
    B, S = 4, 32
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(B, 1, S, S, generator=g)
    masks = (torch.rand(B, S, S, generator=g) > 0.6).float()
    se, si = torch.rand(B, S, S, generator=g), torch.rand(B, S, S, generator=g)
    full = O.paed_binary_step_loss(logits, masks, se, si)
    lo, m, e, i = shard_batch((logits, masks, se, si), rank, world)
    lo = lo.clone().requires_grad_(True)
    p = torch.sigmoid(lo)
    sx = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], dtype=torch.float32).view(1, 1, 3, 3)
    gx = torch.nn.functional.conv2d(p, sx, padding=1)
    gy = torch.nn.functional.conv2d(p, sx.transpose(2, 3), padding=1)
    edge = torch.sqrt(gx ** 2 + gy ** 2 + 1e-6)
    mx = edge.view(lo.shape[0], -1).max(1)[0]
    t = m[:, None]
    sums = torch.stack([
        torch.nn.functional.binary_cross_entropy(p, t, reduction="sum"), (p * t).sum(), p.sum(), t.sum(),
        (i[:, None] * p).sum(), ((e[:, None] * edge).flatten(1).sum(1) / (mx + 1e-6)).sum()])
    sums = _all_reduce_sum(sums, None)
    npix = float(B * S * S)
    loss = sums[0] / npix + 0.1 * (1 - (2 * sums[1] + 1e-6) / (sums[2] + sums[3] + 1e-6)) \
        + 5.0 * torch.abs(sums[5] / npix - 0.5 * sums[4] / npix)
    loss.backward()
    # reference gradient of the full-batch loss w.r.t. this rank's shard
    lf = logits.clone().requires_grad_(True)
    O.paed_binary_step_loss(lf, masks, se, si).backward()
    per = B // world
    return loss.item(), full.item(), (lo.grad - lf.grad[rank * per:(rank + 1) * per]).abs().max().item()


def test_paed_binary_global_batch_semantics():
    out = _run(_paed_global_job)
    for r in (0, 1):
        loss, full, gerr = out[r]
        assert abs(loss - full) < 1e-5 * abs(full)
        assert gerr < 1e-7


def test_shard_batch_even_split():
    x, y = torch.arange(8).view(8, 1), torch.arange(8)
    a = shard_batch((x, y), 1, 2)
    assert a[0].flatten().tolist() == [4, 5, 6, 7] and a[1].tolist() == [4, 5, 6, 7]
    with pytest.raises(ValueError):
        shard_batch((torch.zeros(7, 1),), 0, 2)


def test_bucket_ranges_partition_the_arena():
    """bucket ranges (engine.bucket_ranges) tile the arena exactly, in backward completion order."""
    from visiontransformer_b200.model import ViTSegmentationModel
    m = ViTSegmentationModel(17, 16, 128, 3, 2)
    eng = m.engine
    # emulate ensure_packed's slot layout on CPU (no device needed for the bookkeeping)
    off = 0
    from visiontransformer_b200.engine import _Slot
    params = dict(m.named_parameters())
    for n in param_order(3):
        eng.slots[n] = _Slot(n, off, params[n].numel(), tuple(params[n].shape))
        off += (params[n].numel() + 63) // 64 * 64
    eng.grads = torch.zeros(off)
    r = eng.bucket_ranges()
    assert [n for n, _, _ in r] == ["head", "layer2", "layer1", "layer0", "embed"]
    covered = sorted((s, e) for _, s, e in r)
    assert covered[0][0] == 0 and covered[-1][1] == off
    for (s0, e0), (s1, e1) in zip(covered, covered[1:]):
        assert e0 == s1


def _metric_counts_job(rank, world):
    """Each rank holds the pixel counts of its shard of the batch; the global metrics come from the summed counts."""
    from visiontransformer_b200 import metrics as M
    g = torch.Generator().manual_seed(7)
    pred = torch.randint(0, 5, (4, 24, 24), generator=g)
    tgt = torch.randint(0, 5, (4, 24, 24), generator=g)

    def counts(p, t):
        out = torch.zeros(p.shape[0], 5, 3, dtype=torch.int32)
        for c in range(5):
            out[:, c, 0] = ((p == c) & (t == c)).flatten(1).sum(1)
            out[:, c, 1] = (p == c).flatten(1).sum(1)
            out[:, c, 2] = (t == c).flatten(1).sum(1)
        return out
    (ps, ts) = shard_batch((pred, tgt), rank, world)
    tot = M.all_reduce_sum_counts(counts(ps, ts))
    return tot, M.pixel_accuracy(tot), counts(pred, tgt).sum(0, keepdim=True).to(torch.int64)


def test_metric_counts_all_reduce_to_the_global_batch_counts():
    out = _run(_metric_counts_job)
    for r in (0, 1):
        tot, acc, want = out[r]
        assert torch.equal(tot, want)
        assert abs(acc.item() - (want[0, :, 0].sum() / want[0, :, 1].sum()).item()) < 1e-7
