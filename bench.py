"""bench.py — benchmarks of the ViT-segmentation hot path, one per BASELINE.json config.

  python bench.py --gpus N --steps K --warmup W              # headline: --config ce (BASELINE configs[1])
  python bench.py --config {ce,paed_bin,paed_multi,vitl384,infer512} ...
  python bench.py --impl reference [--config ...] ...        # the reference algorithm on the host CPU cores

  ce          ViT-B/16 cross-entropy segmentation training, batch 64 per GPU, 224x224 (weak scaling)     configs[1]
  paed_bin    ViT-B/16 PAEDTrainer (BCE + 0.1 Dice + 5|PAED|, C=1), GLOBAL batch 512 over N GPUs (strong)  configs[2]
  paed_multi  ViT-B/16 multi-class soft PAED (C=17), GLOBAL batch 512 over N GPUs (strong)                 configs[2]
  vitl384     ViT-L/16 (1024 / 24 layers / 16 heads / MLP 4096) CE training at 384x384, batch 32 per GPU   configs[3]
  infer512    ViT-B/16 inference at 512x512, 256 images sharded over N GPUs, fused uint8 mask (strong)     configs[4]

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same metric through
the public module API with pinned HOST batches (H2D copies and the D2H result read inside the timed region);
`roofline` = achieved TFLOP/s of the GEMM kernel measured with CUDA events inside eager steps of the same workload,
against BOTH measured cuBLAS peaks; `cpu_baseline` = the oracle port of the reference timed on this box's host cores
on a bounded sample; `library_baseline` = the reference's own module structure (transformers.ViTModel + seg head, i.e.
cuBLASLt + SDPA kernels) under torch.autocast(bf16) on the same GPU — a stated comparator, never the target;
`dp_check` (N > 1) = data-parallel equivalence on the NCCL ranks, checked before the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"
VITB = dict(patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
CONFIGS = {
    "ce": dict(kind="train", wrapper="ce", num_classes=17, arch=VITB, inter=3072, image=224, batch_per_gpu=64,
               lr=1e-5, metric="ViT-B/16 seg train images/sec", cpu_batch=8,
               workload="ViT-B/16 CE segmentation training, batch 64/GPU, 224x224, C=17 (BASELINE configs[1])"),
    "paed_bin": dict(kind="train", wrapper="paed_bin", num_classes=1, arch=VITB, inter=3072, image=224,
                     global_batch=512, lr=1e-4, metric="ViT-B/16 PAED-loss train images/sec", cpu_batch=8,
                     workload="ViT-B/16 PAEDTrainer training (BCE + 0.1 Dice + 5|PAED|, C=1), global batch 512 sharded "
                              "over the GPUs, 224x224 (BASELINE configs[2])"),
    "paed_multi": dict(kind="train", wrapper="paed_multi", num_classes=17, arch=VITB, inter=3072, image=224,
                       global_batch=512, lr=1e-4, metric="ViT-B/16 PAED-loss train images/sec", cpu_batch=8,
                       workload="ViT-B/16 multi-class soft PAED training (C=17), global batch 512 sharded over the "
                                "GPUs, 224x224 (BASELINE configs[2], secondary loss)"),
    "vitl384": dict(kind="train", wrapper="ce", num_classes=17,
                    arch=dict(patch_size=16, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16),
                    inter=4096, image=384, batch_per_gpu=32, lr=1e-5, metric="ViT-L/16 @384 seg train images/sec",
                    cpu_batch=2,
                    workload="ViT-L/16 (1024/24L/16h/MLP 4096) CE segmentation training, 384x384 (577 tokens), batch "
                             "32/GPU, C=17 (BASELINE configs[3])"),
    "infer512": dict(kind="infer", wrapper="ce", num_classes=17, arch=VITB, inter=3072, image=512, global_batch=256,
                     micro_batch=32, metric="ViT-B/16 seg inference images/sec", cpu_batch=4,
                     workload="ViT-B/16 batched inference, 512x512 (1025 tokens), 256 images sharded over the GPUs in "
                              "micro-batches of 32, fused uint8 class mask (BASELINE configs[4])"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        sm.append(float(c[1])); mx.append(float(c[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs (host side; shapes as the reference's dataset produces them)
# ------------------------------------------------------------------------------------------------------------------
def _disc_masks(B, S, gen):
    """1-3 random discs per image -> {0,1} masks [B,S,S] (PAEDTrainer targets; SDFs come from the device EDT)."""
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    out = torch.zeros(B, S, S)
    for b in range(B):
        for _ in range(int(torch.randint(1, 4, (1,), generator=gen))):
            cy, cx = (int(v) for v in torch.randint(20, S - 20, (2,), generator=gen))
            r = int(torch.randint(8, 40, (1,), generator=gen))
            out[b] += ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).float()
    return out.clamp_(max=1.0)


def _host_batch(cfg, B, gen, dev=None):
    """one pinned host batch of the config's training inputs."""
    S, C = cfg["image"], cfg["num_classes"]
    x = torch.rand(B, 3, S, S, generator=gen)
    if cfg["wrapper"] == "paed_bin":
        masks = _disc_masks(B, S, gen)
        if dev is not None:
            from visiontransformer_b200.paed.segmentation import compute_sdf_batch
            se, si = compute_sdf_batch(masks.to(dev))
            se, si = se.cpu(), si.cpu()
        else:
            import numpy as np
            from oracle import vitseg_oracle as O
            sdf = [O.compute_sdf(m.numpy().astype(np.uint8)) for m in masks]
            se = torch.stack([torch.from_numpy(a) for a, _ in sdf])
            si = torch.stack([torch.from_numpy(b) for _, b in sdf])
        batch = (x, masks, se, si)
    else:
        batch = (x, torch.randint(0, C, (B, 256, 256), generator=gen))   # dataset masks are 256x256 (CE/classes.py:77)
    if dev is not None:
        batch = tuple(t.pin_memory() for t in batch)
    return batch


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_time(cfg, steps: int, warmup: int):
    """one step of the config's workload (training: fwd + loss + bwd + Adam, fp32, dropout 0.1 as the reference's train
    mode; inference: eval forward + sigmoid + argmax) on cfg['cpu_batch'] images, restated by oracle/vitseg_oracle.py.
    Returns (seconds per step, threads, description of the sample)."""
    from oracle import vitseg_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    B, S = cfg["cpu_batch"], cfg["image"]
    ocfg = O.OracleConfig(num_classes=cfg["num_classes"], image_size=S, intermediate_size=cfg["inter"], **cfg["arch"])
    sd = O.seeded_state_dict(ocfg, seed=0, bf16_representable=False)
    gen = torch.Generator().manual_seed(1234)
    batch = _host_batch(cfg, B, gen)
    times = []
    if cfg["kind"] == "infer":
        with torch.no_grad():
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                O.forward(sd, batch[0], ocfg).sigmoid().argmax(1)     # testViTModel.py:121-126
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        what = f"{steps} timed eval forwards (+ sigmoid + argmax, fp32) of {B} images"
    else:
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        used = [v for k, v in leaves.items() if not k.startswith("backbone.pooler")]
        opt = (torch.optim.AdamW if cfg["wrapper"] == "paed_bin" else torch.optim.Adam)(used, lr=cfg["lr"])
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            logits = O.forward(leaves, batch[0], ocfg, dropout=(0.1, 0.1))
            if cfg["wrapper"] == "ce":
                loss = O.ce_loss(logits, O.resize_target(batch[1], S))
            elif cfg["wrapper"] == "paed_multi":
                loss = O.paed_multiclass_step_loss(logits, O.resize_target(batch[1], S))
            else:
                loss = O.paed_binary_step_loss(logits, O.resize_target(batch[1], S), batch[2], batch[3])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            loss.item()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        what = f"{steps} timed training steps (fwd + loss + bwd + Adam, fp32, dropout 0.1) on {B} images each"
    return statistics.median(times), threads, f"{what}, {warmup} warm-up, median"


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5 if cfg["cpu_batch"] >= 8 else 2))
    warm = 1
    sec, threads, sample = cpu_reference_time(cfg, steps, warm)
    val = cfg["cpu_batch"] / sec
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": _scaling(cfg),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + f" — CPU sample batch {cfg['cpu_batch']}", "name": args.config,
                   "dropout": 0.1 if cfg["kind"] == "train" else 0.0},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _scaling(cfg):
    return "weak" if "batch_per_gpu" in cfg else "strong"


# ------------------------------------------------------------------------------------------------------------------
# library comparator: the reference's module structure on the same GPU under stock PyTorch bf16 autocast
# ------------------------------------------------------------------------------------------------------------------
def library_baseline(cfg, dev, B, steps=5, warmup=3):
    """transformers.ViTModel (SDPA attention, cuBLASLt GEMMs) + the reference's seg head and upsample
    (model/CE/classes.py:222-262), torch.autocast(bf16), torch.optim.Adam(fused=True): what a user of the reference
    gets on this box today.  Losses: nn.CrossEntropyLoss / the oracle's restatement of the PAED losses."""
    import torch.nn.functional as F
    from torch import nn
    try:
        from transformers import ViTConfig, ViTModel
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"transformers not importable: {e}"}
    from oracle import vitseg_oracle as O   # comparator leg only: loss restatements, never on the product path

    S, C = cfg["image"], cfg["num_classes"]
    a = cfg["arch"]
    vc = ViTConfig(image_size=S, patch_size=a["patch_size"], num_channels=3, hidden_size=a["hidden_size"],
                   num_hidden_layers=a["num_hidden_layers"], num_attention_heads=a["num_attention_heads"],
                   intermediate_size=cfg["inter"], qkv_bias=True, hidden_dropout_prob=0.1,
                   attention_probs_dropout_prob=0.1)

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = ViTModel(vc)
            self.seg_head = nn.Sequential(nn.Conv2d(a["hidden_size"], 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, C, 1))

        def forward(self, x):
            f = self.backbone(x).last_hidden_state[:, 1:, :]
            Bq, T, D = f.shape
            g = int(T ** 0.5)
            f = f.transpose(1, 2).reshape(Bq, D, g, g)
            return F.interpolate(self.seg_head(f), size=x.shape[2:], mode="bilinear", align_corners=False)

    try:
        torch.manual_seed(0)
        m = Ref().to(dev)
        gen = torch.Generator().manual_seed(99)
        batch = tuple(t.to(dev) for t in _host_batch(cfg, B, gen, dev))

        def timed(fn):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        if cfg["kind"] == "infer":
            m.eval()

            def step():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    return m(batch[0]).sigmoid().argmax(1)
        else:
            m.train()
            opt = (torch.optim.AdamW if cfg["wrapper"] == "paed_bin" else torch.optim.Adam)(m.parameters(), lr=cfg["lr"], fused=True)

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits = m(batch[0])
                logits = logits.float()
                with torch.device(dev):   # the loss restatements build their small constant kernels with torch.tensor(...)
                    _loss_and_step(logits)

            def _loss_and_step(logits):
                if cfg["wrapper"] == "ce":
                    loss = F.cross_entropy(logits, O.resize_target(batch[1], S))
                elif cfg["wrapper"] == "paed_multi":
                    loss = O.paed_multiclass_step_loss(logits, O.resize_target(batch[1], S))
                else:
                    loss = O.paed_binary_step_loss(logits, O.resize_target(batch[1], S), batch[2], batch[3])
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
        ms = timed(step)
        out = {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": B,
               "what": "transformers.ViTModel (SDPA) + reference seg head under torch.autocast(bf16), "
                       + ("eval forward + sigmoid + argmax" if cfg["kind"] == "infer" else "fwd + loss + bwd + fused torch Adam, dropout 0.1")
                       + "; eager PyTorch (no CUDA graph, no torch.compile), inputs resident on the device"}
    except Exception as e:  # noqa: BLE001  (e.g. out of memory at the full batch)
        out = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    finally:
        m = opt = batch = None
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# data-parallel equivalence on the NCCL ranks (runs before the timed region when N > 1)
# ------------------------------------------------------------------------------------------------------------------
def dp_check(dev, rank, world):
    """N ranks each train a tiny model on their shard of a global batch through DataParallel; every rank also trains
    a single-process replica on the WHOLE batch.  CE (gradients averaged) and PAEDTrainer (6-scalar all-reduce inside
    the loss, gradients summed): losses and post-step weights must agree.  Plain SGD keeps the comparison linear in
    the gradients (Adam's g/sqrt(v) turns bf16-level noise on near-zero gradients into O(lr) weight differences)."""
    import torch.distributed as dist

    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.dp import DataParallel, shard_batch
    from visiontransformer_b200.paed.classes import PAEDTrainer

    gen = torch.Generator().manual_seed(4242)      # same inputs on every rank
    Bg = 4 * world
    tiny = dict(kind="train", image=224)
    ce_batch = tuple(t.to(dev) for t in _host_batch(dict(tiny, wrapper="ce", num_classes=17), Bg, gen, dev))
    pb_batch = tuple(t.to(dev) for t in _host_batch(dict(tiny, wrapper="paed_bin", num_classes=1), Bg, gen, dev))

    def build(cls, C):
        torch.manual_seed(7)                        # same weights on every rank and in both replicas
        m = cls(C, 16, 128, 2, 2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        with torch.no_grad():                        # a head with real signal (torch's default init is tiny)
            for p in m.model.seg_head.parameters():
                p.mul_(4.0)
            for p in m.parameters():                 # non-zero biases: a relative error needs a non-zero reference
                if p.dim() == 1:
                    p.add_(0.02 * torch.randn_like(p))
        return m.to(dev).train()

    def compare(cls, C, batch, steps=2):
        m = build(cls, C)
        dp = DataParallel(m, torch.optim.SGD(m.parameters(), lr=0.05))
        losses = [dp.step(shard_batch(batch, rank, world), i).item() for i in range(steps)]
        if cls is LightningViTModel:                # mean of equal-size shard means == global mean
            t = torch.tensor(losses, device=dev)
            dist.all_reduce(t)
            losses = (t / world).tolist()
        ref = build(cls, C)
        opt = torch.optim.SGD(ref.parameters(), lr=0.05)
        rl = []
        for i in range(steps):
            loss = ref.training_step(batch, i)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            rl.append(loss.item())
        werr = max(((p - q).abs().max() / q.abs().max().clamp_min(1e-12)).item()
                   for (_, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()))
        lerr = max(abs(p - q) / abs(q) for p, q in zip(losses, rl))
        return lerr, werr

    ce_l, ce_w = compare(LightningViTModel, 17, ce_batch)
    pb_l, pb_w = compare(PAEDTrainer, 1, pb_batch)
    # sharded inference: gathered masks equal the single-process masks
    m = build(LightningViTModel, 17).eval()
    got = DataParallel(m).predict_masks(ce_batch[0], gather=True)
    same = bool((got == m.model.predict_mask(ce_batch[0])).all())
    res = torch.tensor([ce_l, ce_w, pb_l, pb_w, 0.0 if same else 1.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)      # worst over ranks
    ce_l, ce_w, pb_l, pb_w, bad = res.tolist()
    ok = ce_l < 2e-3 and pb_l < 2e-3 and max(ce_w, pb_w) < 2e-2 and bad == 0.0
    return {"ce_loss_err": ce_l, "paed_loss_err": pb_l, "weight_err": max(ce_w, pb_w), "sharded_masks_equal": bad == 0.0,
            "ok": ok, "what": f"tiny model (128 wide, 2 layers), global batch {Bg}, 2 SGD steps: DataParallel over {world} "
                              "NCCL ranks vs a single-process replica on the whole batch; worst over ranks"}


# ------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch.distributed as dist

    from visiontransformer_b200 import kernels as K
    from visiontransformer_b200.dp import DataParallel
    from visiontransformer_b200.model import flops_per_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's version / debug banner goes to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    if "batch_per_gpu" in cfg:
        B = cfg["batch_per_gpu"]
    else:
        if cfg["global_batch"] % world != 0:
            raise ValueError(f"global batch {cfg['global_batch']} is not divisible by {world} GPUs")
        B = cfg["global_batch"] // world
    S, C = cfg["image"], cfg["num_classes"]
    train = cfg["kind"] == "train"

    check = dp_check(dev, rank, world) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    torch.manual_seed(0)
    drop = args.dropout if train else 0.0
    kw = dict(image_size=S, intermediate_size=cfg["inter"], hidden_dropout_prob=drop, attention_probs_dropout_prob=drop)
    a = cfg["arch"]
    margs = (C, a["patch_size"], a["hidden_size"], a["num_hidden_layers"], a["num_attention_heads"])
    if cfg["wrapper"] == "ce":
        from visiontransformer_b200.ce.classes import LightningViTModel as Wrapper
    elif cfg["wrapper"] == "paed_multi":
        from visiontransformer_b200.paed.classes import LightningViTModel as Wrapper
    else:
        from visiontransformer_b200.paed.classes import PAEDTrainer as Wrapper
    lm = Wrapper(*margs, **kw).to(dev)
    mcfg = lm.model.backbone.config
    peaks, peak_src = _peaks()
    peak_sus = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    peak_burst = float(peaks.get("bf16_tflops", peak_sus))
    gen = torch.Generator().manual_seed(1234 + rank)
    extra = {}

    if train:
        lm.train()
        # the reference's optimizers (CE/classes.py:296-297 Adam 1e-5; PAED/classes.py:486-487 Adam 1e-4, :536-548
        # AdamW 1e-4) as the one-pass kernel over the flat arenas; --torch-adam uses torch's fused optimizer instead
        if args.torch_adam:
            ocls = torch.optim.AdamW if cfg["wrapper"] == "paed_bin" else torch.optim.Adam
            opt = ocls(lm.parameters(), lr=cfg["lr"], fused=True, capturable=True)
            opt_name = f"torch {ocls.__name__}(lr={cfg['lr']}, fused, capturable)"
        else:
            from visiontransformer_b200.optim import FusedAdam, FusedAdamW
            opt = (FusedAdamW if cfg["wrapper"] == "paed_bin" else FusedAdam)(lm, lr=cfg["lr"])
            opt_name = f"{type(opt).__name__}(lr={cfg['lr']}) = vs_adam_step"
        dp = DataParallel(lm, opt)
        dp.broadcast_parameters()
        host = [_host_batch(cfg, B, gen, dev) for _ in range(2)]
        dbatch = tuple(t.to(dev) for t in host[0])

        from visiontransformer_b200.graph import GraphedTrainStep
        for i in range(2):
            dp.step(dbatch, i)           # eager warm-up (one-time attribute sets, arena build)
        K.reset_launch_count()
        dp.step(dbatch, 0)
        launches_per_step = K.launch_count()   # libvitseg kernels per step (torch glue kernels not counted)
        graphed = None if args.no_graph else GraphedTrainStep(lambda b, i: dp.step(b, i), dbatch, warmup=2,
                                                              engines=[lm.model.engine])

        def resident_step(i):
            if graphed is None:
                dp.step(dbatch, i)
            else:
                graphed.replay()

        # ---- e2e: pinned host batches, prefetched on a copy stream one step ahead, loss read back every step
        copy_stream = torch.cuda.Stream(device=dev)
        dev_b = [tuple(torch.empty_like(t) for t in dbatch) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        results = []

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                for d, h in zip(dev_b[slot], host[slot]):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_step(i):
            slot = i & 1
            if i == 0:
                prefetch(0)
            prefetch(slot ^ 1)  # next step's batch overlaps this step's compute
            torch.cuda.current_stream().wait_event(ready[slot])
            if graphed is None:
                loss = dp.step(dev_b[slot], i)
            else:
                loss = graphed(dev_b[slot])   # D2D into the graph's static inputs, then replay
            consumed[slot].record()
            # device -> host read of the step result, every step: the loss goes to pinned host memory by an async copy
            # queued behind the step, and the host collects it one step later — after it has queued the NEXT step —
            # so that reading the result does not leave the GPU idle for the launch latency of a 260-kernel graph
            # (a blocking loss.item() right here measured 0.10 ms per step, r02 s28).  flush_loss() collects the last one.
            loss_pin[slot].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[slot].record()
            if i > 0:
                loss_ev[slot ^ 1].synchronize()
                results.append(float(loss_pin[slot ^ 1]))
            pending[0] = slot

        loss_pin = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        pending = [None]

        def flush_loss():
            if pending[0] is not None:
                loss_ev[pending[0]].synchronize()
                results.append(float(loss_pin[pending[0]]))
                pending[0] = None

        h2d = sum(t.numel() * t.element_size() for t in host[0])
        d2h = 4
        e2e_note = ("pinned host batch prefetched one step ahead on a copy stream; loss copied to pinned host memory "
                    "behind every step and collected by the host one step later (pipelined read-back)")
        images_per_step = B
    else:
        # ---- inference: this rank's shard of the image set in micro-batches, fused upsample + argmax -> uint8 mask
        lm.eval()
        mb = min(cfg["micro_batch"], B)
        if B % mb != 0:
            raise ValueError(f"per-GPU shard {B} is not a multiple of the micro-batch {mb}")
        nmb = B // mb
        host_x = torch.rand(B, 3, S, S, generator=gen).pin_memory()
        host_out = torch.empty(B, S, S, dtype=torch.uint8).pin_memory()
        dx = host_x.to(dev)
        masks_dev = torch.empty(B, S, S, dtype=torch.uint8, device=dev)
        from visiontransformer_b200.graph import GraphedInference
        with torch.no_grad():
            for _ in range(2):
                lm.model.predict_mask(dx[:mb])
            K.reset_launch_count()
            lm.model.predict_mask(dx[:mb])
            launches_per_step = K.launch_count() * nmb
            graphed = None if args.no_graph else GraphedInference(lm.model.predict_mask, dx[:mb])

        def run_mb(x):
            if graphed is None:
                with torch.no_grad():
                    return lm.model.predict_mask(x)
            return graphed(x)

        def resident_step(i):
            for j in range(nmb):
                masks_dev[j * mb:(j + 1) * mb].copy_(run_mb(dx[j * mb:(j + 1) * mb]))

        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty(mb, 3, S, S, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        results = []

        def e2e_step(i):
            # H2D of micro-batch j+1 overlaps the forward of micro-batch j; masks go back to pinned host memory
            def fetch(j):
                s = j & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[s])
                    stage[s].copy_(host_x[j * mb:(j + 1) * mb], non_blocking=True)
                    ready[s].record(copy_stream)
            fetch(0)
            for j in range(nmb):
                s = j & 1
                if j + 1 < nmb:
                    fetch(j + 1)
                torch.cuda.current_stream().wait_event(ready[s])
                out = run_mb(stage[s])
                consumed[s].record()
                host_out[j * mb:(j + 1) * mb].copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            results.append(int(host_out[0, 0, 0]))

        for s_ in consumed:
            s_.record()
        h2d = host_x.numel() * 4
        d2h = host_out.numel()
        e2e_note = "pinned host images copied per micro-batch on a copy stream (overlapping the previous forward); uint8 masks copied back to pinned host memory"
        images_per_step = B

    # ---- warm-up, then `value`
    for i in range(max(3, args.warmup)):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(resident_step, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    launches = launches_per_step * args.steps
    value = world * images_per_step * args.steps / (ms / 1e3)

    if train:
        for s_ in consumed:
            s_.record()
    for i in range(2):
        e2e_step(i)
    if train:
        flush_loss()
    ms_e2e = timed(e2e_step, args.steps)
    if train:
        flush_loss()
    e2e_value = world * images_per_step * args.steps / (ms_e2e / 1e3)

    # ---- GEMM-kernel timing: the same step run eagerly with CUDA events around every vs_gemm_bf16 launch
    # (events cannot be timed inside a captured graph; same kernels, same shapes, same stream)
    K.enable_gemm_timing(True)
    barrier()
    n_eager = 3
    for i in range(n_eager):
        if train:
            dp.step(dbatch, i)
        else:
            with torch.no_grad():
                for j in range(nmb):
                    lm.model.predict_mask(dx[j * mb:(j + 1) * mb])
    gemm_events = K.collect_gemm_timing()
    K.enable_gemm_timing(False)

    if train and args.config == "ce":
        # ---- inference companions (same model): logits contract and fused mask path
        # Both run as ONE CUDA graph each (graph.GraphedInference, what worker.InferencePipeline replays): launched
        # eagerly, the ~110 launches of a forward cost about as much Python/ctypes enqueue time as the GPU needs for
        # the batch, and the two paths then differ by enqueue noise instead of by their kernels.
        from visiontransformer_b200.graph import GraphedInference
        lm.eval()
        with torch.no_grad():
            for _ in range(3):
                lm(dbatch[0])
            n_inf = max(5, args.steps)
            if args.no_graph:
                f_log, f_mask = (lambda i: lm(dbatch[0])), (lambda i: lm.model.predict_mask(dbatch[0]))
            else:
                g_log = GraphedInference(lm, dbatch[0])
                g_mask = GraphedInference(lm.model.predict_mask, dbatch[0])
                f_log, f_mask = (lambda i: g_log.graph.replay()), (lambda i: g_mask.graph.replay())
            # interleaved (A B A B) after a warm-up of both: the GPU runs power-capped, and whichever path is timed
            # second in a plain A-then-B order sees the lower clocks (r02 s22: the mask path, whose last kernel takes
            # 29 us against 83 us for the logits upsample — tools/head_bench.py — measured 7 % SLOWER that way)
            for i in range(n_inf):
                f_log(i)
                f_mask(i)
            ms_inf = ms_mask = 0.0
            for _ in range(2):
                ms_inf += 0.5 * timed(f_log, n_inf)
                ms_mask += 0.5 * timed(f_mask, n_inf)
        lm.train()
        extra["inference"] = {"logits_images_per_sec": world * B * n_inf / (ms_inf / 1e3),
                              "mask_images_per_sec": world * B * n_inf / (ms_mask / 1e3), "batch": B,
                              "cuda_graph": not args.no_graph,
                              "note": "eval forward of the same model, inputs resident: [B,17,224,224] fp32 logits (module contract) / fused uint8 mask"}
    if not train:
        with torch.no_grad():
            for _ in range(2):
                lm(dx[:mb])
            ms_log = timed(lambda i: [lm(dx[j * mb:(j + 1) * mb]) for j in range(nmb)], max(3, args.steps // 2))
        extra["logits_images_per_sec"] = world * B * max(3, args.steps // 2) / (ms_log / 1e3)
        extra["logits_note"] = "same shard through the module contract: [B,17,512,512] fp32 logits materialised (eager)"

    # ---- roofline of the dominant kernel (tcgen05 GEMM)
    fl_img = flops_per_image(mcfg, train)
    tot_flops = sum(f for f, _ in gemm_events)
    tot_ms = sum(t for _, t in gemm_events)
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic_r02.json")
    if os.path.exists(tpath) and args.config == "ce":  # dram bytes per GEMM launch from one ncu pass over this step
        with open(tpath) as f:
            traffic = float(json.load(f)["traffic_bytes_per_launch"])
    if tot_ms > 0:
        ach = tot_flops / (tot_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus,
                "peak_burst": peak_burst, "frac_burst": ach / peak_burst,
                "traffic": traffic, "traffic_unit": "bytes per launch (mean over the GEMM launches of one step; profiles/gemm_traffic_r02.json)",
                "kernel": "vs::gemm_kernel (tcgen05, all variants)", "launches_timed": len(gemm_events),
                "timing": f"CUDA events around each GEMM launch in {n_eager} eager (non-graph) steps of the same workload",
                "share_of_step": (tot_ms / n_eager) / (ms / args.steps),
                "peak_source": f"{peak_src}: frac vs bf16_tflops_sustained (kernel timed inside a long step), frac_burst vs bf16_tflops"}
    step_tflops = value * fl_img / 1e12

    lib = None
    if rank == 0 and world == 1 and not args.no_library_baseline:
        if train:
            del graphed
        lib = library_baseline(cfg, dev, min(B, 64 if S <= 224 else 32))

    if rank == 0:
        # ---- cpu baseline (oracle port on the host cores, bounded sample)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 3 if cfg["cpu_batch"] >= 8 else 2
            sec, threads, sample = cpu_reference_time(cfg, n_cpu, 1)
            cpu = {"value": cfg["cpu_batch"] / sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        conf = {"workload": cfg["workload"], "name": args.config, "global_batch": B * world, "batch_per_gpu": B,
                "parallelism": f"dp{world}", "cuda_graph": not args.no_graph, "dropout": drop,
                "l2": "working set per step (activations + weights, GBs) >> 126 MB L2; no flush needed"}
        if train:
            conf["optimizer"] = opt_name + " in timed region"
            if world > 1:
                conf["grad_allreduce"] = {"multimem": "in-switch (NVLS) reduction by vs_multimem_allreduce_f32, one co-resident CTA per SM, bucketed and overlapped with backward",
                                          "nccl": "bucketed NCCL all-reduce overlapped with backward"}.get(dp.reduce_backend, dp.reduce_backend)
                if dp.reduce_fallback_reason:
                    conf["grad_allreduce_fallback"] = dp.reduce_fallback_reason[:200]
            if cfg["wrapper"] != "paed_bin":
                conf["label_resize"] = "256 -> image-size nearest inside the step, as LightningViTModel.training_step"
        else:
            conf["micro_batch"] = mb
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": _scaling(cfg), "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": conf,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "note": e2e_note},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "library_baseline": lib,
            "dp_check": check,
            "model_tflops": {"achieved": step_tflops, "frac_of_sustained_peak": step_tflops / (peak_sus * world),
                             "frac_of_burst_peak": step_tflops / (peak_burst * world), "flops_per_image": fl_img,
                             "note": ("whole training step, 3x forward FLOPs" if train else "forward FLOPs") + " (BASELINE.md §4)"},
            "final_result": results[-1] if results else None,
        }
        line.update(extra)
        _emit(line)
    if world > 1:
        # NCCL communicators captured into the CUDA graph: tear down explicitly and leave without the
        # process-group destructor (it can wait forever on the captured work objects)
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


_STDOUT_FD = None


def _emit(line: dict) -> None:
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ce", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="use torch's fused optimizer instead of the arena kernel")
    ap.add_argument("--dropout", type=float, default=0.1, help="hidden/attention dropout (reference default 0.1)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    # stdout carries exactly one JSON line: anything a library writes to file descriptor 1 while the benchmark runs
    # (NCCL's version banner, for one) is sent to stderr instead; _emit() restores the descriptor for the result.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
