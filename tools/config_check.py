"""Runs the other BASELINE.json configs on one GPU (per-GPU shard sizes) and prints images/s:
  cfg3  ViT-B/16 PAED training (PAEDTrainer C=1, and multi-class soft PAED C=17), batch 64/GPU
  cfg4  ViT-L/16 @384 training: reference-L (1024/16L/16h/I=3072) and canonical ViT-L (24L/I=4096), batch 32/GPU
  cfg5  ViT-B/16 @512 batched inference, batch 32/GPU: logits contract and fused uint8 mask"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _synth import binary_targets  # noqa: E402
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402
from visiontransformer_b200.graph import GraphedTrainStep  # noqa: E402
from visiontransformer_b200.model import flops_per_image  # noqa: E402
from visiontransformer_b200.optim import FusedAdam, FusedAdamW  # noqa: E402
from visiontransformer_b200.paed.classes import LightningViTModel as PAEDMulti  # noqa: E402
from visiontransformer_b200.paed.classes import PAEDTrainer  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_case(name, module, opt, batch, B):
    def step(b, i):
        loss = module.training_step(b, i)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss.detach()
    for i in range(2):
        step(batch, i)
    g = GraphedTrainStep(step, batch, warmup=2, engines=[module.model.engine])
    ms = timed(g.replay)
    cfg = module.model.backbone.config
    tf = B / (ms / 1e3) * flops_per_image(cfg, True) / 1e12
    print(f"{name:58s} {ms:8.2f} ms/step  {B/(ms/1e3):9.1f} img/s  {tf:7.1f} TFLOP/s  loss {g.static_loss.item():.4f}",
          flush=True)
    del g


torch.manual_seed(0)
which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
if "cfg3" in which:
    B = 64
    x = torch.rand(B, 3, 224, 224, device=dev)
    masks, se, si = [t.to(dev) for t in binary_targets(B, 224, seed=3)]
    m = PAEDTrainer(1, 16, 768, 12, 12).to(dev).train()
    train_case("cfg3 ViT-B/16 PAEDTrainer (BCE+Dice+|PAED|), B=64", m, FusedAdamW(m, lr=1e-4), (x, masks, se, si), B)
    del m
    y = torch.randint(0, 17, (B, 256, 256), device=dev)
    m = PAEDMulti(17, 16, 768, 12, 12).to(dev).train()
    train_case("cfg3 ViT-B/16 multi-class soft PAED, B=64", m, FusedAdam(m, lr=1e-4), (x, y), B)
    del m
    torch.cuda.empty_cache()
if "cfg4" in which:
    B = 32
    x = torch.rand(B, 3, 384, 384, device=dev)
    y = torch.randint(0, 17, (B, 384, 384), device=dev)
    m = LightningViTModel(17, 16, 1024, 16, 16, image_size=384).to(dev).train()
    train_case("cfg4 reference-L/16 (1024/16L/16h/I=3072) CE @384, B=32", m, FusedAdam(m, lr=1e-5), (x, y), B)
    del m
    torch.cuda.empty_cache()
    m = LightningViTModel(17, 16, 1024, 24, 16, image_size=384, intermediate_size=4096).to(dev).train()
    train_case("cfg4 canonical ViT-L/16 (1024/24L/16h/I=4096) CE @384, B=32", m, FusedAdam(m, lr=1e-5), (x, y), B)
    del m
    torch.cuda.empty_cache()
if "cfg5" in which:
    B = 32
    x = torch.rand(B, 3, 512, 512, device=dev)
    m = LightningViTModel(17, 16, 768, 12, 12, image_size=512).to(dev).eval()
    with torch.no_grad():
        ms = timed(lambda: m(x))
        ms2 = timed(lambda: m.model.predict_mask(x))
    cfg = m.model.backbone.config
    print(f"cfg5 ViT-B/16 @512 inference B=32: logits {ms:.2f} ms ({B/(ms/1e3):.0f} img/s, "
          f"{B/(ms/1e3)*flops_per_image(cfg, False)/1e12:.0f} TFLOP/s); fused mask {ms2:.2f} ms ({B/(ms2/1e3):.0f} img/s)",
          flush=True)
