"""Attention micro-benchmark (ViT-B/16 shapes): forward / backward time with CUDA events, 20 reps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, N, H = int(os.environ.get("AB", 64)), int(os.environ.get("AN", 197)), int(os.environ.get("AH", 12))
D = H * 64
qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
ctx = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
dctx = torch.randn(B, N, H, 64, device=dev).bfloat16()
dqkv = torch.zeros(B, N, 3, H, 64, device=dev, dtype=torch.bfloat16)
dq_acc, delta = torch.empty(B, N, D, device=dev), torch.empty(B, H, N, device=dev)
seed = torch.tensor([1], device=dev, dtype=torch.int32)
x = torch.randn(B * N, D, device=dev)
g = torch.ones(D, device=dev)
y16 = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)


def t(name, fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/reps*1e3:8.1f} us", flush=True)


for p in (0.0, 0.1):
    drop = (p, seed, 1000) if p > 0 else None
    t(f"fwd p={p}", lambda: K.attention_fwd(qkv, ctx, lse, B, N, H, 0.125, dropout=drop))
    t(f"bwd p={p}", lambda: K.attention_bwd(qkv, ctx, dctx, lse, dqkv, dq_acc, delta, B, N, H, 0.125, dropout=drop))
    if N <= 256:
        t(f"bwd (short kernel) p={p}", lambda: K.attention_bwd(qkv, ctx, dctx, lse, dqkv, None, delta, B, N, H, 0.125, dropout=drop))
t("reference: layernorm_fwd (58 MB)", lambda: K.layernorm_fwd(x, g, g, 1e-12, y_bf16=y16))
