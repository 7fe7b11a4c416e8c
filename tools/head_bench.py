"""Micro-benchmark of the inference tail: bilinear upsample to fp32 logits vs fused upsample + argmax (uint8 mask)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")


def t(name, fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1)/reps*1e3:8.1f} us", flush=True)


for B, C, g, S in ((64, 17, 14, 224), (32, 17, 32, 512), (64, 1, 14, 224)):
    low = torch.randn(B, C, g, g, device=dev)
    full = torch.empty(B, C, S, S, device=dev)
    mask = torch.empty(B, S, S, device=dev, dtype=torch.uint8)
    t(f"upsample_fwd   B={B} C={C} g={g} S={S}", lambda: K.upsample_fwd(low, full))
    t(f"upsample_argmax B={B} C={C} g={g} S={S}", lambda: K.upsample_argmax(low, mask))
    ref = full.sigmoid().argmax(1) if C > 1 else (full[:, 0] > 0).long()
    K.upsample_fwd(low, full)
    K.upsample_argmax(low, mask)
    print("   agreement with sigmoid().argmax():", (ref == mask.long()).float().mean().item())

# ---- glue kernels around the head / embeddings (ViT-B/16, B = 64): each timed with a 256 MB L2 flush between reps
B, g, D, P, S, C, F = 64, 14, 768, 16, 224, 17, 256
T1 = g * g + 1
flush = torch.empty(64 * 1024 * 1024, device=dev)


def tf(name, fn, reps=10):
    fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"{name:50s} {tot/reps*1e3:8.1f} us (cold L2)", flush=True)


img = torch.rand(B, 3, S, S, device=dev)
patches = torch.zeros(B * T1, 3 * P * P, device=dev, dtype=torch.bfloat16)
tok = torch.randn(B * T1, D, device=dev).bfloat16()
col = torch.empty(B * g * g, 9 * D, device=dev, dtype=torch.bfloat16)
dtok = torch.empty(B * T1, D, device=dev)
feat = torch.randn(B * g * g, F, device=dev).bfloat16()
w = torch.randn(C, F, device=dev)
bias = torch.randn(C, device=dev)
low = torch.empty(B, C, g, g, device=dev)
dlow = torch.randn(B, C, g, g, device=dev)
dfeat = torch.empty_like(feat)
dw, db = torch.zeros(C, F, device=dev), torch.zeros(C, device=dev)
dx = torch.randn(B * T1, D, device=dev)
dcls, dpos, dpb = torch.zeros(D, device=dev), torch.zeros(T1 * D, device=dev), torch.zeros(D, device=dev)
tf("patchify", lambda: K.patchify(img, patches, P))
tf("head_im2col", lambda: K.head_im2col(tok, col, B, g, D))
tf("head_col2im", lambda: K.head_col2im(col, dtok, B, g, D))
tf("conv1x1_fwd", lambda: K.conv1x1_fwd(feat, w, bias, low, B, g, F, C))
tf("conv1x1_bwd", lambda: K.conv1x1_bwd(dlow, feat, w, dfeat, dw, db, B, g, F, C))
tf("embed_bwd", lambda: K.embed_bwd(dx, dcls, dpos, dpb, B, T1, D))
labels = torch.randint(0, C, (B, 256, 256), device=dev)
loss_sum = torch.zeros(2, device=dev)
dlo = torch.zeros(B, C, g, g, device=dev)
tf("upsample_ce (int64 labels 256x256)", lambda: K.upsample_ce(low, labels, loss_sum, dlo, S))
mask8 = torch.empty(B, S, S, device=dev, dtype=torch.uint8)
tf("upsample_argmax", lambda: K.upsample_argmax(low, mask8))
