"""LayerNorm micro-benchmark (ViT-B/16 token matrix 12608 x 768): forward / backward variants, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
M, D = int(os.environ.get("LM", 12608)), int(os.environ.get("LD", 768))
x = torch.randn(M, D, device=dev)
g = torch.randn(D, device=dev)
b = torch.randn(D, device=dev)
y16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
dy16 = torch.randn(M, D, device=dev).bfloat16()
dxin = torch.randn(M, D, device=dev)
dx = torch.empty(M, D, device=dev)
dx16 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
seed = torch.tensor([1], device=dev, dtype=torch.int32)
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # L2 flush between timed launches


def t(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / reps * 1e3
    print(f"{name:48s} {us:8.1f} us  {nbytes / us / 1e6:6.2f} TB/s", flush=True)


n = M * D
K.layernorm_fwd(x, g, b, 1e-12, y_bf16=y16, mean=mean, rstd=rstd)
t("fwd (x f32 -> y bf16)", lambda: K.layernorm_fwd(x, g, b, 1e-12, y_bf16=y16, mean=mean, rstd=rstd), n * 6)
t("bwd dy16, dx_in, dx f32 + bf16 (per-layer form)", lambda: K.layernorm_bwd(dy16, x, g, mean, rstd, dxin, dx, dx16, dg, db), n * 16)
t("bwd same + dropout 0.1 on the bf16 copy", lambda: K.layernorm_bwd(dy16, x, g, mean, rstd, dxin, dx, dx16, dg, db, dropout=(0.1, seed, 3)), n * 16)
t("bwd dy16, no dx_in, dx f32 + bf16", lambda: K.layernorm_bwd(dy16, x, g, mean, rstd, None, dx, dx16, dg, db), n * 12)
t("bwd dy16, no dx_in, dx f32 only", lambda: K.layernorm_bwd(dy16, x, g, mean, rstd, None, dx, None, dg, db), n * 10)
t("colsum bf16 [M, D]", lambda: K.colsum(dx16, dg, accumulate=True), n * 2)
t("copy f32 (torch) [M, D]", lambda: dx.copy_(dxin), n * 8)
