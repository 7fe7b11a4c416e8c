"""Attribute epilogue cost: fc1-shaped GEMM (12608 x 3072 x 768) with epilogue options toggled."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
M, N, Kd = 12608, 3072, 768
a = (torch.randn(M, Kd, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16)
wt = w.t().contiguous()
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
outf = torch.empty(M, N, device=dev)
res = torch.randn(M, N, device=dev)


def t(name, fn, reps=30):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms*1e3:7.1f} us  {2.0*M*N*Kd/ms/1e9:7.1f} TFLOP/s", flush=True)


for cfg in [int(c) for c in os.environ.get("EPI_CFGS", "0,1,2,3,4").split(",")]:
    print("tile_cfg", cfg)
    t("plain bf16", lambda: K.gemm(a, w, out, tile_cfg=cfg))
    t("bias", lambda: K.gemm(a, w, out, bias=bias, tile_cfg=cfg))
    t("bias+gelu", lambda: K.gemm(a, w, out, bias=bias, act=1, tile_cfg=cfg))
    t("bias+out2", lambda: K.gemm(a, w, out, bias=bias, out2=out2, tile_cfg=cfg))
    t("bias+gelu+out2 (fc1)", lambda: K.gemm(a, w, out, bias=bias, act=1, out2=out2, tile_cfg=cfg))
    t("aux gelu' (fc2 dgrad-like, K-major B)", lambda: K.gemm(a, w, out, aux=aux, aux_mode=1, tile_cfg=cfg))
    t("aux gelu' MN-major B", lambda: K.gemm(a, wt, out, b_mn=True, aux=aux, aux_mode=1, tile_cfg=cfg))
    t("plain MN-major B", lambda: K.gemm(a, wt, out, b_mn=True, tile_cfg=cfg))
    t("fp32 out plain", lambda: K.gemm(a, w, outf, tile_cfg=cfg))
    t("fp32 out bias+res", lambda: K.gemm(a, w, outf, bias=bias, residual=res, tile_cfg=cfg))
