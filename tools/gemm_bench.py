"""GEMM micro-benchmark: TFLOP/s of vs_gemm_bf16 per shape and tile configuration (CUDA events, 20 reps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
CFG = {0: "auto", 1: "pair256", 2: "pair192", 3: "pair128", 4: "single256", 5: "single128"}


def bench(M, N, Kd, a_mn=False, b_mn=False, cfgs=(0, 1, 2, 3), reps=20, **kw):
    a = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    b = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if kw.get("f32") else torch.bfloat16)
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    res = torch.randn(M, N, device=dev) if kw.get("res") else None
    line = f"M{M} N{N} K{Kd} a{int(a_mn)}b{int(b_mn)} {kw}: "
    for cfg in cfgs:
        def run():
            K.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, bias=bias, residual=res, act=kw.get("act", 0),
                   accumulate=kw.get("acc", False), tile_cfg=cfg)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        line += f" {CFG[cfg]}={2.0*M*N*Kd/ms/1e9:7.1f}TF({ms*1e3:6.1f}us)"
    print(line, flush=True)


if __name__ == "__main__":
    if "model" not in sys.argv[1:]:
        # pure main loop: one wave exactly, long K
        bench(18944, 256, 16384, cfgs=(1, 4))
        bench(18944, 512, 8192, cfgs=(1, 4))
        bench(8192, 8192, 8192, cfgs=(1, 2, 4))
    else:
        # the data-gradient GEMMs of the ViT-B/16 step (MN-major weights), all pair configurations
        bench(12608, 768, 768, b_mn=True, cfgs=(0, 1, 2, 3))
        bench(12608, 768, 2304, b_mn=True, cfgs=(0, 1, 2, 3))
        bench(12608, 3072, 768, b_mn=True, cfgs=(0, 1, 2, 3))
    # model shapes
    bench(12608, 2304, 768, bias=True)
    bench(12608, 768, 768, bias=True, res=True, f32=True)
    bench(12608, 3072, 768, bias=True, act=1)
    bench(12608, 768, 3072, bias=True, res=True, f32=True)
    bench(12608, 768, 3072, b_mn=True)
    bench(768, 3072, 12608, a_mn=True, b_mn=True, f32=True, acc=True)
