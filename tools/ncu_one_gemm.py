"""One GEMM shape in isolation for an `ncu --set full` capture:
  python tools/ncu_one_gemm.py M N K [f32] [res] [drop] [gelu] [out2] [aux1] [bmn] [cfg=N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

M, N, Kd = (int(v) for v in sys.argv[1:4])
opts = sys.argv[4:]
cfg = next((int(o[4:]) for o in opts if o.startswith("cfg=")), 0)
dev = torch.device("cuda:0")
a = (torch.randn(M, Kd, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
f32 = "f32" in opts
out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
res = torch.randn(M, N, device=dev) if "res" in opts else None
seed = torch.tensor([7], device=dev, dtype=torch.int32)
drop = (0.1, seed, 3) if "drop" in opts else None
bmn = "bmn" in opts
wb = w.t().contiguous() if bmn else w
out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if "out2" in opts else None
aux = torch.randn(M, N, device=dev).to(torch.bfloat16) if "aux1" in opts else None


def run():
    K.gemm(a, wb, out, b_mn=bmn, bias=None if aux is not None else bias, residual=res, dropout=drop, tile_cfg=cfg,
           act=1 if "gelu" in opts else 0, out2=out2, aux=aux, aux_mode=1 if aux is not None else 0)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"M{M} N{N} K{Kd} {opts}: {e0.elapsed_time(e1) * 100:.1f} us/launch; tuned: {K.tuned_configs()}")
