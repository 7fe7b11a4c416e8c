"""Loss-kernel micro-benchmark at the headline shape (B=64, C=17, 224x224 from a 14x14 grid): fused upsample+CE,
PAED multi-class (fused from low-res logits), PAED binary; forward+backward, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _synth import binary_targets  # noqa: E402
from visiontransformer_b200 import losses as LS  # noqa: E402

dev = torch.device("cuda:0")
B, C, g, S = 64, 17, 14, 224
torch.manual_seed(0)
low = torch.randn(B, C, g, g, device=dev, requires_grad=True)
y = torch.randint(0, C, (B, S, S), device=dev)
low1 = torch.randn(B, 1, g, g, device=dev, requires_grad=True)
masks, se, si = [t.to(dev) for t in binary_targets(B, S, seed=3)]


def t(name, fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:48s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us", flush=True)


def ce():
    low.grad = None
    LS.upsample_cross_entropy(low, y, S).backward()


def pm():
    low.grad = None
    LS.paed_multiclass_soft_fused(low, y, S).backward()


t("upsample + CE fwd+bwd", ce)
t("PAED multi-class soft (fused) fwd+bwd", pm)

from visiontransformer_b200.paed.segmentation import compute_sdf_batch  # noqa: E402

t("compute_sdf_batch (exact EDT, B=64, 224x224)", lambda: compute_sdf_batch(masks))
