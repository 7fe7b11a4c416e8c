"""Inference throughput (ViT-B/16, C=17): eager calls vs one CUDA graph per (batch, size), logits contract and fused mask."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402
from visiontransformer_b200.graph import GraphedInference  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, steps=20, warm=3):
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


for S, B in ((224, 64), (512, 32)):
    m = LightningViTModel(17, 16, 768, 12, 12, image_size=S).to(dev).eval()
    x = torch.rand(B, 3, S, S, device=dev)
    with torch.no_grad():
        e_log = timed(lambda: m(x))
        e_msk = timed(lambda: m.model.predict_mask(x))
        g_log = GraphedInference(m, x)
        g_msk = GraphedInference(m.model.predict_mask, x)
        t_log = timed(lambda: g_log(x))
        t_msk = timed(lambda: g_msk(x))
        assert torch.equal(g_msk(x), m.model.predict_mask(x))
    print(f"S={S} B={B}: logits eager {e_log:.2f} ms ({B / e_log * 1e3:.0f} img/s) graph {t_log:.2f} ms ({B / t_log * 1e3:.0f} img/s); "
          f"mask eager {e_msk:.2f} ms ({B / e_msk * 1e3:.0f} img/s) graph {t_msk:.2f} ms ({B / t_msk * 1e3:.0f} img/s)", flush=True)
    del m, g_log, g_msk
    torch.cuda.empty_cache()
