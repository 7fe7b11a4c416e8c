"""Synthetic PAED targets for the benchmark tools (random discs + SDF maps from the package's own host-side
compute_sdf) — the tools must not import oracle/."""
import numpy as np
import torch

from visiontransformer_b200.paed.segmentation import compute_sdf


def binary_targets(B, S, seed=3):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:S, 0:S]
    masks, se, si = [], [], []
    for _ in range(B):
        m = np.zeros((S, S), dtype=bool)
        for _ in range(rng.randint(1, 4)):
            cy, cx, r = rng.randint(20, S - 20), rng.randint(20, S - 20), rng.randint(8, 40)
            m |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        e, i = compute_sdf(m.astype(np.uint8))
        masks.append(m.astype(np.float32)); se.append(e); si.append(i)
    return tuple(torch.from_numpy(np.stack(a)) for a in (masks, se, si))
