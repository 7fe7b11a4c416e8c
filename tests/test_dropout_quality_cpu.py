"""CPU: statistical quality of the attention dropout generator (drop_keep4, csrc/common.cuh) through its numpy
restatement in tools/dropout_quality.py — the GPU test test_dropout_gpu.py::test_attention_mask_matches_the_numpy_
restatement_of_the_generator pins the CUDA generator to this restatement bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from dropout_quality import quad_fields  # noqa: E402


def test_quad_generator_statistics():
    n = 1 << 21
    thresh = int(0.1 * 65536 + 0.5)
    idx = np.arange(n, dtype=np.uint64)
    se = 1.0 / np.sqrt(n)
    for seed in (0x12345678, 0xDEADBEEF, 1):
        f = quad_fields(idx, seed)
        keep = (f >= thresh).astype(np.float64)
        # keep rate of each of the four fields and overall
        assert abs(keep.mean() - (1 - thresh / 65536)) < 1e-3
        for j in range(4):
            assert abs(keep[:, j].mean() - (1 - thresh / 65536)) < 2e-3
            h = np.bincount(f[:, j] >> 8, minlength=256)
            chi2 = ((h - n / 256) ** 2 / (n / 256)).sum()
            assert chi2 < 400.0, (seed, j, chi2)          # 255 degrees of freedom: mean 255, sd 22.6
            kj = keep[:, j] - keep[:, j].mean()
            assert abs(np.mean(kj[:-1] * kj[1:]) / kj.var()) < 5 * se   # same field, neighbouring quads
        d = f < thresh
        for a in range(4):
            for b in range(a + 1, 4):
                assert abs((d[:, a] & d[:, b]).mean() - (thresh / 65536) ** 2) < 6e-4   # pairwise independence
        flat = keep.reshape(-1)
        k = flat - flat.mean()
        for lag in (1, 2, 3, 4, 8, 197, 200):
            assert abs(np.mean(k[:-lag] * k[lag:]) / k.var()) < 5 * se / 2 + 1e-3
