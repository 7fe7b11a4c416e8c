"""Offline quality check of the four-decisions-per-hash dropout generator of the attention kernels (drop_keep4 in
visiontransformer_b200/csrc/common.cuh), restated in numpy: keep rate, uniformity of the four 16-bit fields, pairwise
joint drop rates, autocorrelation of the mask and of each field across neighbouring quads.  No GPU needed."""
import numpy as np

M = np.uint64(0xFFFFFFFF)


def quad_fields(quad, seed):
    x = (quad ^ np.uint64(seed)) & M
    x = (x * np.uint64(0x9E3779B1)) & M
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & M
    x ^= x >> np.uint64(13)
    w = x * np.uint64(0xC2B2AE35)
    hi = w >> np.uint64(32)
    lo = (w & M) ^ hi
    f3 = ((x >> np.uint64(16)) ^ hi) & np.uint64(0xFFFF)
    return np.stack([lo & np.uint64(0xFFFF), lo >> np.uint64(16), hi & np.uint64(0xFFFF), f3], -1).astype(np.int64)


if __name__ == "__main__":
    n = 1 << 24
    thresh = int(0.1 * 65536 + 0.5)
    idx = np.arange(n, dtype=np.uint64)
    se = 1.0 / np.sqrt(n)
    for seed in (0x12345678, 0xDEADBEEF, 1, 0):
        f = quad_fields(idx, seed)
        keep = (f >= thresh).astype(np.float64)
        flat = keep.reshape(-1)
        k = flat - flat.mean()
        lags = " ".join("%d:%+.4f" % (lag, np.mean(k[:-lag] * k[lag:]) / k.var()) for lag in (1, 2, 3, 4, 8, 16, 197, 200))
        chi = " ".join("%.0f" % (((np.bincount(f[:, j] >> 8, minlength=256) - n / 256) ** 2 / (n / 256)).sum()) for j in range(4))
        d = f < thresh
        joint = " ".join("%d%d:%.5f" % (a, b, (d[:, a] & d[:, b]).mean()) for a in range(4) for b in range(a + 1, 4))
        fld = []
        for j in range(4):
            kj = keep[:, j] - keep[:, j].mean()
            fld.append("%+.5f" % (np.mean(kj[:-1] * kj[1:]) / kj.var()))
        print(f"seed {seed:#x}: keep {flat.mean():.5f} (expect {1 - thresh / 65536:.5f}); mask autocorrelation {lags}")
        print(f"   chi2(255) of the fields: {chi}; joint drop rates {joint} (expect {(thresh / 65536) ** 2:.5f})")
        print(f"   per-field correlation between neighbouring quads: {' '.join(fld)} (standard error {se:.5f})")
