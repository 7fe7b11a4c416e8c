"""CPU: the host side of the worker pre-processing — the restatement of Pillow's resampling coefficient tables
(libImaging/Resample.c precompute_coeffs / normalize_coeffs_8bpc) that the device kernels consume — checked by running
the same two-pass fixed-point arithmetic in numpy against PIL.Image.resize(BILINEAR), which is what
transforms.Resize((224, 224)) does to the PIL image at model/CE/testViTModel.py:92-97.  Bit-exact."""
import numpy as np
import pytest

from visiontransformer_b200.worker import PRECISION_BITS, pillow_bilinear_coeffs

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def numpy_two_pass(arr, oh, ow):
    """arr uint8 [H,W,C] -> uint8 [oh,ow,C] with the integer arithmetic of vs_resample_h_u8 / vs_resample_v_u8."""
    H, W, C = arr.shape
    cur = arr.astype(np.int64)
    if W != ow:
        b, c, _ = pillow_bilinear_coeffs(W, ow)
        out = np.zeros((H, ow, C), dtype=np.int64)
        for xx in range(ow):
            x0, n = b[xx]
            ss = (cur[:, x0:x0 + n, :] * c[xx, :n][None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = np.clip(ss >> PRECISION_BITS, 0, 255)
        cur = out
    if H != oh:
        b, c, _ = pillow_bilinear_coeffs(H, oh)
        out = np.zeros((oh, cur.shape[1], C), dtype=np.int64)
        for yy in range(oh):
            y0, n = b[yy]
            ss = (cur[y0:y0 + n] * c[yy, :n][:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
        cur = out
    return cur.astype(np.uint8)


@pytest.mark.parametrize("hw", [(480, 640), (354, 531), (224, 300), (100, 224), (224, 224), (225, 1000), (37, 53),
                                (1200, 1600)])
def test_coefficient_tables_reproduce_pillow_bilinear_resize(hw):
    rng = np.random.RandomState(hw[0] * 7 + hw[1])
    a = rng.randint(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(numpy_two_pass(a, 224, 224), ref)


def test_coefficient_tables_other_target_sizes():
    rng = np.random.RandomState(5)
    a = rng.randint(0, 256, (300, 411, 3), dtype=np.uint8)
    for S in (384, 512, 96):
        ref = np.asarray(Image.fromarray(a).resize((S, S), Image.BILINEAR))
        assert np.array_equal(numpy_two_pass(a, S, S), ref)
    b, c, k = pillow_bilinear_coeffs(1000, 224)
    assert b.shape == (224, 2) and c.shape == (224, k) and k == 11
    assert np.all(c.sum(1) >= (1 << PRECISION_BITS) - k) and np.all(c.sum(1) <= (1 << PRECISION_BITS) + k)
