"""GPU (-m gpu): the worker-side inference pipeline (model/CE/testViTModel.py:92-97,121-143).

  1. device resize + ToTensor == ToTensor()(PIL.resize((S, S), BILINEAR)) BIT FOR BIT on the same decoded pixels;
  2. nvJPEG decode vs Pillow (libjpeg-turbo) decode of the same files: the stated tolerance of the pipeline;
  3. JPEG bytes -> PNG bytes: the PNGs decode to palette[class map], and the class maps agree with the reference-order
     pipeline (PIL decode -> PIL resize -> ToTensor -> same model) except where the decoders' grey-level differences flip
     a near-tie."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _photo(h, w, seed):
    """smooth photo-like content (low-frequency colour fields + a few hard-edged shapes)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.002, 0.02, 2).tolist() + [rng.uniform(0, 6.28)]
            img[..., c] += np.sin(xx * fx + yy * fy + ph)
    img = (img - img.min()) / (img.max() - img.min())
    for _ in range(5):
        cy, cx, r = rng.randint(0, h), rng.randint(0, w), rng.randint(10, max(11, min(h, w) // 4))
        img[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = rng.uniform(0, 1, 3)
    return (img * 255).astype(np.uint8)


def _jpeg(arr, quality=90, subsampling=2):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()


@pytest.mark.parametrize("hw,S", [((480, 640), 224), ((354, 531), 224), ((1200, 1600), 224), ((100, 224), 224),
                                 ((224, 224), 224), ((300, 411), 512), ((2000, 1500), 384)])
def test_device_resize_is_bit_identical_to_pillow(hw, S):
    from visiontransformer_b200.worker import _CoeffCache, resize_to_tensor
    dev = _dev()
    a = _photo(hw[0], hw[1], seed=hw[0] + hw[1])
    a[::7, ::5] = np.random.RandomState(1).randint(0, 256, a[::7, ::5].shape, dtype=np.uint8)   # hard pixel noise too
    ref = np.asarray(Image.fromarray(a).resize((S, S), Image.BILINEAR)).astype(np.float32) / 255.0   # ToTensor
    img = torch.from_numpy(a).permute(2, 0, 1).contiguous().to(dev)
    out = torch.empty(3, S, S, device=dev)
    resize_to_tensor(img, out, _CoeffCache(dev))
    assert torch.equal(out.cpu(), torch.from_numpy(ref).permute(2, 0, 1))


def test_pipeline_jpeg_to_png_matches_reference_order_pipeline():
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.worker import InferencePipeline
    dev = _dev()
    torch.manual_seed(3)
    m = LightningViTModel(17, 16, 128, 2, 2).to(dev).eval()
    with torch.no_grad():
        for p in m.model.seg_head.parameters():
            p.mul_(6.0)    # class maps with real structure instead of near-ties everywhere
    palette = np.random.RandomState(2).randint(0, 256, (17, 3)).astype(np.uint8)
    arrs = [_photo(480, 640, 11), _photo(354, 531, 12), _photo(768, 1024, 13), _photo(224, 224, 14)]
    jpegs = [_jpeg(arrs[0], 90, 2), _jpeg(arrs[1], 95, 0), _jpeg(arrs[2], 85, 2), _jpeg(arrs[3], 92, 1)]
    pipe = InferencePipeline(m, palette, input_size=224, max_batch=3)   # two chunks
    pngs = pipe.run(jpegs)
    assert len(pngs) == 4 and all(p[:8] == b"\x89PNG\r\n\x1a\n" for p in pngs)
    assert set(pipe.timings) == {"decode", "preprocess", "model", "d2h", "png"}
    # stage 2: decoder difference (nvJPEG vs libjpeg-turbo) on these files
    dec = pipe.decode(jpegs)
    worst, mean_abs, frac4 = 0, 0.0, 0.0
    for d, j in zip(dec, jpegs):
        ref = np.asarray(Image.open(io.BytesIO(j)).convert("RGB"))
        assert tuple(d.shape) == (3, ref.shape[0], ref.shape[1])
        diff = np.abs(d.permute(1, 2, 0).cpu().numpy().astype(int) - ref.astype(int))
        worst = max(worst, int(diff.max()))
        mean_abs = max(mean_abs, float(diff.mean()))
        frac4 = max(frac4, float(np.mean(diff > 4)))
    print(f"nvJPEG vs Pillow decode: max grey-level difference {worst}, worst file: mean |diff| {mean_abs:.3f}, "
          f"{100 * frac4:.2f} % of values differ by more than 4 levels")
    assert mean_abs < 1.5 and frac4 < 0.05   # measured: mean 0.91, 2.0 % (> 4 levels), max 92 at hard chroma edges of 4:2:0 files
    # stage 3: reference-order pipeline on the CPU side (PIL decode, PIL resize, ToTensor), same model
    x_ref = torch.stack([torch.from_numpy(np.asarray(Image.open(io.BytesIO(j)).convert("RGB").resize((224, 224), Image.BILINEAR))
                                          .astype(np.float32) / 255.0).permute(2, 0, 1) for j in jpegs]).to(dev)
    with torch.no_grad():
        ref_mask = m(x_ref).sigmoid().argmax(1).cpu().numpy()            # testViTModel.py:121-126
        ours_mask = m.model.predict_mask(pipe.preprocess(dec)).cpu().numpy()
    agree = float((ref_mask == ours_mask).mean())
    print(f"class-map agreement with the PIL-decoded reference-order pipeline: {agree:.4f}")
    # a RANDOM-weight model turns grey-level differences into class flips far more readily than a trained one: 94.6 %
    # measured here; the resize itself contributes nothing (bit-identical, first test)
    assert agree > 0.90
    for png, mk in zip(pngs, ours_mask):
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(png)).convert("RGB")), palette[mk])
