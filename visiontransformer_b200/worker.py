"""Worker-side inference pipeline: what happens either side of the model in the reference's
inference path, moved onto the GPU.

  reference (CPU)                                                        here
  ---------------------------------------------------------------------  ------------------------------------------
  PIL.Image.open(jpeg)                 model/CE/testViTModel.py:92        nvJPEG batched decode on the device
                                                                           (torchvision.io.decode_jpeg(device='cuda'))
  transforms.Resize((224, 224))        model/CE/testViTModel.py:93-96     vs_resample_h_u8 / vs_resample_v_u8: Pillow's
  transforms.ToTensor()                                                   fixed-point BILINEAR resize, bit-identical for
                                                                           the same pixels, + /255 into the batch tensor
  model(x).sigmoid().argmax()          model/CE/testViTModel.py:121-126   ViTSegmentationModel.predict_mask (fused)
  index_to_color[pred_labels]          model/CE/testViTModel.py:139-143   vs_colorize_mask
  PNG file posted as `mask_image`      backend/core/views.py:116-149      host: one D2H copy of the uint8 RGB batch,
                                                                           PNG (zlib) encoding per image

Numerical contract: the resize is Pillow's arithmetic exactly (tests/test_worker_*): given the same decoded RGB pixels
the fp32 input tensor equals ToTensor()(img.resize(...)) bit for bit.  JPEG decoding is nvJPEG's, not libjpeg-turbo's
(Pillow's): the two IDCT / chroma-upsampling implementations differ by a few grey levels on a small fraction of
pixels (measured in tests/test_worker_gpu.py on synthetic photos with hard-edged shapes: mean |difference| 0.9 grey
levels, 2 % of the values off by more than 4, up to 92 at hard chroma edges of 4:2:0 files, where nvJPEG replicates
chroma samples and libjpeg-turbo interpolates them), which is the only source of
difference between this pipeline's class maps and the reference's.  The HTTP callback itself is out of scope; the
bytes this module returns are what the worker posts."""
from __future__ import annotations

import io
import math
import time
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import kernels as K

PRECISION_BITS = 32 - 8 - 2   # libImaging/Resample.c


def pillow_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for the BILINEAR filter (support 1.0) over the
    full axis: -> (bounds int32 [out, 2] = (first source index, tap count), coefficients int32 [out, ksize], ksize)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size     # box coordinates are C floats in Pillow
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)      # C cast: truncation (the operand is never below -0.5 here)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        x = np.arange(n, dtype=np.float64)
        w = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(w < 1.0, 1.0 - w, 0.0)
        tot = 0.0
        for v in w:                              # Pillow accumulates the normaliser sequentially in double
            tot += float(v)
        if tot != 0.0:
            w = w / tot
        kk[xx, :n] = w
        bounds[xx] = (xmin, n)
    fixed = np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))
    return bounds, np.trunc(fixed).astype(np.int32), ksize


class _CoeffCache:
    def __init__(self, device):
        self.device = device
        self._tab: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, int]] = {}

    def get(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._tab.get(key)
        if t is None:
            b, c, k = pillow_bilinear_coeffs(in_size, out_size)
            t = (torch.from_numpy(b).to(self.device), torch.from_numpy(c).to(self.device), k)
            self._tab[key] = t
        return t


def resize_to_tensor(img_u8: torch.Tensor, out: torch.Tensor, cache: _CoeffCache) -> None:
    """uint8 [3,H,W] device image -> out fp32 [3,S,S] = ToTensor()(PIL.resize((S,S), BILINEAR)) (Pillow pass order:
    horizontal first; a pass whose size does not change is skipped, as ImagingResample does)."""
    Cn, H, W = img_u8.shape
    So_h, So_w = out.shape[-2], out.shape[-1]
    cur = img_u8
    if W != So_w:
        b, c, k = cache.get(W, So_w)
        tmp = torch.empty(Cn, H, So_w, device=img_u8.device, dtype=torch.uint8)
        K.resample_h(cur, b, c, k, So_w, tmp)
        cur = tmp
    if H != So_h:
        b, c, k = cache.get(H, So_h)
        K.resample_v(cur.contiguous(), b, c, k, So_h, dst_f32=out)
    else:
        K.u8_to_f32(cur.contiguous(), out)


class InferencePipeline:
    """JPEG bytes in -> PNG bytes (colour mask) out, everything between decode and colourisation on the GPU.

    model: ViTSegmentationModel (or a wrapper with .model) in eval mode on a CUDA device; palette: uint8 [C,3]
    (index_to_color of model/CE/testViTModel.py:136-138)."""

    def __init__(self, model, palette, input_size: int = 224, max_batch: int = 64):
        self.model = model.model if hasattr(model, "model") and hasattr(model.model, "predict_mask") else model
        self.device = next(self.model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("InferencePipeline needs the model on a CUDA device (no CPU path)")
        self.size = input_size
        self.max_batch = max_batch
        self.palette = torch.as_tensor(np.asarray(palette, dtype=np.uint8)).to(self.device).contiguous()
        self._coef = _CoeffCache(self.device)
        self.timings: Dict[str, float] = {}

    # ---- stages -------------------------------------------------------------------------------------------------
    def decode(self, jpegs: Sequence[bytes]) -> List[torch.Tensor]:
        """nvJPEG batched decode -> list of uint8 [3,H,W] device tensors (RGB)."""
        from torchvision.io import ImageReadMode, decode_jpeg
        datas = [torch.frombuffer(bytearray(j), dtype=torch.uint8) for j in jpegs]
        return decode_jpeg(datas, device=self.device, mode=ImageReadMode.RGB)

    def preprocess(self, images: Sequence[torch.Tensor]) -> torch.Tensor:
        """batch assembly: every decoded image resized (Pillow BILINEAR) and scaled into one fp32 [B,3,S,S] tensor."""
        x = torch.empty(len(images), 3, self.size, self.size, device=self.device, dtype=torch.float32)
        for i, img in enumerate(images):
            resize_to_tensor(img, x[i], self._coef)
        return x

    @torch.no_grad()
    def predict(self, x: torch.Tensor):
        """-> (uint8 class map [B,S,S], uint8 RGB mask image [B,S,S,3])."""
        mask = self.model.predict_mask(x)
        return mask, K.colorize_mask(mask, self.palette)

    @staticmethod
    def encode_png(rgb: np.ndarray) -> bytes:
        from PIL import Image
        buf = io.BytesIO()
        Image.fromarray(rgb, mode="RGB").save(buf, format="PNG", compress_level=1)
        return buf.getvalue()

    # ---- the whole path -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, jpegs: Sequence[bytes]) -> List[bytes]:
        """JPEG files -> PNG files of the colour masks (the `mask_image` payloads).  Stage times (seconds, device stages
        synchronised) are left in self.timings."""
        out: List[bytes] = []
        tm = {"decode": 0.0, "preprocess": 0.0, "model": 0.0, "d2h": 0.0, "png": 0.0}
        for s in range(0, len(jpegs), self.max_batch):
            chunk = jpegs[s:s + self.max_batch]
            t0 = time.perf_counter()
            imgs = self.decode(chunk)
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            x = self.preprocess(imgs)
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            _, rgb = self.predict(x)
            torch.cuda.synchronize(self.device)
            t3 = time.perf_counter()
            host = rgb.cpu().numpy()
            t4 = time.perf_counter()
            out.extend(self.encode_png(host[i]) for i in range(host.shape[0]))
            t5 = time.perf_counter()
            for k, v in zip(tm, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                tm[k] += v
        self.timings = tm
        return out
