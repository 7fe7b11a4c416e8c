"""CPU: the drop-in boundary — module surface, state_dict layout, C-ABI symbols, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from oracle import vitseg_oracle as O
from visiontransformer_b200 import _lib
from visiontransformer_b200.ce.classes import LightningViTModel as CELightning
from visiontransformer_b200.ce.classes import ViTSegmentationModel
from visiontransformer_b200.engine import param_order
from visiontransformer_b200.model import flops_per_image
from visiontransformer_b200.paed.classes import LightningViTModel as PAEDLightning
from visiontransformer_b200.paed.classes import PAEDTrainer, paed_loss_multiclass_soft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_vitb16():
    m = CELightning(17, 16, 768, 12, 12)
    sd = m.state_dict()
    cfg = O.OracleConfig(17, 16, 768, 12, 12)
    expect = {"model." + k: v for k, v in O.param_shapes(cfg).items()}
    assert len(sd) == 204
    assert set(sd.keys()) == set(expect.keys())
    for k, shape in expect.items():
        assert tuple(sd[k].shape) == shape, k
        assert sd[k].dtype == torch.float32
    assert sum(p.numel() for p in m.parameters()) == 88_163_345


def test_state_dict_matches_hf_vit_names():
    """the backbone keys are exactly transformers.ViTModel's (the reference instantiates ViTModel(config))."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.ViTConfig(image_size=224, patch_size=16, num_channels=3, hidden_size=128, num_hidden_layers=2,
                                 num_attention_heads=2, intermediate_size=3072, qkv_bias=True)
    hf = transformers.ViTModel(cfg)
    ours = ViTSegmentationModel(17, 16, 128, 2, 2)
    hf_keys = {"backbone." + k: tuple(v.shape) for k, v in hf.state_dict().items()}
    our_keys = {k: tuple(v.shape) for k, v in ours.state_dict().items() if k.startswith("backbone.")}
    assert hf_keys == our_keys


def test_load_state_dict_strict_roundtrip():
    cfg = O.OracleConfig(17, 16, 128, 2, 2)
    sd = O.seeded_state_dict(cfg, 3)
    m = CELightning(17, 16, 128, 2, 2)
    m.load_state_dict(O.to_module_state_dict(sd, "model."), strict=True)
    out = m.state_dict()
    for k, v in sd.items():
        assert torch.equal(out["model." + k], v)
    m2 = ViTSegmentationModel(17, 16, 128, 2, 2)
    m2.load_state_dict({k[len("model."):]: v for k, v in out.items()}, strict=True)


def test_module_surface():
    m = ViTSegmentationModel(17, 16, 128, 2, 2)
    assert m.backbone.config.hidden_size == 128            # used at model/CE/classes.py:241
    assert isinstance(m.seg_head[0], torch.nn.Conv2d) and m.seg_head[0].kernel_size == (3, 3)
    assert isinstance(m.seg_head[2], torch.nn.Conv2d) and m.seg_head[2].out_channels == 17
    lt = CELightning(17, 16, 128, 2, 2)
    for attr in ("model", "loss_fn", "_resize_target", "training_step", "validation_step", "configure_optimizers"):
        assert hasattr(lt, attr)
    opt = lt.configure_optimizers()
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["lr"] == 1e-5
    assert len(opt.param_groups[0]["params"]) == len(list(lt.parameters()))
    pt = PAEDTrainer(1, 8, 128, 2, 2)
    conf = pt.configure_optimizers()
    assert isinstance(conf["optimizer"], torch.optim.AdamW) and conf["lr_scheduler"]["monitor"] == "val_IoU"
    pm = PAEDLightning(3, 16, 128, 2, 2)
    assert pm.num_classes == 17 and pm.model.seg_head[2].out_channels == 17   # model/PAED/classes.py:418
    assert pm.configure_optimizers().param_groups[0]["lr"] == 1e-4
    y = torch.randint(0, 17, (2, 256, 256))
    assert torch.equal(lt._resize_target(y, (224, 224)), O.resize_target(y, 224))


def test_param_order_covers_all_parameters():
    m = ViTSegmentationModel(17, 16, 128, 3, 2)
    assert sorted(param_order(3)) == sorted(n for n, _ in m.named_parameters())


def test_constructor_validation():
    with pytest.raises(ValueError):
        ViTSegmentationModel(17, 16, 100, 2, 2)      # head_dim != 64
    with pytest.raises(ValueError):
        ViTSegmentationModel(17, 16, 128, 2, 2, image_size=230)


def test_cpu_input_fails_loudly():
    m = ViTSegmentationModel(17, 16, 128, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        paed_loss_multiclass_soft(torch.zeros(1, 2, 8, 8), torch.zeros(1, 2, 8, 8))


def test_flops_formula_matches_baseline_md():
    cfg = ViTSegmentationModel(17, 16, 768, 12, 12).backbone.config
    assert abs(flops_per_image(cfg, False) / 1e9 - 35.82) < 0.01
    assert abs(flops_per_image(cfg, True) / 1e9 - 107.5) < 0.05


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vitseg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 25
    assert os.path.exists(_lib.LIB_PATH), "libvitseg.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vitseg.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(syms)
    lib.vs_abi_version.restype = ctypes.c_int
    assert lib.vs_abi_version() == 3


def test_compute_entry_fails_without_device():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    lib = _lib.load()
    d = _lib.GemmDesc()
    d.M = d.N = d.K = 128
    d.A = d.B = d.out = 64
    d.lda = d.ldb = d.ldo = 128
    rc = lib.vs_gemm_bf16(ctypes.byref(d), None)
    assert rc != 0
    assert b"CUDA" in lib.vs_last_error() or b"device" in lib.vs_last_error()
