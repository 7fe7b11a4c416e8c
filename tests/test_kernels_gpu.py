"""GPU (-m gpu): every libvitseg kernel against a plain PyTorch fp32 computation of the same op (tools/kernel_probe.py
holds the cases; this wrapper turns its report into pass/fail)."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    spec = importlib.util.spec_from_file_location("kernel_probe", os.path.join(ROOT, "tools", "kernel_probe.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("group", ["ln", "gemm", "attn", "head", "loss"])
def test_kernel_group(group):
    mod = _probe()
    mod.RESULTS.clear()
    getattr(mod, "probe_" + group)()
    torch.cuda.synchronize()
    bad = [(n, e, t) for (n, e, t, ok) in mod.RESULTS if not ok]
    assert len(mod.RESULTS) > 0
    assert not bad, bad


def test_native_library_is_loaded():
    """the CUDA path must be the one that ran: libvitseg.so is mapped into this process."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from visiontransformer_b200 import _lib
    _lib.load()
    with open("/proc/self/maps") as f:
        assert "libvitseg.so" in f.read()
