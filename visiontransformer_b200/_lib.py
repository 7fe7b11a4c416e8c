"""ctypes binding of libvitseg.so (include/vitseg.h).  Host code passes raw device pointers and the current
CUDA stream; nothing here computes.  There is no CPU fallback: a missing library or device raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VS_LIB_PATH: load another build of the same library (A/B runs of kernel variants on one box)
LIB_PATH = os.environ.get("VS_LIB_PATH") or os.path.join(_HERE, "lib", "libvitseg.so")

_lib = None

c_void_p, c_int, c_i64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("a_mn_major", c_int), ("b_mn_major", c_int),
        ("A", c_void_p), ("lda", c_i64),
        ("B", c_void_p), ("ldb", c_i64),
        ("out", c_void_p), ("ldo", c_i64),
        ("out_dtype", c_int), ("accumulate", c_int),
        ("bias", c_void_p), ("act", c_int),
        ("out2", c_void_p), ("ldo2", c_i64),
        ("aux", c_void_p), ("ldaux", c_i64), ("aux_mode", c_int),
        ("residual", c_void_p), ("ldr", c_i64),
        ("row_tokens", c_int), ("split_k", c_int), ("tile_cfg", c_int),
        ("dropout_p", c_float), ("dropout_seed", c_void_p), ("dropout_site", C.c_uint32),
        ("out_colsum", c_void_p),
    ]


# name -> argtypes (all return int except vs_last_error)
_SIGS = {
    "vs_abi_version": [],
    "vs_device_sm_count": [],
    "vs_gemm_bf16": [C.POINTER(GemmDesc), c_void_p],
    "vs_colsum_bf16": [c_void_p, c_i64, c_int, c_int, c_void_p, c_int, c_void_p],
    "vs_colsum_cast_bf16": [c_void_p, c_i64, c_int, c_int, c_void_p, c_int, c_void_p, c_i64, c_int, c_void_p],
    "vs_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                         c_void_p],
    "vs_layernorm_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, C.c_uint32, c_void_p],
    "vs_attention_fwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, C.c_uint32,
                         c_void_p],
    "vs_attention_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                         c_float, c_float, c_void_p, C.c_uint32, c_void_p],
    "vs_dropout_rows": [c_void_p, c_void_p, c_i64, c_float, c_void_p, C.c_uint32, c_void_p],
    "vs_dropout_mask": [c_void_p, c_i64, c_int, c_int, c_float, c_void_p, C.c_uint32, c_void_p],
    "vs_patchify": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_cls_rows": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_embed_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_head_im2col": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_head_col2im": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_conv1x1_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_conv1x1_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_upsample_bilinear_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_upsample_bilinear_bwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_upsample_argmax": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_upsample_argmax_stats": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_colorize_mask": [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p],
    "vs_resample_h_u8": [c_void_p, c_i64, c_i64, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "vs_resample_v_u8": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p,
                         c_void_p],
    "vs_u8_to_f32": [c_void_p, c_void_p, c_i64, c_float, c_void_p],
    "vs_multimem_allreduce_f32": [c_void_p, c_i64, c_int, c_int, c_float, c_void_p],
    "vs_sdf_workspace_bytes": [c_int, c_int],
    "vs_sdf_targets": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "vs_upsample_ce": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vs_paed_binary_stats": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vs_paed_binary_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_void_p],
    "vs_paed_multiclass": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_int, c_void_p],
    "vs_paed_multiclass_dense": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_int, c_int, c_void_p],
    "vs_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_float, c_float,
                     c_float, c_float, c_int, c_float, c_int, c_i64, c_i64, c_void_p],
    "vs_cast_f32_bf16": [c_void_p, c_void_p, c_i64, c_void_p],
    "vs_cast_bf16_rows": [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_void_p],
    "vs_pack_conv3x3": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "vs_unpack_conv3x3_grad": [c_void_p, c_void_p, c_int, c_int, c_void_p],
}

EXPORTED_SYMBOLS = tuple(["vs_last_error", *_SIGS.keys()])


def _preload_cudart() -> None:
    """libvitseg.so links the CUDA runtime dynamically (libcudart.so.12).  `import torch` has normally mapped it
    already; if the loader cannot find it by name, take the copy that ships next to torch."""
    try:
        C.CDLL("libcudart.so.12", mode=C.RTLD_GLOBAL)
        return
    except OSError:
        pass
    import glob
    import sys
    for base in sys.path:
        for cand in glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib", "libcudart.so.12*")):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                continue


def load() -> C.CDLL:
    """Loads libvitseg.so (building it is __graft_entry__.build()'s job).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m visiontransformer_b200.build` "
            "(there is no CPU or PyTorch fallback for the ViT-segmentation kernels)")
    _preload_cudart()
    lib = C.CDLL(LIB_PATH)
    lib.vs_last_error.restype = C.c_char_p
    lib.vs_last_error.argtypes = []
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int64 if name.endswith("_bytes") else C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on {t.device}; visiontransformer_b200 runs only on a CUDA sm_100a device "
            "(no CPU fallback is shipped)")
