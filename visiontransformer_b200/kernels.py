"""Tensor-level wrappers over the C ABI (one Python function per vs_* entry point).

These are plumbing: argument checking, pointer extraction, the current stream.  Every function requires CUDA
tensors and raises otherwise — nothing here falls back to PyTorch math."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import GemmDesc, check, ptr, require_cuda, stream

BF16, F32 = torch.bfloat16, torch.float32
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
AUX_NONE, AUX_GELU_GRAD, AUX_RELU_MASK = 0, 1, 2


# --- bookkeeping used by bench.py: kernels launched by this process and CUDA-event timing of the GEMM kernel -------
_LAUNCHES = 0
_GEMM_TIMING = False
_GEMM_EVENTS = []


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


def launch_count() -> int:
    return _LAUNCHES


def enable_gemm_timing(flag: bool) -> None:
    """Bracket every vs_gemm_bf16 launch with CUDA events on the launching stream (bench.py roofline)."""
    global _GEMM_TIMING
    _GEMM_TIMING = flag
    if flag:
        _GEMM_EVENTS.clear()


def collect_gemm_timing():
    """-> list of (algorithmic FLOPs, milliseconds) per GEMM launch recorded since enable_gemm_timing(True)."""
    torch.cuda.synchronize()
    out = [(fl, e0.elapsed_time(e1)) for fl, e0, e1 in _GEMM_EVENTS]
    _GEMM_EVENTS.clear()
    return out


def _rowmajor(t: torch.Tensor, what: str) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, f"{what}: expected a row-major 2-D view, got strides {t.stride()}"
    return t.stride(0)


# --- tile-configuration autotuning ---------------------------------------------------------------------------------
# The library's cost model (gemm.cu: choose_tiles) is a fit over a handful of shapes; which of the five tile
# configurations wins depends on the epilogue (bf16 TMA-store vs fp32 transposed vs atomics) as much as on the shape.
# The first eager call of every (shape, operand majors, epilogue) combination therefore times the five candidates on
# the real operands (CUDA events, 2 warm-up + 4 timed launches each) and caches the winner; calls made while a CUDA
# graph is being captured use the cache (or the cost model if the shape was never seen eagerly).
# VS_GEMM_AUTOTUNE=0 turns it off.
_AUTOTUNE = os.environ.get("VS_GEMM_AUTOTUNE", "1") != "0"
_TUNED = {}


def _autotune(d, out, accumulate):
    lib = _lib.load()
    st = stream()
    saved_out, saved_colsum = d.out, d.out_colsum
    scratch = None
    cs_scratch = None
    if d.out_colsum:   # trial launches must not add into the caller's bias-gradient buffer either
        cs_scratch = torch.zeros(d.N, device=out.device, dtype=F32)
        d.out_colsum = ptr(cs_scratch)
    if accumulate:   # trial launches must not add into the caller's gradient buffer
        scratch = torch.empty_like(out)
        d.out, d.ldo = ptr(scratch), scratch.stride(0)
    best, best_ms = 0, float("inf")
    try:
        for cfg in (1, 2, 3, 4, 5):
            d.tile_cfg = cfg
            ok = True
            for _ in range(3):
                ok = ok and lib.vs_gemm_bf16(C.byref(d), st) == 0
            if not ok:
                continue
            # best of two 8-launch batches: a 4-launch batch mis-ranked configurations 5 % apart (r02 session 13:
            # 21.8 us picked where the forced pair256 ran 20.7 us)
            ms = float("inf")
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(8):
                    lib.vs_gemm_bf16(C.byref(d), st)
                e1.record()
                e1.synchronize()
                ms = min(ms, e0.elapsed_time(e1))
            if ms < best_ms:
                best, best_ms = cfg, ms
    finally:
        d.out = saved_out
        d.out_colsum = saved_colsum
        if accumulate:
            d.ldo = out.stride(0)
        d.tile_cfg = 0
    return best


def tuned_configs():
    """{(M, N, K, a_mn, b_mn, ...): tile_cfg} chosen so far (bench.py / tools report it)."""
    return dict(_TUNED)


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, act=ACT_NONE, out2=None, aux=None, aux_mode=AUX_NONE,
         residual=None, row_tokens=0, accumulate=False, split_k=0, tile_cfg=0, M=None, N=None, K=None,
         dropout=None, colsum=None):
    """out[M,N] = epilogue(sum_k A(m,k) B(n,k)); see include/vitseg.h:vs_gemm_desc.

    a: bf16 [M,K] (or [K,M] when a_mn); b: bf16 [N,K] (or [K,N] when b_mn); out: bf16 or fp32 2-D view."""
    require_cuda(a, "gemm")
    assert a.dtype == BF16 and b.dtype == BF16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.A, d.lda = ptr(a), _rowmajor(a, "A")
    d.B, d.ldb = ptr(b), _rowmajor(b, "B")
    d.out, d.ldo = ptr(out), _rowmajor(out, "out")
    assert out.dtype in (BF16, F32)
    d.out_dtype = 1 if out.dtype == F32 else 0
    d.accumulate = int(accumulate)
    if bias is not None:
        assert bias.dtype == F32 and bias.is_contiguous()
    d.bias = ptr(bias)
    d.act = act
    if out2 is not None:
        assert out2.dtype == BF16
        d.out2, d.ldo2 = ptr(out2), _rowmajor(out2, "out2")
    if aux is not None:
        assert aux.dtype == BF16
        d.aux, d.ldaux = ptr(aux), _rowmajor(aux, "aux")
    d.aux_mode = aux_mode
    if residual is not None:
        assert residual.dtype == F32
        d.residual, d.ldr = ptr(residual), _rowmajor(residual, "residual")
    if colsum is not None:   # out_colsum[n] += sum_m out[m, n] (bias gradient of the producing Linear), bf16 outputs
        assert colsum.dtype == F32 and colsum.is_contiguous() and colsum.numel() == N and out.dtype == BF16
        d.out_colsum = ptr(colsum)
    d.row_tokens = row_tokens
    d.split_k = split_k
    d.tile_cfg = tile_cfg
    if dropout is not None and dropout[0] > 0.0:   # (p, seed tensor (uint32/int32 device), site)
        d.dropout_p, d.dropout_seed, d.dropout_site = dropout[0], ptr(dropout[1]), dropout[2]
    if tile_cfg == 0 and _AUTOTUNE:
        key = (M, N, K, bool(a_mn), bool(b_mn), d.out_dtype, bool(accumulate), bias is not None, act, out2 is not None,
               aux_mode, residual is not None, row_tokens > 0, d.dropout_p > 0.0, split_k, colsum is not None)
        cfg = _TUNED.get(key)
        if cfg is None and not torch.cuda.is_current_stream_capturing():
            cfg = _TUNED[key] = _autotune(d, out, accumulate)
        d.tile_cfg = cfg or 0
    if _GEMM_TIMING:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.load().vs_gemm_bf16(C.byref(d), stream()), "vs_gemm_bf16")
        e1.record()
        _GEMM_EVENTS.append((2.0 * M * N * K, e0, e1))
    else:
        check(_lib.load().vs_gemm_bf16(C.byref(d), stream()), "vs_gemm_bf16")
    _count(1)
    return out


def colsum(x, out, accumulate=True):
    require_cuda(x, "colsum")
    assert x.dtype == BF16 and out.dtype == F32
    _count(1)
    check(_lib.load().vs_colsum_bf16(ptr(x), _rowmajor(x, "x"), x.shape[0], x.shape[1], ptr(out), int(accumulate),
                                    stream()), "vs_colsum_bf16")
    return out


def colsum_cast(x, x_f32, out, accumulate=True):
    """column sums of bf16 x [M,N] whose first x_f32.shape[1] columns are taken from fp32 x_f32 (rounded and stored
    into x on the way): dQKV bias gradient + dQ cast in one pass."""
    require_cuda(x, "colsum_cast")
    assert x.dtype == BF16 and x_f32.dtype == F32 and out.dtype == F32 and x_f32.shape[0] == x.shape[0]
    _count(1)
    check(_lib.load().vs_colsum_cast_bf16(ptr(x), _rowmajor(x, "x"), x.shape[0], x.shape[1], ptr(out), int(accumulate),
                                         ptr(x_f32), _rowmajor(x_f32, "x_f32"), x_f32.shape[1], stream()),
          "vs_colsum_cast_bf16")
    return out


def layernorm_fwd(x, gamma, beta, eps, y_bf16=None, y_f32=None, mean=None, rstd=None):
    require_cuda(x, "layernorm_fwd")
    assert x.dtype == F32 and x.is_contiguous()
    M, D = x.shape
    _count(1)
    check(_lib.load().vs_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), eps, M, D, ptr(y_bf16), ptr(y_f32), ptr(mean),
                                      ptr(rstd), stream()), "vs_layernorm_fwd")


def _drop(dropout):
    if dropout is None or dropout[0] <= 0.0:
        return 0.0, None, 0
    return float(dropout[0]), ptr(dropout[1]), int(dropout[2])


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_in, dx_out, dx_bf16, dgamma, dbeta, dropout=None, dbias=None):
    require_cuda(x, "layernorm_bwd")
    M, D = x.shape
    assert dy.is_contiguous() and dy.dtype in (BF16, F32)
    _count(1)
    check(_lib.load().vs_layernorm_bwd(ptr(dy), int(dy.dtype == F32), ptr(x), ptr(gamma), ptr(mean), ptr(rstd),
                                      ptr(dx_in), M, D, ptr(dx_out), ptr(dx_bf16), ptr(dgamma), ptr(dbeta), ptr(dbias),
                                      *_drop(dropout),
                                      stream()),
          "vs_layernorm_bwd")


def attention_fwd(qkv, ctx, lse, B, N, H, scale, dropout=None):
    require_cuda(qkv, "attention_fwd")
    assert qkv.dtype == BF16 and qkv.is_contiguous() and ctx.is_contiguous()
    _count(1)
    check(_lib.load().vs_attention_fwd(ptr(qkv), ptr(ctx), ptr(lse), B, N, H, scale, *_drop(dropout), stream()),
          "vs_attention_fwd")


def attention_bwd(qkv, ctx, dctx, lse, dqkv, dq_accum, delta, B, N, H, scale, dropout=None):
    require_cuda(qkv, "attention_bwd")
    _count(2)
    check(_lib.load().vs_attention_bwd(ptr(qkv), ptr(ctx), ptr(dctx), ptr(lse), ptr(dqkv), ptr(dq_accum), ptr(delta), B,
                                      N, H, scale, *_drop(dropout), stream()), "vs_attention_bwd")


def dropout_rows(x, x_bf16, dropout):
    """in-place dropout of a contiguous fp32 tensor (+ optional bf16 copy); dropout = (p, seed tensor, site)."""
    require_cuda(x, "dropout_rows")
    assert x.dtype == F32 and x.is_contiguous()
    _count(1)
    check(_lib.load().vs_dropout_rows(ptr(x), ptr(x_bf16), x.numel(), *_drop(dropout), stream()), "vs_dropout_rows")


def dropout_mask(out, scheme, dropout, row_len=0):
    assert out.dtype == torch.uint8 and out.is_contiguous()
    check(_lib.load().vs_dropout_mask(ptr(out), out.numel(), scheme, row_len, *_drop(dropout), stream()),
          "vs_dropout_mask")
    return out


def patchify(img, out, P):
    require_cuda(img, "patchify")
    assert img.dtype == F32 and img.is_contiguous()
    B, _, S, _ = img.shape
    _count(1)
    check(_lib.load().vs_patchify(ptr(img), ptr(out), B, S, P, stream()), "vs_patchify")


def cls_rows(cls, pos, x, B, T1, D):
    _count(1)
    check(_lib.load().vs_cls_rows(ptr(cls), ptr(pos), ptr(x), B, T1, D, stream()), "vs_cls_rows")


def embed_bwd(dx, dcls, dpos, dbias, B, T1, D):
    _count(1)
    check(_lib.load().vs_embed_bwd(ptr(dx), ptr(dcls), ptr(dpos), ptr(dbias), B, T1, D, stream()), "vs_embed_bwd")


def head_im2col(tokens, col, B, g, D):
    _count(1)
    check(_lib.load().vs_head_im2col(ptr(tokens), ptr(col), B, g, D, stream()), "vs_head_im2col")


def head_col2im(dcol, dtokens, B, g, D):
    _count(1)
    check(_lib.load().vs_head_col2im(ptr(dcol), ptr(dtokens), B, g, D, stream()), "vs_head_col2im")


def conv1x1_fwd(feat, w, b, logits, B, g, F, Cn):
    _count(1)
    check(_lib.load().vs_conv1x1_fwd(ptr(feat), ptr(w), ptr(b), ptr(logits), B, g, F, Cn, stream()), "vs_conv1x1_fwd")


def conv1x1_bwd(dlogits, feat, w, dfeat, dw, db, B, g, F, Cn):
    _count(1)
    check(_lib.load().vs_conv1x1_bwd(ptr(dlogits), ptr(feat), ptr(w), ptr(dfeat), ptr(dw), ptr(db), B, g, F, Cn,
                                    stream()), "vs_conv1x1_bwd")


def upsample_fwd(low, full):
    require_cuda(low, "upsample_fwd")
    B, Cn, g, _ = low.shape
    S = full.shape[-1]
    assert low.is_contiguous() and full.is_contiguous() and low.dtype == F32 and full.dtype == F32
    _count(1)
    check(_lib.load().vs_upsample_bilinear_fwd(ptr(low), ptr(full), B, Cn, g, S, stream()), "vs_upsample_bilinear_fwd")
    return full


def upsample_bwd(dfull, dlow):
    B, Cn, g, _ = dlow.shape
    S = dfull.shape[-1]
    assert dfull.is_contiguous() and dfull.dtype == F32
    _count(1)
    check(_lib.load().vs_upsample_bilinear_bwd(ptr(dfull), ptr(dlow), B, Cn, g, S, stream()), "vs_upsample_bilinear_bwd")
    return dlow


def upsample_argmax(low, mask):
    require_cuda(low, "upsample_argmax")
    B, Cn, g, _ = low.shape
    S = mask.shape[-1]
    assert mask.dtype == torch.uint8 and mask.is_contiguous() and low.is_contiguous()
    _count(1)
    check(_lib.load().vs_upsample_argmax(ptr(low), ptr(mask), B, Cn, g, S, stream()), "vs_upsample_argmax")
    return mask


def upsample_argmax_stats(low, labels, size, mask=None):
    """-> int32 counts [B, NC, 3] = {intersection, predicted, target} per image and class (NC = C, or 2 for C == 1);
    optionally also writes the uint8 class map."""
    require_cuda(low, "upsample_argmax_stats")
    B, Cn, g, _ = low.shape
    assert low.is_contiguous() and labels.dtype == torch.int64 and labels.is_contiguous()
    assert tuple(labels.shape) == (B, size, size), f"labels must be [B,{size},{size}], got {tuple(labels.shape)}"
    if mask is not None:
        assert mask.dtype == torch.uint8 and mask.is_contiguous() and tuple(mask.shape) == (B, size, size)
    counts = torch.empty(B, 2 if Cn == 1 else Cn, 3, device=low.device, dtype=torch.int32)
    _count(1)
    check(_lib.load().vs_upsample_argmax_stats(ptr(low), ptr(labels), ptr(mask), ptr(counts), B, Cn, g, size,
                                               stream()), "vs_upsample_argmax_stats")
    return counts


def colorize_mask(mask, palette):
    """uint8 class map [...], uint8 palette [C,3] -> uint8 RGB image [..., 3]."""
    require_cuda(mask, "colorize_mask")
    assert mask.dtype == torch.uint8 and mask.is_contiguous()
    assert palette.dtype == torch.uint8 and palette.is_contiguous() and palette.dim() == 2 and palette.shape[1] == 3
    rgb = torch.empty(*mask.shape, 3, device=mask.device, dtype=torch.uint8)
    _count(1)
    check(_lib.load().vs_colorize_mask(ptr(mask), ptr(palette), ptr(rgb), mask.numel(), palette.shape[0], stream()),
          "vs_colorize_mask")
    return rgb


def resample_h(src, bounds, coef, ksize, wout, dst):
    """uint8 [C,H,W] (last dim contiguous) -> uint8 [C,H,wout]: Pillow's horizontal resampling pass."""
    require_cuda(src, "resample_h")
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.stride(2) == 1
    Cn, H, W = src.shape
    _count(1)
    check(_lib.load().vs_resample_h_u8(ptr(src), src.stride(1), src.stride(0), Cn, H, W, ptr(bounds), ptr(coef), ksize,
                                       wout, ptr(dst), stream()), "vs_resample_h_u8")
    return dst


def resample_v(src, bounds, coef, ksize, hout, dst_f32=None, scale=255.0, dst_u8=None):
    """uint8 [C,H,W] contiguous -> [C,hout,W]: Pillow's vertical pass; fp32 (value / scale) and/or uint8 output."""
    require_cuda(src, "resample_v")
    assert src.dtype == torch.uint8 and src.is_contiguous()
    Cn, H, W = src.shape
    _count(1)
    check(_lib.load().vs_resample_v_u8(ptr(src), Cn, H, W, ptr(bounds), ptr(coef), ksize, hout, ptr(dst_f32), scale,
                                       ptr(dst_u8), stream()), "vs_resample_v_u8")


def u8_to_f32(src, dst, scale=255.0):
    require_cuda(src, "u8_to_f32")
    assert src.dtype == torch.uint8 and src.is_contiguous() and dst.dtype == F32 and dst.is_contiguous()
    _count(1)
    check(_lib.load().vs_u8_to_f32(ptr(src), ptr(dst), src.numel(), scale, stream()), "vs_u8_to_f32")


def multimem_allreduce(multicast_ptr: int, n: int, rank: int, world: int, scale: float):
    """in-switch all-reduce of n fp32 elements at a multicast (NVLS) address: this rank's 1/world slice (dp.py)."""
    _count(1)
    check(_lib.load().vs_multimem_allreduce_f32(multicast_ptr, n, rank, world, scale, stream()), "vs_multimem_allreduce_f32")


def sdf_targets(mask):
    """mask fp32 [B,S,S] (object > 0.5) -> (sdf_ext, sdf_int) fp32 [B,S,S]: compute_sdf of every image, on the device."""
    require_cuda(mask, "sdf_targets")
    assert mask.dtype == F32 and mask.dim() == 3 and mask.shape[1] == mask.shape[2] and mask.is_contiguous()
    B, S, _ = mask.shape
    lib = _lib.load()
    work = torch.empty(int(lib.vs_sdf_workspace_bytes(B, S)), device=mask.device, dtype=torch.uint8)
    ext, inn = torch.empty_like(mask), torch.empty_like(mask)
    _count(3)
    check(lib.vs_sdf_targets(ptr(mask), ptr(ext), ptr(inn), ptr(work), B, S, stream()), "vs_sdf_targets")
    return ext, inn


def upsample_ce(low, labels, loss_sum, dlow, size):
    """labels int64 or uint8 [B, LH, LW]; LH x LW != size x size = nearest-resized on the fly (CE/classes.py:273-274)."""
    require_cuda(low, "upsample_ce")
    B, Cn, g, _ = low.shape
    assert labels.dtype in (torch.int64, torch.uint8) and labels.is_contiguous() and low.is_contiguous()
    assert labels.dim() == 3 and labels.shape[0] == B
    _count(1)
    check(_lib.load().vs_upsample_ce(ptr(low), ptr(labels), 0 if labels.dtype == torch.int64 else 1, labels.shape[1],
                                     labels.shape[2], ptr(loss_sum), ptr(dlow), B, Cn, g, size, stream()),
          "vs_upsample_ce")


def paed_binary_stats(low, mask, sdf_ext, sdf_int, stats, keys):
    require_cuda(low, "paed_binary_stats")
    B, _, g, _ = low.shape
    S = mask.shape[-1]
    for t in (low, mask, sdf_ext, sdf_int):
        assert t.dtype == F32 and t.is_contiguous()
    _count(1)
    check(_lib.load().vs_paed_binary_stats(ptr(low), ptr(mask), ptr(sdf_ext), ptr(sdf_int), ptr(stats), ptr(keys), B, g,
                                          S, stream()), "vs_paed_binary_stats")


def paed_binary_bwd(low, mask, sdf_ext, sdf_int, coef, keys, dlow):
    B, _, g, _ = low.shape
    S = mask.shape[-1]
    assert coef.dtype == F32 and coef.is_contiguous()
    _count(1)
    check(_lib.load().vs_paed_binary_bwd(ptr(low), ptr(mask), ptr(sdf_ext), ptr(sdf_int), ptr(coef), ptr(keys),
                                        ptr(dlow), B, g, S, stream()), "vs_paed_binary_bwd")


def paed_multiclass(low, labels, t1, t2, t3, loss_sum, dlow):
    require_cuda(low, "paed_multiclass")
    B, Cn, g, _ = low.shape
    S = labels.shape[-1]
    assert labels.dtype == torch.int64 and labels.is_contiguous() and low.is_contiguous()
    _count(7 if dlow is not None else 4)
    check(_lib.load().vs_paed_multiclass(ptr(low), ptr(labels), ptr(t1), ptr(t2), ptr(t3), ptr(loss_sum), ptr(dlow), B,
                                        Cn, g, S, stream()), "vs_paed_multiclass")


def paed_multiclass_dense(msk, prob, t1, t2, t3, loss_sum, dprob, class_penalty=True):
    require_cuda(prob, "paed_multiclass_dense")
    B, Cn, S, S2 = prob.shape
    assert S == S2 and msk.shape == prob.shape and msk.dtype == F32 and prob.dtype == F32
    assert msk.is_contiguous() and prob.is_contiguous()
    _count(7 if dprob is not None else 4)
    check(_lib.load().vs_paed_multiclass_dense(ptr(msk), ptr(prob), ptr(t1), ptr(t2), ptr(t3), ptr(loss_sum),
                                              ptr(dprob), B, Cn, S, int(class_penalty), stream()),
          "vs_paed_multiclass_dense")


def cast_bf16(src, dst):
    require_cuda(src, "cast_bf16")
    assert src.dtype == F32 and dst.dtype == BF16 and src.is_contiguous() and dst.is_contiguous()
    _count(1)
    check(_lib.load().vs_cast_f32_bf16(ptr(src), ptr(dst), src.numel(), stream()), "vs_cast_f32_bf16")


def cast_bf16_rows(src, dst):
    M, D = src.shape
    _count(1)
    check(_lib.load().vs_cast_bf16_rows(ptr(src), _rowmajor(src, "src"), ptr(dst), _rowmajor(dst, "dst"), M, D,
                                       stream()), "vs_cast_bf16_rows")


def pack_conv3x3(w, out):
    O, I = w.shape[0], w.shape[1]
    assert w.is_contiguous()
    _count(1)
    check(_lib.load().vs_pack_conv3x3(ptr(w), ptr(out), O, I, stream()), "vs_pack_conv3x3")


def unpack_conv3x3_grad(g, dw):
    O, I = dw.shape[0], dw.shape[1]
    _count(1)
    check(_lib.load().vs_unpack_conv3x3_grad(ptr(g), ptr(dw), O, I, stream()), "vs_unpack_conv3x3_grad")


def adam_step(param, grad, exp_avg, exp_avg_sq, shadow, lr_dev, step_dev, beta1, beta2, eps, weight_decay, decoupled,
              grad_scale, zero_grad, skip_begin, skip_end):
    require_cuda(param, "adam_step")
    _count(1)
    check(_lib.load().vs_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(shadow), param.numel(),
                                  ptr(lr_dev), ptr(step_dev), beta1, beta2, eps, weight_decay, int(decoupled),
                                  grad_scale, int(zero_grad), skip_begin, skip_end, stream()), "vs_adam_step")
