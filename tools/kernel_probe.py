"""Diagnostic probe: runs every libvitseg kernel against a plain PyTorch fp32 computation on the GPU and prints
error statistics without stopping at the first failure.  Development aid for gpurun sessions; the pass/fail
versions of these checks live in tests/test_kernels_gpu.py.

usage: python tools/kernel_probe.py [gemm] [ln] [attn] [head] [loss] ..."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
RESULTS = []


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def report(name, err, tol):
    ok = err <= tol and err == err
    RESULTS.append((name, err, tol, ok))
    print(f"{'PASS' if ok else 'FAIL'} {name:58s} err={err:.3e} tol={tol:.1e}", flush=True)


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        RESULTS.append((name, float("nan"), 0, False))
        print(f"FAIL {name}: EXCEPTION {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def bf(x):
    return x.to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------
def probe_gemm():
    def case(M, N, Kd, a_mn, b_mn, **kw):
        def f():
            a = bf(torch.randn(M, Kd, device=dev))
            b = bf(torch.randn(N, Kd, device=dev) * 0.5)
            ref = a.float() @ b.float().t()
            A = a.t().contiguous() if a_mn else a
            Bm = b.t().contiguous() if b_mn else b
            out_f32 = kw.get("f32", False)
            out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
            bias = torch.randn(N, device=dev) if kw.get("bias") else None
            res = torch.randn(M, N, device=dev) if kw.get("res") else None
            act = kw.get("act", 0)
            if bias is not None:
                ref = ref + bias
            pre = ref.clone()
            if act == 1:
                ref = F.gelu(ref)
            elif act == 2:
                ref = F.relu(ref)
            aux = None
            aux_mode = kw.get("aux_mode", 0)
            if aux_mode:
                aux = bf(torch.randn(M, N, device=dev))
                if aux_mode == 1:
                    x = aux.float().requires_grad_(True)
                    (gg,) = torch.autograd.grad(F.gelu(x).sum(), x)
                    ref = ref * gg
                else:
                    ref = ref * (aux.float() > 0)
            if res is not None:
                ref = ref + res
            out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if kw.get("out2") else None
            accumulate = kw.get("acc", False)
            if accumulate:
                base = torch.randn(M, N, device=dev)
                out.copy_(base)
                ref = ref + base
            K.gemm(A, Bm, out, a_mn=a_mn, b_mn=b_mn, bias=bias, act=act, out2=out2, aux=aux, aux_mode=aux_mode,
                   residual=res, accumulate=accumulate, split_k=kw.get("split_k", 0), tile_cfg=kw.get("cfg", 0))
            torch.cuda.synchronize()
            report(f"gemm M{M} N{N} K{Kd} a_mn{int(a_mn)} b_mn{int(b_mn)} {kw}", rel(out, ref), 1e-2)
            if out2 is not None:
                report("   out2 (pre-activation)", rel(out2, pre), 1e-2)
        run(f"gemm {M}x{N}x{Kd} {a_mn} {b_mn} {kw}", f)

    case(128, 256, 64, False, False, f32=True)
    case(128, 256, 256, False, False, f32=True)
    case(256, 512, 768, False, False)
    case(1000, 768, 768, False, False, bias=True, act=1, out2=True)
    case(1000, 768, 3072, False, False, bias=True, res=True, f32=True)
    case(128, 128, 128, False, False, f32=True)
    case(300, 64, 256, False, False, f32=True)
    case(128, 256, 64, False, True, f32=True)
    case(1000, 768, 3072, False, True, aux_mode=1)
    case(1000, 256, 6912, False, True, aux_mode=2)
    case(128, 256, 64, True, False, f32=True)
    case(128, 256, 64, True, True, f32=True)
    # patch-4 embedding shapes: K = N = 3 * 4 * 4 = 48 (partial k-block, 16-column tail)
    case(6274, 128, 48, False, False, bias=True, res=True, f32=True)
    case(128, 48, 6274, True, True, f32=True, acc=True)
    case(300, 48, 128, False, True)
    case(768, 3072, 1000, True, True, f32=True, acc=True)
    case(768, 768, 12608, True, True, f32=True, acc=True)
    case(2304, 768, 4000, True, True, f32=True, acc=True, split_k=3)
    case(12608, 768, 768, False, False, bias=True)
    case(12608, 3072, 768, False, False, bias=True, act=1)
    # every tile configuration x operand-major combination (cfg 1-3: cta_group::2 pairs; 4-5: single CTA)
    for cfg in (1, 2, 3, 4, 5):
        for a_mn in (False, True):
            for b_mn in (False, True):
                case(1000, 768, 832, a_mn, b_mn, f32=True, cfg=cfg)
        case(12608, 768, 768, False, False, bias=True, res=True, f32=True, cfg=cfg)
        case(768, 3072, 4000, True, True, f32=True, acc=True, split_k=2, cfg=cfg)

    # bf16 outputs (shared-memory staged TMA-store epilogue on the 256- / 128-column tiles, register-direct on 192):
    # every tile configuration x {plain, bias + GELU + pre-activation copy, GELU' aux, ReLU-mask aux}, ragged M, a
    # 16-column tail, several tiles per CTA (aux prefetch / buffer reuse across tiles)
    for cfg in (1, 2, 3, 4, 5):
        case(1000, 768, 832, False, False, cfg=cfg)
        case(5000, 1024, 256, False, False, bias=True, act=1, out2=True, cfg=cfg)
        case(5000, 1024, 256, False, True, aux_mode=1, cfg=cfg)
        case(1000, 256, 512, False, True, aux_mode=2, bias=True, cfg=cfg)
        case(300, 48, 128, False, True, cfg=cfg)
        case(40000, 512, 64, False, False, bias=True, act=1, out2=True, cfg=cfg)

    # partial last wave cut into column sub-tiles (VS_GEMM_TAIL): the headline data-gradient / forward shapes, whose tile
    # counts leave 2 - 8 tiles for the last wave, in every tile configuration
    for cfg in (1, 2, 3, 4, 5):
        case(12608, 768, 3072, False, True, cfg=cfg)
        case(12608, 768, 768, False, True, cfg=cfg)
        case(12608, 2304, 768, False, False, bias=True, cfg=cfg)
        case(12608, 3072, 768, False, True, aux_mode=1, cfg=cfg)
        case(12608, 3072, 768, False, False, bias=True, act=1, out2=True, cfg=cfg)
        case(12608, 768, 3072, False, False, bias=True, res=True, f32=True, cfg=cfg)

    def colsum_case():
        # out_colsum: fused into the TMA-store epilogue (cfg 1, 3, 4, 5), a pass after the GEMM otherwise (cfg 2)
        for cfg in (0, 1, 2, 3, 4, 5):
            M, N, Kd = 5000, 1024, 256
            a = bf(torch.randn(M, Kd, device=dev))
            b = bf(torch.randn(N, Kd, device=dev) * 0.5)
            aux = bf(torch.randn(M, N, device=dev))
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            cs = torch.randn(N, device=dev)
            cs0 = cs.clone()
            K.gemm(a, b.t().contiguous(), out, b_mn=True, aux=aux, aux_mode=1, colsum=cs, tile_cfg=cfg)
            report(f"gemm fused column sums cfg{cfg}", rel(cs, cs0 + out.float().sum(0)), 2e-5)
            x = aux.float().requires_grad_(True)
            (gg,) = torch.autograd.grad(F.gelu(x).sum(), x)
            report(f"   output (column-persistent tile order) cfg{cfg}", rel(out, (a.float() @ b.float().t()) * gg), 1e-2)
        # ragged shapes: 13 row tiles on 6 units per column tile, a 16-column tail, fewer row tiles than units per column
        for (M, N, Kd) in ((3200, 3072, 128), (300, 48, 128), (12608, 3072, 768)):
            a = bf(torch.randn(M, Kd, device=dev))
            b = bf(torch.randn(N, Kd, device=dev) * 0.5)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            cs = torch.zeros(N, device=dev)
            K.gemm(a, b, out, colsum=cs)
            report(f"gemm fused column sums {M}x{N}x{Kd}", rel(cs, out.float().sum(0)), 2e-5)
            report(f"   output {M}x{N}x{Kd}", rel(out, a.float() @ b.float().t()), 1e-2)
    run("gemm colsum", colsum_case)

    def strided_case():
        # bf16 output / aux that are column slices of wider matrices (row stride != N), as the engine's views can be
        M, N, Kd = 777, 256, 192
        a = bf(torch.randn(M, Kd, device=dev))
        b = bf(torch.randn(N, Kd, device=dev) * 0.5)
        wide = torch.zeros(M, 3 * N, device=dev, dtype=torch.bfloat16)
        auxw = bf(torch.randn(M, 2 * N, device=dev))
        K.gemm(a, b, wide[:, N:2 * N], aux=auxw[:, N:], aux_mode=2)
        ref = (a.float() @ b.float().t()) * (auxw[:, N:].float() > 0)
        report("gemm bf16 strided out / aux views", rel(wide[:, N:2 * N], ref), 1e-2)
        report("   neighbours untouched", float(wide[:, :N].abs().max() + wide[:, 2 * N:].abs().max()), 1e-30)
    run("gemm strided", strided_case)

    def patch_case():
        Bn, T1, D, Kd = 3, 197, 768, 768
        a = bf(torch.randn(Bn * T1, Kd, device=dev))
        w = bf(torch.randn(D, Kd, device=dev) * 0.1)
        bias = torch.randn(D, device=dev)
        pos = torch.randn(T1, D, device=dev)
        x = torch.zeros(Bn * T1, D, device=dev)
        K.gemm(a, w, x, bias=bias, residual=pos, row_tokens=T1)
        ref = (a.float() @ w.float().t() + bias).view(Bn, T1, D) + pos
        report("gemm patch-embed broadcast residual (row % T1)", rel(x.view(Bn, T1, D), ref), 1e-2)
    run("gemm patch", patch_case)


def probe_colsum():
    def f():
        for M, N, nf in ((12608, 2304, 768), (197, 384, 128), (1000, 3072, 1024)):
            x = bf(torch.randn(M, N, device=dev))
            xf = torch.randn(M, nf, device=dev)
            want_x = x.clone()
            want_x[:, :nf] = bf(xf)
            out = torch.randn(N, device=dev)
            want = out + want_x.float().sum(0)
            K.colsum_cast(x, xf, out, accumulate=True)
            report(f"colsum_cast {M}x{N} (first {nf} fp32): stored bf16 copy", float((x.float() - want_x.float()).abs().max()), 0.0)
            report(f"colsum_cast {M}x{N}: sums", rel(out, want), 1e-5)
            out2 = torch.empty(N, device=dev)
            K.colsum(x, out2, accumulate=False)
            report(f"colsum {M}x{N}", rel(out2, want_x.float().sum(0)), 1e-5)
    run("colsum", f)


def probe_ln():
    probe_colsum()
    def f():
        # M = 12605 / 5003: more row groups than SMs x stages (ring reuse) and a ragged last group
        for D, M in ((768, 1000), (1024, 1000), (512, 1000), (768, 12605), (128, 5003), (1024, 4099)):
            x = torch.randn(M, D, device=dev) * 2 + 0.5
            g = torch.randn(D, device=dev)
            b = torch.randn(D, device=dev)
            y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
            yf = torch.empty(M, D, device=dev)
            mean = torch.empty(M, device=dev)
            rstd = torch.empty(M, device=dev)
            K.layernorm_fwd(x, g, b, 1e-12, y, yf, mean, rstd)
            ref = F.layer_norm(x, (D,), g, b, 1e-12)
            report(f"ln fwd f32 D{D} M{M}", rel(yf, ref), 1e-5)
            report(f"ln fwd bf16 D{D}", rel(y, ref), 1e-2)
            xr = x.clone().requires_grad_(True)
            gr = g.clone().requires_grad_(True)
            br = b.clone().requires_grad_(True)
            dy = torch.randn(M, D, device=dev)
            F.layer_norm(xr, (D,), gr, br, 1e-12).backward(dy)
            dxin = torch.randn(M, D, device=dev)
            dx = torch.empty(M, D, device=dev)
            dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
            dg = torch.zeros(D, device=dev)
            db = torch.zeros(D, device=dev)
            dbias = torch.zeros(D, device=dev)
            K.layernorm_bwd(dy, x, g, mean, rstd, dxin, dx, dxb, dg, db, dbias=dbias)
            report(f"ln bwd dx D{D} M{M}", rel(dx, xr.grad + dxin), 1e-4)
            report(f"ln bwd bf16 copy D{D}", rel(dxb, xr.grad + dxin), 1e-2)
            report(f"ln bwd fused colsum D{D} M{M}", rel(dbias, (xr.grad + dxin).sum(0)), 1e-4)
            report(f"ln bwd dgamma D{D}", rel(dg, gr.grad), 1e-4)
            report(f"ln bwd dbeta D{D}", rel(db, br.grad), 1e-4)
            dyb = bf(dy)
            dg.zero_(); db.zero_()
            K.layernorm_bwd(dyb, x, g, mean, rstd, None, dx, None, dg, db)
            xr.grad = None
            F.layer_norm(xr, (D,), g, b, 1e-12).backward(dyb.float())
            report(f"ln bwd dx (bf16 dy) D{D}", rel(dx, xr.grad), 1e-4)
    run("ln", f)


def probe_attn():
    def case(Bn, N, H):
        def f():
            D = H * 64
            qkv = bf(torch.randn(Bn, N, 3, H, 64, device=dev))
            ctx = torch.empty(Bn, N, H, 64, device=dev, dtype=torch.bfloat16)
            lse = torch.empty(Bn, H, N, device=dev)
            K.attention_fwd(qkv, ctx, lse, Bn, N, H, 0.125)
            q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).float().requires_grad_(True) for i in range(3)]
            s = (q @ k.transpose(-1, -2)) * 0.125
            ref = torch.softmax(s, -1) @ v
            report(f"attn fwd B{Bn} N{N} H{H}", rel(ctx.permute(0, 2, 1, 3), ref), 1e-2)
            report(f"attn lse B{Bn} N{N} H{H}", rel(lse, torch.logsumexp(s, -1)), 1e-3)
            dctx = bf(torch.randn(Bn, N, H, 64, device=dev))
            ref.backward(dctx.permute(0, 2, 1, 3).float())
            dqkv = torch.zeros(Bn, N, 3, H, 64, device=dev, dtype=torch.bfloat16)
            dq_acc = torch.empty(Bn, N, D, device=dev)
            delta = torch.empty(Bn, H, N, device=dev)
            K.attention_bwd(qkv, ctx, dctx, lse, dqkv, dq_acc, delta, Bn, N, H, 0.125)
            report(f"attn bwd dK B{Bn} N{N} H{H}", rel(dqkv[:, :, 1].permute(0, 2, 1, 3), k.grad), 2e-2)
            report(f"attn bwd dV B{Bn} N{N} H{H}", rel(dqkv[:, :, 2].permute(0, 2, 1, 3), v.grad), 2e-2)
            report(f"attn bwd dQ B{Bn} N{N} H{H}", rel(dq_acc.view(Bn, N, H, 64).permute(0, 2, 1, 3), q.grad), 2e-2)
            if N <= 256:
                # short-sequence kernel: whole (batch, head) items per CTA, all three gradients written as bf16
                dqkv2 = torch.full_like(dqkv, float("nan"))
                K.attention_bwd(qkv, ctx, dctx, lse, dqkv2, None, delta, Bn, N, H, 0.125)
                for i, (nm, ref_g) in enumerate((("dQ", q.grad), ("dK", k.grad), ("dV", v.grad))):
                    report(f"attn bwd (short) {nm} B{Bn} N{N} H{H}", rel(dqkv2[:, :, i].permute(0, 2, 1, 3), ref_g), 2e-2)
        run(f"attn {Bn} {N} {H}", f)

    case(1, 128, 1)
    case(2, 197, 12)
    case(40, 197, 12)    # 480 items on 148 SMs: several items per CTA (buffer hand-over between items), ragged last round
    case(7, 130, 3)
    case(5, 100, 2)      # one query tile
    case(1, 64, 2)
    case(2, 256, 2)      # largest single-pass forward
    case(1, 17, 1)
    case(3, 250, 3)
    case(1, 257, 1)      # first streaming-forward length
    case(2, 577, 4)
    case(1, 1025, 2)


def probe_head():
    def f():
        Bn, g, D, Fd, Cn = 2, 14, 768, 256, 17
        T = g * g
        tok = bf(torch.randn(Bn, T + 1, D, device=dev))
        col = torch.empty(Bn * T, 9 * D, device=dev, dtype=torch.bfloat16)
        K.head_im2col(tok, col, Bn, g, D)
        feat = tok[:, 1:].float().transpose(1, 2).reshape(Bn, D, g, g)
        w = torch.randn(Fd, D, 3, 3, device=dev) * 0.05
        wp = torch.empty(Fd, 9 * D, device=dev, dtype=torch.bfloat16)
        K.pack_conv3x3(w, wp)
        ref = F.conv2d(feat, bf(w).float(), padding=1)  # [B,F,g,g]
        got = (col.float() @ wp.float().t()).view(Bn, g, g, Fd).permute(0, 3, 1, 2)
        report("head im2col+pack == conv3x3", rel(got, ref), 1e-3)
        # col2im adjoint
        dcol = bf(torch.randn(Bn * T, 9 * D, device=dev))
        dtok = torch.empty(Bn, T + 1, D, device=dev)
        K.head_col2im(dcol, dtok, Bn, g, D)
        fr = feat.clone().requires_grad_(True)
        colr = F.unfold(fr, 3, padding=1)  # [B, D*9, T] with (c,ky,kx) order
        colr = colr.view(Bn, D, 9, T).permute(0, 3, 2, 1).reshape(Bn * T, 9 * D)
        colr.backward(dcol.float())
        refd = fr.grad.reshape(Bn, D, T).transpose(1, 2)
        report("head col2im adjoint", rel(dtok[:, 1:], refd), 1e-2)
        report("head col2im cls row zero", dtok[:, 0].abs().max().item(), 0.0)
        # conv1x1
        ft = bf(torch.relu(torch.randn(Bn * T, Fd, device=dev)))
        w2 = torch.randn(Cn, Fd, device=dev) * 0.1
        b2 = torch.randn(Cn, device=dev)
        lg = torch.empty(Bn, Cn, g, g, device=dev)
        K.conv1x1_fwd(ft, w2, b2, lg, Bn, g, Fd, Cn)
        ref1 = (ft.float() @ w2.t() + b2).view(Bn, T, Cn).transpose(1, 2).reshape(Bn, Cn, g, g)
        report("conv1x1 fwd", rel(lg, ref1), 1e-5)
        dl = torch.randn(Bn, Cn, g, g, device=dev)
        dfeat = torch.empty(Bn * T, Fd, device=dev, dtype=torch.bfloat16)
        dw = torch.zeros(Cn, Fd, device=dev)
        db = torch.zeros(Cn, device=dev)
        K.conv1x1_bwd(dl, ft, w2, dfeat, dw, db, Bn, g, Fd, Cn)
        dl2 = dl.view(Bn, Cn, T).transpose(1, 2).reshape(Bn * T, Cn)
        report("conv1x1 bwd dfeat", rel(dfeat, (dl2 @ w2) * (ft.float() > 0)), 1e-2)
        report("conv1x1 bwd dw", rel(dw, dl2.t() @ ft.float()), 1e-4)
        report("conv1x1 bwd db", rel(db, dl2.sum(0)), 1e-4)
        # other classifier widths (partial / multiple 256-feature chunks) and the binary head
        for F2, C2 in ((128, 1), (512, 4), (264, 17)):
            ft2 = bf(torch.relu(torch.randn(Bn * T, F2, device=dev)))
            w3 = torch.randn(C2, F2, device=dev) * 0.1
            b3 = torch.randn(C2, device=dev)
            lg2 = torch.empty(Bn, C2, g, g, device=dev)
            K.conv1x1_fwd(ft2, w3, b3, lg2, Bn, g, F2, C2)
            ref2 = (ft2.float() @ w3.t() + b3).view(Bn, T, C2).transpose(1, 2).reshape(Bn, C2, g, g)
            report(f"conv1x1 fwd F{F2} C{C2}", rel(lg2, ref2), 1e-5)
        # patchify
        img = torch.rand(Bn, 3, 224, 224, device=dev)
        pm = torch.zeros(Bn * (T + 1), 768, device=dev, dtype=torch.bfloat16)
        K.patchify(img, pm, 16)
        refp = F.unfold(img, 16, stride=16).transpose(1, 2)
        report("patchify", rel(pm.view(Bn, T + 1, 768)[:, 1:], refp), 4e-3)
        report("patchify cls rows untouched", pm.view(Bn, T + 1, 768)[:, 0].abs().max().item(), 0.0)
        # colsum / embed_bwd / casts
        x = bf(torch.randn(1000, 768, device=dev))
        cs = torch.zeros(768, device=dev)
        K.colsum(x, cs, accumulate=False)
        report("colsum", rel(cs, x.float().sum(0)), 1e-4)
        dx = torch.randn(Bn, T + 1, D, device=dev)
        dcls = torch.zeros(D, device=dev)
        dpos = torch.zeros(T + 1, D, device=dev)
        dbp = torch.zeros(D, device=dev)
        K.embed_bwd(dx, dcls, dpos, dbp, Bn, T + 1, D)
        report("embed_bwd dpos", rel(dpos, dx.sum(0)), 1e-5)
        report("embed_bwd dcls", rel(dcls, dx[:, 0].sum(0)), 1e-5)
        report("embed_bwd dbias (patch rows only)", rel(dbp, dx[:, 1:].sum((0, 1))), 1e-5)
        # rewritten glue kernels at other shapes: batch above / not a multiple of the 8-way unrolled loop of embed_bwd,
        # patch 8 / patch 4 / 384-pixel patchify, im2col with 128 / 1024 channels and a 24 x 24 grid
        for Bq, Tq, Dq in ((19, 50, 256), (8, 197, 128)):
            dxq = torch.randn(Bq, Tq, Dq, device=dev)
            dc, dp, dbq = torch.zeros(Dq, device=dev), torch.zeros(Tq, Dq, device=dev), torch.zeros(Dq, device=dev)
            K.embed_bwd(dxq, dc, dp, dbq, Bq, Tq, Dq)
            report(f"embed_bwd dpos B{Bq}", rel(dp, dxq.sum(0)), 1e-5)
            report(f"embed_bwd dcls B{Bq}", rel(dc, dxq[:, 0].sum(0)), 1e-5)
            report(f"embed_bwd dbias B{Bq}", rel(dbq, dxq[:, 1:].sum((0, 1))), 1e-5)
        for Sq, Pq in ((224, 8), (224, 4), (384, 16), (64, 32)):
            gq = Sq // Pq
            imq = torch.rand(3, 3, Sq, Sq, device=dev)
            pq = torch.zeros(3 * (gq * gq + 1), 3 * Pq * Pq, device=dev, dtype=torch.bfloat16)
            K.patchify(imq, pq, Pq)
            refq = F.unfold(imq, Pq, stride=Pq).transpose(1, 2)
            report(f"patchify S{Sq} P{Pq}", rel(pq.view(3, gq * gq + 1, -1)[:, 1:], refq), 4e-3)
            report(f"patchify S{Sq} P{Pq} cls rows untouched", pq.view(3, gq * gq + 1, -1)[:, 0].abs().max().item(), 0.0)
        for Bq, gq, Dq in ((3, 24, 128), (1, 7, 1024), (5, 14, 768)):
            Tq = gq * gq
            tq = bf(torch.randn(Bq, Tq + 1, Dq, device=dev))
            cq = torch.full((Bq * Tq, 9 * Dq), 7.0, device=dev, dtype=torch.bfloat16)
            K.head_im2col(tq, cq, Bq, gq, Dq)
            fq = tq[:, 1:].float().transpose(1, 2).reshape(Bq, Dq, gq, gq)
            rq = F.unfold(fq, 3, padding=1).view(Bq, Dq, 9, Tq).permute(0, 3, 2, 1).reshape(Bq * Tq, 9 * Dq)
            report(f"head im2col B{Bq} g{gq} D{Dq} (exact)", (cq.float() - rq).abs().max().item(), 0.0)
        # conv1x1 backward over enough pixels for two blocks per SM
        Bq = 48
        ftq = bf(torch.relu(torch.randn(Bq * T, Fd, device=dev)))
        dlq = torch.randn(Bq, Cn, g, g, device=dev)
        dfq = torch.empty(Bq * T, Fd, device=dev, dtype=torch.bfloat16)
        dwq, dbq2 = torch.zeros(Cn, Fd, device=dev), torch.zeros(Cn, device=dev)
        K.conv1x1_bwd(dlq, ftq, w2, dfq, dwq, dbq2, Bq, g, Fd, Cn)
        dlq2 = dlq.view(Bq, Cn, T).transpose(1, 2).reshape(Bq * T, Cn)
        report("conv1x1 bwd dfeat B48", rel(dfq, (dlq2 @ w2) * (ftq.float() > 0)), 1e-2)
        report("conv1x1 bwd dw B48", rel(dwq, dlq2.t() @ ftq.float()), 1e-4)
        report("conv1x1 bwd db B48", rel(dbq2, dlq2.sum(0)), 1e-4)
        gpk = torch.randn(Fd, 9 * D, device=dev)
        dwc = torch.zeros(Fd, D, 3, 3, device=dev)
        K.unpack_conv3x3_grad(gpk, dwc)
        report("unpack_conv3x3_grad", rel(dwc, gpk.view(Fd, 3, 3, D).permute(0, 3, 1, 2)), 0.0)
    run("head", f)


def probe_loss():
    def f():
        # (5, 14: class counts below the kernels' template sizes = the runtime-predicate instantiations; g = 56: patch 4;
        #  S = 1280: more 4-column groups per row than threads per block in the upsample forward)
        for (Bn, Cn, g, S) in ((2, 17, 14, 224), (1, 17, 32, 512), (2, 1, 28, 224), (2, 5, 14, 224), (1, 14, 24, 384),
                               (1, 8, 56, 224), (1, 3, 80, 1280)):
            low = torch.randn(Bn, Cn, g, g, device=dev)
            full = torch.empty(Bn, Cn, S, S, device=dev)
            K.upsample_fwd(low, full)
            ref = F.interpolate(low, size=(S, S), mode="bilinear", align_corners=False)
            report(f"upsample fwd {Bn},{Cn},{g},{S}", (full - ref).abs().max().item(), 1e-5)
            dfull = torch.randn(Bn, Cn, S, S, device=dev)
            dlow = torch.empty_like(low)
            K.upsample_bwd(dfull, dlow)
            lr = low.clone().requires_grad_(True)
            F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False).backward(dfull)
            report(f"upsample bwd {Bn},{Cn},{g},{S}", rel(dlow, lr.grad), 1e-4)
            mask = torch.empty(Bn, S, S, device=dev, dtype=torch.uint8)
            K.upsample_argmax(low, mask)
            refm = (ref[:, 0] > 0) if Cn == 1 else ref.argmax(1)
            report(f"upsample argmax mismatch frac {Bn},{Cn},{g},{S}", (mask.long() != refm.long()).float().mean().item(), 1e-5)
            if Cn > 1:
                labels = torch.randint(0, Cn, (Bn, S, S), device=dev)
                labels[0, :3, :5] = -100
                ls = torch.zeros(2, device=dev)
                dl = torch.zeros_like(low)
                K.upsample_ce(low, labels, ls, dl, S)
                lr.grad = None
                lossr = F.cross_entropy(F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False), labels,
                                        reduction="sum")
                lossr.backward()
                report(f"upsample_ce loss {Bn},{Cn},{g},{S}", abs(ls[0].item() - lossr.item()) / lossr.item(), 1e-5)
                report(f"upsample_ce count", abs(ls[1].item() - (labels != -100).sum().item()), 0.0)
                report(f"upsample_ce dlow {Bn},{Cn},{g},{S}", rel(dl, lr.grad), 1e-4)
                # labels at their stored resolution: the legacy-'nearest' resize (model/CE/classes.py:273-274) inside
                # the kernel's label read, int64 and uint8
                for LS, dt in ((256, torch.int64), (256, torch.uint8), (100, torch.int64)):
                    lab = torch.randint(0, Cn, (Bn, LS, LS), device=dev)
                    want = F.interpolate(lab[:, None].float(), size=(S, S), mode="nearest")[:, 0].long()
                    ls2, dl2 = torch.zeros(2, device=dev), torch.zeros_like(low)
                    K.upsample_ce(low, lab.to(dt), ls2, dl2, S)
                    lr.grad = None
                    l2 = F.cross_entropy(F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False), want,
                                         reduction="sum")
                    l2.backward()
                    report(f"upsample_ce labels {LS}->{S} {dt} loss", abs(ls2[0].item() - l2.item()) / l2.item(), 1e-5)
                    report(f"upsample_ce labels {LS}->{S} {dt} dlow", rel(dl2, lr.grad), 1e-4)
    run("loss/upsample", f)

    def paed_bin():
        Bn, g, S = 3, 14, 224
        low = torch.randn(Bn, 1, g, g, device=dev) * 2
        mask = (torch.rand(Bn, S, S, device=dev) > 0.7).float()
        sdf_e = torch.rand(Bn, S, S, device=dev)
        sdf_i = torch.rand(Bn, S, S, device=dev)
        stats = torch.zeros(Bn, 8, device=dev)
        keys = torch.zeros(Bn, dtype=torch.int64, device=dev)
        K.paed_binary_stats(low, mask, sdf_e, sdf_i, stats, keys)
        lr = low.clone().requires_grad_(True)
        z = F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False)
        p = torch.sigmoid(z)
        t = mask[:, None]
        sx = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], device=dev, dtype=torch.float32).view(1, 1, 3, 3)
        gx = F.conv2d(p, sx, padding=1)
        gy = F.conv2d(p, sx.transpose(2, 3), padding=1)
        edge = torch.sqrt(gx ** 2 + gy ** 2 + 1e-6)
        mx = edge.view(Bn, -1).max(1)[0]
        ref_stats = torch.stack([
            F.binary_cross_entropy(p, t, reduction="none").view(Bn, -1).sum(1),
            (p * t).view(Bn, -1).sum(1), p.view(Bn, -1).sum(1), t.view(Bn, -1).sum(1),
            (sdf_i[:, None] * p).view(Bn, -1).sum(1), (sdf_e[:, None] * edge).view(Bn, -1).sum(1)], 1)
        report("paed_binary stats", rel(stats[:, :6], ref_stats), 1e-4)
        got_max = (keys >> 32).to(torch.int32).view(torch.float32)
        report("paed_binary max edge", rel(got_max, mx), 1e-6)
        coef = torch.randn(Bn, 8, device=dev)
        coef[:, 3] = 0
        coef[:, 7] = 0
        L = (ref_stats * coef[:, :6]).sum() + (mx * coef[:, 6]).sum()
        L.backward()
        dlow = torch.zeros_like(low)
        K.paed_binary_bwd(low, mask, sdf_e, sdf_i, coef, keys, dlow)
        report("paed_binary bwd dlow", rel(dlow, lr.grad), 1e-3)
    run("paed binary", paed_bin)

    def paed_multi():
        Bn, Cn, g, S = 2, 17, 14, 224
        low = torch.randn(Bn, Cn, g, g, device=dev)
        labels = torch.randint(0, Cn, (Bn, S, S), device=dev)
        t1, t2, t3 = [torch.empty(Bn, Cn, S, S, device=dev) for _ in range(3)]
        ls = torch.zeros(1, device=dev)
        dlow = torch.zeros_like(low)
        K.paed_multiclass(low, labels, t1, t2, t3, ls, dlow)
        lr = low.clone().requires_grad_(True)
        p = torch.softmax(F.interpolate(lr, size=(S, S), mode="bilinear", align_corners=False), 1)
        m = F.one_hot(labels, Cn).permute(0, 3, 1, 2).float()
        x = torch.arange(19, device=dev).float() - 9
        gk = torch.exp(-(x ** 2) / 18)
        k2 = gk[:, None] * gk[None, :]
        k2 = (k2 / k2.sum())[None, None].repeat(Cn, 1, 1, 1)
        base = (F.conv2d(m, k2, padding=9, groups=Cn) - F.conv2d(p, k2, padding=9, groups=Cn)).abs()
        loss = (m * (1 - p) * base * 2).sum()
        loss.backward()
        report("paed_multiclass loss", abs(ls.item() - loss.item()) / abs(loss.item()), 1e-4)
        report("paed_multiclass dlow", rel(dlow, lr.grad), 1e-3)
    run("paed multiclass", paed_multi)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ln", "gemm", "attn", "head", "loss"]
    print("device:", torch.cuda.get_device_name(0), "SMs:", torch.cuda.get_device_properties(0).multi_processor_count)
    for w in which:
        {"gemm": probe_gemm, "ln": probe_ln, "attn": probe_attn, "head": probe_head, "loss": probe_loss}[w]()
    nfail = sum(1 for r in RESULTS if not r[3])
    print(f"SUMMARY: {len(RESULTS) - nfail} pass, {nfail} fail")
