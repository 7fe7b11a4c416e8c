"""Timeline of one CTA of the short-sequence attention backward (library built with -DVS_ATTN_TRACE, loaded through
VS_LIB_PATH): prints, per global step, when each role reached its hand-over points (SM clock cycles relative to the
first event).  Used to find which hand-over bounds the step period.

  VS_LIB_PATH=.../libvitseg_trace.so python tools/attn_trace.py [dropout_p]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, N, H = 64, 197, 12
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
D = H * 64
qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
ctx = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
dctx = torch.randn(B, N, H, 64, device=dev).bfloat16()
dqkv = torch.zeros(B, N, 3, H, 64, device=dev, dtype=torch.bfloat16)
delta = torch.empty(B, H, N, device=dev)
seed = torch.tensor([1], device=dev, dtype=torch.int32)
drop = (p, seed, 1000) if p > 0 else None
K.attention_fwd(qkv, ctx, lse, B, N, H, 0.125, dropout=drop)
lib = _lib.load()
lib.vs_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
lib.vs_debug_attn_trace.restype = C.c_int
buf = (C.c_ulonglong * 49152)()
for _ in range(3):
    K.attention_bwd(qkv, ctx, dctx, lse, dqkv, None, delta, B, N, H, 0.125, dropout=drop)
    n = lib.vs_debug_attn_trace(buf, 49152)
ev = sorted(((buf[i] & 0xFFFFFFFFFF, buf[i] >> 56, (buf[i] >> 40) & 0xFFFF) for i in range(n) if buf[i] != 0))
n = len(ev)
t0 = ev[0][0]
names = {1: "grad:pfull", 2: "mma:ahead", 3: "grad:dq+kvfree", 4: "grad:issued", 5: "drain:kvfull", 6: "drain:kvfree",
         12: "sdp:begin", 13: "sdp:ready", 14: "sdp:issued", 15: "sdp:commit", 7: "cmp:wait_s", 8: "cmp:got_s", 9: "cmp:chunk0", 10: "cmp:got_buf", 11: "cmp:pfull"}
print(f"{n} events")
for t, e, s in ev:
    print(f"{t - t0:8d} {names.get(e, e):14s} step {s}")
