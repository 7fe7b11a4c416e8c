"""Per-step timing breakdown on one GPU: host enqueue time vs device time for the bench workload, plus a per-kernel
CUDA-event table for the GEMM family (shape -> TFLOP/s)."""
import argparse
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200 import kernels as K  # noqa: E402
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
DROP = args.dropout
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = LightningViTModel(17, 16, 768, 12, 12, hidden_dropout_prob=DROP, attention_probs_dropout_prob=DROP).to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-5, fused=True)
x = torch.rand(args.batch, 3, 224, 224, device=dev)
y = torch.randint(0, 17, (args.batch, 256, 256), device=dev)


def step(i):
    loss = m.training_step((x, y), i)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(3):
    step(i)
torch.cuda.synchronize()
# host enqueue time (GPU queue is deep enough that the host never blocks within a few steps)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    step(i)
t_host = (time.perf_counter() - t0) / args.steps
e1.record()
torch.cuda.synchronize()
print(f"host enqueue {t_host*1e3:.2f} ms/step; device {e0.elapsed_time(e1)/args.steps:.2f} ms/step")

# phase split with events
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
evs[0].record()
loss = m.training_step((x, y), 0)
evs[1].record()
loss.backward()
evs[2].record()
opt.step()
opt.zero_grad(set_to_none=True)
evs[3].record()
torch.cuda.synchronize()
print(f"fwd+loss {evs[0].elapsed_time(evs[1]):.2f} ms, bwd {evs[1].elapsed_time(evs[2]):.2f} ms, "
      f"adam {evs[2].elapsed_time(evs[3]):.2f} ms")

# GEMM table
K.enable_gemm_timing(True)
orig = K.gemm
shapes = []


def traced(a, b, out, **kw):
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    M = a.shape[1] if a_mn else a.shape[0]
    Kd = a.shape[0] if a_mn else a.shape[1]
    N = b.shape[1] if b_mn else b.shape[0]
    tag = f"M{M} N{N} K{Kd} a{int(a_mn)}b{int(b_mn)} act{kw.get('act', 0)} aux{kw.get('aux_mode', 0)} " \
          f"res{int(kw.get('residual') is not None)} {'f32' if out.dtype == torch.float32 else 'bf16'}" \
          f"{' acc' if kw.get('accumulate') else ''}{' out2' if kw.get('out2') is not None else ''}"
    shapes.append(tag)
    return orig(a, b, out, **kw)


K.gemm = traced
import visiontransformer_b200.engine as E  # noqa: E402
E.K.gemm = traced
step(0)
ev = K.collect_gemm_timing()
K.enable_gemm_timing(False)
agg = collections.OrderedDict()
for tag, (fl, ms) in zip(shapes, ev):
    a = agg.setdefault(tag, [0, 0.0, 0.0])
    a[0] += 1; a[1] += fl; a[2] += ms
tot = sum(v[2] for v in agg.values())
print(f"GEMM total {tot:.2f} ms over {len(ev)} launches")
for tag, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"  {ms:7.3f} ms n={n:3d} avg {ms/n*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {tag}")

# ---- every libvitseg wrapper timed with CUDA events (one eager step)
K.gemm = orig
E.K.gemm = orig
import visiontransformer_b200.losses as LS  # noqa: E402
names = [n for n in dir(K) if callable(getattr(K, n)) and not n.startswith("_") and n not in
         ("check", "ptr", "stream", "require_cuda", "enable_gemm_timing", "collect_gemm_timing", "launch_count",
          "reset_launch_count", "GemmDesc")]
records = []
origs = {}


def wrap(name, fn):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        records.append((name, e0, e1))
        return r
    return inner


for n in names:
    origs[n] = getattr(K, n)
    setattr(K, n, wrap(n, origs[n]))
step(0)
torch.cuda.synchronize()
agg2 = collections.OrderedDict()
for n, e0, e1 in records:
    a = agg2.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot2 = sum(v[1] for v in agg2.values())
print(f"all libvitseg calls: {tot2:.2f} ms")
for n, (c, ms) in sorted(agg2.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:7.3f} ms n={c:3d} avg {ms/c*1e3:7.1f} us  {n}")
