"""Kernel timeline of ONE data-parallel training step (CUPTI through torch.profiler; works under torchrun):

  python -m torch.distributed.run --nproc-per-node N tools/dp_timeline.py --out gpurun_out/timeline_nN.csv
  python tools/dp_timeline.py --out gpurun_out/timeline_n1.csv                      # single GPU, same step

Rank 0 writes one CSV row per GPU kernel of the profiled step: launch index, name, start (us, relative), duration (us) and
whether the kernel's interval overlaps an NCCL kernel on the same GPU.  Comparing the N = 1 and N > 1 files launch by
launch shows WHICH kernels stretch while the gradient all-reduce is resident ("the loss is not exposed
communication, it is the compute kernels slowing under NCCL co-residency").  tools/summarize_timeline.py prints the table."""
import argparse
import csv
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visiontransformer_b200.ce.classes import LightningViTModel  # noqa: E402
from visiontransformer_b200.dp import DataParallel  # noqa: E402
from visiontransformer_b200.graph import GraphedTrainStep  # noqa: E402
from visiontransformer_b200.optim import FusedAdam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--eager", action="store_true", help="profile eager launches instead of a CUDA-graph replay")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = LightningViTModel(17, 16, 768, 12, 12).to(dev).train()
opt = FusedAdam(m, lr=1e-5)
dp = DataParallel(m, opt)
dp.broadcast_parameters()
g = torch.Generator().manual_seed(rank)
x = torch.rand(args.batch, 3, 224, 224, generator=g).to(dev)
y = torch.randint(0, 17, (args.batch, 256, 256), generator=g).to(dev)
for i in range(3):
    dp.step((x, y), i)
graphed = None if args.eager else GraphedTrainStep(lambda b, i: dp.step(b, i), (x, y), warmup=2, engines=[m.model.engine])


def step(i):
    if graphed is None:
        dp.step((x, y), i)
    else:
        graphed.replay()


for i in range(5):
    step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
           and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    # keep the middle step: kernels between the 1st and 2nd occurrence boundary of the optimizer kernel
    adam = [i for i, e in enumerate(evs) if "adam_kernel" in e.name]
    lo, hi = (adam[0] + 1, adam[1] + 1) if len(adam) >= 2 else (0, len(evs))
    step_evs = evs[lo:hi]
    def is_comm(name):
        n = name.lower()
        return "nccl" in n or "multimem_allreduce" in n or "barrier" in n

    nccl = [(e.time_range.start, e.time_range.end) for e in step_evs if is_comm(e.name)]
    t0 = step_evs[0].time_range.start
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["idx", "name", "start_us", "dur_us", "overlaps_nccl"])
        k = 0
        for e in step_evs:
            s, t = e.time_range.start, e.time_range.end
            is_nccl = is_comm(e.name)
            ov = any(s < b and a < t for a, b in nccl) and not is_nccl
            w.writerow([-1 if is_nccl else k, e.name[:120], f"{s - t0:.1f}", f"{t - s:.2f}", int(ov)])
            k += 0 if is_nccl else 1
    # bubble between consecutive graph replays: end of a step's optimizer kernel -> start of the next step's first kernel
    gaps = [evs[i + 1].time_range.start - evs[i].time_range.end for i in adam if i + 1 < len(evs)]
    intra = sum(max(0.0, b.time_range.start - a.time_range.end) for a, b in zip(step_evs[:-1], step_evs[1:]) if not is_comm(a.name) and not is_comm(b.name))
    print(f"rank 0: gap between steps (optimizer end -> next step's first kernel): {[round(g, 1) for g in gaps]} us; sum of the "
          f"gaps between consecutive compute kernels inside the step: {intra:.0f} us", file=sys.stderr)
    span = step_evs[-1].time_range.end - t0
    print(f"rank 0: {len(step_evs)} kernels in the profiled step, {len(nccl)} NCCL kernels, span {span / 1e3:.3f} ms -> {args.out}",
          file=sys.stderr)
if world > 1:
    del graphed
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)
