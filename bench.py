"""bench.py — headline benchmark of the ViT-segmentation hot path (BASELINE.json configs[1]):
ViT-B/16 cross-entropy segmentation training, batch 64 per GPU, 224x224 synthetic RGB, bf16 tensor-core operands.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same metric through
the public module API with pinned HOST batches (H2D copies and the D2H loss read inside the timed region);
`roofline` = achieved TFLOP/s of the GEMM kernel measured with CUDA events inside the timed steps;
`cpu_baseline` = the oracle port of the reference timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ViT-B/16 seg train images/sec"
UNIT = "images/s"
MODEL = dict(num_classes=17, patch_size=16, hidden_size=768, num_hidden_layers=12, num_attention_heads=12)
IMAGE = 224
BATCH_PER_GPU = 64
CPU_SAMPLE_BATCH = 8


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        sm.append(float(c[1])); mx.append(float(c[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(steps: int, warmup: int, batch: int = CPU_SAMPLE_BATCH):
    """CE training step (fwd + bwd + Adam, fp32, dropout 0.1 as the reference's train mode) of ViT-B/16 on `batch` images — the reference's
    createViTmodel.py hot loop restated by oracle/vitseg_oracle.py.  Returns (seconds per step, threads)."""
    from oracle import vitseg_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = O.OracleConfig(**MODEL, image_size=IMAGE)
    sd = O.seeded_state_dict(cfg, seed=0, bf16_representable=False)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    used = [v for k, v in leaves.items() if not k.startswith("backbone.pooler")]
    opt = torch.optim.Adam(used, lr=1e-5)
    x = O.synthetic_images(batch, IMAGE, seed=1234, bf16_representable=False)
    y = O.resize_target(O.synthetic_labels(batch, MODEL["num_classes"], seed=1235), IMAGE)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = O.ce_loss(O.forward(leaves, x, cfg, dropout=(0.1, 0.1)), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = 1
    sec, threads = cpu_reference_step_time(steps, warm)
    val = CPU_SAMPLE_BATCH / sec
    sample = (f"{steps} timed CE training steps (fwd+bwd+Adam, fp32) on {CPU_SAMPLE_BATCH} images each, "
              f"{warm} warm-up, median")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 CE segmentation training, 224x224, C=17 (BASELINE configs[1] model; CPU sample "
                               f"batch {CPU_SAMPLE_BATCH})", "dropout": 0.1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from oracle import vitseg_oracle as O  # cpu_baseline leg + synthetic-input helpers only
    from visiontransformer_b200 import kernels as K
    from visiontransformer_b200.ce.classes import LightningViTModel
    from visiontransformer_b200.dp import DataParallel
    from visiontransformer_b200.model import flops_per_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's version / debug banner goes to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B = BATCH_PER_GPU
    C = MODEL["num_classes"]

    torch.manual_seed(0)
    # reference defaults (model/CE/classes.py:233-234): hidden and attention dropout 0.1, active in train()
    lm = LightningViTModel(**MODEL, image_size=IMAGE, hidden_dropout_prob=args.dropout, attention_probs_dropout_prob=args.dropout)
    lm = lm.to(dev).train()
    # Adam(lr=1e-5) of model/CE/classes.py:296-297, as the one-pass kernel over the flat arenas (update + bf16 weight
    # shadow + gradient zeroing); --torch-adam uses torch.optim.Adam(fused, capturable) instead
    if args.torch_adam:
        opt = torch.optim.Adam(lm.parameters(), lr=1e-5, fused=True, capturable=True)
    else:
        from visiontransformer_b200.optim import FusedAdam
        opt = FusedAdam(lm, lr=1e-5)
    dp = DataParallel(lm, opt)
    dp.broadcast_parameters()

    # synthetic data: device-resident copy (for `value`) and pinned host double buffer (for `e2e`)
    gen = torch.Generator().manual_seed(1234 + rank)
    host_x = [torch.rand(B, 3, IMAGE, IMAGE, generator=gen).pin_memory() for _ in range(2)]
    host_y = [torch.randint(0, C, (B, 256, 256), generator=gen).pin_memory() for _ in range(2)]
    dx, dy = host_x[0].to(dev), host_y[0].to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- the whole step (fwd + fused loss + bwd + bucketed all-reduce + Adam) as ONE CUDA graph
    from visiontransformer_b200.graph import GraphedTrainStep
    for i in range(2):
        dp.step((dx, dy), i)           # eager warm-up (one-time attribute sets, arena build)
    K.reset_launch_count()
    dp.step((dx, dy), 0)
    launches_per_step = K.launch_count()   # libvitseg kernels per step (torch's Adam / label-resize kernels not counted)
    if args.no_graph:
        graphed = None
    else:
        graphed = GraphedTrainStep(lambda b, i: dp.step(b, i), (dx, dy), warmup=2, engines=[lm.model.engine])

    def resident_step(i):
        if graphed is None:
            dp.step((dx, dy), i)
        else:
            graphed.replay()

    # ---- warm-up, then `value`
    for i in range(max(3, args.warmup)):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(resident_step, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    launches = launches_per_step * args.steps
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: pinned host batches, prefetched on a copy stream one step ahead, loss read back every step
    copy_stream = torch.cuda.Stream(device=dev)
    dev_x = [torch.empty_like(dx) for _ in range(2)]
    dev_y = [torch.empty_like(dy) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            dev_x[slot].copy_(host_x[slot], non_blocking=True)
            dev_y[slot].copy_(host_y[slot], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step(i):
        slot = i & 1
        if i == 0:
            prefetch(0)
        prefetch(slot ^ 1)  # next step's batch overlaps this step's compute
        torch.cuda.current_stream().wait_event(ready[slot])
        if graphed is None:
            loss = dp.step((dev_x[slot], dev_y[slot]), i)
        else:
            loss = graphed((dev_x[slot], dev_y[slot]))   # D2D into the graph's static inputs, then replay
        consumed[slot].record()
        losses.append(loss.item())  # device -> host read of the step result

    for s in consumed:
        s.record()
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8

    # ---- GEMM-kernel timing: the same step run eagerly with CUDA events around every vs_gemm_bf16 launch
    # (events cannot be timed inside a captured graph; same kernels, same shapes, same stream)
    K.enable_gemm_timing(True)
    barrier()
    for i in range(3):
        dp.step((dx, dy), i)
    gemm_events = K.collect_gemm_timing()
    K.enable_gemm_timing(False)
    ms_gemm_steps = None

    # ---- inference companions (same model): logits contract and fused mask path
    lm.eval()
    with torch.no_grad():
        for _ in range(3):
            lm(dx)
        ms_inf = timed(lambda i: lm(dx), max(5, args.steps))
        ms_mask = timed(lambda i: lm.model.predict_mask(dx), max(5, args.steps))
    lm.train()
    n_inf = max(5, args.steps)

    # ---- roofline of the dominant kernel (tcgen05 GEMM), CUDA events recorded inside the timed steps
    peaks, peak_src = _peaks()
    cfg = lm.model.backbone.config
    fl_img = flops_per_image(cfg, True)
    tot_flops = sum(f for f, _ in gemm_events)
    tot_ms = sum(t for _, t in gemm_events)
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    roof = None
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gemm_traffic_r01.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, one ncu pass over this step
        with open(tpath) as f:
            traffic = float(json.load(f)["traffic_bytes_per_launch"])
    if tot_ms > 0:
        ach = tot_flops / (tot_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_unit": "bytes per launch (mean over the 149 GEMM launches of one step; profiles/gemm_traffic_r01.json)", "kernel": "vs::gemm_kernel (tcgen05, all variants)", "launches_timed": len(gemm_events), "timing": "CUDA events around each GEMM launch in 3 eager (non-graph) steps of the same workload",
                "share_of_step": (tot_ms / 3) / (ms / args.steps), "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)"}
    step_tflops = value * fl_img / 1e12

    if rank == 0:
        # ---- cpu baseline (oracle port on the host cores, bounded sample)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, threads = cpu_reference_step_time(3, 1)
            cpu = {"value": CPU_SAMPLE_BATCH / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"3 timed CE training steps (fwd+bwd+Adam, fp32) on {CPU_SAMPLE_BATCH} images, 1 warm-up, median"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ViT-B/16 CE segmentation training, batch 64/GPU, 224x224, C=17 (BASELINE configs[1])",
                       "global_batch": B * world, "parallelism": f"dp{world}", "optimizer": ("torch Adam(lr=1e-5, fused, capturable)" if args.torch_adam else "FusedAdam(lr=1e-5) = vs_adam_step") + " in timed region", "cuda_graph": graphed is not None,
                       "dropout": args.dropout, "l2": "working set (4.2 GB activations + 0.9 GB weights/grads per step) >> 126 MB L2; no flush needed",
                       "label_resize": "256->224 nearest inside the step, as LightningViTModel.training_step"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "note": "pinned host batch prefetched one step ahead on a copy stream"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "model_tflops": {"achieved": step_tflops, "frac_of_sustained_peak": step_tflops / (peak_tf * world),
                             "flops_per_image": fl_img, "note": "whole training step, 3x forward FLOPs (BASELINE.md §4)"},
            "inference": {"logits_images_per_sec": world * B * n_inf / (ms_inf / 1e3),
                          "mask_images_per_sec": world * B * n_inf / (ms_mask / 1e3), "batch": B,
                          "note": "eval forward of the same model: [B,17,224,224] fp32 logits (module contract) / fused uint8 mask"},
            "final_loss": losses[-1] if losses else None,
        }
        _emit(line)
    if world > 1:
        # NCCL communicators captured into the CUDA graph: tear down explicitly and leave without the
        # process-group destructor (it can wait forever on the captured work objects)
        del graphed
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


_STDOUT_FD = None


def _emit(line: dict) -> None:
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="use torch.optim.Adam instead of the fused arena kernel")
    ap.add_argument("--dropout", type=float, default=0.1, help="hidden/attention dropout (reference default 0.1)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library writes to file descriptor 1 while the benchmark runs
    # (NCCL's version banner, for one) is sent to stderr instead; _emit() restores the descriptor for the result.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
