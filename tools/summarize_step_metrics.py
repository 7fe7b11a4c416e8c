"""Aggregates an `ncu --metrics ... --csv` log of one training step (tools/profile_step.py) per kernel family and
prints a markdown table: share of the step, DRAM GB/s, tensor-pipe %, issue %, occupancy, registers."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
d = collections.OrderedDict()
for x in csv.DictReader(lines):
    d.setdefault(x["ID"], {"k": x["Kernel Name"]})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
for v in d.values():
    k = re.sub(r"\(.*", "", v["k"]).replace("void ", "")
    k = re.sub(r"^vs::", "", k)
    if k.startswith("gemm_kernel"):
        key = "gemm_kernel (all variants)"
    elif k.startswith("at::") or "elementwise" in k or "reduce" in k:
        key = "torch glue (at::)"
    else:
        key = re.sub(r"<.*", "", k)
    a = agg.setdefault(key, dict(n=0, t=0.0, rd=0.0, wr=0.0, tens=0.0, issue=0.0, warps=0.0, regs=0))
    t = v["gpu__time_duration.sum"]
    a["n"] += 1
    a["t"] += t
    a["rd"] += v["dram__bytes_read.sum"]
    a["wr"] += v["dram__bytes_write.sum"]
    a["tens"] += v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * t
    a["issue"] += v["smsp__issue_active.avg.pct_of_peak_sustained_active"] * t
    a["warps"] += v["sm__warps_active.avg.pct_of_peak_sustained_active"] * t
    a["regs"] = max(a["regs"], int(v["launch__registers_per_thread"]))
T = sum(a["t"] for a in agg.values())
print(f"{len(d)} launches, {T / 1e6:.3f} ms serialised (cold-cache, one launch at a time: compare shares, not absolutes)\n")
print("| kernel | launches | time (us) | share | avg us | DRAM GB/s | tensor pipe % | issue % | warps active % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"]
    print(f"| `{k}` | {a['n']} | {t / 1e3:.0f} | {100 * t / T:.1f} % | {t / 1e3 / a['n']:.1f} | {(a['rd'] + a['wr']) / t:.0f} | "
          f"{a['tens'] / t:.1f} | {a['issue'] / t:.1f} | {a['warps'] / t:.1f} | {a['regs']} |")
