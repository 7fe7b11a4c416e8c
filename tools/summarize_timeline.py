"""Compares two kernel timelines of the same training step written by tools/dp_timeline.py (N = 1 vs N > 1):
per kernel family, the time spent in launches that overlapped an NCCL kernel and how much longer those launches took
than the same launch (same index in the step) on a single GPU.

  python tools/summarize_timeline.py gpurun_out/timeline_n1.csv gpurun_out/timeline_n8.csv"""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        for r in csv.DictReader(f):
            rows.append((int(r["idx"]), r["name"], float(r["start_us"]), float(r["dur_us"]), int(r["overlaps_nccl"])))
    return rows


def family(name):
    m = re.search(r"vs::(\w+)", name)
    if m:
        return m.group(1)
    n = name.lower()
    return "comm" if ("nccl" in n or "barrier" in n) else "torch glue"


a, b = load(sys.argv[1]), load(sys.argv[2])
base = {i: (n, d) for i, n, _, d, _ in a if i >= 0}
span_a = max(s + d for _, _, s, d, _ in a)
span_b = max(s + d for _, _, s, d, _ in b)
nccl_time = sum(d for i, _, _, d, _ in b if i < 0)
print(f"step span: {span_a / 1e3:.3f} ms ({sys.argv[1]}) vs {span_b / 1e3:.3f} ms ({sys.argv[2]}); NCCL kernels busy "
      f"{nccl_time / 1e3:.3f} ms in the latter\n")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0, 0.0, 0.0])
mismatch = 0
for i, n, _, d, ov in b:
    if i < 0 or i not in base:
        continue
    if family(base[i][0]) != family(n):
        mismatch += 1
        continue
    g = agg[family(n)]
    if ov:
        g[0] += 1; g[1] += base[i][1]; g[2] += d
    else:
        g[3] += 1; g[4] += base[i][1]; g[5] += d
print("| kernel family | launches overlapping NCCL | their time alone (us) | with NCCL resident (us) | stretch | other launches | alone (us) | multi-GPU run (us) | stretch |")
print("|---|---|---|---|---|---|---|---|---|")
tot = [0.0, 0.0, 0.0, 0.0]
for fam, g in sorted(agg.items(), key=lambda kv: -(kv[1][2] + kv[1][5])):
    s1 = g[2] / g[1] if g[1] else float("nan")
    s2 = g[5] / g[4] if g[4] else float("nan")
    print(f"| `{fam}` | {g[0]} | {g[1]:.0f} | {g[2]:.0f} | {s1:.2f}x | {g[3]} | {g[4]:.0f} | {g[5]:.0f} | {s2:.2f}x |")
    tot[0] += g[1]; tot[1] += g[2]; tot[2] += g[4]; tot[3] += g[5]
print(f"| total | | {tot[0]:.0f} | {tot[1]:.0f} | {tot[1] / max(tot[0], 1e-9):.2f}x | | {tot[2]:.0f} | {tot[3]:.0f} | "
      f"{tot[3] / max(tot[2], 1e-9):.2f}x |")
if mismatch:
    print(f"\n({mismatch} launches did not line up by index and were skipped)")
