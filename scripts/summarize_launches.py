#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: isolates ONE training
step (the launches between two consecutive adam_kernel launches) and prints per-kernel totals and shares.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_step_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
    adam = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k or "adam_state_kernel" in k]
    if len(adam) >= 2:
        step = rows[adam[0] + 1:adam[1] + 1]     # the first complete step of the capture window
    else:
        step = rows
    tot = sum(t for _, t in step)
    agg = OrderedDict()
    for k, t in step:
        k = re.sub(r"\s+", " ", k)
        a = agg.setdefault(k, [0.0, 0])
        a[0] += t; a[1] += 1
    print(f"# launches in the step: {len(step)}   sum of kernel durations: {tot:.1f} us")
    print(f"# {'us':>9} {'share':>6} {'n':>4} {'avg us':>8}  kernel")
    tc = 0.0
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t:9.1f} {100 * t / tot:5.1f}% {n:4d} {t / n:8.1f}  {k[:150]}")
        if "gemm_tc_kernel" in k or "chain_kernel" in k:
            tc += t
    print(f"# tcgen05 kernels (gemm_tc_kernel + chain kernels) share of the step: {100 * tc / tot:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
