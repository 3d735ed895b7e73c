#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] [--top 25] > profiles/<name>.md
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)(<.*>)?\(", name)
    if not m:
        return name[:90]
    base, targs = m.group(1), m.group(2) or ""
    return (base + targs)[:110]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = defaultdict(lambda: [0.0, 0])
    for _, k, us, _, _ in rows:
        a = agg[short(k)]
        a[0] += us
        a[1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"launches: {len(rows)}  total device time: {total / 1e3:.3f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| share | total us | calls | avg us | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, (us, n) in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
        print(f"| {100 * us / total:5.1f}% | {us:10.1f} | {n:4d} | {us / n:8.1f} | `{k}` |")


if __name__ == "__main__":
    main()
