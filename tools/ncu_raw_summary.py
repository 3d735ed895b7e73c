#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a markdown table (one row per profiled launch).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv
    python tools/ncu_raw_summary.py gpurun_out/x_raw.csv [--peak-gbs 6540] > profiles/<name>.md
Columns: duration, DRAM read / write bytes and achieved DRAM GB/s (= (read + write) / duration), DRAM / L2 / tensor-pipe /
issue-slot utilisation, and the three largest warp-stall reasons (average stalled warps per issue-active cycle).
"""
import csv
import re
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    path = sys.argv[1]
    peak = float(sys.argv[sys.argv.index("--peak-gbs") + 1]) if "--peak-gbs" in sys.argv else 6539.9
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]

    def scale(name, v):   # bytes columns come in K/M/Gbyte
        u = units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def dur_us(r):
        u = units[col["gpu__time_duration.sum"]].lower()
        return num(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

    print("| kernel | grid x block | time us | dram rd MB | dram wr MB | dram GB/s | % of HBM peak | dram % | L2 % | tensor pipe % | issue active % | top stalls (warps / issue) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r[col["Kernel Name"]])
        name = re.sub(r"\(.*", "", name)[:60]
        t = dur_us(r)
        rd, wr = scale("dram__bytes_read.sum", num(r[col["dram__bytes_read.sum"]])), scale("dram__bytes_write.sum", num(r[col["dram__bytes_write.sum"]]))
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t else 0.0
        g = lambda k: num(r[col[k]]) if k in col else 0.0
        st = sorted(((num(r[col[s]]), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:3]
        print(f"| `{name}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | "
              f"{g('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")


if __name__ == "__main__":
    main()
