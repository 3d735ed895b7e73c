#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` output.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > /tmp/src.csv
    python tools/ncu_top_stalls.py /tmp/src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {name: i for i, name in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if not r or not r[0].startswith("0x"):
        break  # next kernel block
    if len(r) == len(hdr):
        data.append(r)
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
S = idx["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
print(rows[0][1] if rows[0] else "", "\ntotal samples", tot)
agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[S] or 0))[:n]:
    st = {c: int(r[idx[c]] or 0) for c in stall_cols}
    main = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda t: -t[1])[:2])
    print(f"{int(r[S]):7d} {100 * int(r[S]) / max(tot, 1):5.1f}%  exec={r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:64]:64s} {main}")
