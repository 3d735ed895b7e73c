#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in libhulc2_b200.so (runs without a GPU):

    python tools/sass_summary.py > profiles/sass_summary.md
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor loads / stores,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, UCGABAR / CGA = cluster barriers / DSMEM addressing."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hulc2_b200", "libhulc2_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
pats = OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"),
                    ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("cluster", r"\bUCGABAR|\bCGAERRBAR|MAPA|\.CLUSTER")])
kern, rows = None, OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = {k: 0 for k in pats}
        rows[kern]["instr"] = 0
        continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        rows[kern]["instr"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                rows[kern][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
print("# SASS summary of `hulc2_b200/libhulc2_b200.so` (sm_100a; `cuobjdump -sass`, counted per kernel by tools/sass_summary.py)\n")
print("Only kernels that contain tcgen05 / TMEM / TMA / cluster instructions are listed; the totals row covers every kernel.\n")
print("| kernel | instructions | " + " | ".join(pats) + " |")
print("|---|---:|" + "---:|" * len(pats))
tot = {k: 0 for k in list(pats) + ["instr"]}
for (k, r), name in zip(rows.items(), dem):
    for c in tot:
        tot[c] += r[c]
    if not any(r[c] for c in ("UTC*MMA", "LDTM", "STTM", "UTMALDG", "cluster")):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    print(f"| `{name[:70]}` | {r['instr']} | " + " | ".join(str(r[c]) for c in pats) + " |")
print(f"| **all {len(rows)} kernels** | {tot['instr']} | " + " | ".join(str(tot[c]) for c in pats) + " |")
