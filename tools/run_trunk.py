#!/usr/bin/env python
"""One forward+backward of the static (and gripper) conv trunk at bench size, for ncu captures and quick timing.

    python tools/run_trunk.py [--frames 2048] [--reps 3] [--hw 200 200]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hulc2_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2048)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--hw", type=int, nargs=2, default=[200, 200])
a = ap.parse_args()
ops.set_precision("bf16")
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(a.frames, 3, *a.hw, device=dev, generator=g) * 2 - 1
ws = [torch.randn(32, 3, 8, 8, device=dev) * 0.05, torch.randn(64, 32, 4, 4, device=dev) * 0.04, torch.randn(64, 64, 3, 3, device=dev) * 0.04]
bs = [torch.zeros(32, device=dev), torch.zeros(64, device=dev), torch.zeros(64, device=dev)]
for rep in range(a.reps):
    if rep == a.reps - 1:
        _lib.profile_begin()
    xs, y1, y2, y3, bits = ops._convb_trunk_fwd(x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])
    dz3 = (torch.randn_like(y3, dtype=torch.float32) * (y3 > 0)).bfloat16()
    gr = ops._convb_trunk_bwd(xs, y1, y2, dz3, *ws, bits=bits)
torch.cuda.synchronize()
recs = _lib.profile_end()
tot = 0.0
for r in sorted(recs.values(), key=lambda r: -r["ms"]):
    tot += r["ms"]
    print(f"{r['ms']:8.3f} ms  {r['flops'] / max(r['ms'], 1e-9) / 1e9:8.1f} TF/s  {r['key']}")
print(f"total {tot:.3f} ms for {a.frames} frames")
