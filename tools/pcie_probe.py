#!/usr/bin/env python
"""H2D bandwidth from pinned host memory: one stream vs the transfer split over two streams (is bench.py's e2e leg,
0.58 GB per step, at the PCIe roof?)."""
import time

import torch

dev = torch.device("cuda")
n = 290 * (1 << 20)
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
dst = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
s = [torch.cuda.Stream() for _ in range(2)]


def run(two_streams: bool, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(2):
            with torch.cuda.stream(s[i if two_streams else 0]):
                dst[i].copy_(host[i], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return 2 * n / dt / 1e9


for _ in range(2):
    run(False, 2), run(True, 2)
print(f"one stream : {run(False):.1f} GB/s")
print(f"two streams: {run(True):.1f} GB/s")
# many small tensors like a batch dict (64 x 9 MB)
hs = [torch.empty(9 << 20, dtype=torch.uint8).pin_memory() for _ in range(64)]
ds = [torch.empty(9 << 20, dtype=torch.uint8, device=dev) for _ in range(64)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    for h, d in zip(hs, ds):
        d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print(f"64 x 9 MB, one stream: {5 * 64 * (9 << 20) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
