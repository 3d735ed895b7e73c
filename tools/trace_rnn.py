"""Phase timeline of the TMA-fed cluster recurrence kernel (rnn_cluster2_sm100.cu), CTA 0: clock64 stamps at the phase
boundaries of every step (hulc2_rnn_set_trace) -> mean cycles per phase over steps 2..S-1, forward and backward.
python tools/trace_rnn.py [S B H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hulc2_b200 import ops, _lib
from hulc2_b200._lib import call

S, B, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 128, 2048)))
dev = torch.device("cuda")
lib = _lib.load_library()
g = torch.Generator().manual_seed(0)
pre = torch.randn(S, B, H, generator=g).to(dev)
w = (torch.randn(H, H, generator=g) * 0.02).to(dev)
dh = torch.randn(S, B, H, generator=g).to(dev)
ws = ops.workspace(dev)
h = torch.empty(S, B, H, device=dev)
names = ["step start", "flag of k-tile 0 seen", "flag of last k-tile seen (box issued)", "box 0 landed (issuer 1)", "last box landed",
         "last issuer committed", "accumulator complete", "TMEM read + issuer sums", "strips exchanged (push + barrier)",
         "reduced, epilogue math done", "state stored + warp synced", "previous strips consumed by the cluster", "strips pushed (stores issued)",
         "push barrier arrived", "strips added", "consumed-arrive issued"]
for label in ("fwd", "bwd"):
    trace = torch.zeros((S + 1) * 16, dtype=torch.int64, device=dev)
    for rep in range(3):
        lib.hulc2_rnn_set_trace(trace.data_ptr() if rep == 2 else None)
        if label == "fwd":
            call("hulc2_rnn_relu_fwd", pre.data_ptr(), w.data_ptr(), None, h.data_ptr(), S, B, H, 1, ws.data_ptr(), ws.numel())
        else:
            d = dh.clone()
            call("hulc2_rnn_relu_bwd", d.data_ptr(), w.data_ptr(), h.data_ptr(), None, S, B, H, 1, ws.data_ptr(), ws.numel())
        torch.cuda.synchronize()
    lib.hulc2_rnn_set_trace(None)
    t = trace.view(S + 1, 16).cpu()
    steps = range(2, S - 1)
    print(f"--- {label}: mean cycles since the step's start (steps 2..{S - 2}), path {lib.hulc2_rnn_last_path() & 255}")
    period = float(sum(int(t[i + 1, 0] - t[i, 0]) for i in steps)) / len(steps)
    for k, name in enumerate(names):
        v = [int(t[i, k] - t[i, 0]) for i in steps if int(t[i, k]) > 0]
        if v:
            print(f"  {k:2d} {name:42s} {sum(v) / len(v):8.0f}   (min {min(v)}, max {max(v)})")
    print(f"     step period {period:8.0f} cycles")
