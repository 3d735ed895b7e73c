"""Times the decoder recurrence (S=32, B=128, H=2048: both modalities of a B=64 step) on each kernel generation:
python tools/bench_rnn.py  -> per kernel: us per layer call fwd / bwd, us per step, TFLOP/s, streaming-model GB/s.
Streaming model: per step sizeof(W_hh bf16) + 3 B H 4 bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hulc2_b200 import ops, _lib
from hulc2_b200._lib import call

S, B, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 128, 2048)))
dev = torch.device("cuda")
lib = _lib.load_library()
print("co-resident clusters: size 8:", lib.hulc2_rnn_cluster_capacity(8), " size 4:", lib.hulc2_rnn_cluster_capacity(4), "| last error:", lib.hulc2_last_error().decode())
g = torch.Generator().manual_seed(0)
pre = torch.randn(S, B, H, generator=g).to(dev)
w = (torch.randn(H, H, generator=g) * 0.02).to(dev)
dh = torch.randn(S, B, H, generator=g).to(dev)
ws = ops.workspace(dev)
h = torch.empty(S, B, H, device=dev)
flops = 2.0 * S * B * H * H
stream_bytes = S * (H * H * 2 + 3 * B * H * 4)
ref = None
for which, name in ((1, "1-D persistent (rnn_persistent_sm100.cu)"), (-1, "cluster split-K (rnn_cluster_sm100.cu)"),
                    (0, "cluster split-K, TMA-fed (rnn_cluster2_sm100.cu)")):
    lib.hulc2_rnn_select_kernel(which)
    res = {}
    for label in ("fwd", "bwd"):
        def run():
            if label == "fwd":
                call("hulc2_rnn_relu_fwd", pre.data_ptr(), w.data_ptr(), None, h.data_ptr(), S, B, H, 1, ws.data_ptr(), ws.numel())
            else:
                d.copy_(dh)
                call("hulc2_rnn_relu_bwd", d.data_ptr(), w.data_ptr(), h.data_ptr(), None, S, B, H, 1, ws.data_ptr(), ws.numel())
        d = dh.clone()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the copy_ of the bwd input is timed separately and subtracted
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            d.copy_(dh)
        c1.record()
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3 - (c0.elapsed_time(c1) / reps * 1e3 if label == "bwd" else 0.0)
        res[label] = us
    path = lib.hulc2_rnn_last_path()
    name += f" [path {path & 255}, v2 reject {path >> 8}]"
    out = h.clone()
    if ref is None:
        ref = out
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"{name}: fwd {res['fwd']:.1f} us ({res['fwd']/S:.2f} us/step, {flops/res['fwd']/1e6:.1f} TFLOP/s, "
          f"{stream_bytes/res['fwd']/1e3:.0f} GB/s streaming-model) | bwd {res['bwd']:.1f} us ({res['bwd']/S:.2f} us/step) | max rel diff vs 1-D {err:.2e}")
lib.hulc2_rnn_select_kernel(0)
