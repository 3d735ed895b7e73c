"""Times the dense contraction shapes of the policy step (B=128 windows) on the gather kernel (fp32 operands converted on
the fly) and on the TMA kernel (bf16 mirrors): python tools/bench_gemm.py  -> one line per shape."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hulc2_b200 import ops

ops.set_precision("bf16")
dev = "cuda"
# (name, M, N, K, a_major, b_major): NT = fwd, NN = dgrad (B MN-major), TN = wgrad (both MN-major)
SHAPES = [
    ("rnn_l1_in fwd", 4096, 2048, 2048, "k", "k"), ("rnn_l1_in dgrad", 4096, 2048, 2048, "k", "mn"),
    ("rnn wgrad", 2048, 2048, 4096, "mn", "mn"), ("ffn1 fwd", 4096, 2048, 128, "k", "k"), ("ffn2 fwd", 4096, 128, 2048, "k", "k"),
    ("ffn1 wgrad", 2048, 128, 4096, "mn", "mn"), ("pp layer fwd", 128, 2048, 2048, "k", "k"), ("pp layer dgrad", 128, 2048, 2048, "k", "mn"),
    ("pp wgrad", 2048, 2048, 128, "mn", "mn"), ("heads fwd", 4096, 182, 2048, "k", "k"), ("heads dgrad", 4096, 2048, 184, "k", "mn"),
    ("heads wgrad", 182, 2048, 4096, "mn", "mn"), ("emb in fwd", 4096, 2048, 64, "k", "k"), ("grip fc fwd", 4096, 128, 3136, "k", "k"),
    ("grip fc wgrad", 128, 3136, 4096, "mn", "mn"), ("pr fc mean", 128, 4096, 128, "k", "k"), ("fc_state", 128, 1024, 4096, "k", "k"),
]


def run(name, M, N, K, am, bm, reps=20):
    A = torch.randn((M, K) if am == "k" else (K, M), device=dev)
    B = torch.randn((N, K) if bm == "k" else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    a_rs, a_ks = (K, 1) if am == "k" else (1, M)
    b_rs, b_ks = (K, 1) if bm == "k" else (1, N)
    A16, B16 = ops.to_bf16(A), ops.to_bf16(B)
    out = {}
    for label, kw in (("gather", {}), ("tma", {"A16": A16, "B16": B16})):
        for _ in range(3):
            ops.gemm(M, N, K, A, a_rs, a_ks, B, b_rs, b_ks, C, N, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.gemm(M, N, K, A, a_rs, a_ks, B, b_rs, b_ks, C, N, **kw)
        e1.record()
        torch.cuda.synchronize()
        out[label] = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * M * N * K
    print(f"{name:18s} M={M:5d} N={N:5d} K={K:5d} {am}/{bm}: gather {out['gather']:8.1f} us ({fl/out['gather']/1e6:7.1f} TF)   tma {out['tma']:8.1f} us ({fl/out['tma']/1e6:7.1f} TF)")


for s in SHAPES:
    run(*s)
