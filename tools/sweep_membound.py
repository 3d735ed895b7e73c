#!/usr/bin/env python
"""Size sweep of the memory-bound kernels (SURVEY.md 7 "Roofline on tiny tensors": report config size AND a >= 256 MB size).

    python tools/sweep_membound.py > profiles/r02_membound_sweep.md

Every kernel is called through the C-ABI on synthetic operands, `REPS` times; an L2 flush (256 MB write) precedes every call
and CUDA events bracket the call alone; the median is reported.  `alg MB` = algorithmic bytes (SURVEY.md 8d per-unit figures
x units), GB/s = alg bytes / median time, % = of the measured HBM peak (MEASURED_PEAKS.json, else 6540 GB/s)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from hulc2_b200 import _lib, ops  # noqa: E402
from hulc2_b200._lib import call  # noqa: E402

REPS = int(os.environ.get("HULC2_SWEEP_REPS", "15"))      # 1 under ncu (HULC2_SWEEP_ONLY=large restricts the run to the >= 256 MB rows)
ONLY = os.environ.get("HULC2_SWEEP_ONLY", "")
dev = torch.device("cuda")
peak = 6539.9
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
rows = []


def timed(name, size, alg_bytes, fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    rows.append((name, size, alg_bytes / 1e6, ms * 1e3, gbs, 100 * gbs / peak))


def R(*shape, dtype=torch.float32, lo=-1.0, hi=1.0):
    return (torch.rand(*shape, device=dev) * (hi - lo) + lo).to(dtype)


def logistic(B, S, tag):
    A, M, LD = 6, 10, ops.HEAD_LD
    heads = R(S * B, LD)
    act = R(B, S, 7)
    act[..., 6] = torch.where(act[..., 6] > 0, 1.0, -1.0)
    amin, amax = torch.full((1, 1, A, M), -1.0, device=dev), torch.full((1, 1, A, M), 1.0, device=dev)
    out, g, dheads = torch.empty(3, device=dev), torch.ones(1, device=dev), torch.empty_like(heads)
    ws = ops.workspace(dev)
    n = B * S * A
    timed("logistic-mixture NLL fwd (+ gripper CE)", tag, 124.0 * n, lambda: call(
        "hulc2_logistic_loss_seg_fwd", heads.data_ptr(), LD, act.data_ptr(), amin.data_ptr(), amax.data_ptr(), out.data_ptr(), B, S, A, M, 10, -7.0, 1.0, 1, B,
        ws.data_ptr(), ws.numel()))
    timed("logistic-mixture NLL bwd", tag, 244.0 * n, lambda: call(
        "hulc2_logistic_loss_seg_bwd", heads.data_ptr(), LD, act.data_ptr(), amin.data_ptr(), amax.data_ptr(), g.data_ptr(), dheads.data_ptr(), B, S, A, M, 10, -7.0, 1.0, 1, B))
    u1, u2 = R(B, S, A, M, lo=0.0), R(B, S, A, lo=0.0)
    gb = torch.tensor([-1.0, 1.0], device=dev)
    o = torch.empty(B, S, 7, device=dev)
    timed("logistic-mixture sampling (Gumbel argmax + inverse CDF)", tag, 168.0 * n + 12.0 * B * S, lambda: call(
        "hulc2_logistic_sample", heads.data_ptr(), LD, u1.data_ptr(), u2.data_ptr(), gb.data_ptr(), o.data_ptr(), B, S, A, M, -7.0, 1))


def kl(B, tag):
    pp, pr = R(B, 1024), R(B, 1024)
    loss, g = torch.empty(1, device=dev), torch.ones(1, device=dev)
    dpp, dpr = torch.empty_like(pp), torch.empty_like(pr)
    timed("categorical KL fwd (32 x 32, balanced)", tag, 8192.0 * B, lambda: call("hulc2_kl_fwd", pp.data_ptr(), pr.data_ptr(), loss.data_ptr(), B, 32, 32, 0.8, 0.01))
    timed("categorical KL bwd", tag, 16384.0 * B, lambda: call("hulc2_kl_bwd", pp.data_ptr(), pr.data_ptr(), g.data_ptr(), dpp.data_ptr(), dpr.data_ptr(), B, 32, 32, 0.8, 0.01))


def ssm(F, tag):
    HW, C = 441, 64
    y3 = R(F, HW, C, dtype=torch.bfloat16)
    lin = torch.linspace(-1, 1, 21, device=dev)
    xm, ym = lin.repeat_interleave(21).contiguous(), lin.repeat(21).contiguous()
    temp = torch.ones(1, device=dev)
    out, dout, dz = torch.empty(F, 128, device=dev), R(F, 128), torch.empty_like(y3)
    stats = torch.empty(F, 128, device=dev)
    # the calls the train step makes: forward saving the softmax statistics, backward consuming them (one pass each)
    timed("SpatialSoftmax fwd (bf16 NHWC in, saves statistics)", tag, 2.0 * y3.numel() + 8.0 * out.numel(), lambda: call(
        "hulc2_spatial_softmax_fwd_bf16_stats", y3.data_ptr(), xm.data_ptr(), ym.data_ptr(), temp.data_ptr(), out.data_ptr(), stats.data_ptr(), F, HW, C))
    timed("SpatialSoftmax bwd (from statistics, + conv3 ReLU mask)", tag, 4.0 * y3.numel() + 12.0 * dout.numel(), lambda: call(
        "hulc2_spatial_softmax_bwd_bf16_stats", y3.data_ptr(), xm.data_ptr(), ym.data_ptr(), temp.data_ptr(), out.data_ptr(), stats.data_ptr(), dout.data_ptr(),
        dz.data_ptr(), None, F, HW, C, 1))


def adam(n, tag):
    pr, g, m, v = R(n), R(n), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    timed("fused Adam", tag, 28.0 * n, lambda: call("hulc2_adam_step", pr.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-4, 0.9, 0.999, 1e-8, 0.0, 1, 1.0))


def layernorm(rows_, D, tag):
    x, res = R(rows_, D), R(rows_, D)
    gam, bet = R(D), R(D)
    y, t = torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(rows_, device=dev), torch.empty(rows_, device=dev)
    timed("LayerNorm(+residual) fwd", tag, 16.0 * rows_ * D, lambda: call(
        "hulc2_layernorm_fwd", x.data_ptr(), D, res.data_ptr(), D, None, 1.0, gam.data_ptr(), bet.data_ptr(), y.data_ptr(), D, t.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows_, D, 1e-5))
    dy, dx, dres = R(rows_, D), torch.empty_like(x), torch.empty_like(x)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    timed("LayerNorm(+residual) bwd", tag, 16.0 * rows_ * D, lambda: call(
        "hulc2_layernorm_bwd", dy.data_ptr(), D, t.data_ptr(), D, gam.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), D, dres.data_ptr(), None, 1.0, dg.data_ptr(), db.data_ptr(), rows_, D))


def cells(B, H, tag):
    gi, gh, hp = R(B, 3 * H), R(B, 3 * H), R(B, H)
    h, sv = torch.empty(B, H, device=dev), torch.empty(B, 4 * H, device=dev)
    timed("GRU cell fwd", tag, 48.0 * B * H, lambda: call("hulc2_gru_cell_fwd", gi.data_ptr(), 3 * H, gh.data_ptr(), hp.data_ptr(), h.data_ptr(), sv.data_ptr(), B, H))
    dh, dgi, dgh, dhp = R(B, H), torch.empty(B, 3 * H, device=dev), torch.empty(B, 3 * H, device=dev), torch.empty(B, H, device=dev)
    timed("GRU cell bwd", tag, 52.0 * B * H, lambda: call("hulc2_gru_cell_bwd", dh.data_ptr(), None, sv.data_ptr(), hp.data_ptr(), dgi.data_ptr(), 3 * H, dgh.data_ptr(), dhp.data_ptr(), B, H))
    gi4, gh4, cp = R(B, 4 * H), R(B, 4 * H), R(B, H)
    c = torch.empty(B, H, device=dev)
    timed("LSTM cell fwd", tag, 60.0 * B * H, lambda: call("hulc2_lstm_cell_fwd", gi4.data_ptr(), 4 * H, gh4.data_ptr(), cp.data_ptr(), h.data_ptr(), c.data_ptr(), sv.data_ptr(), B, H))
    dg4, dc = torch.empty(B, 4 * H, device=dev), torch.empty(B, H, device=dev)
    timed("LSTM cell bwd", tag, 56.0 * B * H, lambda: call("hulc2_lstm_cell_bwd", dh.data_ptr(), None, None, sv.data_ptr(), c.data_ptr(), cp.data_ptr(), dg4.data_ptr(), 4 * H, dc.data_ptr(), B, H))


def frames(F, tag):
    u8 = torch.randint(0, 256, (F, 200, 200, 3), dtype=torch.uint8, device=dev)
    sh = torch.randint(-10, 11, (F, 2), dtype=torch.int32, device=dev)
    fr = ops.U8Frames(u8, sh)
    xs = torch.empty(F * 50 * 50 * 48 + 64, dtype=torch.bfloat16, device=dev)
    timed("uint8 frame pack (shift + normalise + space-to-depth)", tag, 3.0 * u8.numel(), lambda: fr.pack_into(xs.data_ptr()))


def tcp(rows_, tag):
    a, r, o = R(rows_, 7), R(rows_, 15), torch.empty(rows_, 7, device=dev)
    timed("world -> tcp frame", tag, 88.0 * rows_, lambda: call("hulc2_world_to_tcp", a.data_ptr(), r.data_ptr(), 15, o.data_ptr(), rows_))


def main():
    cfg, big = "config (B=128 windows)", ">= 256 MB"
    small = ONLY != "large"
    if small:
        logistic(128, 32, cfg)
    logistic(128 * 96, 32, big)            # 12288 windows: 293 MB fwd
    if small:
        kl(128, cfg)
    kl(32768, big)                         # 268 MB fwd
    ssm(4096, cfg + ": 4096 frames")       # already 231 MB
    if small:
        ssm(8192, "2x config")
    adam(47053840, cfg + ": 47.05 M parameters")
    if small:
        layernorm(4096, 128, cfg)
    layernorm(4096 * 128, 128, big)
    if small:
        cells(128, 2048, cfg)
    cells(128 * 24, 2048, big)
    if small:
        frames(2048, cfg + ": 2048 frames")
    frames(4096, "2x config")
    if small:
        tcp(4096, cfg)
    tcp(4096 * 1024, big)
    print("# r02 size sweep of the memory-bound kernels (C-ABI calls, L2 flushed before every call, median of %d, CUDA events)\n" % REPS)
    print(f"HBM peak = {peak:.0f} GB/s (MEASURED_PEAKS.json). `alg MB` = algorithmic bytes (SURVEY.md 8d).\n")
    print("| kernel | size | alg MB | time us | GB/s | % of HBM peak |")
    print("|---|---|---:|---:|---:|---:|")
    for name, size, mb, us, gbs, pct in rows:
        print(f"| {name} | {size} | {mb:.2f} | {us:.1f} | {gbs:.0f} | {pct:.1f} |")
    print("\nReading: at config size the loss / KL / sampling / frame-transform kernels move 0.1-3 MB: a launch is a few microseconds of fixed latency "
          "around a sub-microsecond transfer, so their roofline fraction at config size says nothing about the kernel; the >= 256 MB rows show what the same "
          "code does when the transfer dominates.")


if __name__ == "__main__":
    main()
