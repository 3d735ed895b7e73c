#!/usr/bin/env python
"""bench.py -- HULC++ low-level policy train step on N B200s (BASELINE.json metric: train windows/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

A "step" = zero_grad -> Hulc2.training_step on {vis: B, lang: B} windows -> backward -> bucketed gradient
all-reduce -> fused Adam, i.e. SURVEY.md 8d config 2 (B=64 per modality, window 32, static 200x200 +
gripper 84x84, 7-dof actions, language as [B,384] embeddings, dropout 0.1 active).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

WINDOW = 32
FLOPS_PER_WINDOW_FWD_BWD = 13.01e9  # SURVEY.md 8d (torch FlopCounterMode on the reference graph)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the conv trunk at
# the bench shape (profiles/r01_j_ncu_conv_trunk.md); keyed like the per-call profile
NCU_TRAFFIC_BYTES = {
    "convb_fwd[c1,F=4096,48x50x50->32,k2s1]": 1.581e9, "convb_fwd[c2,F=4096,32x49x49->64,k4s2]": 0.868e9,
    "convb_fwd[c3,F=4096,64x23x23->64,k3s1]": 0.468e9, "convb_wgrad[c3,F=4096,64x23x23->64,k3s1]": 0.514e9,
    "convb_dgrad[c3,F=4096,64x23x23<-64,k3s1]": 0.756e9, "convb_wgrad[c2,F=4096,32x49x49->64,k4s2]": 0.900e9,
    "convb_dgrad[c2,F=4096,32x49x49<-64,k4s2]": 1.499e9, "convb_wgrad[c1,F=4096,48x50x50->32,k2s1]": 1.617e9,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="windows per modality per rank")
    ap.add_argument("--precision", default=os.environ.get("HULC2_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--cpu-batch", type=int, default=None,
                    help="windows per modality of the bounded CPU sample (default: 8 for the in-run cpu_baseline, 32 = configs[0] for --impl reference)")
    ap.add_argument("--profile-passes", type=int, default=3, help="eager per-call profiling passes behind the timed region (median per call)")
    ap.add_argument("--no-fp32-frames", action="store_true", help="skip the secondary value_fp32_frames measurement")
    ap.add_argument("--no-store-e2e", action="store_true", help="e2e only with all frames of every window crossing PCIe (no HBM-resident frame store)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true", help="drive every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--frames", default="uint8", choices=["uint8", "fp32"],
                    help="camera frames in the batch: uint8 HWC + RandomShiftsAug draw, scaled/normalised/shifted on the device "
                         "(datamodule kernel, SURVEY 8f-1), or the reference batch contract's fp32 NCHW tensors")
    ap.add_argument("--workload", default="train", choices=["train", "rollout", "validation"],
                    help="train = configs[1] (the headline metric); rollout = configs[4]: batched policy inference, --envs parallel "
                         "environments, one re-plan + 29 plain control steps per bench step (a separate metric, never the default)")
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--variant", default="calvin", choices=["calvin", "real_world", "real_world_rgbd"],
                    help="calvin = configs[1] (default, the headline); real_world[_rgbd] = configs[3]: cfg_low_level_rw shape, "
                         "150x200 static RGB (+ depth_static), no clip loss, decoder slice [0,128]")
    ap.add_argument("--no-graph-profile", action="store_true", help="skip the in-graph per-call timeline (roofline from the cold eager profile)")
    ap.add_argument("--dump-profile", default=None, help="write the per-call CUDA-event profile of one step to this JSON file")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_step_factory(B, hidden_size=2048):
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch
    from oracle import hulc2_oracle as O

    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.1, hidden_size=hidden_size))
    names = {n for n, _ in m.named_parameters()}
    P = {k: v.detach().clone().requires_grad_(k in names) for k, v in m.state_dict().items()}
    cfg = hulc2_config(pkg="x", dropout_p=0.1, hidden_size=hidden_size)
    batch = synthetic_batch(B, seed=1)
    g = torch.Generator().manual_seed(3)
    S, E, H, FF, p = WINDOW, 128, 8, 2048, 0.1
    leaves = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=2e-4)

    def step():
        noise = {}
        for mod in batch:
            masks = {"emb": torch.rand(B, S, E, generator=g) > p}
            for i in range(2):
                masks[f"attn{i}"] = torch.rand(B, H, S, S, generator=g) > p
                masks[f"sa{i}"] = torch.rand(B, S, E, generator=g) > p
                masks[f"ff1{i}"] = torch.rand(B, S, FF, generator=g) > p
                masks[f"ff2{i}"] = torch.rand(B, S, E, generator=g) > p
            noise[mod] = {"plan_idx": torch.randint(0, 32, (B, 32), generator=g), "masks": masks}
        opt.zero_grad()
        out = O.training_step(batch, noise, P, cfg)
        out["loss"].backward()
        opt.step()
        return float(out["loss"])

    return step


def time_cpu(B, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_factory(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return 2 * B / dt, dt, cores


def bench_config(args, world, h2d_bytes=None, cuda_graph=True):
    """`config` of the JSON line: names the WORKLOAD (BASELINE.json configs[1] / configs[3]); both arms print the same dict."""
    B = args.batch
    rw, rgbd = args.variant != "calvin", args.variant == "real_world_rgbd"
    per_win = 32 * ((3 * (150 if rw else 200) * 200 + 3 * 84 * 84) * (1 if args.frames == "uint8" else 4) + ((150 * 200 * 4) if rgbd else 0))
    frames_gb = 2 * B * per_win / 1e9
    return {"workload": (f"configs[1]: Hulc2 policy train step (fwd+bwd+allreduce+Adam), B={B}/modality/GPU, window 32, static 200x200 + gripper 84x84 RGB, 7-dof, lang [B,384], dropout 0.1"
                         if not rw else
                         f"configs[3]: cfg_low_level_rw-shaped Hulc2 train step, B={B}/modality/GPU, window 32, static 150x200 RGB{' + depth_static' if rgbd else ''} + gripper 84x84, 7-dof, no clip loss, dropout 0.1"),
            "frames": ("uint8 HWC frames + per-frame RandomShiftsAug draw in the batch; scale/normalise/shift run on the device inside the step (fused into the trunk's pack kernel)"
                       if args.frames == "uint8" else "fp32 NCHW frames in [-1,1] (reference batch contract; transforms already applied)"),
            "windows_per_step_per_gpu": 2 * B, "parallelism": f"dp{world}", "l2": f"inputs ({frames_gb:.2f} GB of frames/step) exceed the 126 MB L2",
            "precision": args.precision, "cuda_graph": bool(cuda_graph)}


def cpu_baseline(args, B, steps, warmup, check_config_batch):
    """The oracle port on the host cores: `steps` timed steps of {vis:B, lang:B} windows after `warmup`; optionally ONE timed
    step (after one warm-up) at the reference's own CPU config batch B=32 (configs[0]) so the bias of the small sample is stated."""
    wps, dt, cores = time_cpu(B, steps, warmup)
    cpu = {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
           "sample": f"oracle port of Hulc2.training_step+backward+Adam (CPU restatement pinned to the unmodified reference), fp32, {{vis:{B}, lang:{B}}} windows/step "
                     f"(bounded sample of the B={args.batch} workload), {warmup} warm-up + {steps} timed steps ({dt:.2f} s/step), torch {torch.__version__}, {cores} threads"}
    if check_config_batch and B != 32:
        wps32, dt32, _ = time_cpu(32, 1, 1)
        cpu["config_batch_check"] = {"windows_per_step": 64, "value": wps32, "s_per_step": dt32, "bias_of_sample": wps / wps32,
                                     "note": "one timed step (after one warm-up) at the reference's own CPU config, configs[0] = {vis:32, lang:32}"}
    return cpu


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch or 32          # configs[0]: the reference's own CPU-runnable case
    cpu = cpu_baseline(args, B, args.steps, args.warmup, False)
    wps = cpu["value"]
    line = {
        "impl": "reference", "metric": "train windows/sec", "value": wps, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 2 * B / wps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": cpu,
        "e2e": {"value": wps, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- main arm
def nbytes(x):
    if isinstance(x, dict):
        return sum(nbytes(v) for v in x.values())
    return x.numel() * x.element_size() if isinstance(x, torch.Tensor) else 0


def measure_e2e_store(args, dev, world, rank, hw, barrier, steps):
    """End to end through the datamodule path: the episode frames live uint8 in an HBM-resident ring
    (`DeviceEpisodeStore`, filled once from pinned host memory OUTSIDE the timed region, like loading a dataset split); a step's
    host inputs are its window descriptors (starts, RandomShiftsAug draws), the language embeddings and the frames that are NEW
    in this step -- one frame per window and camera, as when consecutive windows of play data advance by one step and share the
    other 31 -- all copied from pinned host memory inside the timed region, the loss read back every step.  The window gather is
    part of the captured step (`PolicyTrainer(collate=store.batch_from_descriptors)`)."""
    import torch.distributed as dist

    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.datamodule import DeviceEpisodeStore
    from hulc2_b200.synthetic import tree_map
    from hulc2_b200.trainer import PolicyTrainer

    B, S, N = args.batch, WINDOW, 8192
    new = 2 * B                                      # frames ingested per step and camera: one per window
    g = torch.Generator().manual_seed(11 + rank)
    host_rgb = {"rgb_static": torch.randint(0, 256, (N, hw[0], hw[1], 3), generator=g, dtype=torch.uint8).pin_memory(),
                "rgb_gripper": torch.randint(0, 256, (N, 84, 84, 3), generator=g, dtype=torch.uint8).pin_memory()}
    rel = torch.rand(N, 7, generator=g) * 2 - 1
    rel[:, 6] = torch.where(torch.rand(N, generator=g) < 0.5, -1.0, 1.0)
    robot = torch.rand(N, 15, generator=g) * 2 - 1
    robot[:, 3:6] *= 0.9 * 3.14159265 / 2
    store = DeviceEpisodeStore(host_rgb, rel, robot, torch.rand(N, 24, generator=g) * 2 - 1, device=dev)     # one-time dataset upload

    def descriptors(i):
        head = (i * new) % N                         # ring position the step's new frames are written to
        lo = N // 2 if head < N // 2 else 0          # windows come from the half of the ring that is not being written
        d = {}
        for mod in ("vis", "lang"):
            dd = {"win_start": torch.randint(lo, lo + N // 2 - S, (B,), generator=g, dtype=torch.int64),
                  "shift_rgb_static": torch.randint(-10, 11, (B, S, 2), generator=g, dtype=torch.int32),
                  "shift_rgb_gripper": torch.randint(-4, 5, (B, S, 2), generator=g, dtype=torch.int32)}
            if mod == "lang":
                dd["lang"] = torch.randn(B, 384, generator=g)
                dd["use_for_aux_lang_loss"] = torch.ones(B, dtype=torch.bool)
            d[mod] = dd
        return tree_map(lambda t: t.pin_memory(), d)

    n_desc = 8
    host = [descriptors(i) for i in range(n_desc)]

    def ingest(i):                                   # runs on fit_host's copy stream, inside the timed region
        head = (i * new) % N
        store.write_frames(head, {k: v[head : head + new] for k, v in host_rgb.items()})

    torch.manual_seed(0)
    model = instantiate(hulc2_config(dropout_p=0.1, static_hw=hw)).to(dev).train()
    trainer = PolicyTrainer(model, use_graph=not args.no_graph, collate=lambda d: store.batch_from_descriptors(d, S))
    trainer.fit_host((host[i % n_desc] for i in range(4)), pre_copy=ingest)          # warm-up: 2 eager steps, capture, replay
    barrier()
    t0 = time.perf_counter()
    losses = trainer.fit_host((host[i % n_desc] for i in range(steps)), pre_copy=ingest)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    h2d = nbytes(host[0]) + new * sum(v[0].numel() for v in host_rgb.values())
    out = {"value": 2 * B * world / (float(ms) * 1e-3), "unit": "windows/s", "ms_per_step": float(ms), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
           "steps": steps, "new_frames_per_step_per_camera": new, "store_frames": N, "cuda_graph": bool(trainer._graph is not None),
           "loss_finite": bool(all(l == l for l in losses))}
    del trainer, model, store
    torch.cuda.empty_cache()
    return out


def family_of(key: str) -> str:
    """Kernel of a per-call profile key: one entry point AT ONE PROBLEM SHAPE (`gemm16[M=4096,N=2048,K=128]` and
    `gemm16[M=128,N=2048,K=2048]` are different tilings / split-K plans of the contraction kernel and sit at different points of
    the roofline: adding their bytes and times up describes no kernel), with the two directions of the recurrence merged (the
    same kernel, run forward and reversed)."""
    base, _, shape = key.partition("[")
    if base in ("rnn_relu_fwd", "rnn_relu_bwd"):
        return "rnn_relu[" + shape
    return key


def entry_point_of(key: str) -> str:
    base = key.split("[")[0]
    return {"rnn_relu_fwd": "rnn_relu", "rnn_relu_bwd": "rnn_relu"}.get(base, base)


def median_profile(passes):
    """{key: rec} per pass -> {key: rec} with the MEDIAN time over passes (calls / flops / bytes are identical per pass)."""
    out = {}
    for key in passes[0]:
        ms = sorted(p[key]["ms"] for p in passes if key in p)
        out[key] = dict(passes[0][key], ms=ms[len(ms) // 2])
    return out


COLD_TIMING = ("CUDA events per C-ABI call on the launch stream, eager step behind the timed region, L2 flushed (256 MB write) before every call, "
               "median of {n} passes; kernel = one entry point at one problem shape (recurrence: both directions); achieved = sum of algorithmic bytes (or FLOPs) / sum of durations")
GRAPH_TIMING = ("CUDA event-record nodes (cudaEventRecordExternal) on either side of every C-ABI call INSIDE a replayed CUDA graph of the whole step "
                "(the path the timed region runs: no host launch latency, caches as in the real step; activations and frames exceed L2), median of {n} "
                "replays; kernel = one entry point at one problem shape (recurrence: both directions); achieved = sum of algorithmic bytes (or FLOPs) / sum of durations")


def roofline_from_profile(recs, peaks, step_tflops, n_passes, timing=COLD_TIMING):
    """Groups the per-call records into kernel families, places every family that has an algorithmic work model against the
    roof that bounds it (FLOP/byte vs the measured ridge), and reports the family with the largest share of the step."""
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    fams = {}
    for r in recs.values():
        f = fams.setdefault(family_of(r["key"]), {"family": family_of(r["key"]), "ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0, "members": []})
        f["ms"] += r["ms"]
        f["calls"] += r["calls"]
        f["flops"] += r["flops"]
        f["bytes"] += r.get("bytes", 0.0)
        f["members"].append(r["key"])
    total_ms = sum(f["ms"] for f in fams.values())
    for f in fams.values():
        sec = f["ms"] * 1e-3
        f["share_of_step"] = f["ms"] / total_ms if total_ms else None
        if (f["bytes"] <= 0 and f["flops"] <= 0) or sec <= 0:
            f["bound"] = None                                   # no work model: never a roofline candidate
            continue
        intensity = f["flops"] / f["bytes"] if f["bytes"] > 0 else float("inf")
        if intensity < ridge:
            f.update(bound="hbm", achieved=f["bytes"] / sec / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
        else:
            f.update(bound="tensor", achieved=f["flops"] / sec / 1e12, peak=peaks["tflops"], unit="TFLOP/s")
        f["frac"] = f["achieved"] / f["peak"]
        f["flop_per_byte"] = None if intensity == float("inf") else intensity
    ranked = sorted((f for f in fams.values() if f["bound"]), key=lambda f: -f["ms"])
    if not ranked:
        return None
    top = ranked[0]
    traffic = [NCU_TRAFFIC_BYTES.get(k) for k in top["members"]]
    roof = {"bound": top["bound"], "kernel": top["family"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
            "frac": top["frac"], "traffic": (sum(traffic) / top["calls"]) if traffic and all(t is not None for t in traffic) else None,
            "algorithmic_bytes_per_launch": top["bytes"] / top["calls"], "algorithmic_flops_per_launch": top["flops"] / top["calls"],
            "flop_per_byte": top["flop_per_byte"], "ridge_flop_per_byte": ridge, "peak_source": peaks["src"],
            "share_of_step": top["share_of_step"], "calls_per_step": top["calls"], "avg_ms": top["ms"] / top["calls"], "members": top["members"],
            "step_tflops": step_tflops,
            "timing": timing.format(n=n_passes)}
    roof["top_kernels"] = [{"family": f["family"], "ms": round(f["ms"], 4), "calls": f["calls"], "share_of_step": round(f["share_of_step"], 4), "bound": f["bound"],
                            "achieved": round(f["achieved"], 1), "unit": f["unit"], "frac": round(f["frac"], 4)} for f in ranked[:8]]
    # the same records added up per ENTRY POINT (all shapes of a kernel together): where the step's time goes
    eps = {}
    for r in recs.values():
        e = eps.setdefault(entry_point_of(r["key"]), {"entry_point": entry_point_of(r["key"]), "ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        e["ms"] += r["ms"]; e["calls"] += r["calls"]; e["flops"] += r["flops"]; e["bytes"] += r.get("bytes", 0.0)
    roof["by_entry_point"] = [{"entry_point": e["entry_point"], "ms": round(e["ms"], 4), "calls": e["calls"], "share_of_step": round(e["ms"] / total_ms, 4),
                               "gbs": round(e["bytes"] / (e["ms"] * 1e-3) / 1e9, 1) if e["ms"] > 0 else None,
                               "tflops": round(e["flops"] / (e["ms"] * 1e-3) / 1e12, 1) if e["ms"] > 0 else None}
                              for e in sorted(eps.values(), key=lambda e: -e["ms"])[:10]]
    roof["profiled_step_ms"] = total_ms
    return roof


def measure_fp32_frames(args, dev, world, rank, rw, rgbd, hw, barrier, steps=20, warmup=4):
    """The same step with the reference batch contract (fp32 NCHW frames, transforms already applied) in the batch."""
    import torch.distributed as dist

    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch_fast
    from hulc2_b200.trainer import PolicyTrainer

    torch.manual_seed(0)
    model = instantiate(hulc2_config(dropout_p=0.1, variant="real_world" if rw else "calvin", static_hw=hw, depth_static=rgbd)).to(dev).train()
    trainer = PolicyTrainer(model, use_graph=not args.no_graph)
    batch = synthetic_batch_fast(args.batch, seed=1 + rank, device=dev, frames="fp32", static_hw=hw, depth_static=rgbd)
    for i in range(warmup):
        trainer.train_step(batch, i)
    if trainer.static_batch is not None:
        batch = trainer.static_batch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        trainer.train_step(batch, i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    del trainer, model, batch
    torch.cuda.empty_cache()
    return {"value": 2 * args.batch * world / (float(ms) * 1e-3), "unit": "windows/s", "ms_per_step": float(ms), "steps": steps, "warmup": warmup,
            "frames": "fp32 NCHW frames in [-1,1] in the batch (reference batch contract), inputs resident in HBM"}


def e2e_line(args, store, host_val, host_h2d, host_steps, host_ms, blocking):
    """`e2e`: through the datamodule path when it was measured (frames resident in an HBM ring, a step uploads its descriptors and
    its NEW frames), with the every-frame-over-PCIe measurement beside it; else the latter alone."""
    host = {"value": host_val, "unit": "windows/s", "h2d_bytes_per_step": int(host_h2d), "d2h_bytes_per_step": 4, "steps": host_steps,
            "api": f"PolicyTrainer.fit_host(pinned host batches, frames {args.frames}): ALL 32 frames of every window cross PCIe every step "
                   f"(H2D of step i+1 overlapped with step i)",
            "pcie_gbs": host_h2d / (host_ms * 1e-3) / 1e9, "blocking_call_value": blocking}
    if store is None:
        return host
    return {"value": store["value"], "unit": "windows/s", "h2d_bytes_per_step": store["h2d_bytes_per_step"], "d2h_bytes_per_step": 4,
            "steps": store["steps"], "ms_per_step": store["ms_per_step"],
            "api": "PolicyTrainer(collate=DeviceEpisodeStore.batch_from_descriptors).fit_host(pinned window descriptors, pre_copy=store.write_frames): "
                   "episode frames live uint8 in an HBM ring (filled once from pinned host memory before the timed region); every step copies, from "
                   "pinned host memory inside the timed region, its window descriptors + language embeddings + the frames that are new in this step "
                   f"({store['new_frames_per_step_per_camera']} per camera = one per window: consecutive windows share 31 of 32 frames), and reads its loss back",
            "store": {k: store[k] for k in ("new_frames_per_step_per_camera", "store_frames", "cuda_graph", "loss_finite")},
            "all_frames_over_pcie": host}


def run_b200(args):
    import torch.distributed as dist

    from hulc2_b200 import _lib, ops
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch_fast, tree_map
    from hulc2_b200.trainer import PolicyTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(args.precision)
    B = args.batch

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_batch or 8, 5, 3, True)      # 3 warm-up + 5 timed steps

    torch.manual_seed(0)
    rw, rgbd = args.variant != "calvin", args.variant == "real_world_rgbd"
    hw = (150, 200) if rw else (200, 200)
    model = instantiate(hulc2_config(dropout_p=0.1, variant="real_world" if rw else "calvin", static_hw=hw, depth_static=rgbd)).to(dev).train()
    trainer = PolicyTrainer(model, use_graph=not args.no_graph)
    batch = synthetic_batch_fast(B, seed=1 + rank, device=dev, frames=args.frames, static_hw=hw, depth_static=rgbd)
    h2d = nbytes(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.train_step(batch, i)
    if trainer.static_batch is not None:
        # the captured graph reads the trainer's own static input buffers (a clone of `batch` made at capture): time the step
        # on inputs resident THERE, as a device-side producer (datamodule) would fill them -- no per-step device-to-device copy
        batch = trainer.static_batch
    barrier()
    l0 = _lib.load_library().hulc2_launch_count()
    r0 = trainer.replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        h0 = time.perf_counter()
        for i in range(args.steps):
            loss = trainer.train_step(batch, i)
        host_ms = (time.perf_counter() - h0) * 1e3 / args.steps   # host time to ENQUEUE a step (no sync inside the loop)
        e1.record()
        torch.cuda.synchronize()
    barrier()
    # kernels of this library launched in the timed region: direct C-ABI launches + (graph replays x kernels captured per replay)
    launches = _lib.load_library().hulc2_launch_count() - l0 + (trainer.replays - r0) * trainer.launches_per_replay
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    value = 2 * B * world / (ms * 1e-3)

    # end-to-end: K steps through PolicyTrainer.fit_host -- every step copies ITS pinned host batch to the device and reads
    # ITS loss back, inside the timed region; the copy of step i+1 overlaps the kernels of step i (two host batches alternate)
    host = [tree_map(lambda t: t.cpu().pin_memory(), batch), tree_map(lambda t: t.cpu().pin_memory(), batch)]
    trainer.fit_host([host[i % 2] for i in range(2)])
    barrier()
    e2e_steps = max(args.e2e_steps, 2)
    t0 = time.perf_counter()
    trainer.fit_host(host[i % 2] for i in range(e2e_steps))
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = 2 * B * world / (float(e2e_ms) * 1e-3)
    e2e_store = None
    if args.variant == "calvin" and args.frames == "uint8" and not args.no_store_e2e:
        e2e_store = measure_e2e_store(args, dev, world, rank, hw, barrier, max(e2e_steps, 50))     # (pipeline fill / drain amortised over >= 50 steps)
    # the same without the software pipeline: one blocking call per step (copy, then step, then read)
    t0 = time.perf_counter()
    for i in range(2):
        trainer.train_step_from_host(host[i % 2], i)
    torch.cuda.synchronize()
    e2e_blocking = 2 * B * world / ((time.perf_counter() - t0) / 2)

    # secondary number on the reference batch contract: fp32 NCHW frames in the batch instead of uint8 (+ on-device transform)
    value_fp32 = None
    if args.frames == "uint8" and not args.no_fp32_frames:
        value_fp32 = measure_fp32_frames(args, dev, world, rank, rw, rgbd, hw, barrier)

    # roofline: per-call CUDA-event timing of the same step driven eagerly behind the timed region.  Every profiled call is
    # preceded by a write of a buffer larger than L2 (cold cache as under ncu, and the GPU stays behind the host so no launch
    # latency lands inside an event pair); median per call over `--profile-passes` passes; kernels grouped into families.
    roof = None
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)          # 256 MB > 126 MB L2
    passes = []
    for _ in range(max(args.profile_passes, 1)):
        # every rank runs the step (it contains the gradient all-reduce); only rank 0 records the per-call events
        if rank == 0:
            _lib.profile_begin(flush)
        trainer.eager_step(batch, 0)
        torch.cuda.synchronize()
        if rank == 0:
            passes.append(_lib.profile_end())
    del flush
    # ... and the same per-call table measured INSIDE a replayed graph of the step (event-record nodes around every call): the
    # primary roofline table, because it is the path the timed region runs; the cold-cache eager table is kept next to it
    graph_passes = None
    if trainer._graph is not None and not args.no_graph_profile:
        graph_passes = trainer.profile_replay(batch, passes=max(args.profile_passes, 1))
    if rank == 0:
        recs = median_profile(passes)
        step_tflops = 2 * B * FLOPS_PER_WINDOW_FWD_BWD / (ms * 1e-3) / 1e12
        cold = roofline_from_profile(recs, load_peaks(), step_tflops, len(passes))
        roof = cold
        recs_g = None
        if graph_passes:
            recs_g = median_profile(graph_passes)
            roof = roofline_from_profile(recs_g, load_peaks(), step_tflops, len(graph_passes), timing=GRAPH_TIMING)
            if roof is not None:
                # where the replayed step WAITS: the largest intervals between two consecutive library calls (at N > 1 the exposed
                # part of the gradient all-reduce shows up here, in front of the optimizer step)
                roof["largest_gaps_between_calls"] = [{"ms": round(g["ms"], 4), "after": g["after"], "before": g["before"]}
                                                      for g in getattr(trainer, "last_profile_gaps", [])[:8]]
            if roof is not None and cold is not None:
                roof["cold_cache_eager_profile"] = {"top_kernels": cold["top_kernels"], "profiled_step_ms": cold["profiled_step_ms"],
                                                    "timing": cold["timing"]}
        if args.dump_profile:
            with open(args.dump_profile, "w") as f:
                json.dump({"cold_eager": sorted(recs.values(), key=lambda r: -r["ms"]),
                           "graph_replay": sorted(recs_g.values(), key=lambda r: -r["ms"]) if recs_g else None}, f, indent=1)

    if rank == 0:
        line = {
            "metric": "train windows/sec", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": bench_config(args, world, cuda_graph=trainer._graph is not None),
            "value_fp32_frames": value_fp32,
            "clocks": clk.summary(), "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "e2e": e2e_line(args, e2e_store, e2e_val, h2d, e2e_steps, float(e2e_ms), e2e_blocking),
            "roofline": roof, "cpu_baseline": cpu, "loss": float(loss),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured step holds NCCL work; tear down without the collective destroy (observed to hang at exit)
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


# ----------------------------------------------------------------------------- configs[4]: batched rollout inference
def run_rollout(args):
    """N envs, S=1, replan_freq 30 -> one bench step = 1 re-plan + 29 plain control steps for all
    envs through hulc2_b200.rollout.RolloutServer (two CUDA graphs).  `value`: observations already on the device;
    `e2e`: every control step copies its uint8 camera frames + proprioception from pinned host memory and reads the
    [N,1,7] action back (what an environment loop does)."""
    from hulc2_b200 import _lib, ops
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.rollout import RolloutServer
    from hulc2_b200.synthetic import synthetic_obs, tree_map

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ops.set_precision(args.precision)
    N, CYCLE = args.envs, 30

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import hulc2_oracle as O

        torch.manual_seed(0)
        n_cpu = 16
        mc = instantiate(hulc2_config(dropout_p=0.0))
        P = {k: v.detach().clone() for k, v in mc.state_dict().items()}
        obs_c, goal_c = synthetic_obs(n_cpu, seed=2)
        ro = O.OracleRollout(P, hulc2_config(pkg="x", dropout_p=0.0))
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        g = torch.Generator().manual_seed(1)
        t0 = time.perf_counter()
        for s in range(6):
            ro.step(obs_c, goal_c, torch.randint(0, 32, (n_cpu, 32), generator=g), torch.rand(n_cpu, 1, 6, 10, generator=g),
                    torch.rand(n_cpu, 1, 6, generator=g))
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu * 6 / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"oracle port of Hulc2.step, fp32, {n_cpu} envs x 6 control steps (1 re-plan + 5 plain)"}

    torch.manual_seed(0)
    model = instantiate(hulc2_config(dropout_p=0.1)).to(dev).eval()
    srv = RolloutServer(model, use_graph=not args.no_graph)
    g = torch.Generator().manual_seed(3)
    obs_f, goal = synthetic_obs(N, seed=2)
    host = [dict(obs_f, rgb_obs={"rgb_static": torch.randint(0, 256, (N, 1, 200, 200, 3), generator=g, dtype=torch.uint8),
                                 "rgb_gripper": torch.randint(0, 256, (N, 1, 84, 84, 3), generator=g, dtype=torch.uint8)})
            for _ in range(2)]
    host = [tree_map(lambda t: t.pin_memory(), h) for h in host]
    goal_h = tree_map(lambda t: t.pin_memory(), goal)
    devobs = [tree_map(lambda t: t.to(dev), h) for h in host]
    goal_d = tree_map(lambda t: t.to(dev), goal)
    h2d = nbytes(host[0])

    def cycle(obs_list, gl, read_back):
        srv.reset()
        for s in range(CYCLE):
            a = srv.step(obs_list[s % 2], gl)
            if read_back:
                a_host.copy_(a, non_blocking=False)

    a_host = torch.empty(N, 1, 7).pin_memory()
    for _ in range(max(args.warmup, 1)):
        cycle(devobs, goal_d, False)
    torch.cuda.synchronize()
    l0 = _lib.load_library().hulc2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        e0.record()
        for _ in range(args.steps):
            cycle(devobs, goal_d, False)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (_lib.load_library().hulc2_launch_count() - l0) + args.steps * (srv.launches["replan"] + CYCLE * srv.launches["act"])
    value = N * CYCLE / (ms * 1e-3)
    cycle(host, goal_h, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(max(args.e2e_steps // 2, 2)):
        cycle(host, goal_h, True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(args.e2e_steps // 2, 2)
    # roofline of the dominant kernel family over one re-plan + one plain control step, driven eagerly with per-call events
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    eager = RolloutServer(model, use_graph=False)
    passes = []
    for _ in range(max(args.profile_passes, 1)):
        eager.reset()
        _lib.profile_begin(flush)
        eager.step(devobs[0], goal_d)
        eager.step(devobs[1], goal_d)
        torch.cuda.synchronize()
        passes.append(_lib.profile_end())
    roof = roofline_from_profile(median_profile(passes), load_peaks(), None, len(passes))
    line = {
        "metric": "rollout env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": f"configs[4]: policy rollout inference, {N} parallel envs, S=1, language goal, replan_freq 30; one bench step = "
                               f"1 re-plan + 29 plain control steps ({CYCLE} x {N} env-steps)",
                   "frames": "uint8 HWC camera frames, scaled/normalised on the device", "precision": args.precision,
                   "cuda_graph": bool(srv._graphs), "ms_per_control_step": ms / CYCLE,
                   "l2": f"observations ({h2d / 1e6:.0f} MB per control step) exceed the 126 MB L2"},
        "clocks": clk.summary(), "gpu_launches": int(launches),
        "e2e": {"value": N * CYCLE / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d * CYCLE),
                "d2h_bytes_per_step": N * 7 * 4 * CYCLE, "api": "RolloutServer.step(pinned host obs) + action read-back, every control step",
                "ms_per_control_step": e2e_ms / CYCLE},
        "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- validation_step
def cpu_validation_factory(B):
    """Oracle port of Hulc2.validation_step (hulc2.py:510-598): per modality encoders + goal + lmp_val, no_grad."""
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch
    from oracle import hulc2_oracle as O

    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.0))
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cfg = hulc2_config(pkg="x", dropout_p=0.0)
    batch = synthetic_batch(B, seed=1)
    g = torch.Generator().manual_seed(3)

    def step():
        with torch.no_grad():
            for mod, db in batch.items():
                nz = {"plan_idx_pp": torch.randint(0, 32, (B, 32), generator=g), "plan_idx_pr": torch.randint(0, 32, (B, 32), generator=g),
                      "u1_pp": torch.rand(B, WINDOW, 6, 10, generator=g), "u2_pp": torch.rand(B, WINDOW, 6, generator=g),
                      "u1_pr": torch.rand(B, WINDOW, 6, 10, generator=g), "u2_pr": torch.rand(B, WINDOW, 6, generator=g)}
                emb = O.perceptual_encoder(db["rgb_obs"], db["depth_obs"], P)
                goal = O.language_goal(db["lang"], P) if "lang" in mod else O.visual_goal(emb[:, -1], P)
                out = O.lmp_val(emb, goal, db["actions"], db["state_info"]["robot_obs"], nz, P, cfg)
                if "lang" in mod:
                    O.clip_loss(out[-1], goal, db["use_for_aux_lang_loss"], P)
        return float(out[1])

    return step


def run_validation(args):
    """`validation_step` on {vis: B, lang: B} windows of the configs[1] shape through PolicyValidator (one captured graph):
    a separate metric (validation windows/sec), never the default bench line."""
    from hulc2_b200 import _lib, ops
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch_fast, tree_map
    from hulc2_b200.trainer import PolicyValidator, _zip_copy

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ops.set_precision(args.precision)
    B = args.batch
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Bc = args.cpu_batch or 8
        step = cpu_validation_factory(Bc)
        for _ in range(2):
            step()
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        dt = (time.perf_counter() - t0) / 5
        cpu = {"value": 2 * Bc / dt, "unit": "windows/s", "cores": cores, "kind": "port",
               "sample": f"oracle port of Hulc2.validation_step, fp32, no_grad, {{vis:{Bc}, lang:{Bc}}} windows/step, 2 warm-up + 5 timed steps ({dt:.2f} s/step)"}
    torch.manual_seed(0)
    model = instantiate(hulc2_config(dropout_p=0.1)).to(dev).eval()
    val = PolicyValidator(model, use_graph=not args.no_graph)
    batch = synthetic_batch_fast(B, seed=1, device=dev, frames=args.frames)
    h2d = nbytes(batch)
    for i in range(max(args.warmup, 3)):
        val.validate(batch, i)
    if val.static_batch is not None:
        batch = val.static_batch
    torch.cuda.synchronize()
    l0, r0 = _lib.load_library().hulc2_launch_count(), val.replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        e0.record()
        for i in range(args.steps):
            out, logged = val.validate(batch, i)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.load_library().hulc2_launch_count() - l0 + (val.replays - r0) * val.launches_per_replay
    # end to end: pinned host batch -> static device batch -> replay -> one logged scalar read back, every step
    host = [tree_map(lambda t: t.cpu().pin_memory(), batch) for _ in range(2)]
    key = "val_act/action_loss_pp"

    def e2e_step(i):
        with torch.cuda.stream(val.stream):
            _zip_copy(val.static_batch if val.static_batch is not None else batch, host[i % 2])
        _, lg = val.validate(val.static_batch if val.static_batch is not None else batch, i)
        return float(lg[key])

    e2e_step(0)
    torch.cuda.synchronize()
    n_e2e = max(args.e2e_steps, 2)
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    # roofline of the dominant kernel family: eager validation passes with per-call events (see run_b200)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    eager = PolicyValidator(model, use_graph=False)
    passes = []
    for _ in range(max(args.profile_passes, 1)):
        _lib.profile_begin(flush)
        eager.validate(batch, 0)
        torch.cuda.synchronize()
        passes.append(_lib.profile_end())
    roof = roofline_from_profile(median_profile(passes), load_peaks(), None, len(passes))
    line = {
        "metric": "validation windows/sec", "value": 2 * B / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16",
        "data": "synthetic",
        "config": {"workload": f"Hulc2.validation_step (lmp_val: proposal + recognition plans, two decoder passes with action sampling, KL, MAE / gripper metrics), "
                               f"configs[1] shape, B={B}/modality, window 32, static 200x200 + gripper 84x84",
                   "frames": args.frames, "precision": args.precision, "cuda_graph": bool(val._graph is not None),
                   "l2": f"inputs ({h2d / 1e9:.2f} GB of frames/step) exceed the 126 MB L2"},
        "clocks": clk.summary(), "gpu_launches": int(launches),
        "e2e": {"value": 2 * B / (e2e_ms * 1e-3), "unit": "windows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "steps": n_e2e,
                "api": "PolicyValidator.validate(pinned host batch) + read-back of val_act/action_loss_pp, blocking per step"},
        "roofline": roof, "cpu_baseline": cpu,
        "logged": {k: float(v) for k, v in logged.items()},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "rollout":
        run_rollout(a)
    elif a.workload == "validation":
        run_validation(a)
    else:
        run_b200(a)
