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
    ap.add_argument("--cpu-batch", type=int, default=8, help="windows per modality of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true", help="drive every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--frames", default="uint8", choices=["uint8", "fp32"],
                    help="camera frames in the batch: uint8 HWC + RandomShiftsAug draw, scaled/normalised/shifted on the device "
                         "(datamodule kernel, SURVEY 8f-1), or the reference batch contract's fp32 NCHW tensors")
    ap.add_argument("--workload", default="train", choices=["train", "rollout"],
                    help="train = configs[1] (the headline metric); rollout = configs[4]: batched policy inference, --envs parallel "
                         "environments, one re-plan + 29 plain control steps per bench step (a separate metric, never the default)")
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--variant", default="calvin", choices=["calvin", "real_world", "real_world_rgbd"],
                    help="calvin = configs[1] (default, the headline); real_world[_rgbd] = configs[3]: cfg_low_level_rw shape, "
                         "150x200 static RGB (+ depth_static), no clip loss, decoder slice [0,128]")
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wps, dt, cores = time_cpu(args.cpu_batch, args.steps, args.warmup)
    sample = f"oracle port of Hulc2.training_step+backward+Adam, fp32, {{vis:{args.cpu_batch}, lang:{args.cpu_batch}}} windows per step (bounded sample of the B={args.batch} workload), torch {torch.__version__}"
    line = {
        "impl": "reference", "metric": "train windows/sec", "value": wps, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: Hulc2 policy train step, B=64/modality, window 32, static 200x200 + gripper 84x84, 7-dof",
                   "windows_per_step_timed": 2 * args.cpu_batch},
        "cpu_baseline": {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port", "sample": sample},
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
        wps, dt, cores = time_cpu(args.cpu_batch, 2, 1)
        cpu = {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
               "sample": f"oracle port, fp32, {{vis:{args.cpu_batch}, lang:{args.cpu_batch}}} windows/step, 1 warm-up + 2 timed steps ({dt:.2f} s/step)"}

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
    # the same without the software pipeline: one blocking call per step (copy, then step, then read)
    t0 = time.perf_counter()
    for i in range(2):
        trainer.train_step_from_host(host[i % 2], i)
    torch.cuda.synchronize()
    e2e_blocking = 2 * B * world / ((time.perf_counter() - t0) / 2)

    # roofline of the dominant kernel: per-call CUDA-event timing over one extra step (outside the timed region)
    roof = None
    # every rank runs the step (it contains the gradient all-reduce); only rank 0 records the per-call events
    if rank == 0:
        _lib.profile_begin()
    trainer.eager_step(batch, 0)           # eager pass of the same step (per-call events cannot be taken inside a graph replay)
    torch.cuda.synchronize()
    if rank == 0:
        recs = _lib.profile_end()
        peaks = load_peaks()
        if args.dump_profile:
            with open(args.dump_profile, "w") as f:
                json.dump(sorted(recs.values(), key=lambda r: -r["ms"]), f, indent=1)
        top = max(recs.values(), key=lambda r: r["ms"]) if recs else None
        total_ms = sum(r["ms"] for r in recs.values())
        if top:
            sec = top["ms"] / top["calls"] * 1e-3
            tflops = (top["flops"] / top["calls"]) / sec / 1e12 if top["flops"] else 0.0
            gbs = (top.get("bytes", 0.0) / top["calls"]) / sec / 1e9
            # which roof bounds this kernel: algorithmic FLOP/byte against the machine's ridge point (measured peaks)
            ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
            intensity = top["flops"] / top["bytes"] if top.get("bytes") else float("inf")
            if intensity < ridge:
                roof = {"bound": "hbm", "kernel": top["key"], "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC_BYTES.get(top["key"]),
                        "algorithmic_bytes_per_launch": top["bytes"] / top["calls"], "flop_per_byte": intensity,
                        "ridge_flop_per_byte": ridge, "tflops": tflops}
            else:
                roof = {"bound": "tensor", "kernel": top["key"], "achieved": tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": tflops / peaks["tflops"], "traffic": None}
            roof.update({"peak_source": peaks["src"], "share_of_step": top["ms"] / total_ms if total_ms else None,
                         "calls_per_step": top["calls"], "avg_ms": top["ms"] / top["calls"],
                         "step_tflops": 2 * B * FLOPS_PER_WINDOW_FWD_BWD / (ms * 1e-3) / 1e12})
            tops = sorted(recs.values(), key=lambda r: -r["ms"])[:8]
            roof["top_kernels"] = [{"key": r["key"], "ms": round(r["ms"], 3), "calls": r["calls"]} for r in tops]

    if rank == 0:
        line = {
            "metric": "train windows/sec", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": (f"configs[1]: Hulc2 policy train step (fwd+bwd+allreduce+Adam), B={B}/modality/GPU, window 32, static 200x200 + gripper 84x84 RGB, 7-dof, lang [B,384], dropout 0.1"
                                    if not rw else
                                    f"configs[3]: cfg_low_level_rw-shaped Hulc2 train step, B={B}/modality/GPU, window 32, static 150x200 RGB{' + depth_static' if rgbd else ''} + gripper 84x84, 7-dof, no clip loss, dropout 0.1"),
                       "frames": ("uint8 HWC frames + per-frame RandomShiftsAug draw in the batch; scale/normalise/shift run on the device inside the step (fused into the trunk's pack kernel)"
                                  if args.frames == "uint8" else "fp32 NCHW frames in [-1,1] (reference batch contract; transforms already applied)"),
                       "windows_per_step_per_gpu": 2 * B, "parallelism": f"dp{world}", "l2": f"inputs ({h2d / 1e9:.2f} GB of frames/step) exceed the 126 MB L2",
                       "precision": args.precision, "cuda_graph": bool(trainer._graph is not None)},
            "clocks": clk.summary(), "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "e2e": {"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "api": f"PolicyTrainer.fit_host(pinned host batches, frames {args.frames}): H2D of step i+1 overlapped with step i",
                    "pcie_gbs": h2d * world / (float(e2e_ms) * 1e-3) / 1e9 / world, "blocking_call_value": e2e_blocking},
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
        "roofline": None, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "rollout":
        run_rollout(a)
    else:
        run_b200(a)
