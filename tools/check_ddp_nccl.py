#!/usr/bin/env python
"""Multi-rank parity on hardware: N-rank gradients through the CAPTURED step graph (bucketed NCCL
all-reduce inside the CUDA graph, 1/world folded into Adam) == the mean of the ranks' single-GPU gradients
(torch DDP semantics, hulc2/training.py:72-75).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tools/check_ddp_nccl.py [--precision fp32|bf16] [--out gpurun_out/ddp_nccl_check.json]

Per rank: (1) its own batch (seed 1 + rank), a plain single-GPU eager pass -> local gradients; all ranks exchange them with
an ordinary all_gather and form the expected mean.  (2) A PolicyTrainer(use_graph=True) with lr = 0 (parameters stay put, so
every step has the same gradient) runs 2 eager + capture + 2 replays; after each step the gradient arena (sum over ranks)
times grad_scale must equal the expected mean.  (3) With lr = 2e-4 three more replays: parameters stay bit-identical across
ranks (every rank applied the same reduced gradient).  Rank 0 prints one JSON line; exit code 1 on mismatch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--hidden", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    from helpers import build_model, to_device
    from hulc2_b200 import noise, ops
    from hulc2_b200.synthetic import synthetic_batch
    from hulc2_b200.trainer import PolicyTrainer

    ops.set_precision(args.precision)
    B = args.batch
    batch = to_device(synthetic_batch(B, seed=1 + rank, aux="all"), dev)
    g = torch.Generator().manual_seed(100 + rank)
    draws = [torch.randint(0, 32, (B, 32), generator=g).to(dev) for _ in batch]

    # (1) single-GPU gradients of this rank, plain eager autograd (no optimizer arena, no reducer)
    m0 = build_model("calvin", hidden_size=args.hidden).to(dev).train()
    with noise.supplied(categories=[d.clone() for d in draws]):
        loss0 = m0.training_step(batch, 0)
    loss0.backward()
    torch.cuda.synchronize()
    names = [n for n, p in m0.named_parameters()]
    local_flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for _, p in m0.named_parameters()])
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    expected = torch.stack(gathered).double().mean(0)
    sizes = [p.numel() for _, p in m0.named_parameters()]
    del m0

    # (2) the captured step with the NCCL all-reduce inside the graph, lr = 0
    m = build_model("calvin", hidden_size=args.hidden).to(dev).train()
    tr = PolicyTrainer(m, use_graph=True, bucket_mb=25.0)      # several buckets launched from inside backward (the default is one + head)
    tr.scheduler = None                    # a scheduler step would write its base lr back into the group after step 1
    for grp in tr.optimizer.param_groups:
        grp["lr"] = 0.0
    worst = []
    import contextlib

    for step in range(4):
        ctx = noise.supplied(categories=draws) if tr._graph is None else contextlib.nullcontext()
        with ctx:
            loss_t = tr.train_step(batch, step)
        torch.cuda.synchronize()
        if rank == 0 and os.environ.get("HULC2_DDP_CHECK_VERBOSE"):
            ar = tr.optimizer._arenas[0]
            print(f"[step {step + 1}] loss {float(loss_t):.6f} (single-GPU eager loss of this rank {float(loss0):.6f}) param checksum "
                  f"{float(ar['p'].double().abs().sum()):.9f} m checksum {float(ar['m'].double().abs().sum()):.6e} lr_dev {tr.optimizer._lr_dev.tolist()} "
                  f"step_dev {int(tr.optimizer._step_dev)} draws checksum {[int(d.sum()) for d in draws]}", flush=True)
        got = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for _, p in m.named_parameters()]).double()
        got = got * tr.optimizer.grad_scale
        errs, o = [], 0
        for n, sz in zip(names, sizes):
            e, r = got[o:o + sz], expected[o:o + sz]
            errs.append((float((e - r).abs().max() / (r.abs().max() + 1e-30)), n, e[:3].tolist(), r[:3].tolist()))
            o += sz
        worst.append(max(errs))
        if rank == 0 and os.environ.get("HULC2_DDP_CHECK_VERBOSE"):
            bad = sorted(errs, reverse=True)[:6]
            print(f"[step {step + 1}] " + "; ".join(f"{n}: err {x:.2e} got {g_} want {w_}" for x, n, g_, w_ in bad), flush=True)
            print(f"[step {step + 1}] params with err > 1e-3: {sum(1 for x in errs if x[0] > 1e-3)} of {len(errs)}", flush=True)
    assert tr._graph is not None and tr.replays == 2

    # (3) real updates: parameters must stay bit-identical across ranks
    for grp in tr.optimizer.param_groups:
        grp["lr"] = 2e-4
    for step in range(3):
        tr.train_step(batch, 4 + step)
    torch.cuda.synchronize()
    arena = tr.optimizer._arenas[0]["p"]
    chk = torch.stack([arena.double().sum(), arena.double().abs().sum(), arena.view(torch.int32).to(torch.int64).sum().double()])
    chks = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(chks, chk)
    identical = all(bool((c == chks[0]).all()) for c in chks)

    tol = 2e-5 if args.precision == "fp32" else 2e-2     # fp32: sum order of the ring reduction + LayerNorm atomics; bf16: operand rounding
    ok = identical and all(w[0] <= tol for w in worst)
    if rank == 0:
        line = {"check": "ddp_nccl_graph_gradients", "world": world, "precision": args.precision, "hidden": args.hidden, "batch_per_modality": B,
                "worst_rel_err_per_step": [{"step": i + 1, "mode": "eager" if i < 2 else "graph replay", "err": w[0], "param": w[1]} for i, w in enumerate(worst)],
                "tolerance": tol, "params_bit_identical_across_ranks": identical, "replays": tr.replays, "buckets": len(tr.reducer.buckets), "ok": ok}
        print(json.dumps(line), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(line, f, indent=1)
    sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0 if ok else 1)      # a captured step holds NCCL work: skip the collective teardown (see bench.py)


if __name__ == "__main__":
    main()
