"""World-size-2 CPU (gloo) tests of the data-parallel host logic: the bucketed gradient reducer over the
FusedAdam gradient arena must reproduce DDP's mean-of-rank-gradients (hulc2/training.py:72-75, torch DDP semantics),
launch buckets tail-first while backward is still running, not wait for parameters that receive no gradient, and
bench.py's sharding rule (each rank its own {vis:B, lang:B}, value = all windows / max-over-ranks time) must hold.
No CUDA kernel is called: the arena / bucket / hook code is device-agnostic torch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(24, 40)
        self.b = torch.nn.Linear(40, 40)
        self.unused = torch.nn.Linear(8, 8)     # never reached by forward (cf. plan_recognition.layernorm in the default config)
        self.c = torch.nn.Linear(40, 6)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


def _rank_batch(rank: int):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(16, 24, generator=g), torch.randn(16, 6, generator=g)


def _worker(rank: int, world: int, port: int, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hulc2_b200.ddp import GradBucketReducer
        from hulc2_b200.optim import FusedAdam

        net = _Net()
        opt = FusedAdam(net.parameters(), lr=2e-4)
        red = GradBucketReducer(opt, bucket_mb=300 * 4 / (1 << 20))        # 300-element cap -> buckets [c, unused], [b], [a]
        assert red.world == world and opt.grad_scale == 1.0 / world
        assert len(red.buckets) >= 3
        # buckets tile the arena exactly once, tail first
        spans = sorted((b["view"].data_ptr(), b["view"].numel()) for b in red.buckets)
        arena = opt._arenas[0]
        assert spans[0][0] == arena["g"].data_ptr() and sum(n for _, n in spans) == arena["n"]
        assert red.buckets[0]["view"].data_ptr() > red.buckets[-1]["view"].data_ptr()
        launched = []
        orig = red._launch
        red._launch = lambda b: (launched.append(b["index"]), orig(b))[1]

        for step in range(3):                       # step 0 learns which parameters are used; steps 1, 2 overlap
            if step < 2:
                arena["g"].zero_()                  # param.grad stays the arena view: autograd accumulates in place
            else:
                opt.zero_grad()                     # param.grad = None: autograd adopts its own tensors, the hook lands them
            x, y = _rank_batch(rank)
            loss = ((net(x) - y) ** 2).mean()
            red.prepare()
            launched.clear()
            loss.backward()
            during_backward = list(launched)
            red.finish()
            assert all(b["launched"] for b in red.buckets)
            assert sorted(launched) == list(range(len(red.buckets)))
            if step >= 1:
                # every bucket without an unused parameter fired from its hook, the decoder-side bucket first
                assert during_backward and during_backward[0] == 0, during_backward
                assert len(during_backward) == len(red.buckets), (during_backward, len(red.buckets))
        avg = {n: (p.grad * opt.grad_scale).clone() for n, p in net.named_parameters()}
        # every param.grad is still a view of the arena (the all-reduce ran in place on contiguous buckets)
        lo, hi = arena["g"].data_ptr(), arena["g"].data_ptr() + 4 * arena["n"]
        assert all(lo <= p.grad.data_ptr() < hi for p in net.parameters())
        out[rank] = {k: v.numpy() for k, v in avg.items()}
    finally:
        dist.destroy_process_group()


def test_bucketed_reducer_matches_mean_of_rank_gradients():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        got = [dict(out[r]) for r in range(world)]
    # reference: plain autograd on each rank's batch, averaged (DDP semantics)
    want = None
    for r in range(world):
        net = _Net()
        x, y = _rank_batch(r)
        ((net(x) - y) ** 2).mean().backward()
        g = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) / world for n, p in net.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for r in range(world):
        for k, v in want.items():
            torch.testing.assert_close(torch.from_numpy(got[r][k]), v, rtol=1e-6, atol=1e-7, msg=lambda m, k=k, r=r: f"rank {r} {k}: {m}")
    # both ranks hold identical averaged gradients
    for k in want:
        assert (got[0][k] == got[1][k]).all()


def _bench_worker(rank: int, world: int, port: int, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hulc2_b200.synthetic import synthetic_batch

        b = synthetic_batch(2, seed=1 + rank)          # bench.py: seed = 1 + rank -> every rank its own windows
        ms = torch.tensor([10.0 + 5.0 * rank])
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)       # bench.py: the step time is the max over ranks
        out[rank] = (float(b["vis"]["actions"].sum()), float(ms))
    finally:
        dist.destroy_process_group()


def test_bench_sharding_rule_and_max_over_ranks_timing():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_bench_worker, args=(world, port, out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    assert res[0][0] != res[1][0]                      # different windows per rank (weak scaling, no data-path collective)
    assert res[0][1] == res[1][1] == 15.0              # max over ranks
