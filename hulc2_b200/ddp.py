"""Data-parallel gradient exchange (the reference's only collective: DDP's bucketed all-reduce,
hulc2/training.py:72-75).

Windows are independent and every loss term is a local mean, so ranks only exchange gradients.  The
gradients already live in one contiguous fp32 arena (``FusedAdam``), laid out in module order; backward
produces them from the tail (decoder) to the head (encoders).  The arena is cut into ~``bucket_mb``
contiguous buckets; a per-parameter post-accumulate hook counts arrivals and, when a bucket is complete,
launches ``all_reduce`` on it (NCCL over NVLink/NVSwitch, asynchronous on NCCL's own stream) while
backward keeps running.  ``finish()`` waits for the outstanding buckets before the optimizer step; the
1/world_size averaging is folded into the Adam kernel (``FusedAdam.grad_scale``), reproducing DDP's
mean-of-rank-gradients exactly.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, optimizer, bucket_mb: float = 25.0, process_group=None, head_mb: float = 4.0):
        self.opt = optimizer
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        optimizer.grad_scale = 1.0 / self.world
        optimizer.attach()
        self.buckets: List[dict] = []
        self._handles = []
        self._hooks = []
        self._fired = set()
        self.used = None  # learned on the first step: ids of parameters that actually receive gradients
        cap = int(bucket_mb * (1 << 20) / 4)
        head_cap = int(head_mb * (1 << 20) / 4)
        for arena in optimizer._arenas:
            if not arena["n"]:
                continue
            params, offs = arena["params"], arena["offs"]
            # walk from the tail so bucket 0 is the first one backward completes
            end = arena["n"]
            cur_params, start = [], end
            for p, o in reversed(list(zip(params, offs))):
                cur_params.append(p)
                start = o
                if end - start >= cap:
                    self._add_bucket(arena, start, end, cur_params)
                    cur_params, end = [], start
            if cur_params:
                # The head-of-arena bucket holds the parameters whose gradients arrive LAST (the perceptual encoders: conv1's weight
                # gradient is the final kernel of backward), so its all-reduce is the one transfer nothing can overlap.  Keep it
                # small: everything but the first ~head_mb of the arena goes into a bucket of its own that completes earlier.
                ordered = list(reversed(cur_params))            # arena order
                off_of = {id(p): o for p, o in zip(params, offs)}
                cut = 0                                           # parameters that fit in the first head_cap elements
                while cut + 1 < len(ordered) and off_of[id(ordered[cut + 1])] - start <= head_cap:
                    cut += 1
                if head_cap > 0 and end - start > head_cap and 0 < cut < len(ordered):
                    split = off_of[id(ordered[cut])]
                    self._add_bucket(arena, split, end, list(reversed(ordered[cut:])))
                    self._add_bucket(arena, start, split, list(reversed(ordered[:cut])))
                else:
                    self._add_bucket(arena, start, end, cur_params)

    def _add_bucket(self, arena, start, end, params):
        b = {"view": arena["g"][start:end], "params": list(params), "pending": 0, "index": len(self.buckets)}
        self.buckets.append(b)
        for p in params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def _make_hook(self, bucket):
        def hook(param):
            self.opt.land_grad(param)          # the gradient must sit in the arena before its bucket is reduced
            self._fired.add(id(param))
            bucket["pending"] -= 1
            if bucket["pending"] == 0:
                self._launch(bucket)

        return hook

    def _launch(self, bucket):
        if self.world > 1:
            from . import ops

            ops.join_side()            # weight gradients written on the side stream (ops.wgrad_section) must be complete
            self._handles.append(dist.all_reduce(bucket["view"], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        bucket["launched"] = True

    def prepare(self, used_params=None) -> None:
        """Call before backward: arms the per-bucket arrival counters.  Parameters that receive no gradient in
        this graph (e.g. plan_recognition.layernorm, unused by the default config) are not waited for."""
        self._handles = []
        used_params = used_params if used_params is not None else self.used
        self._fired = set()
        for b in self.buckets:
            ps = [p for p in b["params"] if used_params is None or id(p) in used_params]
            b["pending"] = len(ps)
            b["launched"] = False

    def finish(self) -> None:
        """Call after backward: reduce buckets whose hooks never completed, then wait for all of them."""
        for b in self.buckets:
            if not b.get("launched"):
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        self.opt.attach()                      # parameters without a gradient this step read as (reduced) zeros
        if self.used is None:
            self.used = set(self._fired)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
