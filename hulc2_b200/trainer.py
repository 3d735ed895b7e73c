"""Native train-step driver: what Lightning's loop does around ``Hulc2.training_step`` (SURVEY.md 3.1:
AMP-less backward -> bucketed gradient all-reduce -> Adam.step), without Lightning.  Used by bench.py and
usable as a minimal trainer; under real Lightning the module is driven by ``Trainer.fit`` instead.

``use_graph=True`` captures the WHOLE step (zero_grad, forward of both modalities, backward with the bucketed
all-reduce, fused Adam, counter bump) in one CUDA graph after two eager warm-up steps and replays it afterwards:
the ~600 C-ABI launches and the autograd bookkeeping of a step cost ~21 ms of host time, about as much as the GPU
needs for the step, so replaying removes the host from the critical path.  Per-step scalars that would be frozen
into the graph live in device counters instead (noise epoch: ``noise.epoch_tensor``; Adam step number:
``FusedAdam.step_counter``).  The batch is read from static input tensors (``static_batch``); a different batch
is copied into them (device->device or pinned host->device) before the replay.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import noise
from ._lib import call
from .ddp import GradBucketReducer
from .synthetic import tree_map


def _zip_copy(dst, src, non_blocking=True):
    if isinstance(dst, dict):
        for k in dst:
            _zip_copy(dst[k], src[k], non_blocking)
    elif isinstance(dst, torch.Tensor):
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src, non_blocking=non_blocking)


class PolicyTrainer:
    def __init__(self, model, bucket_mb: float = 25.0, use_graph: bool = False, graph_warmup: int = 2):
        self.model = model
        self.optimizer = model.configure_optimizers()["optimizer"]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.reducer: Optional[GradBucketReducer] = None
        if hasattr(self.optimizer, "grad_arenas"):
            self.reducer = GradBucketReducer(self.optimizer, bucket_mb)
        self.device = next(model.parameters()).device
        self.use_graph = bool(use_graph) and hasattr(self.optimizer, "step_counter")
        self.graph_warmup = graph_warmup
        self._eager_steps = 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_batch = None
        self._static_loss = None
        self.launches_per_replay = 0
        self.replays = 0
        # Every step runs on ONE side stream: autograd's AccumulateGrad nodes remember the stream they were created on,
        # and a node created on the legacy default stream cannot take part in a stream capture.
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        if self.use_graph:
            self.optimizer.capturable = True

    # ------------------------------------------------------------------ one step (eager body; also what gets captured)
    def _step_body(self, batch: Dict[str, dict], batch_idx: int) -> torch.Tensor:
        noise.begin_step() if self.use_graph else None
        self.optimizer.zero_grad()
        loss = self.model.training_step(batch, batch_idx)
        if self.reducer is not None:
            self.reducer.prepare()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        if self.use_graph:
            # completed-step counters on the device: Adam bias correction and the noise epoch of the NEXT step
            call("hulc2_counter_add", self.optimizer.step_counter(self.device).data_ptr(), 1)
            call("hulc2_counter_add", noise.epoch_tensor(self.device).data_ptr(), 1)
        return loss.detach()

    def train_step(self, batch: Dict[str, dict], batch_idx: int = 0) -> torch.Tensor:
        """zero_grad -> forward -> backward (+ overlapped all-reduce) -> optimizer step; returns the loss tensor.
        Stream semantics are those of an ordinary call: the work is ordered after the caller's current stream and the
        caller's stream waits for it."""
        if self.stream is None:
            return self._step_body(batch, batch_idx)
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream):
            loss = self._train_step_on_stream(batch, batch_idx)
        caller.wait_stream(self.stream)
        return loss

    def _train_step_on_stream(self, batch, batch_idx):
        if not self.use_graph:
            return self._step_body(batch, batch_idx)
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._step_body(batch, batch_idx)
            self._capture(batch, batch_idx)
        else:
            _zip_copy(self.static_batch, batch)
        self._graph.replay()
        self.replays += 1
        self.optimizer.note_replayed_step()
        return self._static_loss

    def eager_step(self, batch, batch_idx: int = 0) -> torch.Tensor:
        """One step driven call by call on the trainer's stream (used for per-kernel event profiling)."""
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream):
            loss = self._step_body(batch, batch_idx)
        caller.wait_stream(self.stream)
        return loss

    def _capture(self, batch, batch_idx):
        self.static_batch = batch               # the tensors the graph reads; later batches are copied into them
        from ._lib import load_library

        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = load_library().hulc2_launch_count()
        with torch.cuda.graph(g, stream=self.stream):
            self._static_loss = self._step_body(batch, batch_idx)
        self.launches_per_replay = int(load_library().hulc2_launch_count() - n0)   # library kernels inside one replay
        # the capture did not execute anything and bumped the host-side step number once: undo that, replay counts it
        for a in self.optimizer._arenas:
            if a["n"]:
                a["step"] -= 1
        self._graph = g

    def train_step_from_host(self, host_batch: Dict[str, dict], batch_idx: int = 0) -> float:
        """End-to-end step: pinned host batch -> device (inside the call) -> step -> loss read back to host."""
        if self.use_graph and self.static_batch is not None:
            with torch.cuda.stream(self.stream):
                _zip_copy(self.static_batch, host_batch)
            loss = self.train_step(self.static_batch, batch_idx)
        else:
            batch = tree_map(lambda t: t.to(self.device, non_blocking=True), host_batch)
            loss = self.train_step(batch, batch_idx)
        return float(loss)  # device->host read of the step's result

    # ------------------------------------------------------------------ pipelined host loop (what Trainer.fit does)
    def fit_host(self, host_batches: Iterable[Dict[str, dict]]) -> List[float]:
        """Trains on an iterable of PINNED host batches and returns every step's loss (read back to the host).

        Software pipeline over three streams: the H2D copy of batch i+1 (copy stream, into one of two device staging
        buffers) runs while the trainer stream computes step i; the loss of step i is copied to a pinned slot right
        behind the step and read by the host one step later, so neither PCIe nor the host read stalls the kernels.
        Every step still moves its own inputs host->device and its own result device->host."""
        if self.stream is None:
            return [self.train_step_from_host(b, i) for i, b in enumerate(host_batches)]
        dev = self.device
        copy = self._copy_stream = getattr(self, "_copy_stream", None) or torch.cuda.Stream(device=dev)
        it = iter(host_batches)
        first = next(it, None)
        if first is None:
            return []
        if getattr(self, "_staging", None) is None or not _same_shapes(self._staging[0], first):
            self._staging = [tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), first) for _ in range(2)]
            self._pipe_static = tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), first)
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        loss_slots = torch.empty(2, dtype=torch.float32).pin_memory()
        loss_done = [torch.cuda.Event() for _ in range(2)]
        losses: List[float] = []

        def upload(i, hb):
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(free[i % 2])          # the step that consumed this staging buffer has read it
                _zip_copy(self._staging[i % 2], hb)
                ready[i % 2].record(copy)

        upload(0, first)
        i, cur = 0, first
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                upload(i + 1, nxt)                        # overlaps with step i below
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready[i % 2])
                # the tensors the step reads: the captured graph's static inputs once a graph exists
                target = self.static_batch if self._graph is not None else self._pipe_static
                _zip_copy(target, self._staging[i % 2])
                free[i % 2].record(self.stream)
                loss = self._train_step_on_stream(target, i)
                loss_slots[i % 2 : i % 2 + 1].copy_(loss.reshape(1), non_blocking=True)
                loss_done[i % 2].record(self.stream)
            if i >= 1:
                loss_done[(i - 1) % 2].synchronize()
                losses.append(float(loss_slots[(i - 1) % 2]))
            i, cur = i + 1, nxt
        loss_done[(i - 1) % 2].synchronize()
        losses.append(float(loss_slots[(i - 1) % 2]))
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        return losses


def _same_shapes(a, b) -> bool:
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same_shapes(a[k], b[k]) for k in a)
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.shape == b.shape and a.dtype == b.dtype
    return True
