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

from typing import Dict, Optional

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
