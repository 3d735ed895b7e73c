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

from . import noise, ops
from ._lib import call
from .ddp import GradBucketReducer
from .synthetic import tree_map


def _zip_copy(dst, src, non_blocking=True):
    """Copies the leaves of batch ``src`` into the same-shaped static batch ``dst`` (the tensors a captured step reads).
    Camera frames delivered as ``ops.U8Frames`` (datamodule batches) keep pointing at the same resident frame store; their
    window starts / lengths / shift draws are index tensors and are copied like any other leaf.  Any leaf that cannot be
    refreshed in place raises: a replay must never silently pair old frames with new actions."""
    if isinstance(dst, dict):
        if not isinstance(src, dict) or dst.keys() != src.keys():
            raise KeyError(f"batch layout changed under a captured step: {sorted(dst)} vs {sorted(src) if isinstance(src, dict) else type(src)}")
        for k in dst:
            _zip_copy(dst[k], src[k], non_blocking)
    elif isinstance(dst, torch.Tensor):
        if not isinstance(src, torch.Tensor) or dst.shape != src.shape or dst.dtype != src.dtype:
            raise ValueError(f"batch leaf changed shape/dtype under a captured step: {tuple(dst.shape)} {dst.dtype} vs "
                             f"{tuple(src.shape) if isinstance(src, torch.Tensor) else type(src)}")
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src, non_blocking=non_blocking)
    elif isinstance(dst, ops.U8Frames):
        if not isinstance(src, ops.U8Frames) or src.u8.data_ptr() != dst.u8.data_ptr() or src.u8.shape != dst.u8.shape or src.S != dst.S:
            raise ValueError("U8Frames of a captured step must view the same resident frame store (pointer, shape, window length)")
        for name in ("win_start", "win_len", "shift"):
            d, s_ = getattr(dst, name), getattr(src, name)
            if (d is None) != (s_ is None):
                raise ValueError(f"U8Frames.{name} appeared/disappeared under a captured step")
            if d is not None:
                _zip_copy(d, s_, non_blocking)
    elif dst is None and src is None:
        pass
    elif isinstance(dst, (int, float, bool, str)) and dst == src:
        pass
    else:
        raise TypeError(f"cannot refresh a batch leaf of type {type(dst).__name__} in place for a captured step")


class PolicyTrainer:
    def __init__(self, model, bucket_mb: Optional[float] = None, use_graph: bool = False, graph_warmup: int = 2,
                 device_counters: Optional[bool] = None, collate=None):
        import os

        if bucket_mb is None:
            # Default: ONE bucket behind the small head-of-arena bucket.  Measured at 4 x B200 (tools/scale_ab_buckets.sh): 12 / 25 /
            # 50 / 100 / 200 MB buckets -> 7.20 / 7.09 / 7.14 / 7.04 / 6.89 ms per step (6.7 at N = 1).  The bucket that holds everything
            # but the first 4 MB completes when the encoders' MLP gradients arrive, ~2 ms before the end of backward, so its 184 MB
            # all-reduce still hides behind the conv trunk's backward, while every additional captured NCCL launch costs fork / join
            # edges in the step graph and SMs taken from the kernels running beside it.  HULC2_BUCKET_MB overrides (A/B switch).
            bucket_mb = float(os.environ.get("HULC2_BUCKET_MB", "256"))
        self.model = model
        # collate: optional device-side batch builder run INSIDE the step (and therefore inside the captured graph): maps
        # the tensors handed to train_step (e.g. window descriptors: starts, lengths, augmentation draws) to the model's batch
        # dict (datamodule.DeviceEpisodeStore.batch_from_descriptors gathers windows out of a frame store resident in HBM)
        self.collate = collate
        opt_cfg = model.configure_optimizers()
        self.optimizer = opt_cfg["optimizer"]
        sched = opt_cfg.get("lr_scheduler")
        # Lightning steps {"scheduler", "interval": "step", "frequency": 1} after every optimizer step (hulc2.py:185-198)
        self.scheduler = sched["scheduler"] if isinstance(sched, dict) else sched
        self._sched_every = int(sched.get("frequency", 1)) if isinstance(sched, dict) else 1
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.reducer: Optional[GradBucketReducer] = None
        if hasattr(self.optimizer, "grad_arenas"):
            self.reducer = GradBucketReducer(self.optimizer, bucket_mb)
        self.device = next(model.parameters()).device
        self.use_graph = bool(use_graph) and hasattr(self.optimizer, "step_counter")
        # device_counters: drive the eager step exactly like the captured body (noise epoch + Adam step number on the
        # device, host noise counter restarted per step) -- what the graph-vs-eager parity tests compare against
        self.device_counters = self.use_graph if device_counters is None else (bool(device_counters) and hasattr(self.optimizer, "step_counter"))
        self.graph_warmup = graph_warmup
        self.steps_done = 0
        self.recaptures = 0
        self._captured_scalars = None
        self._eager_steps = 0
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_batch = None
        self._static_loss = None
        self.launches_per_replay = 0
        self.replays = 0
        # Every step runs on ONE side stream: autograd's AccumulateGrad nodes remember the stream they were created on,
        # and a node created on the legacy default stream cannot take part in a stream capture.
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        if self.device_counters:
            self.optimizer.capturable = True

    # ------------------------------------------------------------------ one step (eager body; also what gets captured)
    def _step_body(self, batch: Dict[str, dict], batch_idx: int) -> torch.Tensor:
        noise.begin_step() if self.device_counters else None
        self.optimizer.zero_grad()
        if self.collate is not None:
            batch = self.collate(batch)
        loss = self.model.training_step(batch, batch_idx)
        if self.reducer is not None:
            self.reducer.prepare()
        loss.backward()
        ops.join_side()                     # (the autograd final callback has joined already; no-op then)
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        if self.device_counters:
            # completed-step counters on the device: Adam bias correction and the noise epoch of the NEXT step
            call("hulc2_counter_add", self.optimizer.step_counter(self.device).data_ptr(), 1)
            call("hulc2_counter_add", noise.epoch_tensor(self.device).data_ptr(), 1)
        return loss.detach()

    def train_step(self, batch: Dict[str, dict], batch_idx: int = 0) -> torch.Tensor:
        """zero_grad -> forward -> backward (+ overlapped all-reduce) -> optimizer step; returns the loss tensor.
        Stream semantics are those of an ordinary call: the work is ordered after the caller's current stream and the
        caller's stream waits for it."""
        if self.stream is None:
            return self._train_step_on_stream(batch, batch_idx)
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream):
            loss = self._train_step_on_stream(batch, batch_idx)
        caller.wait_stream(self.stream)
        return loss

    def _host_scalars(self) -> tuple:
        """Host-side scalars that a capture freezes into kernel arguments: the KL / clip loss weights (``set_kl_beta`` is
        called by the KL-annealing callbacks once per epoch, utils/kl_callbacks.py:19-22) and Adam's betas / eps / weight
        decay / gradient scale.  The learning rate is NOT among them (device-resident, ``FusedAdam.sync_lr``)."""
        m = self.model
        opt = self.optimizer.captured_scalars() if hasattr(self.optimizer, "captured_scalars") else ()
        return (float(getattr(m, "kl_beta", 0.0)), float(getattr(m, "kl_balancing_mix", 0.0)),
                float(getattr(m, "clip_auxiliary_loss_beta", 0.0)), bool(m.training), ops.get_precision(), opt)

    def _after_step(self) -> None:
        self.steps_done += 1
        if self.scheduler is not None and self.steps_done % self._sched_every == 0:
            self.scheduler.step()

    def _train_step_on_stream(self, batch, batch_idx):
        loss = self._train_step_inner(batch, batch_idx)
        self._after_step()
        return loss

    def _train_step_inner(self, batch, batch_idx):
        if not self.use_graph:
            return self._step_body(batch, batch_idx)
        self.optimizer.sync_lr()                        # group["lr"] -> device (a scheduler may have moved it)
        if self._graph is not None and self._captured_scalars != self._host_scalars():
            # a frozen scalar changed (e.g. set_kl_beta at an epoch boundary): capture again over the same static batch
            _zip_copy(self.static_batch, batch)
            batch = self.static_batch
            self._graph = None
            self.recaptures += 1
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

    def profile_replay(self, batch, passes: int = 3) -> List[dict]:
        """Per-call timeline of a REPLAYED step: captures the step body once more with an external timing event on either
        side of every C-ABI call (event-record nodes of the graph, ``_lib.profile_begin_graph``), replays that graph
        ``passes`` times and returns one {key: record} dict per replay.  What bench.py's roofline table is built from: the
        durations are those of the kernels inside the real replay (no host launch latency, no cold-cache artefact)."""
        from . import _lib

        assert self.use_graph and self.static_batch is not None, "profile_replay needs a trainer that has captured its step"
        _zip_copy(self.static_batch, batch)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        _lib.profile_begin_graph()
        try:
            with torch.cuda.graph(g, stream=self.stream):
                self._step_body(self.static_batch, 0)
            for a in self.optimizer._arenas:       # the capture executed nothing (see _capture)
                if a["n"]:
                    a["step"] -= 1
            out = []
            for _ in range(passes):
                g.replay()
                self.optimizer.note_replayed_step()
                out.append(_lib.profile_read_graph())
            self.last_profile_gaps = _lib.profile_read_gaps()      # of the last replay
        finally:
            _lib.profile_end_graph()
        return out

    def _capture(self, batch, batch_idx):
        # The graph reads its inputs from the trainer's OWN static buffers; every later batch is copied into them (a caller
        # that wants zero-copy steps fills / passes ``trainer.static_batch`` itself).  The caller's tensors are never written.
        if self.static_batch is None or not _same_shapes(self.static_batch, batch):
            self.static_batch = tree_map(torch.clone, batch)
        else:
            _zip_copy(self.static_batch, batch)
        batch = self.static_batch
        from ._lib import load_library

        self._captured_scalars = self._host_scalars()
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
    def fit_host(self, host_batches: Iterable[Dict[str, dict]], pre_copy=None) -> List[float]:
        """Trains on an iterable of PINNED host batches and returns every step's loss (read back to the host).

        Software pipeline over three streams: the H2D copy of batch i+1 (copy stream, into one of two device staging
        buffers) runs while the trainer stream computes step i; the loss of step i is copied to a pinned slot right
        behind the step and read by the host one step later, so neither PCIe nor the host read stalls the kernels.
        Every step still moves its own inputs host->device and its own result device->host.

        ``pre_copy(i)``: optional callable run on the copy stream right before batch i's copy -- the hook through which a
        device-resident episode store (``datamodule.DeviceEpisodeStore.write_frames``) ingests the frames that are new in step
        i, so that batches only carry window descriptors (``ops.U8Frames``) instead of 32 frames per window."""
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
        if getattr(self, "_pipe_sync", None) is None:
            # events and the pinned loss slots are made once per trainer: cudaHostAlloc costs milliseconds, a whole step's worth
            self._pipe_sync = ([torch.cuda.Event() for _ in range(2)], [torch.cuda.Event() for _ in range(2)],
                               torch.empty(2, dtype=torch.float32).pin_memory(), [torch.cuda.Event() for _ in range(2)])
        ready, free, loss_slots, loss_done = self._pipe_sync
        losses: List[float] = []

        def upload(i, hb):
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(free[i % 2])          # the step that consumed this staging buffer has read it
                if pre_copy is not None:
                    pre_copy(i)
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


class PolicyValidator:
    """``Hulc2.validation_step`` (hulc2/models/hulc2.py:510-598) driven as Lightning's validation loop does (no_grad, eval
    mode), with the whole step -- encoders, both plan networks, the two decoder passes with action sampling, KL, the MAE /
    gripper-success reductions -- captured in ONE CUDA graph over a static batch and replayed.  Sampling noise comes from
    the Philox kernels keyed by the device epoch counter, bumped once per replay.  ``validate`` returns the step's output
    dict and the logged scalars as device tensors (static buffers, overwritten by the next call)."""

    def __init__(self, model, use_graph: bool = True):
        self.model = model
        self.device = next(model.parameters()).device
        self.use_graph = use_graph
        self.stream = torch.cuda.Stream(device=self.device)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_batch = None
        self._out = None
        self._logged = None
        self.launches_per_replay = 0
        self.replays = 0
        self._keepalive: list = []

    def _body(self, batch, batch_idx):
        noise.begin_step()
        out = self.model.validation_step(batch, batch_idx)
        call("hulc2_counter_add", noise.epoch_tensor(self.device).data_ptr(), 1)
        logged = {k: v for k, v in getattr(self.model, "logged", {}).items() if k.startswith("val") and isinstance(v, torch.Tensor)}
        return out, logged

    def validate(self, batch: Dict[str, dict], batch_idx: int = 0):
        was_training = self.model.training
        self.model.eval()
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)
        try:
            with torch.cuda.stream(self.stream), torch.no_grad():
                if not self.use_graph:
                    res = self._body(batch, batch_idx)
                elif self._graph is None:
                    res = self._body(batch, batch_idx)         # the first call runs eagerly (fills the bf16 weight mirrors) ...
                    torch.cuda.synchronize()
                    self.static_batch = tree_map(torch.clone, batch)   # ... then the same body is recorded over own static inputs
                    from ._lib import load_library

                    g = torch.cuda.CUDAGraph()
                    n0 = load_library().hulc2_launch_count()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._out, self._logged = self._body(self.static_batch, batch_idx)
                    self.launches_per_replay = int(load_library().hulc2_launch_count() - n0)
                    self._graph = g
                    self._keepalive.append((dict(ops._w16), [a[1] for a in ops._arenas]))
                else:
                    _zip_copy(self.static_batch, batch)
                    self._graph.replay()
                    self.replays += 1
                    res = (self._out, self._logged)
        finally:
            self.model.train(was_training)
        caller.wait_stream(self.stream)
        return res

    def refresh(self) -> None:
        """Parameters changed since the capture (training continued): the captured bf16 weight mirrors are stale."""
        ops.invalidate_weight_mirrors()
        self._graph = None
        self._keepalive.clear()
