"""Native train-step driver: what Lightning's loop does around ``Hulc2.training_step`` (SURVEY.md 3.1:
AMP-less backward -> bucketed gradient all-reduce -> Adam.step), without Lightning.  Used by bench.py and
usable as a minimal trainer; under real Lightning the module is driven by ``Trainer.fit`` instead.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .ddp import GradBucketReducer
from .synthetic import tree_map


class PolicyTrainer:
    def __init__(self, model, bucket_mb: float = 25.0):
        self.model = model
        self.optimizer = model.configure_optimizers()["optimizer"]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.reducer: Optional[GradBucketReducer] = None
        if hasattr(self.optimizer, "grad_arenas"):
            self.reducer = GradBucketReducer(self.optimizer, bucket_mb)
        self._dev_batch = None

    def train_step(self, batch: Dict[str, dict], batch_idx: int = 0) -> torch.Tensor:
        """zero_grad -> forward -> backward (+ overlapped all-reduce) -> optimizer step; returns the loss tensor."""
        self.optimizer.zero_grad()
        loss = self.model.training_step(batch, batch_idx)
        if self.reducer is not None:
            self.reducer.prepare()
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return loss.detach()

    def train_step_from_host(self, host_batch: Dict[str, dict], batch_idx: int = 0) -> float:
        """End-to-end step: pinned host batch -> device (inside the call) -> step -> loss read back to host."""
        dev = next(self.model.parameters()).device
        batch = tree_map(lambda t: t.to(dev, non_blocking=True), host_batch)
        loss = self.train_step(batch, batch_idx)
        return float(loss)  # device->host read of the step's result
