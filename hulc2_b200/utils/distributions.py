"""Latent-plan distribution (mirror of hulc2/utils/distributions.py:15-60).

Same ``Distribution(**kwargs)`` surface (``get_dist``, ``detach_state``, ``sample_latent_plan``,
``build_state``, ``forward_dist``).  ``get_dist`` returns a light ``PlanDist`` instead of a
``torch.distributions`` object: its ``rsample``/``sample`` run the library's one-hot /
straight-through kernels with indices drawn by :mod:`hulc2_b200.noise`.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional, Union

import torch
import torch.nn as nn

from .. import noise, ops

DiscState = namedtuple("DiscState", ["logit"])
ContState = namedtuple("ContState", ["mean", "std"])
State = Union[DiscState, ContState]


class PlanDist:
    """Independent(OneHotCategoricalStraightThrough(logits=[B,cat,cls]), 1) restricted to what the policy uses."""

    def __init__(self, logits: torch.Tensor, category_size: int, class_size: int):
        self.logits = logits  # [B, cat*cls] unnormalised
        self.category_size, self.class_size = category_size, class_size

    def rsample(self, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """one-hot(sample) + (probs - probs.detach()); returns [B, cat, cls] like torch."""
        if idx is None:
            idx = noise.categories(self.logits, self.category_size, self.class_size)
        plan = ops.PlanRSampleFunction.apply(self.logits, idx, self.category_size, self.class_size)
        return plan.view(-1, self.category_size, self.class_size)

    def sample(self, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        if idx is None:
            idx = noise.categories(self.logits, self.category_size, self.class_size)
        return ops.onehot(idx, self.category_size, self.class_size).view(-1, self.category_size, self.class_size)


class Distribution:
    def __init__(self, **kwargs):
        self.dist = kwargs.get("dist")
        assert self.dist == "discrete" or self.dist == "continuous"
        if self.dist == "discrete":
            self.category_size = kwargs.get("category_size")
            self.class_size = kwargs.get("class_size")
        else:
            raise NotImplementedError(
                "continuous latent plans (conf/model/distribution/continuous.yaml) are a SURVEY 8f 'next' row"
            )

    def get_dist(self, state):
        return PlanDist(state.logit, self.category_size, self.class_size)

    def detach_state(self, state):
        return DiscState(state.logit.detach())

    def sample_latent_plan(self, distribution: PlanDist) -> torch.Tensor:
        return torch.flatten(distribution.sample(), start_dim=-2, end_dim=-1)

    def build_state(self, hidden_size, plan_features):
        return nn.Sequential(nn.Linear(hidden_size, plan_features))

    def forward_dist(self, x):
        return DiscState(x)
