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


class GaussPlanDist:
    """Independent(Normal(mean, std), 1) restricted to what the policy uses (distributions.py:28-29)."""

    def __init__(self, mean: torch.Tensor, std: torch.Tensor):
        self.mean, self.stddev = mean, std

    def rsample(self, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        if eps is None:
            eps = noise.normal(self.mean.shape, self.mean.device)
        return ops.GaussRSampleFunction.apply(self.mean, self.stddev, eps)

    def sample(self, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        with torch.no_grad():
            return self.rsample(eps)


class Distribution:
    def __init__(self, **kwargs):
        self.dist = kwargs.get("dist")
        assert self.dist == "discrete" or self.dist == "continuous"
        if self.dist == "discrete":
            self.category_size = kwargs.get("category_size")
            self.class_size = kwargs.get("class_size")

    def get_dist(self, state):
        if self.dist == "discrete":
            return PlanDist(state.logit, self.category_size, self.class_size)
        return GaussPlanDist(state.mean, state.std)

    def detach_state(self, state):
        if self.dist == "discrete":
            return DiscState(state.logit.detach())
        return ContState(state.mean.detach(), state.std.detach())

    def sample_latent_plan(self, distribution) -> torch.Tensor:
        sampled_plan = distribution.sample()
        if self.dist == "discrete":
            sampled_plan = torch.flatten(sampled_plan, start_dim=-2, end_dim=-1)
        return sampled_plan

    def build_state(self, hidden_size, plan_features):
        if self.dist == "discrete":
            return nn.Sequential(nn.Linear(hidden_size, plan_features))
        return nn.Sequential(nn.Linear(hidden_size, 2 * plan_features))

    def forward_dist(self, x):
        if self.dist == "discrete":
            return DiscState(x)
        mean, std = ops.GaussStateFunction.apply(x)
        return ContState(mean, std)
