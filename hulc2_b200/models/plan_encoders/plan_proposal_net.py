"""PlanProposalNetwork (mirror of hulc2/models/plan_encoders/plan_proposal_net.py:8-47):
cat[emb_0, goal] -> 4 x (Linear 2048 + ReLU) -> fc_state -> DiscState(logit)."""
import torch
import torch.nn as nn

from ... import ops
from ...utils.distributions import Distribution, State


class PlanProposalNetwork(nn.Module):
    def __init__(
        self,
        perceptual_features: int,
        latent_goal_features: int,
        plan_features: int,
        activation_function: str,
        hidden_size: int,
        dist: Distribution,
    ):
        super().__init__()
        if activation_function != "ReLU":
            raise NotImplementedError("the CUDA path fuses ReLU epilogues; conf default is activation_function: ReLU")
        self.perceptual_features = perceptual_features
        self.latent_goal_features = latent_goal_features
        self.plan_features = plan_features
        self.hidden_size = hidden_size
        self.in_features = self.perceptual_features + self.latent_goal_features
        self.act_fn = getattr(nn, activation_function)()
        self.dist = dist
        self.fc_model = nn.Sequential(
            nn.Linear(in_features=self.in_features, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=hidden_size),
            self.act_fn,
        )
        self.fc_state = self.dist.build_state(self.hidden_size, self.plan_features)

    def forward(self, initial_percep_emb: torch.Tensor, latent_goal: torch.Tensor) -> State:
        x = ops.concat_cols(initial_percep_emb, latent_goal)
        layers = [(self.fc_model[i].weight, self.fc_model[i].bias) for i in (0, 2, 4, 6)]
        layers.append((self.fc_state[0].weight, self.fc_state[0].bias))
        my_state = ops.mlp(x, layers, [True, True, True, True, False])
        return self.dist.forward_dist(my_state)
