"""Goal encoders (mirror of hulc2/models/encoders/goal_encoders.py:8-71): 3-layer MLP + LayerNorm -> 32-d."""
import torch
import torch.nn as nn

from ... import ops


class VisualGoalEncoder(nn.Module):
    def __init__(
        self,
        hidden_size: int,
        latent_goal_features: int,
        in_features: int,
        l2_normalize_goal_embeddings: bool,
        activation_function: str,
    ):
        super().__init__()
        if activation_function != "ReLU" or l2_normalize_goal_embeddings:
            raise NotImplementedError("conf defaults: activation_function ReLU, l2_normalize_goal_embeddings False")
        self.l2_normalize_output = l2_normalize_goal_embeddings
        self.act_fn = getattr(nn, activation_function)()
        self.mlp = nn.Sequential(
            nn.Linear(in_features=in_features, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=latent_goal_features),
        )
        self.ln = nn.LayerNorm(latent_goal_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.mlp
        x = ops.mlp(x, [(m[0].weight, m[0].bias), (m[2].weight, m[2].bias), (m[4].weight, m[4].bias)], [True, True, False])
        return ops.layer_norm(x, self.ln.weight, self.ln.bias, eps=self.ln.eps)


class LanguageGoalEncoder(nn.Module):
    def __init__(
        self,
        lang_net,
        in_features: int,
        hidden_size: int,
        latent_goal_features: int,
        l2_normalize_goal_embeddings: bool,
        word_dropout_p: float,
        activation_function: str,
    ):
        super().__init__()
        if activation_function != "ReLU" or l2_normalize_goal_embeddings or word_dropout_p != 0.0:
            raise NotImplementedError("conf defaults: ReLU, l2_normalize_goal_embeddings False, word_dropout_p 0.0")
        self.lang_net = lang_net
        self.l2_normalize_output = l2_normalize_goal_embeddings
        self.act_fn = getattr(nn, activation_function)()
        self.mlp = nn.Sequential(
            nn.Dropout(word_dropout_p),
            nn.Linear(in_features=in_features, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=hidden_size),
            self.act_fn,
            nn.Linear(in_features=hidden_size, out_features=latent_goal_features),
        )
        self.ln = nn.LayerNorm(latent_goal_features)

    def forward(self, x) -> torch.Tensor:
        # a frozen sentence encoder (SBERT is run under no_grad and detached in the reference,
        # sbert_lang_encoder.py:51-54) may be injected as lang_net; otherwise x is the [B,384] embedding
        if self.lang_net is not None:
            x = self.lang_net(x)
        m = self.mlp
        x = ops.mlp(x, [(m[1].weight, m[1].bias), (m[3].weight, m[3].bias), (m[5].weight, m[5].bias)], [True, True, False])
        return ops.layer_norm(x, self.ln.weight, self.ln.bias, eps=self.ln.eps)
