"""PlanRecognitionTransformersNetwork (mirror of hulc2/models/plan_encoders/plan_recognition_net.py:77-148).

The ``nn.TransformerEncoder`` only owns the parameters (``transformer_encoder.layers.{i}.self_attn.in_proj_weight``
etc.); ``forward`` runs the post-LN encoder layers on the CUDA kernels in batch-major ``[B,S,E]`` order
(row order is irrelevant to the per-token ops; attention groups by window).  ``fc`` is applied to the
sequence mean -- ``mean_s(fc(x_s)) == fc(mean_s x_s)`` -- which is what ``x = torch.mean(self.fc(x), dim=1)``
computes with 32x fewer FLOPs.  The BiLSTM/BiRNN alternates of the reference are non-default and not mirrored.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ... import noise, ops
from ...utils.distributions import Distribution, State


class PlanRecognitionTransformersNetwork(nn.Module):
    def __init__(
        self,
        num_heads: int,
        num_layers: int,
        encoder_hidden_size: int,
        fc_hidden_size: int,
        plan_features: int,
        in_features: int,
        action_space: int,
        encoder_normalize: bool,
        positional_normalize: bool,
        position_embedding: bool,
        max_position_embeddings: int,
        dropout_p: bool,
        dist: Distribution,
    ):
        super().__init__()
        self.in_features = in_features
        self.plan_features = plan_features
        self.action_space = action_space
        self.padding = False
        self.dist = dist
        self.hidden_size = fc_hidden_size
        self.position_embedding = position_embedding
        self.encoder_normalize = encoder_normalize
        self.positional_normalize = positional_normalize
        self.num_heads, self.num_layers, self.dropout_p = num_heads, num_layers, float(dropout_p)
        if self.in_features % num_heads != 0:
            raise NotImplementedError("feature padding to a multiple of num_heads is never triggered (128 % 8 == 0)")
        if not position_embedding or encoder_normalize or positional_normalize:
            raise NotImplementedError("only position_embedding=true, encoder/positional_normalize=false (transformers.yaml)")
        self.position_embeddings = nn.Embedding(max_position_embeddings, self.in_features)
        encoder_layer = nn.TransformerEncoderLayer(
            self.in_features, num_heads, dim_feedforward=encoder_hidden_size, dropout=dropout_p
        )
        self.layernorm = nn.LayerNorm(self.in_features)
        self.dropout = nn.Dropout(p=dropout_p)
        self.transformer_encoder = nn.TransformerEncoder(
            encoder_layer, num_layers=num_layers, norm=None, enable_nested_tensor=False
        )
        self.fc = nn.Linear(in_features=self.in_features, out_features=fc_hidden_size)
        self.fc_state = self.dist.build_state(fc_hidden_size, self.plan_features)

    def _keep(self, shape, device):
        if not (self.training and self.dropout_p > 0.0):
            return None
        return noise.keep_mask(shape, self.dropout_p, device)

    def forward(self, perceptual_emb: torch.Tensor) -> Tuple[State, torch.Tensor]:
        B, S, E = perceptual_emb.shape
        H = self.num_heads
        dev = perceptual_emb.device
        scale = 1.0 / (1.0 - self.dropout_p) if self.dropout_p < 1.0 else 0.0
        x = ops.AddPosFunction.apply(perceptual_emb, self.position_embeddings.weight, self._keep((B, S, E), dev), scale)
        x = x.view(B * S, E)
        for layer in self.transformer_encoder.layers:
            sa = layer.self_attn
            qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias)
            att = ops.AttentionFunction.apply(qkv, B, S, H, self._keep((B, H, S, S), dev), scale)
            att = ops.linear(att, sa.out_proj.weight, sa.out_proj.bias)
            x = ops.layer_norm(x, layer.norm1.weight, layer.norm1.bias, res=att, keep=self._keep((B * S, E), dev),
                               keep_scale=scale, eps=layer.norm1.eps)
            ff_keep = self._keep((B * S, layer.linear1.out_features), dev)
            ff = ops.mlp(x, [(layer.linear1.weight, layer.linear1.bias), (layer.linear2.weight, layer.linear2.bias)],
                         [True, False], keeps=(ff_keep, None), keep_scale=scale)
            x = ops.layer_norm(x, layer.norm2.weight, layer.norm2.bias, res=ff, keep=self._keep((B * S, E), dev),
                               keep_scale=scale, eps=layer.norm2.eps)
        xm = ops.MeanSeqFunction.apply(x.view(B, S, E))
        seq_feat = ops.linear(xm, self.fc.weight, self.fc.bias, fp32=ops.clip_fp32 and ops.clip_fp32_level >= 2)
        my_state = ops.linear(seq_feat, self.fc_state[0].weight, self.fc_state[0].bias)
        return self.dist.forward_dist(my_state), seq_feat
