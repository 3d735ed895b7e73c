"""ProjVisLang (mirror of hulc2/models/auxiliary_loss_networks/proj_vis_lang.py:7-27)."""
from typing import Tuple

import torch
import torch.nn as nn

from ... import ops


class ProjVisLang(nn.Module):
    def __init__(self, im_dim: int, lang_dim: int, output_dim: int, proj_lang: bool = True):
        super().__init__()
        self.mlp_im = nn.Sequential(
            nn.Linear(in_features=im_dim, out_features=128), nn.ReLU(), nn.Linear(in_features=128, out_features=output_dim)
        )
        self.mlp_lang = None
        if proj_lang:
            self.mlp_lang = nn.Sequential(
                nn.Linear(in_features=lang_dim, out_features=128), nn.ReLU(), nn.Linear(in_features=128, out_features=output_dim)
            )

    def forward(self, vis_emb: torch.Tensor, lang_emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        m = self.mlp_im
        vis_emb = ops.mlp(vis_emb, [(m[0].weight, m[0].bias), (m[2].weight, m[2].bias)], [True, False], fp32=ops.clip_fp32)
        if self.mlp_lang is not None:
            m = self.mlp_lang
            lang_emb = ops.mlp(lang_emb, [(m[0].weight, m[0].bias), (m[2].weight, m[2].bias)], [True, False], fp32=ops.clip_fp32)
        return vis_emb, lang_emb
