"""Gripper-camera encoder (mirror of hulc2/models/perceptual_encoders/vision_network_gripper.py:11-95).

Only ``conv_encoder: nature_cnn`` is selectable (the only value any shipped config uses); parameters
live under ``conv_model.{0,2,4,7}``, ``fc1.0``, ``fc2``, ``ln``.
"""
from typing import Tuple

import torch
import torch.nn as nn

from ... import ops


def nature_cnn(act_fn, num_c):
    return nn.Sequential(
        nn.Conv2d(num_c, 32, 8, stride=4),
        act_fn,
        nn.Conv2d(32, 64, 4, stride=2),
        act_fn,
        nn.Conv2d(64, 64, 3, stride=1),
        act_fn,
        nn.Flatten(start_dim=1),
        nn.Linear(64 * 7 * 7, 128),
        act_fn,
    )


_CONV_ENCODERS = {"nature_cnn": nature_cnn}


class VisionNetwork(nn.Module):
    def __init__(
        self,
        input_width: int,
        input_height: int,
        conv_encoder: str,
        activation_function: str,
        dropout_vis_fc: float,
        l2_normalize_output: bool,
        visual_features: int,
        num_c: int,
    ):
        super().__init__()
        if activation_function != "ReLU":
            raise NotImplementedError("the CUDA path fuses ReLU epilogues; conf default is activation_function: ReLU")
        if l2_normalize_output or dropout_vis_fc != 0.0:
            raise NotImplementedError("l2_normalize_output / dropout_vis_fc are off in every shipped config")
        if conv_encoder not in _CONV_ENCODERS:
            raise NotImplementedError(f"conv_encoder={conv_encoder!r}: only nature_cnn is selected by any config")
        self.l2_normalize_output = l2_normalize_output
        self.act_fn = getattr(nn, activation_function)()
        self.conv_model = _CONV_ENCODERS[conv_encoder](self.act_fn, num_c)
        self.fc1 = nn.Sequential(nn.Linear(in_features=128, out_features=512), self.act_fn, nn.Dropout(dropout_vis_fc))
        self.fc2 = nn.Linear(in_features=512, out_features=visual_features)
        self.ln = nn.LayerNorm(visual_features)

    def features(self, x: torch.Tensor) -> torch.Tensor:
        c = self.conv_model
        flat = ops.GripperConvFlatten.apply(x, c[0].weight, c[0].bias, c[2].weight, c[2].bias, c[4].weight, c[4].bias)
        return ops.mlp(
            flat,
            [(c[7].weight, c[7].bias), (self.fc1[0].weight, self.fc1[0].bias), (self.fc2.weight, self.fc2.bias)],
            [True, True, False],
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(self.features(x), self.ln.weight, self.ln.bias, eps=self.ln.eps)

    @staticmethod
    def calc_out_size(w: int, h: int, kernel_size: int, padding: int, stride: int) -> Tuple[int, int]:
        width = (w - kernel_size + 2 * padding) // stride + 1
        height = (h - kernel_size + 2 * padding) // stride + 1
        return width, height
