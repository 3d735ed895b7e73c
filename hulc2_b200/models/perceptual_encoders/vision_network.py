"""Static-camera encoder (mirror of hulc2/models/perceptual_encoders/vision_network.py:11-108).

The torch ``nn`` sub-modules only own the parameters under the reference's state_dict names
(``conv_model.{0,2,4}``, ``fc1.0``, ``fc2``, ``ln``, ``spatial_softmax.{x_map,y_map,temperature}``);
``forward`` runs the CUDA kernels: implicit-GEMM conv trunk + SpatialSoftmax, fc1/fc2 MLP, LayerNorm.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from ... import ops


class VisionNetwork(nn.Module):
    def __init__(
        self,
        input_width: int,
        input_height: int,
        activation_function: str,
        dropout_vis_fc: float,
        l2_normalize_output: bool,
        visual_features: int,
        num_c: int,
        use_sinusoid: bool,
        spatial_softmax_temp: float,
    ):
        super().__init__()
        if activation_function != "ReLU":
            raise NotImplementedError("the CUDA path fuses ReLU epilogues; conf default is activation_function: ReLU")
        if use_sinusoid or l2_normalize_output or dropout_vis_fc != 0.0:
            raise NotImplementedError("use_sinusoid / l2_normalize_output / dropout_vis_fc are off in every shipped config")
        self.l2_normalize_output = l2_normalize_output
        self.act_fn = getattr(nn, activation_function)()
        w, h = self.calc_out_size(input_width, input_height, 8, 0, 4)
        w, h = self.calc_out_size(w, h, 4, 0, 2)
        w, h = self.calc_out_size(w, h, 3, 0, 1)
        self.use_sinusoid = use_sinusoid
        temp = None if not isinstance(spatial_softmax_temp, float) else spatial_softmax_temp
        self.spatial_softmax = SpatialSoftmax(num_rows=w, num_cols=h, temperature=temp)
        self.conv_model = nn.Sequential(
            nn.Conv2d(in_channels=num_c, out_channels=32, kernel_size=8, stride=4),
            self.act_fn,
            nn.Conv2d(in_channels=32, out_channels=64, kernel_size=4, stride=2),
            self.act_fn,
            nn.Conv2d(in_channels=64, out_channels=64, kernel_size=3, stride=1),
            self.act_fn,
        )
        self.fc1 = nn.Sequential(nn.Linear(in_features=128, out_features=512), self.act_fn, nn.Dropout(dropout_vis_fc))
        self.fc2 = nn.Linear(in_features=512, out_features=visual_features)
        self.ln = nn.LayerNorm(visual_features)

    def features(self, x: torch.Tensor) -> torch.Tensor:
        """Everything up to (not including) the final LayerNorm: [N, visual_features]."""
        c = self.conv_model
        ss = self.spatial_softmax
        kp = ops.StaticConvSSM.apply(
            x, c[0].weight, c[0].bias, c[2].weight, c[2].bias, c[4].weight, c[4].bias, ss.x_map, ss.y_map, ss.temperature
        )
        return ops.mlp(kp, [(self.fc1[0].weight, self.fc1[0].bias), (self.fc2.weight, self.fc2.bias)], [True, False])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(self.features(x), self.ln.weight, self.ln.bias, eps=self.ln.eps)

    @staticmethod
    def calc_out_size(w: int, h: int, kernel_size: int, padding: int, stride: int) -> Tuple[int, int]:
        width = (w - kernel_size + 2 * padding) // stride + 1
        height = (h - kernel_size + 2 * padding) // stride + 1
        return width, height


class SpatialSoftmax(nn.Module):
    """Parameter/buffer holder + standalone NCHW forward (vision_network.py:68-108)."""

    def __init__(self, num_rows: int, num_cols: int, temperature: Optional[float] = None):
        super().__init__()
        self.num_rows = num_rows
        self.num_cols = num_cols
        grid_x, grid_y = torch.meshgrid(
            torch.linspace(-1.0, 1.0, num_cols), torch.linspace(-1.0, 1.0, num_rows), indexing="ij"
        )
        self.register_buffer("x_map", grid_x.reshape(-1))
        self.register_buffer("y_map", grid_y.reshape(-1))
        if temperature:
            self.register_buffer("temperature", torch.ones(1) * temperature)
        else:
            self.temperature = Parameter(torch.ones(1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N,C,H,W] (the reference's layout) -> [N, 2C]; inference helper, no autograd."""
        n, c, h, w = x.shape
        from ..._lib import call

        x = x.detach().contiguous()
        nhwc = torch.empty(n, h * w, c, device=x.device, dtype=torch.float32)
        call("hulc2_nchw_to_nhwc", x.data_ptr(), nhwc.data_ptr(), n, h * w, c, None)
        out = torch.empty(n, 2 * c, device=x.device, dtype=torch.float32)
        call("hulc2_spatial_softmax_fwd", nhwc.data_ptr(), self.x_map.data_ptr(), self.y_map.data_ptr(),
             self.temperature.data_ptr(), out.data_ptr(), n, h * w, c)
        self.coords = out
        return out
