"""world <-> tcp frame transforms (mirror of hulc2/models/decoders/utils/gripper_control.py:16-63).

Closed-form on the GPU: R = Rx Ry Rz (pytorch3d "XYZ"), the 3x3 inverse of a rotation is its transpose
(no ``torch.inverse`` / LU), no ``assert not isnan`` host sync.  Computed in double internally; forced
fp32 in/out like the reference's ``autocast(dtype=torch.float32)`` block.  No autograd: the reference
only applies them to ground-truth actions (inside the loss) and to sampled actions.
"""
import torch

from .... import ops


def world_to_tcp_frame(action: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    return ops.world_to_tcp(action.detach().float(), robot_obs.detach().float())


def tcp_to_world_frame(action: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    b, s, _ = action.shape
    return ops.tcp_to_world(action.detach().float(), robot_obs.detach().float().reshape(b, s, -1))
