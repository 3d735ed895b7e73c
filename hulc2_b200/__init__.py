"""hulc2_b200 -- B200-native HULC++ low-level policy step.

Mirror of the reference's model surface (``hulc2.models.*`` / ``hulc2.utils.distributions``) whose math
runs in hand-written CUDA kernels for sm_100a behind the C-ABI in ``include/hulc2_b200.h``.

Switching a reference config over is a ``_target_`` change only, e.g.
``hulc2.models.hulc2.Hulc2`` -> ``hulc2_b200.models.hulc2.Hulc2``; :func:`install_as_hulc2` registers the
mirror modules under the reference's own import paths so unmodified configs resolve to them.

Importing this package never touches the GPU; the CUDA library is loaded on the first op call and its
absence is a hard error (no CPU / PyTorch fallback).
"""
from __future__ import annotations

import importlib
import sys
import types

__version__ = "0.1.0"

_MIRRORED = [
    "models.hulc2",
    "models.perceptual_encoders.concat_encoders",
    "models.perceptual_encoders.vision_network",
    "models.perceptual_encoders.vision_network_gripper",
    "models.plan_encoders.plan_proposal_net",
    "models.plan_encoders.plan_recognition_net",
    "models.encoders.goal_encoders",
    "models.auxiliary_loss_networks.proj_vis_lang",
    "models.decoders.action_decoder",
    "models.decoders.logistic_decoder_rnn",
    "models.decoders.utils.rnn",
    "models.decoders.utils.gripper_control",
    "utils.distributions",
]


def install_as_hulc2(force: bool = False) -> None:
    """Registers the mirror modules as ``hulc2.<path>`` in ``sys.modules`` so the reference's Hydra
    ``_target_`` strings (conf/model/**.yaml) instantiate the CUDA-backed classes."""
    def ensure_pkg(name):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []  # type: ignore[attr-defined]
            sys.modules[name] = m
        return sys.modules[name]

    for path in _MIRRORED:
        parts = ("hulc2." + path).split(".")
        for i in range(1, len(parts)):
            ensure_pkg(".".join(parts[:i]))
        target = "hulc2." + path
        if target in sys.modules and not force:
            continue
        mod = importlib.import_module("hulc2_b200." + path)
        sys.modules[target] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)


def set_precision(p: str) -> None:
    from . import ops

    ops.set_precision(p)
