"""Lightning-format checkpoints of the policy (hulc2/evaluation/manager_lmp.py:91-109, hulc2/utils/utils.py:36-45).

The reference saves through ``pytorch_lightning.callbacks.ModelCheckpoint``: a pickled dict with ``state_dict`` (the
names of SURVEY.md 8b), ``hyper_parameters`` (the constructor arguments, DictConfigs with ``_target_`` strings under
``hulc2.models.*``), ``optimizer_states``, ``epoch`` and ``global_step``.  This module reads and writes that layout
without Lightning, and re-targets the reference's class paths at this package's mirror classes.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Optional

import torch

from .. import _MIRRORED
from .._compat import HAVE_OMEGACONF, as_config

_REF_PREFIXES = tuple("hulc2." + m + "." for m in _MIRRORED)


def retarget(node: Any) -> Any:
    """Recursively rewrites ``_target_: hulc2.<mirrored module>.<Class>`` to ``hulc2_b200.<...>`` (other targets --
    torch.optim.*, transformers.* -- stay) and returns DictConfigs for mappings."""
    if isinstance(node, dict):
        out = {}
        for k, v in node.items():
            if k == "_target_" and isinstance(v, str) and v.startswith(_REF_PREFIXES):
                v = "hulc2_b200." + v[len("hulc2."):]
            out[k] = retarget(v)
        return as_config(out)
    if isinstance(node, (list, tuple)):
        return type(node)(retarget(v) for v in node) if type(node) in (list, tuple) else [retarget(v) for v in node]
    return node


def to_plain(node: Any) -> Any:
    """DictConfig / ListConfig (real or stand-in) -> plain dict / list, so a checkpoint unpickles without omegaconf."""
    if HAVE_OMEGACONF:  # pragma: no cover - not installed in the build image
        from omegaconf import OmegaConf  # type: ignore

        if OmegaConf.is_config(node):
            return OmegaConf.to_container(node, resolve=True)
    if isinstance(node, dict):
        return {k: to_plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [to_plain(v) for v in node]
    return node


def read_checkpoint(path, map_location=None) -> dict:
    ckpt = torch.load(str(path), map_location=map_location or "cpu", weights_only=False)
    if "state_dict" not in ckpt:
        raise KeyError(f"{path}: not a Lightning-format checkpoint (no 'state_dict')")
    return ckpt


def save_checkpoint(model, path, optimizer=None, epoch: int = 0, global_step: int = 0) -> None:
    """Writes ``model`` (and optionally its optimizer state) in the layout ``Hulc2.load_from_checkpoint`` -- the
    reference's or this package's -- reads."""
    ckpt = {
        "epoch": int(epoch), "global_step": int(global_step), "pytorch-lightning_version": "1.8.6",
        "state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
        "hyper_parameters": to_plain(dict(getattr(model, "hparams", {}) or {})),
    }
    if optimizer is not None:
        ckpt["optimizer_states"] = [optimizer.state_dict()]
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save(ckpt, str(path))


def initialize_pretrained_weights(model, cfg) -> None:
    """hulc2/utils/utils.py:36-45: warm-start from ``cfg.pretrain_chk``; the checkpoint's position embeddings are cut to this
    model's window (``plan_recognition.position_embeddings.weight[:max_position_embeddings]``), ``pretrain_exclude_pr``
    drops the plan-recognition weights, and the rest is loaded non-strictly."""
    get = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    ckpt = read_checkpoint(get("pretrain_chk"))
    sd = dict(ckpt["state_dict"])
    batch_size = model.plan_recognition.position_embeddings.weight.shape[0]
    weight = "plan_recognition.position_embeddings.weight"
    sd[weight] = sd[weight][:batch_size]
    if get("pretrain_exclude_pr", False):
        for key in list(sd.keys()):
            if key.startswith("plan_recognition"):
                del sd[key]
    model.load_state_dict(sd, strict=False)
