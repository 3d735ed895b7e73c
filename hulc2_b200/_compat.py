"""Hydra / OmegaConf / Lightning compatibility.

The reference builds every sub-module with ``hydra.utils.instantiate`` on DictConfigs and derives
``Hulc2`` from ``pytorch_lightning.LightningModule`` (hulc2/models/hulc2.py:27,71-99).  When those
packages are installed the real ones are used, so ``training.py`` / ``evaluate_policy.py`` drive the
classes unchanged; when they are absent (this image) minimal stand-ins keep the same call surface.
"""
from __future__ import annotations

import importlib
from typing import Any

import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    from omegaconf import DictConfig, ListConfig, OmegaConf  # type: ignore

    HAVE_OMEGACONF = True
except ImportError:
    HAVE_OMEGACONF = False

    class DictConfig(dict):  # type: ignore
        """attr-dict stand-in: nested dicts are wrapped; empty config is falsy."""

        def __init__(self, *a, **kw):
            super().__init__()
            for k, v in dict(*a, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, DictConfig):
                v = DictConfig(v)
            elif isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
                v = ListConfig(v)
            super().__setitem__(k, v)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    class ListConfig(list):  # type: ignore
        pass

    class OmegaConf:  # type: ignore
        @staticmethod
        def load(path):
            raise FileNotFoundError(path)

        @staticmethod
        def create(obj):
            return DictConfig(obj)


def as_config(cfg: Any):
    """dict -> DictConfig (recursively); DictConfig / None pass through."""
    if cfg is None or isinstance(cfg, DictConfig):
        return cfg
    if isinstance(cfg, dict):
        return OmegaConf.create(cfg) if HAVE_OMEGACONF else DictConfig(cfg)
    return cfg


try:  # pragma: no cover
    from hydra.utils import instantiate as _hydra_instantiate  # type: ignore

    HAVE_HYDRA = True
except ImportError:
    HAVE_HYDRA = False
    _hydra_instantiate = None


def instantiate(cfg, *args, **kwargs):
    """``hydra.utils.instantiate`` (real one when available)."""
    if not cfg:
        return None
    if HAVE_HYDRA and HAVE_OMEGACONF and isinstance(cfg, DictConfig):  # pragma: no cover
        return _hydra_instantiate(cfg, *args, **kwargs)
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    cfg.update(kwargs)
    mod, _, name = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    return fn(*args, **cfg)


try:  # pragma: no cover
    import pytorch_lightning as pl  # type: ignore
    from pytorch_lightning.utilities import rank_zero_info, rank_zero_only  # type: ignore

    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except ImportError:
    HAVE_LIGHTNING = False

    def rank_zero_only(fn):
        return fn

    def rank_zero_info(*a, **k):
        pass

    class LightningModule(nn.Module):  # type: ignore
        """Just enough of pl.LightningModule for the policy step: device, log, save_hyperparameters."""

        def __init__(self, *a, **kw):
            super().__init__()
            self.logged = {}
            self.trainer = None
            self.current_epoch = 0
            self.global_step = 0
            self.hparams = {}

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, **kw):
            self.logged[name] = value

        def save_hyperparameters(self, *a, **kw):
            """Records the constructor arguments of the calling ``__init__`` in ``self.hparams`` (what Lightning stores under
            the checkpoint key ``hyper_parameters`` and feeds back to ``cls(**hparams)`` in ``load_from_checkpoint``)."""
            import inspect

            frame = inspect.currentframe().f_back
            try:
                while frame is not None and frame.f_code.co_name != "__init__":
                    frame = frame.f_back
                if frame is None:
                    return
                info = inspect.getargvalues(frame)
                self.hparams = {k: info.locals[k] for k in info.args if k not in ("self", "__class__")}
            finally:
                del frame

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        def unfreeze(self):
            for p in self.parameters():
                p.requires_grad = True
            self.train()
