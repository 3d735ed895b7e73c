"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Imports ``hulc2.models.hulc2.Hulc2`` and friends straight from ``/root/reference``
with the five import shims SURVEY.md section 8c lists (omegaconf, hydra,
pytorch_lightning, pytorch3d.transforms, and a pre-registered empty
``hulc2.models`` package so ``hulc2/models/__init__.py:2-11`` -- which pulls
CLIP / SBERT / R3M -- is skipped).

``/root/reference`` only exists in the build container, never on the GPU box:
this module is used by ``tests/golden/make_golden.py`` (fixture generation) and
by the CPU tests that pin ``oracle/hulc2_oracle.py`` against the real reference
(those tests skip when the directory is absent).  Nothing in the product
package may import it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("HULC2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "hulc2", "models", "hulc2.py"))


# ----------------------------------------------------------------------------- omegaconf shim
class DictConfig(dict):
    """attr-dict; nested dicts are converted on construction/set; falsy when empty."""

    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, DictConfig):
            return DictConfig(v)
        if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
            return ListConfig(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ListConfig(list):
    pass


class _OmegaConf:
    @staticmethod
    def load(path):
        raise FileNotFoundError(path)

    @staticmethod
    def create(obj):
        return DictConfig(obj)


def _instantiate(cfg, *args, **kwargs):
    """hydra.utils.instantiate: resolve ``_target_``, drop ``_recursive_``, pass kwargs."""
    if not cfg:
        return None
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    cfg.update(kwargs)
    mod, _, name = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    return fn(*args, **cfg)


class _LightningModule(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        self.logged = {}
        self.trainer = None
        self.current_epoch = 0

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, name, value, **kw):
        self.logged[name] = value

    def save_hyperparameters(self, *a, **kw):
        pass


# ----------------------------------------------------------------------------- pytorch3d shim
# pytorch3d is an unpinned third-party dependency (requirements.txt:20) absent from
# /root/reference.  Restated from its published algorithm (pytorch3d/transforms/
# rotation_conversions.py): R = Rx(a) Ry(b) Rz(c) for convention "XYZ";
# inverse (atan2(-M12, M22), asin(M02), atan2(-M01, M00)).
def _axis_rot(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == "Y":
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    else:
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    assert convention == "XYZ"
    ms = [_axis_rot(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return torch.matmul(torch.matmul(ms[0], ms[1]), ms[2])


def matrix_to_euler_angles(matrix, convention):
    assert convention == "XYZ"
    a = torch.atan2(-matrix[..., 1, 2], matrix[..., 2, 2])
    b = torch.asin(matrix[..., 0, 2])
    c = torch.atan2(-matrix[..., 0, 1], matrix[..., 0, 0])
    return torch.stack((a, b, c), -1)


def _quat(name):  # NaN-fallback path only (gripper_control.py:51-55): pytorch3d's algorithm as restated in the oracle
    def fn(x):
        from oracle import hulc2_oracle as O

        return getattr(O, name)(x)

    return fn


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    def _mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import omegaconf  # noqa: F401
    except ImportError:
        _mod("omegaconf", DictConfig=DictConfig, ListConfig=ListConfig, OmegaConf=_OmegaConf)
    try:
        import hydra  # noqa: F401
    except ImportError:
        hu = _mod("hydra.utils", instantiate=_instantiate)
        _mod("hydra", utils=hu)
    try:
        import pytorch_lightning  # noqa: F401
    except ImportError:
        plu = _mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f, rank_zero_info=lambda *a, **k: None)
        _mod("pytorch_lightning", LightningModule=_LightningModule, utilities=plu)
    try:
        import pytorch3d.transforms  # noqa: F401
    except ImportError:
        p3t = _mod(
            "pytorch3d.transforms",
            euler_angles_to_matrix=euler_angles_to_matrix,
            matrix_to_euler_angles=matrix_to_euler_angles,
            matrix_to_quaternion=_quat("matrix_to_quaternion"),
            quaternion_to_matrix=_quat("quaternion_to_matrix"),
        )
        _mod("pytorch3d", transforms=p3t)

    import hulc2  # noqa: F401  (top-level __init__ is metadata only)

    if "hulc2.models" not in sys.modules:
        pkg = types.ModuleType("hulc2.models")
        pkg.__path__ = [os.path.join(REF_ROOT, "hulc2", "models")]
        sys.modules["hulc2.models"] = pkg
    _installed = True


def load_reference_hulc2():
    """Returns the unmodified reference ``Hulc2`` class."""
    install_shims()
    from hulc2.models.hulc2 import Hulc2

    return Hulc2


def make_reference_model(cfg: dict):
    """Instantiates the reference Hulc2 from a plain-dict config (see oracle/configs.py)."""
    Hulc2 = load_reference_hulc2()
    import copy

    cfg = copy.deepcopy(cfg)
    cfg.pop("_target_", None)
    cfg.pop("_recursive_", None)
    om = sys.modules["omegaconf"]
    wrapped = {k: (om.DictConfig(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    return Hulc2(**wrapped)
