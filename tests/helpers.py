"""Shared test helpers: golden fixtures, model/oracle construction, comparisons."""
import os
import warnings

import numpy as np
import torch

from hulc2_b200._compat import instantiate
from hulc2_b200.config import hulc2_config
from hulc2_b200.synthetic import synthetic_state_dict

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hulc2_golden.npz")
NON_LEARNED = ("x_map", "y_map", "temperature", "one_hot_embedding_eye", ".ones", "gripper_bounds", "action_max_bound", "action_min_bound")
_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = np.load(GOLDEN_PATH)
    return _golden


def gt(name) -> torch.Tensor:
    return torch.from_numpy(np.asarray(golden()[name]))


def build_model(variant="calvin", static_hw=(200, 200), dropout_p=0.0, device="cpu", **kw):
    """hulc2_b200 Hulc2 with the deterministic numpy weights used for the golden fixtures (seed 0)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = instantiate(hulc2_config(pkg="hulc2_b200", variant=variant, static_hw=static_hw, dropout_p=dropout_p, **kw))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    m.load_state_dict(synthetic_state_dict(shapes, seed=0, skip=NON_LEARNED), strict=False)
    return m.to(device)


def oracle_params(model, requires_grad=True):
    """Flat dict of CPU leaf tensors under state_dict names for oracle/hulc2_oracle.py."""
    names = {n for n, _ in model.named_parameters()}
    return {k: v.detach().cpu().clone().requires_grad_(requires_grad and k in names) for k, v in model.state_dict().items()}


def to_device(x, device):
    if isinstance(x, dict):
        return {k: to_device(v, device) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        return x.to(device)
    return x


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (scale-relative max-norm error)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
