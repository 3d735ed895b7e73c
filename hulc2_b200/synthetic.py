"""Synthetic CALVIN-shaped batches (SURVEY.md section 8d "Synthetic inputs").

The batch dict layout is the reference's (docstring of ``Hulc2.training_step``,
hulc2/models/hulc2.py:341-360; producers hulc2/datasets/npz_dataset.py:117-143).
``numpy`` PCG64 streams are used for the small deterministic test batches (stable
across torch versions); ``fast=True`` draws with a torch generator on the target
device for the full-size benchmark batches.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch


def _modality(rng, B, S, static_hw, lang: bool, depth_static: bool, aux: str) -> dict:
    H, W = static_hw
    f32 = np.float32
    actions = rng.uniform(-1, 1, (B, S, 7)).astype(f32)
    edge = rng.uniform(0, 1, (B, S, 6))
    actions[..., :6] = np.where(edge < 0.01, -1.0, np.where(edge > 0.99, 1.0, actions[..., :6]))
    actions[..., 6] = np.where(rng.uniform(0, 1, (B, S)) < 0.5, -1.0, 1.0)
    raw = rng.uniform(-1, 1, (B, S, 15)).astype(f32)
    raw[..., 3:6] = rng.uniform(-0.9 * math.pi / 2, 0.9 * math.pi / 2, (B, S, 3))
    d = {
        "rgb_obs": {
            "rgb_static": rng.uniform(-1, 1, (B, S, 3, H, W)).astype(f32),
            "rgb_gripper": rng.uniform(-1, 1, (B, S, 3, 84, 84)).astype(f32),
        },
        "depth_obs": {},
        "robot_obs": rng.uniform(-1, 1, (B, S, 8)).astype(f32),
        "actions": actions,
        "state_info": {"robot_obs": raw, "scene_obs": rng.uniform(-1, 1, (B, S, 24)).astype(f32)},
        "idx": np.arange(B, dtype=np.int64),
    }
    if depth_static:
        d["depth_obs"]["depth_static"] = rng.uniform(0, 1, (B, S, H, W)).astype(f32)
    if lang:
        d["lang"] = rng.standard_normal((B, 384)).astype(f32)
        if aux == "all":
            d["use_for_aux_lang_loss"] = np.ones((B,), dtype=bool)
        else:
            m = rng.uniform(0, 1, (B,)) < 0.5
            m[0] = True
            d["use_for_aux_lang_loss"] = m
    return d


def _to_torch(x, device):
    if isinstance(x, dict):
        return {k: _to_torch(v, device) for k, v in x.items()}
    return torch.from_numpy(x).to(device)


def synthetic_batch(
    B: int,
    S: int = 32,
    seed: int = 1,
    static_hw: Tuple[int, int] = (200, 200),
    device="cpu",
    modalities=("vis", "lang"),
    depth_static: bool = False,
    aux: str = "all",
) -> Dict[str, dict]:
    """Deterministic train batch ``{"vis": {...}, "lang": {...}}`` with B windows per modality."""
    rng = np.random.default_rng(seed)
    out = {}
    for mod in modalities:
        out[mod] = _to_torch(_modality(rng, B, S, static_hw, "lang" in mod, depth_static, aux), device)
    return out


def synthetic_batch_fast(B, S=32, seed=1, static_hw=(200, 200), device="cuda", modalities=("vis", "lang"), pin=False, frames="fp32",
                         depth_static=False):
    """Same shapes/ranges drawn with a torch generator (used for the full-size benchmark).  ``frames="uint8"`` gives the
    cameras as the dataset stores them (uint8 ``[B,S,H,W,3]``) plus the RandomShiftsAug draw ``<cam>_shift`` int32
    ``[B,S,2]`` (pad 10 static / 4 gripper, rand_shift.yaml); scale + normalise + shift then run on the device."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    H, W = static_hw

    def U(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, device=dev) * (hi - lo) + lo

    out = {}
    for mod in modalities:
        actions = U(B, S, 7)
        e = torch.rand(B, S, 6, generator=g, device=dev)
        actions[..., :6] = torch.where(e < 0.01, -1.0, torch.where(e > 0.99, 1.0, actions[..., :6]))
        actions[..., 6] = torch.where(torch.rand(B, S, generator=g, device=dev) < 0.5, -1.0, 1.0)
        raw = U(B, S, 15)
        raw[..., 3:6] = U(B, S, 3, lo=-0.9 * math.pi / 2, hi=0.9 * math.pi / 2)
        d = {
            "rgb_obs": {"rgb_static": U(B, S, 3, H, W), "rgb_gripper": U(B, S, 3, 84, 84)} if frames == "fp32" else {
                "rgb_static": torch.randint(0, 256, (B, S, H, W, 3), generator=g, device=dev, dtype=torch.uint8),
                "rgb_static_shift": torch.randint(-10, 11, (B, S, 2), generator=g, device=dev, dtype=torch.int32),
                "rgb_gripper": torch.randint(0, 256, (B, S, 84, 84, 3), generator=g, device=dev, dtype=torch.uint8),
                "rgb_gripper_shift": torch.randint(-4, 5, (B, S, 2), generator=g, device=dev, dtype=torch.int32),
            },
            "depth_obs": {"depth_static": U(B, S, H, W, lo=0.0, hi=1.0)} if depth_static else {},
            "robot_obs": U(B, S, 8),
            "actions": actions,
            "state_info": {"robot_obs": raw, "scene_obs": U(B, S, 24)},
            "idx": torch.arange(B, device=dev),
        }
        if "lang" in mod:
            d["lang"] = torch.randn(B, 384, generator=g, device=dev)
            d["use_for_aux_lang_loss"] = torch.ones(B, dtype=torch.bool, device=dev)
        out[mod] = d
    if pin:
        out = tree_map(lambda t: t.pin_memory(), out)
    return out


def tree_map(fn, x):
    """Applies ``fn`` to every tensor leaf of a batch dict.  ``ops.U8Frames`` leaves (datamodule batches) keep viewing their
    resident frame store; ``fn`` maps their index tensors (window starts / lengths / shift draws)."""
    if isinstance(x, dict):
        return {k: tree_map(fn, v) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        return fn(x)
    if type(x).__name__ == "U8Frames":
        from .ops import U8Frames

        opt = lambda t: None if t is None else fn(t)  # noqa: E731
        return U8Frames(x.u8, opt(x.shift), opt(x.win_start), opt(x.win_len), x.S)
    return x


def synthetic_obs(N: int, seed: int = 2, static_hw=(200, 200), device="cpu") -> Tuple[dict, dict]:
    """Rollout observation + language goal for N parallel envs, S=1 (hulc2_wrapper.py:47-62)."""
    rng = np.random.default_rng(seed)
    H, W = static_hw
    f32 = np.float32
    raw = rng.uniform(-1, 1, (N, 1, 15)).astype(f32)
    raw[..., 3:6] = rng.uniform(-0.9 * math.pi / 2, 0.9 * math.pi / 2, (N, 1, 3))
    obs = {
        "rgb_obs": {
            "rgb_static": rng.uniform(-1, 1, (N, 1, 3, H, W)).astype(f32),
            "rgb_gripper": rng.uniform(-1, 1, (N, 1, 3, 84, 84)).astype(f32),
        },
        "depth_obs": {},
        "robot_obs": rng.uniform(-1, 1, (N, 1, 8)).astype(f32),
        "robot_obs_raw": raw,
    }
    goal = {"lang": rng.standard_normal((N, 384)).astype(f32)}
    return _to_torch(obs, device), _to_torch(goal, device)


def synthetic_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, skip=()) -> Dict[str, torch.Tensor]:
    """Deterministic weights for every floating parameter in ``shapes`` (name -> shape), drawn
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) in key order with numpy; LayerNorm weights ~ 1 + 0.1 U,
    biases small.  Used so golden fixtures do not depend on torch's init RNG."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if any(name.endswith(s) for s in skip):
            continue
        if name == "logit_scale":
            out[name] = torch.tensor(float(np.log(1 / 0.07)), dtype=torch.float32)
            continue
        fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else max(int(shp[0]), 1)
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        v = rng.uniform(-bound, bound, shp).astype(np.float32)
        if (".ln." in name or ".norm" in name or ".layernorm." in name) and name.endswith("weight"):
            v = (1.0 + 0.1 * rng.uniform(-1, 1, shp)).astype(np.float32)
        out[name] = torch.from_numpy(v)
    return out
