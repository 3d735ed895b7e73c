"""Randomness of the policy step as explicit inputs.

The reference draws noise with ``torch.multinomial`` / ``torch.rand`` / ``nn.Dropout``
(distributions.py:38, logistic_decoder_rnn.py:236,248, plan_recognition_net.py:142 + the encoder
layers).  Here every draw goes through this module: by default from the library's Philox4x32-10
kernels (counter-based, seeded per process), or -- for parity tests and reproducible rollouts --
from caller-supplied tensors queued with :func:`supplied`.
"""
from __future__ import annotations

import contextlib
from collections import deque
from typing import Deque, Dict, Optional

import torch

from . import ops

_state = {"seed": 0x5EED, "counter": 0}
_epochs: Dict[tuple, torch.Tensor] = {}   # per device: u64 step counter read by the Philox kernels (CUDA-graph replays)
_queues: Dict[str, Deque[torch.Tensor]] = {"categories": deque(), "uniforms": deque(), "masks": deque(), "normals": deque()}


def manual_seed(seed: int) -> None:
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state["counter"] = 0


def epoch_tensor(device) -> torch.Tensor:
    """Device-resident noise epoch (int64[1], starts at 0).  The Philox kernels add ``epoch << 40`` to their counter,
    so a train step captured in a CUDA graph draws fresh noise on every replay; ``trainer`` bumps it once per step."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _epochs:
        _epochs[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _epochs[key]


def begin_step() -> None:
    """Restart the host-side counter: with the device epoch distinguishing steps, every step issues the same sequence
    of (seed, offset) launch arguments -- what a captured graph replays."""
    _state["counter"] = 0


def _next_offset(n: int) -> int:
    off = _state["counter"]
    _state["counter"] += (n + 3) // 4 + 1
    return off


@contextlib.contextmanager
def supplied(categories=(), uniforms=(), masks=(), normals=()):
    """Queue tensors to be consumed (in call order) instead of fresh draws."""
    _queues["normals"].extend(normals)
    _queues["categories"].extend(categories)
    _queues["uniforms"].extend(uniforms)
    _queues["masks"].extend(masks)
    try:
        yield
    except BaseException:
        for q in _queues.values():      # the body failed: report ITS error, not the draws it left behind
            q.clear()
        raise
    left = {k: len(v) for k, v in _queues.items() if len(v)}
    for q in _queues.values():
        q.clear()
    if left:
        raise AssertionError(f"unused supplied noise: {left}")


def fuse_supplied(n: int) -> None:
    """A step that runs the windows of ``n`` modalities as one batch consumes one draw where the per-modality loop
    consumed ``n``.  Supplied category/mask/uniform tensors are queued per modality (modality 0's draws, then modality 1's, ...):
    regroup them into batch-concatenated tensors in draw order."""
    for key in ("categories", "masks", "normals", "uniforms"):
        q = _queues[key]
        if not q:
            continue
        items = list(q)
        assert len(items) % n == 0, f"supplied {key}: {len(items)} tensors cannot be split over {n} modalities"
        per = len(items) // n
        q.clear()
        q.extend(torch.cat([items[m * per + i] for m in range(n)], 0) for i in range(per))


def uniform(shape, device) -> torch.Tensor:
    if _queues["uniforms"]:
        t = _queues["uniforms"].popleft()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=device, dtype=torch.float32)
    n = 1
    for s in shape:
        n *= int(s)
    return ops.uniform(tuple(shape), device, _state["seed"], _next_offset(n), epoch=epoch_tensor(device))


def normal(shape, device) -> torch.Tensor:
    """Standard normals (the eps of Normal.rsample/sample, distributions.py:28-29): the next supplied tensor, or
    Box-Muller over two Philox uniform draws."""
    if _queues["normals"]:
        t = _queues["normals"].popleft()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=device, dtype=torch.float32)
    n = 1
    for s in shape:
        n *= int(s)
    u1 = ops.uniform(tuple(shape), device, _state["seed"], _next_offset(n), epoch=epoch_tensor(device))
    u2 = ops.uniform(tuple(shape), device, _state["seed"], _next_offset(n), epoch=epoch_tensor(device))
    return ops.box_muller(u1, u2)


def categories(logits: torch.Tensor, cats: int, classes: int) -> torch.Tensor:
    """Category indices [B,cats] ~ Categorical(softmax(logits)) (or the next supplied tensor)."""
    if _queues["categories"]:
        t = _queues["categories"].popleft()
        assert tuple(t.shape) == (logits.shape[0], cats), (tuple(t.shape), (logits.shape[0], cats))
        return t.to(device=logits.device, dtype=torch.int64)
    u = uniform((logits.shape[0], cats), logits.device)
    return ops.categorical_sample(logits.detach(), u, cats, classes)


def keep_mask(shape, p: float, device) -> Optional[torch.Tensor]:
    """uint8 dropout keep mask with P(keep) = 1-p (or the next supplied mask)."""
    if _queues["masks"]:
        t = _queues["masks"].popleft()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=device, dtype=torch.uint8).contiguous()
    n = 1
    for s in shape:
        n *= int(s)
    return ops.dropout_mask(tuple(shape), p, device, _state["seed"], _next_offset(n), epoch=epoch_tensor(device))
