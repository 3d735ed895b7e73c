"""TEST INFRASTRUCTURE ONLY -- noise injection for the unmodified reference.

The reference draws its randomness with ``torch.rand`` (logistic_decoder_rnn.py:236,248)
and ``Categorical.sample`` (through OneHotCategorical.sample, distributions.py:38 /
hulc2.py:235).  These context managers make it consume caller-supplied tensors instead,
in call order, so reference, oracle and CUDA path see identical noise.
"""
from __future__ import annotations

import contextlib
from typing import List

import torch


@contextlib.contextmanager
def supplied_uniforms(queue: List[torch.Tensor]):
    """``torch.rand(shape, ...)`` pops tensors from ``queue`` (shapes must match)."""
    orig = torch.rand
    q = list(queue)

    def fake_rand(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        t = q.pop(0)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = orig
    assert not q, "unused uniforms"


@contextlib.contextmanager
def supplied_categories(queue: List[torch.Tensor]):
    """``Categorical.sample()`` returns the queued index tensors instead of drawing."""
    from torch.distributions import Categorical

    orig = Categorical.sample
    q = list(queue)

    def fake_sample(self, sample_shape=torch.Size()):
        t = q.pop(0)
        assert tuple(t.shape) == tuple(self._batch_shape), (t.shape, self._batch_shape)
        return t.clone().long()

    Categorical.sample = fake_sample
    try:
        yield
    finally:
        Categorical.sample = orig
    assert not q, "unused category draws"


@contextlib.contextmanager
def supplied_normals(queue: List[torch.Tensor]):
    """``Normal.rsample()`` / ``sample()`` (continuous plan, distributions.py:28-29) consume queued standard-normal
    tensors instead of drawing (``torch.distributions.normal._standard_normal`` for rsample, ``torch.normal`` for sample)."""
    import torch.distributions.normal as N

    orig_std, orig_normal = N._standard_normal, torch.normal
    q = list(queue)

    def fake_standard_normal(shape, dtype, device):
        t = q.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t.clone().to(dtype=dtype, device=device)

    def fake_normal(mean, std, *a, **kw):
        t = q.pop(0)
        return mean + std * t

    N._standard_normal = fake_standard_normal
    torch.normal = fake_normal
    try:
        yield
    finally:
        N._standard_normal = orig_std
        torch.normal = orig_normal
    assert not q, "unused normal draws"
