"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of what the reference's dataloader workers do to one window
before ``Hulc2.training_step`` sees it.  The product (``hulc2_b200/csrc/frames.cu``) never imports this file; only
``tests/`` do, as the checker.

Pinned against the reference itself: ``tests/golden/make_golden_datamodule.py`` runs the unmodified
``process_rgb`` / ``RandomShiftsAug`` / ``ScaleImageTensor`` / ``Normalize`` / ``BaseDataset.pad_sequence`` from
``/root/reference`` on seeded uint8 episodes and commits inputs + outputs to ``tests/golden/datamodule_golden.npz``.

What is restated (file:line in /root/reference):
  * window slice                       hulc2/datasets/npz_dataset.py:117-143 (get_sequences: frames [start, start+len))
  * pad to max_window_size             hulc2/datasets/base_dataset.py:121-163 (pad_sequence, pad_with_repetition/zeros)
  * HWC uint8 -> CHW                   hulc2/datasets/utils/episode_utils.py:61-86 (process_rgb)
  * RandomShiftsAug(pad)               hulc2/utils/transforms.py:85-106
  * ScaleImageTensor, Normalize(.5,.5) hulc2/utils/transforms.py:8-19; conf/datamodule/transforms/rand_shift.yaml:2-20

RandomShiftsAug: the base grid is linspace(-1+1/P, 1-1/P, P)[:h] with P = h + 2*pad, i.e. coordinate i has normalised
position -1 + (2i+1)/P; the integer draw s in [0, 2*pad] adds 2s/P; grid_sample(align_corners=False) un-normalises with
((g+1)*P-1)/2 = i + s.  So the sample point is the CENTRE of padded pixel (i+s): the bilinear blend degenerates to a copy
and the augmentation is an integer crop of the replicate-padded frame,
        out[y, x] = in[clamp(y + sy - pad, 0, h-1), clamp(x + sx - pad, 0, w-1)].
The reference evaluates the grid in fp32, so its sample points are off-centre by a few 1e-6 pixels and its outputs
deviate from the exact crop by up to ~1e-2 grey levels (6e-5 after normalisation; measured by the generating script and
stored in the fixture as ``aug_ref_max_dev``).  ``shift`` below is the signed offset (dx, dy) = (sx - pad, sy - pad);
``shift[..., 0]`` moves along width (grid[..., 0] is x), ``shift[..., 1]`` along height.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def window_frame_indices(win_start: np.ndarray, win_len: Optional[np.ndarray], S: int) -> np.ndarray:
    """[B,S] source frame of every (window, step): frames start..start+len-1, then the last one repeated
    (base_dataset.py:152-156 pad_with_repetition)."""
    win_start = np.asarray(win_start, dtype=np.int64)
    L = np.full(win_start.shape, S, dtype=np.int64) if win_len is None else np.maximum(np.asarray(win_len, dtype=np.int64), 1)
    t = np.arange(S, dtype=np.int64)[None, :]
    return win_start[:, None] + np.minimum(t, L[:, None] - 1)


def normalize_u8(u8: np.ndarray) -> np.ndarray:
    """ScaleImageTensor then Normalize(mean .5, std .5) in fp32, reference operation order (transforms.py:8-19)."""
    x = u8.astype(np.float32) / np.float32(255.0)
    return (x - np.float32(0.5)) / np.float32(0.5)


def frames_u8_to_f32(store: np.ndarray, win_start=None, win_len=None, shift=None, S: int = 1) -> np.ndarray:
    """uint8 HWC store [N,H,W,C] -> fp32 NCHW [F,C,H,W] in [-1,1].  ``win_start`` None = identity (frame f = store[f]);
    ``shift`` int [F,2] = (dx,dy) or None."""
    N, H, W, C = store.shape
    src = np.arange(N, dtype=np.int64) if win_start is None else window_frame_indices(win_start, win_len, S).reshape(-1)
    F = src.shape[0]
    out = np.empty((F, C, H, W), dtype=np.float32)
    ys, xs = np.arange(H), np.arange(W)
    for f in range(F):
        dx, dy = (0, 0) if shift is None else (int(shift[f][0]), int(shift[f][1]))
        yy = np.clip(ys + dy, 0, H - 1)
        xx = np.clip(xs + dx, 0, W - 1)
        out[f] = normalize_u8(store[src[f]][yy][:, xx]).transpose(2, 0, 1)   # HWC -> CHW (episode_utils.py:78-82)
    return out


def pack_frames(x: np.ndarray) -> np.ndarray:
    """fp32 NCHW [F,C,H,W] -> fp32 [F, H/4, W/4, 16C] in the trunk's packed-frames order, channel (ci,a,b) =
    x[f,ci,4I+a,4J+b] (include/hulc2_b200.h, pack_frames); the caller rounds to bf16."""
    F, C, H, W = x.shape
    H4, W4 = H // 4, W // 4
    v = x[:, :, : 4 * H4, : 4 * W4].reshape(F, C, H4, 4, W4, 4)        # f ci I a J b
    return np.ascontiguousarray(v.transpose(0, 2, 4, 1, 3, 5)).reshape(F, H4, W4, 16 * C)


def window_gather(store: np.ndarray, win_start, win_len, S: int, mode: int) -> np.ndarray:
    """Per-step vectors [N,D] -> [B,S,D].  mode 0: pad_with_repetition (robot_obs, state_info, absolute actions);
    mode 1: pad_with_zeros (joint actions); mode 2: relative actions -- zeros except the last (gripper) component,
    which repeats (base_dataset.py:129-150)."""
    idx = window_frame_indices(win_start, win_len, S)
    out = store[idx].astype(np.float32).copy()
    if mode in (1, 2):
        L = np.full(idx.shape[0], S) if win_len is None else np.maximum(np.asarray(win_len), 1)
        pad = np.arange(S)[None, :] >= L[:, None]
        if mode == 1:
            out[pad] = 0.0
        else:
            out[..., :-1][pad] = 0.0
    return out


def random_shifts_aug_fp32(x: np.ndarray, shift_draw: np.ndarray, pad: int) -> np.ndarray:
    """Faithful (non-closed-form) restatement of RandomShiftsAug.forward for ONE sequence x [n,c,h,h] float32 with the
    integer draw ``shift_draw`` [n,2] in [0, 2*pad]: replicate pad, fp32 grid, bilinear grid_sample with zero padding
    and align_corners=False -- used to show where the reference's ~1e-2 grey-level deviation from the crop comes from."""
    n, c, h, w = x.shape
    P = h + 2 * pad
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)), mode="edge").astype(np.float32)
    eps = np.float32(1.0) / np.float32(P)
    ar = np.linspace(np.float32(-1.0) + eps, np.float32(1.0) - eps, P, dtype=np.float32)[:h]
    out = np.zeros((n, c, h, w), dtype=np.float32)
    for i in range(n):
        sx = np.float32(shift_draw[i][0]) * (np.float32(2.0) / np.float32(P))
        sy = np.float32(shift_draw[i][1]) * (np.float32(2.0) / np.float32(P))
        gx = ar + sx
        gy = ar + sy
        px = ((gx + np.float32(1.0)) * np.float32(P) - np.float32(1.0)) / np.float32(2.0)
        py = ((gy + np.float32(1.0)) * np.float32(P) - np.float32(1.0)) / np.float32(2.0)
        x0 = np.floor(px).astype(np.int64)
        y0 = np.floor(py).astype(np.int64)
        wx1 = (px - x0).astype(np.float32)
        wy1 = (py - y0).astype(np.float32)

        def tap(yi, xi):
            ok = ((yi >= 0) & (yi < P))[:, None] & ((xi >= 0) & (xi < P))[None, :]
            v = xp[i][:, np.clip(yi, 0, P - 1)][:, :, np.clip(xi, 0, P - 1)]
            return v * ok[None].astype(np.float32)

        out[i] = (tap(y0, x0) * ((1 - wy1)[:, None] * (1 - wx1)[None, :])[None] + tap(y0, x0 + 1) * ((1 - wy1)[:, None] * wx1[None, :])[None]
                  + tap(y0 + 1, x0) * (wy1[:, None] * (1 - wx1)[None, :])[None] + tap(y0 + 1, x0 + 1) * (wy1[:, None] * wx1[None, :])[None])
    return out
