"""CPU: the datamodule oracle (oracle/datamodule_oracle.py) against outputs of the UNMODIFIED reference data pipeline
(tests/golden/datamodule_golden.npz, written by tests/golden/make_golden_datamodule.py), and live against
/root/reference's RandomShiftsAug when that directory exists (build container only)."""
import os

import numpy as np
import pytest

from oracle import datamodule_oracle as D

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datamodule_golden.npz"))
# The reference evaluates RandomShiftsAug's sampling grid in fp32: its sample points sit a few 1e-6 pixels off the pixel
# centres, which blends up to ~1e-2 grey levels of a neighbour into each output (measured by the generating script:
AUG_TOL = 1e-4  # > G["aug_ref_max_dev"] = 5.6e-5 in normalised [-1,1] units).  Source-pixel selection itself is exact.
CASES = ["train24", "train84", "edge24", "val24"]


def _case(tag):
    start, length, S, pad = (int(v) for v in G[f"{tag}/meta"])
    draw = G[f"{tag}/shift_draw"] if f"{tag}/shift_draw" in G else None
    shift = None
    if draw is not None:
        shift = np.zeros((S, 2), dtype=np.int64)       # padded steps repeat the LAST augmented frame: same shift as step len-1
        shift[:length] = draw - pad
        shift[length:] = draw[length - 1] - pad
    return start, length, S, pad, draw, shift


@pytest.mark.parametrize("tag", CASES)
def test_frames_match_reference_pipeline(tag):
    start, length, S, pad, draw, shift = _case(tag)
    out = D.frames_u8_to_f32(G[f"{tag}/store"], np.array([start]), np.array([length]), shift, S)
    ref = G[f"{tag}/out/rgb_static"]
    assert out.shape == ref.shape and out.dtype == np.float32
    if draw is None:
        np.testing.assert_array_equal(out, ref)       # no augmentation: scale + normalise must be bit-exact
    else:
        assert np.abs(out - ref).max() <= AUG_TOL
        # exact source-pixel selection: rounding the reference back to grey levels reproduces the crop bit for bit
        back = lambda v: np.rint((v * 0.5 + 0.5) * 255.0).astype(np.int64)
        np.testing.assert_array_equal(back(out), back(ref))


@pytest.mark.parametrize("tag", CASES)
def test_window_vectors_match_reference_padding(tag):
    start, length, S, _, _, _ = _case(tag)
    ws, wl = np.array([start]), np.array([length])
    np.testing.assert_array_equal(D.window_gather(G[f"{tag}/rel_actions"], ws, wl, S, 2)[0], G[f"{tag}/out/actions"])
    np.testing.assert_array_equal(D.window_gather(G[f"{tag}/robot_obs_raw"], ws, wl, S, 0)[0], G[f"{tag}/out/state_robot_obs"])
    np.testing.assert_array_equal(D.window_gather(G[f"{tag}/scene_obs"], ws, wl, S, 0)[0], G[f"{tag}/out/state_scene_obs"])
    np.testing.assert_array_equal(D.window_gather(G[f"{tag}/robot_obs_raw"][:, [0, 1, 2, 3, 4, 5, 6, 14]], ws, wl, S, 0)[0], G[f"{tag}/out/robot_obs"])


def test_faithful_grid_sample_restatement_explains_the_deviation():
    """The non-closed-form restatement (fp32 grid + bilinear taps) lands on the reference to fp32 rounding, i.e. the
    residual between reference and integer crop is the reference's own grid arithmetic, not a different algorithm."""
    start, length, S, pad, draw, _ = _case("train24")
    x = G["train24/store"][start : start + length].transpose(0, 3, 1, 2).astype(np.float32)
    y = D.random_shifts_aug_fp32(x, draw, pad)
    ref = G["train24/out/rgb_static"][:length]
    assert np.abs(((y / np.float32(255.0)) - np.float32(0.5)) / np.float32(0.5) - ref).max() <= 2e-5


def test_window_indices_and_pack_layout():
    idx = D.window_frame_indices(np.array([3, 10]), np.array([2, 5]), 5)
    np.testing.assert_array_equal(idx, [[3, 4, 4, 4, 4], [10, 11, 12, 13, 14]])
    np.testing.assert_array_equal(D.window_frame_indices(np.array([7]), None, 3), [[7, 8, 9]])
    x = np.arange(2 * 3 * 8 * 12, dtype=np.float32).reshape(2, 3, 8, 12)
    p = D.pack_frames(x)
    assert p.shape == (2, 2, 3, 48)
    for f, I, J, ci, a, b in [(0, 0, 0, 0, 0, 0), (1, 1, 2, 2, 3, 1), (0, 1, 0, 1, 2, 3)]:
        assert p[f, I, J, ci * 16 + a * 4 + b] == x[f, ci, 4 * I + a, 4 * J + b]


@pytest.mark.skipif(not os.path.isdir("/root/reference/hulc2"), reason="reference tree only exists in the build container")
def test_live_against_reference_random_shifts_aug():
    import importlib.util

    import torch

    spec = importlib.util.spec_from_file_location("_ref_transforms", "/root/reference/hulc2/utils/transforms.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(3)
    for hw, pad in ((84, 4), (200, 10)):
        u8 = rng.integers(0, 256, (3, hw, hw, 3), dtype=np.uint8)
        draw = rng.integers(0, 2 * pad + 1, (3, 2))
        orig = torch.randint
        torch.randint = lambda *a, **k: torch.from_numpy(draw.astype(np.float32)).reshape(3, 1, 1, 2)
        try:
            y = m.RandomShiftsAug(pad)(torch.from_numpy(u8).permute(0, 3, 1, 2))
        finally:
            torch.randint = orig
        y = ((m.ScaleImageTensor()(y) - 0.5) / 0.5).numpy()
        assert np.abs(D.frames_u8_to_f32(u8, shift=draw - pad) - y).max() <= AUG_TOL


# ----------------------------------------------------------------------------- host-side index logic of hulc2_b200.datamodule
def _ref_dataset_module():
    import sys
    import types

    from oracle import ref_import

    ref_import.install_shims()
    if "pyhash" not in sys.modules:
        ph = types.ModuleType("pyhash")
        ph.fnv1_32 = lambda: (lambda s: 0)
        sys.modules["pyhash"] = ph
    from hulc2.datasets import npz_dataset

    return npz_dataset


def test_lookup_logic_small_cases():
    from hulc2_b200 import datamodule as M

    look = M.build_episode_lookup([(0, 49), (60, 99)], 16, 32)
    np.testing.assert_array_equal(look, list(range(0, 34)) + list(range(60, 84)))
    assert M.max_window_at(look, 0, 16, 32) == 32
    assert M.max_window_at(look, 33, 16, 32) == 16          # 16 lookups from the end of episode 0 -> only the minimum fits
    assert M.max_window_at(look, 25, 16, 32) == 24
    assert M.max_window_at(look, len(look) - 1, 16, 32) == 16
    ep, ll = M.build_lang_lookup([(0, 40), (50, 100)], 20, 32, skip_frames=2)
    np.testing.assert_array_equal(ep, list(range(0, 21, 2)) + list(range(50, 81, 2)))
    np.testing.assert_array_equal(ll, [0] * 11 + [1] * 16)
    use = M.use_for_aux_lang_loss(ll, np.arange(len(ll)), 1)
    assert use[10] and use[-1] and not use[0] and use.sum() == 2
    with pytest.raises(ValueError):
        M.build_episode_lookup([(0, 20)], 16, 32)


@pytest.mark.skipif(not os.path.isdir("/root/reference/hulc2"), reason="reference tree only exists in the build container")
def test_lookup_logic_live_against_reference():
    import types

    from hulc2_b200 import datamodule as M

    nd = _ref_dataset_module()
    eps = [(0, 70), (100, 140), (141, 230)]
    look = M.build_episode_lookup(eps, 16, 32)
    fake = types.SimpleNamespace(min_window_size=16, max_window_size=32, episode_lookup=list(look), validation=False)
    import numpy.random as npr

    for idx in list(range(0, len(look), 3)) + [len(look) - 1, len(look) - 17]:
        captured = {}
        orig = npr.randint
        npr.randint = lambda lo, hi: captured.setdefault("hi", hi) and lo      # get_window_size draws randint(min, max_window+1)
        try:
            nd.NpzDataset.get_window_size(fake, idx)
        finally:
            npr.randint = orig
        assert captured["hi"] - 1 == M.max_window_at(look, idx, 16, 32), idx
    ll = np.array([0, 0, 0, 1, 1, 2, 2, 2, 2])
    for w in (1, 2):
        fake = types.SimpleNamespace(with_lang=True, aux_lang_loss_window=w, lang_lookup=list(ll))
        ref = [nd.NpzDataset.add_language_info(fake, {}, i)["use_for_aux_lang_loss"] for i in range(len(ll))]
        np.testing.assert_array_equal(M.use_for_aux_lang_loss(ll, np.arange(len(ll)), w), ref)
