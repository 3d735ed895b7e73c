"""GPU parity, datamodule row: the uint8-frame kernels of csrc/frames.cu through the C-ABI against
oracle/datamodule_oracle.py (bit-exact: byte gather + index work + IEEE fp32 scale/normalise), against the committed
outputs of the reference's own data pipeline (tests/golden/datamodule_golden.npz), and at step level: a train step fed
uint8 frames must equal the same step fed the fp32 frames the reference's dataloader would have produced."""
import os

import numpy as np
import pytest
import torch

from helpers import build_model, to_device
from oracle import datamodule_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datamodule_golden.npz"))
AUG_TOL = 1e-4   # reference fp32 grid_sample vs exact integer crop, see tests/test_oracle_datamodule.py


def _u8frames(store, start=None, length=None, shift=None, S=1):
    from hulc2_b200 import ops

    t = lambda a, dt: None if a is None else torch.as_tensor(np.asarray(a)).to(DEV, dt)
    return ops.U8Frames(torch.from_numpy(store).to(DEV), t(shift, torch.int32), t(start, torch.int64), t(length, torch.int32), S)


def _bf16_bits(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy()


CASES = [
    # (N_store, H, W, C, B, S, ragged, pad)
    (9, 24, 24, 3, 3, 8, True, 3),
    (6, 84, 84, 3, 2, 4, True, 4),
    (5, 200, 200, 3, 2, 3, False, 10),
    (4, 150, 200, 3, 1, 4, True, 10),     # real-world config geometry (H != W)
    (7, 36, 52, 1, 2, 5, True, 2),        # single-channel, odd row bytes (scalar load path)
]


@pytest.mark.parametrize("N,H,W,C,B,S,ragged,pad", CASES)
def test_frames_u8_kernels_bit_exact(N, H, W, C, B, S, ragged, pad):
    rng = np.random.default_rng(N * 1000 + H)
    store = rng.integers(0, 256, (N, H, W, C), dtype=np.uint8)
    store[0, :2, :3] = 0
    store[1, -2:, -3:] = 255
    length = rng.integers(1, S + 1, B) if ragged else np.full(B, S)
    start = np.array([rng.integers(0, N - l + 1) for l in length])
    shift = rng.integers(-pad, pad + 1, (B * S, 2))
    shift[0] = (-pad, pad)
    shift[-1] = (pad, -pad)
    for sh in (shift, None):
        fr = _u8frames(store, start, length, sh, S)
        ref = D.frames_u8_to_f32(store, start, length, sh, S)
        np.testing.assert_array_equal(fr.to_f32().cpu().numpy(), ref)
        if C * 16 % 8 == 0 and H >= 8 and W >= 8:
            from hulc2_b200 import ops

            xs = ops.pack_frames(fr)
            np.testing.assert_array_equal(xs.view(torch.int16).cpu().numpy(), _bf16_bits(D.pack_frames(ref)))
            # and the same operand the fp32-frame pack kernel builds from the oracle's frames
            xs2 = ops.pack_frames(torch.from_numpy(ref).to(DEV))
            assert torch.equal(xs.view(torch.int16), xs2.view(torch.int16))
    # identity window map (frames of a plain batch tensor)
    np.testing.assert_array_equal(_u8frames(store).to_f32().cpu().numpy(), D.frames_u8_to_f32(store))


def test_all_256_grey_levels_exact():
    store = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1)
    np.testing.assert_array_equal(_u8frames(store).to_f32().cpu().numpy().reshape(-1), D.normalize_u8(np.arange(256, dtype=np.uint8)))


@pytest.mark.parametrize("tag", ["train24", "train84", "edge24", "val24"])
def test_against_reference_pipeline_fixture(tag):
    start, length, S, pad = (int(v) for v in G[f"{tag}/meta"])
    shift = None
    if f"{tag}/shift_draw" in G:
        draw = G[f"{tag}/shift_draw"]
        shift = np.concatenate([draw, np.repeat(draw[length - 1 : length], S - length, 0)]) - pad
    out = _u8frames(G[f"{tag}/store"], [start], [length], shift, S).to_f32().cpu().numpy()
    ref = G[f"{tag}/out/rgb_static"]
    if shift is None:
        np.testing.assert_array_equal(out, ref)
    else:
        assert np.abs(out - ref).max() <= AUG_TOL
        back = lambda v: np.rint((v * 0.5 + 0.5) * 255.0).astype(np.int64)
        np.testing.assert_array_equal(back(out), back(ref))
    from hulc2_b200.datamodule import DeviceEpisodeStore

    st = DeviceEpisodeStore({"rgb_static": G[f"{tag}/store"]}, G[f"{tag}/rel_actions"], G[f"{tag}/robot_obs_raw"], G[f"{tag}/scene_obs"], device=DEV)
    b = st.window_batch(torch.tensor([start]), torch.tensor([length]), S)
    np.testing.assert_array_equal(b["actions"][0].cpu().numpy(), G[f"{tag}/out/actions"])
    np.testing.assert_array_equal(b["robot_obs"][0].cpu().numpy(), G[f"{tag}/out/robot_obs"])
    np.testing.assert_array_equal(b["state_info"]["robot_obs"][0].cpu().numpy(), G[f"{tag}/out/state_robot_obs"])
    np.testing.assert_array_equal(b["state_info"]["scene_obs"][0].cpu().numpy(), G[f"{tag}/out/state_scene_obs"])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_window_gather_modes(mode):
    from hulc2_b200._lib import call

    rng = np.random.default_rng(mode)
    N, D_, B, S = 50, 7, 9, 6
    table = rng.standard_normal((N, D_)).astype(np.float32)
    length = rng.integers(1, S + 1, B)
    length[0], length[1] = S, 1
    start = np.array([rng.integers(0, N - l + 1) for l in length])
    out = torch.empty(B, S, D_, device=DEV)
    t_d, s_d, l_d = torch.from_numpy(table).to(DEV), torch.from_numpy(start).to(DEV), torch.from_numpy(length.astype(np.int32)).to(DEV)
    call("hulc2_window_gather_f32", t_d.data_ptr(), s_d.data_ptr(), l_d.data_ptr(), out.data_ptr(), B, S, D_, mode)
    np.testing.assert_array_equal(out.cpu().numpy(), D.window_gather(table, start, length, S, mode))


def test_empty_and_bad_arguments():
    from hulc2_b200 import ops
    from hulc2_b200._lib import call

    e = torch.empty(0, 8, 8, 3, dtype=torch.uint8, device=DEV)
    assert ops.U8Frames(e).to_f32().shape == (0, 3, 8, 8)
    with pytest.raises(ValueError):
        ops.U8Frames(torch.zeros(2, 8, 8, 3, device=DEV))                       # not uint8
    with pytest.raises(ValueError):
        ops.U8Frames(torch.zeros(2, 8, 8, 3, dtype=torch.uint8, device=DEV), torch.zeros(3, 2, dtype=torch.int32, device=DEV))
    with pytest.raises(RuntimeError):
        call("hulc2_window_gather_f32", None, None, None, None, 1, 1, 1, 0)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-6), ("bf16", 1e-6)])
def test_train_step_from_uint8_frames_equals_fp32_frames(precision, tol):
    """Step level: uint8 HWC frames (+ shift draw) in the batch dict vs the fp32 NCHW frames the reference dataloader
    would have built from them -- same loss and same gradients (the trunk operand is bit-identical)."""
    from hulc2_b200 import noise, ops
    from hulc2_b200.synthetic import synthetic_batch

    B, S = 2, 32
    rng = np.random.default_rng(5)
    batch = to_device(synthetic_batch(B, seed=1), DEV)
    u8, f32 = {}, {}
    for mod in batch:
        u8[mod], f32[mod] = dict(batch[mod]), dict(batch[mod])
        u8[mod]["rgb_obs"], f32[mod]["rgb_obs"] = {}, {}
        for cam, hw, pad in (("rgb_static", 200, 10), ("rgb_gripper", 84, 4)):
            frames = rng.integers(0, 256, (B, S, hw, hw, 3), dtype=np.uint8)
            shift = rng.integers(-pad, pad + 1, (B, S, 2)).astype(np.int32)
            u8[mod]["rgb_obs"][cam] = torch.from_numpy(frames).to(DEV)
            u8[mod]["rgb_obs"][cam + "_shift"] = torch.from_numpy(shift).to(DEV)
            ref = D.frames_u8_to_f32(frames.reshape(-1, hw, hw, 3), shift=shift.reshape(-1, 2))
            f32[mod]["rgb_obs"][cam] = torch.from_numpy(ref).reshape(B, S, 3, hw, hw).to(DEV)
    idx = [torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)) for _ in batch]
    ops.set_precision(precision)
    try:
        res = []
        for b in (u8, f32):
            m = build_model(device=DEV, hidden_size=256).train()
            with noise.supplied(categories=[i.clone() for i in idx]):
                loss = m.training_step(b, 0)
            loss.backward()
            res.append((float(loss), m.perceptual_encoder.rgb_static_encoder.conv_model[0].weight.grad.clone(),
                        m.perceptual_encoder.rgb_gripper_encoder.conv_model[0].weight.grad.clone()))
        (la, ga, ha), (lb, gb, hb) = res
        assert abs(la - lb) <= tol * abs(lb)
        assert float((ga - gb).abs().max()) <= 1e-5 * float(gb.abs().max())
        assert float((ha - hb).abs().max()) <= 1e-5 * float(hb.abs().max())
    finally:
        ops.set_precision("fp32")


def test_datamodule_batches_train_step():
    """Hulc2DeviceDataModule -> {"vis","lang"} batches described by index tensors -> training_step; checked against the
    same windows materialised by the oracle as fp32 reference-contract batches."""
    from hulc2_b200 import noise, ops
    from hulc2_b200.datamodule import Hulc2DeviceDataModule, synthetic_store

    store = synthetic_store(160, device=DEV, seed=3)
    rng = np.random.default_rng(0)
    lang_emb = torch.from_numpy(rng.standard_normal((3, 384)).astype(np.float32))
    split = {"store": store, "ep_start_end_ids": [(0, 79), (80, 159)], "lang_start_end": [(0, 40), (50, 100), (100, 150)], "lang_emb": lang_emb}
    cfg = {"vis": dict(batch_size=2, min_window_size=16, max_window_size=32), "lang": dict(batch_size=2, min_window_size=20, max_window_size=32)}
    dm = Hulc2DeviceDataModule(cfg, split, split, seed=1)
    dm.setup()
    assert dm.modalities == ["vis", "lang"]
    batch = next(iter(dm.train_dataloader()))
    assert set(batch) == {"vis", "lang"} and batch["lang"]["lang"].shape == (2, 384) and batch["lang"]["use_for_aux_lang_loss"].dtype == torch.bool
    # materialise the same windows with the oracle
    mat = {}
    for mod, b in batch.items():
        mb = {k: v for k, v in b.items()}
        mb["rgb_obs"] = {}
        for cam, fr in b["rgb_obs"].items():
            ref = D.frames_u8_to_f32(fr.u8.cpu().numpy(), fr.win_start.cpu().numpy(), fr.win_len.cpu().numpy(), fr.shift.cpu().numpy(), fr.S)
            mb["rgb_obs"][cam] = torch.from_numpy(ref).reshape(2, fr.S, *ref.shape[1:]).to(DEV)
            assert int(fr.win_len.min()) >= 16 and int(fr.win_len.max()) <= 32
        acts = D.window_gather(store.rel_actions.cpu().numpy(), b["rgb_obs"]["rgb_static"].win_start.cpu().numpy(),
                               b["rgb_obs"]["rgb_static"].win_len.cpu().numpy(), 32, 2)
        np.testing.assert_array_equal(b["actions"].cpu().numpy(), acts)
        mat[mod] = mb
    idx = [torch.randint(0, 32, (2, 32), generator=torch.Generator().manual_seed(5)) for _ in batch]
    ops.set_precision("bf16")
    try:
        losses = []
        for bt in (batch, mat):
            m = build_model(device=DEV, hidden_size=256).train()
            with noise.supplied(categories=[i.clone() for i in idx]):
                losses.append(float(m.training_step(bt, 0)))
        assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[1])
    finally:
        ops.set_precision("fp32")
    # validation loader: no augmentation, deterministic order
    vb = next(iter(dm.val_dataloader()))
    assert vb["vis"]["rgb_obs"]["rgb_static"].shift is None


def test_loaders_shard_windows_across_ranks():
    """Data parallel: rank r of `world` takes entries r, r + world, ... of the SAME epoch permutation (what Lightning's injected
    DistributedSampler does for the reference), padded to equal length; augmentation draws differ per rank."""
    from hulc2_b200.datamodule import WindowLoader, build_episode_lookup, synthetic_store

    store = synthetic_store(200, device=DEV, seed=3)
    look = build_episode_lookup([(0, 99), (100, 199)], 16, 32)
    loaders = [WindowLoader(store, look, 4, 16, 32, True, {"rgb_static": 10}, seed=7, rank=r, world=2) for r in range(2)]
    assert len(loaders[0]) == len(loaders[1]) == ((len(look) + 1) // 2) // 4
    seen, shifts = [], []
    for ld in loaders:
        ids = []
        for b in ld:
            assert b["actions"].shape == (4, 32, 7)               # drop_last: static shapes for the captured step
            ids += b["idx"].tolist()
            shifts.append(b["rgb_obs"]["rgb_static"].shift[:8].clone())
        seen.append(ids)
    assert not set(seen[0]) & set(seen[1])
    assert len(set(seen[0]) | set(seen[1])) >= len(look) - 8        # all but the dropped ragged tail of each shard
    assert not torch.equal(shifts[0], shifts[len(shifts) // 2])
    with pytest.raises(ValueError):
        WindowLoader(store, np.asarray([0, 195]), 2, 16, 32)        # a 16-step window from frame 195 would leave the 200-frame store
