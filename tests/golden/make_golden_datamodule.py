"""Generates tests/golden/datamodule_golden.npz from the UNMODIFIED reference data pipeline.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_datamodule.py
Executes, on seeded uint8 episodes, the reference functions a dataloader worker runs for one window
(npz_dataset.py:117-143): ``process_rgb`` with the train transform chain of conf/datamodule/transforms/rand_shift.yaml
(Resize is the identity at the dataset's native size and is left out; RandomShiftsAug -> ScaleImageTensor ->
Normalize(0.5,0.5)), ``process_actions``, ``get_state_info_dict`` and ``BaseDataset.pad_sequence``.  ``torch.randint``
is replaced for the duration of the augmentation so the integer shift draw is a recorded input.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

ref_import.install_shims()
if "pyhash" not in sys.modules:          # base_dataset.py:9,13 builds a hasher at import; only validation window sizes use it
    ph = types.ModuleType("pyhash")
    ph.fnv1_32 = lambda: (lambda s: 0)
    sys.modules["pyhash"] = ph

import torchvision  # noqa: E402
from hulc2.datasets.base_dataset import BaseDataset  # noqa: E402
from hulc2.datasets.utils.episode_utils import get_state_info_dict, process_actions, process_rgb, process_state  # noqa: E402
from hulc2.utils.transforms import RandomShiftsAug, ScaleImageTensor  # noqa: E402

G = {}


def chain(pad):
    return torchvision.transforms.Compose([RandomShiftsAug(pad), ScaleImageTensor(), torchvision.transforms.Normalize(mean=[0.5], std=[0.5])])


def val_chain():
    return torchvision.transforms.Compose([ScaleImageTensor(), torchvision.transforms.Normalize(mean=[0.5], std=[0.5])])


def window(tag, rng, hw, pad, n_store, start, length, S, draw):
    """One reference window [start, start+length) of an episode store, padded to S."""
    H = W = hw
    store = rng.integers(0, 256, (n_store, H, W, 3), dtype=np.uint8)
    rel = rng.uniform(-1, 1, (n_store, 7)).astype(np.float32)
    robot = rng.uniform(-1, 1, (n_store, 15)).astype(np.float32)
    scene = rng.uniform(-1, 1, (n_store, 24)).astype(np.float32)
    episode = {k: v[start : start + length] for k, v in (("rgb_static", store), ("rel_actions", rel), ("robot_obs", robot), ("scene_obs", scene))}
    obs_space = {"rgb_obs": ["rgb_static"], "actions": ["rel_actions"], "state_obs": ["robot_obs"]}
    om = sys.modules["omegaconf"]
    proprio = om.DictConfig(normalize=True, normalize_robot_orientation=True, keep_indices=[[0, 7], [14, 15]], robot_orientation_idx=[3, 6])
    orig = torch.randint
    if draw is not None:
        d = torch.from_numpy(draw.astype(np.float32)).reshape(length, 1, 1, 2)
        torch.randint = lambda *a, **k: d.clone()
    try:
        rgb = process_rgb(episode, obs_space, {"rgb_static": chain(pad) if draw is not None else val_chain()})
    finally:
        torch.randint = orig
    seq = {**rgb, "depth_obs": {}, **process_actions(episode, obs_space, {}), **get_state_info_dict(episode),
           **process_state(episode, obs_space, {}, proprio)}
    fake = types.SimpleNamespace(save_format="npz", relative_actions=True, pad_with_repetition=BaseDataset.pad_with_repetition,
                                 pad_with_zeros=BaseDataset.pad_with_zeros)
    if S > length:
        seq = BaseDataset.pad_sequence(fake, seq, S - length)
    G[f"{tag}/store"] = store
    G[f"{tag}/rel_actions"] = rel
    G[f"{tag}/robot_obs_raw"] = robot
    G[f"{tag}/scene_obs"] = scene
    G[f"{tag}/meta"] = np.array([start, length, S, pad], dtype=np.int64)
    if draw is not None:
        G[f"{tag}/shift_draw"] = draw.astype(np.int64)          # the reference's randint in [0, 2*pad]: (sx, sy) per frame
    G[f"{tag}/out/rgb_static"] = seq["rgb_obs"]["rgb_static"].numpy()
    G[f"{tag}/out/actions"] = seq["actions"].numpy()
    G[f"{tag}/out/robot_obs"] = seq["robot_obs"].numpy()
    G[f"{tag}/out/state_robot_obs"] = seq["state_info"]["robot_obs"].numpy()
    G[f"{tag}/out/state_scene_obs"] = seq["state_info"]["scene_obs"].numpy()


def main():
    rng = np.random.default_rng(11)
    # train transforms, ragged window (5 valid steps padded to 8), 24x24 frames, pad 3
    window("train24", rng, 24, 3, n_store=9, start=2, length=5, S=8, draw=rng.integers(0, 7, (5, 2)))
    # train transforms at the gripper camera's real geometry (84x84, pad 4), full window of 2
    window("train84", rng, 84, 4, n_store=3, start=1, length=2, S=2, draw=rng.integers(0, 9, (2, 2)))
    # extreme draws: (0,0) and (2*pad, 2*pad) hit the replicate border on both sides
    window("edge24", rng, 24, 3, n_store=4, start=0, length=4, S=4, draw=np.array([[0, 0], [6, 6], [0, 6], [3, 3]]))
    # validation transforms (no augmentation), single valid step padded to 4
    window("val24", rng, 24, 3, n_store=3, start=2, length=1, S=4, draw=None)
    # how far the reference's fp32 grid_sample is from the exact integer crop, at the static camera's geometry (200x200, pad 10)
    x = torch.from_numpy(rng.integers(0, 256, (4, 3, 200, 200), dtype=np.uint8))
    draw = rng.integers(0, 21, (4, 2))
    orig = torch.randint
    torch.randint = lambda *a, **k: torch.from_numpy(draw.astype(np.float32)).reshape(4, 1, 1, 2)
    try:
        y = chain(10)(x).numpy()
    finally:
        torch.randint = orig
    from oracle.datamodule_oracle import frames_u8_to_f32

    exact = frames_u8_to_f32(x.permute(0, 2, 3, 1).contiguous().numpy(), shift=draw - 10)
    G["aug_ref_max_dev"] = np.array(np.abs(y - exact).max(), dtype=np.float64)
    out = os.path.join(ROOT, "tests", "golden", "datamodule_golden.npz")
    np.savez_compressed(out, **G)
    print(f"wrote {out}: {len(G)} arrays, {os.path.getsize(out) / 1e3:.0f} KB; reference-vs-crop max deviation {float(G['aug_ref_max_dev']):.3e}")


if __name__ == "__main__":
    main()
