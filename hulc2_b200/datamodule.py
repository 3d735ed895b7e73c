"""On-device datamodule: the window producer either side of the policy step.

The reference feeds ``Hulc2.training_step`` from CPU dataloader workers: each worker loads ``window`` npz files, stacks
uint8 HWC frames, runs the transform chain in fp32 and pads the window (``hulc2/datasets/npz_dataset.py:117-143``,
``base_dataset.py:93-163``, ``utils/episode_utils.py:61-86``, ``hulc2_sim_data_module.py:23-137``); the fp32 batch
(18 MB per window) then crosses PCIe.  Here an episode store stays **uint8 in HBM** (4.5 MB per window-equivalent, a
2.4 M-frame CALVIN split D is 340 GB over 8 GPUs = 42 GB each) and a batch is *described*, not materialised: window
starts/lengths and the augmentation draw are tiny index tensors, camera frames are ``ops.U8Frames`` views of the store
that the conv trunk's pack kernel gathers, shifts, normalises and re-tiles in one pass (``csrc/frames.cu``).  Per-step
vectors (actions, robot/scene state) are gathered and padded by ``hulc2_window_gather_f32``.

Host-side index logic mirrors the reference: ``episode_lookup`` construction (``npz_dataset.py:187-231``), the window
size draw bounded by the episode end (``:61-84``), language annotations -> (lookup, ``use_for_aux_lang_loss``)
(``:146-197,233-241``).  Batches are the SURVEY 8b dict, so ``training_step`` / ``validation_step`` consume them unchanged.
There is no CPU path: the store lives on a CUDA device.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import call

PAD_REPEAT, PAD_ZEROS, PAD_REL_ACTIONS = 0, 1, 2


# ----------------------------------------------------------------------------- host-side index logic
def build_episode_lookup(ep_start_end_ids: Sequence[Tuple[int, int]], min_window_size: int, max_window_size: int) -> np.ndarray:
    """npz_dataset.py:199-231 (load_file_indices): every frame that can start a window of at least ``min_window_size``."""
    out: List[int] = []
    for start_idx, end_idx in ep_start_end_ids:
        if not end_idx > max_window_size:
            raise ValueError("episode shorter than max_window_size (reference asserts end_idx > max_window_size)")
        out.extend(range(int(start_idx), int(end_idx) + 1 - min_window_size))
    return np.asarray(out, dtype=np.int64)


def build_lang_lookup(ann_start_end: Sequence[Tuple[int, int]], min_window_size: int, max_window_size: int, skip_frames: int = 1,
                      pretrain: bool = False, aux_lang_loss_window: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """npz_dataset.py:146-197 (load_file_indices_lang): (episode_lookup, lang_lookup) for annotated segments."""
    episode_lookup: List[int] = []
    lang_lookup: List[int] = []
    for i, (start_idx, end_idx) in enumerate(ann_start_end):
        start_idx, end_idx = int(start_idx), int(end_idx)
        if pretrain:
            start_idx = max(start_idx, end_idx + 1 - min_window_size - aux_lang_loss_window)
        if not end_idx >= max_window_size:
            raise ValueError("annotation ends before max_window_size (reference asserts end_idx >= max_window_size)")
        cnt = 0
        for idx in range(start_idx, end_idx + 1 - min_window_size):
            if cnt % skip_frames == 0:
                lang_lookup.append(i)
                episode_lookup.append(idx)
            cnt += 1
    return np.asarray(episode_lookup, dtype=np.int64), np.asarray(lang_lookup, dtype=np.int64)


def max_window_at(episode_lookup: np.ndarray, idx: int, min_window_size: int, max_window_size: int) -> int:
    """npz_dataset.py:61-79 (get_window_size, before the random draw): the longest window starting at lookup entry
    ``idx`` that stays inside its episode."""
    window_diff = max_window_size - min_window_size
    n = len(episode_lookup)
    if n <= idx + window_diff:
        return min_window_size + n - idx - 1
    if episode_lookup[idx + window_diff] != episode_lookup[idx] + window_diff:
        seg = episode_lookup[idx : idx + window_diff + 1] - (episode_lookup[idx] + np.arange(window_diff + 1))
        steps_to_next_episode = min_window_size + int(np.nonzero(seg)[0][0]) - 1
        return min(max_window_size, steps_to_next_episode)
    return max_window_size


def use_for_aux_lang_loss(lang_lookup: np.ndarray, idx: np.ndarray, aux_lang_loss_window: int = 1) -> np.ndarray:
    """npz_dataset.py:233-241 (add_language_info)."""
    idx = np.asarray(idx, dtype=np.int64)
    n = len(lang_lookup)
    nxt = np.minimum(idx + aux_lang_loss_window, n - 1)
    return (idx + aux_lang_loss_window >= n) | (lang_lookup[idx] < lang_lookup[nxt])


# ----------------------------------------------------------------------------- the store
class DeviceEpisodeStore:
    """Frames and per-step vectors of a dataset split, resident on one CUDA device.

    ``rgb``: camera name -> uint8 ``[N,H,W,3]`` (HWC, as stored in the episode files); ``rel_actions [N,7]``,
    ``robot_obs [N,15]`` (raw), ``scene_obs [N,24]`` fp32.  ``proprio_keep`` are the reference's
    ``proprio_state.keep_indices`` slices of the (optionally normalised) robot state that form ``batch["robot_obs"]``."""

    def __init__(self, rgb: Dict[str, torch.Tensor], rel_actions, robot_obs, scene_obs=None, device="cuda",
                 proprio_keep: Sequence[Tuple[int, int]] = ((0, 7), (14, 15)), robot_mean=None, robot_std=None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("hulc2_b200: the episode store lives in HBM; there is no CPU datamodule path")

        def dv(x, dtype):
            t = torch.as_tensor(x)
            return t.to(device=dev, dtype=dtype).contiguous()

        self.device = dev
        self.rgb = {k: dv(v, torch.uint8) for k, v in rgb.items()}
        for k, v in self.rgb.items():
            if v.dim() != 4:
                raise ValueError(f"camera '{k}' must be uint8 [N,H,W,C]")
        self.N = next(iter(self.rgb.values())).shape[0]
        self.rel_actions = dv(rel_actions, torch.float32)
        self.robot_obs = dv(robot_obs, torch.float32)
        self.scene_obs = None if scene_obs is None else dv(scene_obs, torch.float32)
        prop = self.robot_obs
        if robot_mean is not None:                      # NormalizeVector (transforms.py:37-50): std 0 -> 1
            std = dv(robot_std, torch.float32).clone()
            std[std == 0.0] = 1.0
            prop = (prop - dv(robot_mean, torch.float32)) / std
        self.proprio = torch.cat([prop[:, a:b] for a, b in proprio_keep], dim=1).contiguous()

    def write_frames(self, start: int, rgb: Dict[str, torch.Tensor], rel_actions=None, robot_obs=None, scene_obs=None) -> None:
        """Streaming ingestion: overwrite store rows ``[start, start + n)`` with new steps (uint8 HWC frames per camera and,
        optionally, their per-step vectors) from (pinned) host memory, asynchronously on the CURRENT stream.  With a ring of
        frames resident in HBM a training step only has to move the frames that are new since the previous step across PCIe
        -- consecutive windows of play data share 31 of their 32 frames -- while batches stay index descriptors."""
        n = None
        for k, v in rgb.items():
            n = v.shape[0] if n is None else n
            if start < 0 or start + v.shape[0] > self.N or v.shape[1:] != self.rgb[k].shape[1:] or v.dtype != torch.uint8:
                raise ValueError(f"write_frames: camera '{k}' rows [{start}, {start + v.shape[0]}) do not fit the store")
            self.rgb[k][start : start + v.shape[0]].copy_(v, non_blocking=True)
        for name, v in (("rel_actions", rel_actions), ("robot_obs", robot_obs), ("scene_obs", scene_obs)):
            if v is not None:
                getattr(self, name)[start : start + v.shape[0]].copy_(v, non_blocking=True)

    def gather(self, table: torch.Tensor, win_start: torch.Tensor, win_len: Optional[torch.Tensor], S: int, mode: int) -> torch.Tensor:
        B, D = win_start.numel(), table.shape[1]
        out = torch.empty(B, S, D, device=self.device, dtype=torch.float32)
        call("hulc2_window_gather_f32", table.data_ptr(), win_start.data_ptr(), None if win_len is None else win_len.data_ptr(),
             out.data_ptr(), B, S, D, mode)
        return out

    def window_batch(self, win_start: torch.Tensor, win_len: Optional[torch.Tensor], S: int,
                     shifts: Optional[Dict[str, torch.Tensor]] = None) -> dict:
        """One modality's batch dict for windows ``[win_start[b], win_start[b]+win_len[b])`` padded to S."""
        win_start = win_start.to(device=self.device, dtype=torch.int64).contiguous()
        win_len = None if win_len is None else win_len.to(device=self.device, dtype=torch.int32).contiguous()
        shifts = shifts or {}
        rgb = {k: ops.U8Frames(v, shifts.get(k), win_start, win_len, S) for k, v in self.rgb.items()}
        d = {
            "rgb_obs": rgb,
            "depth_obs": {},
            "robot_obs": self.gather(self.proprio, win_start, win_len, S, PAD_REPEAT),
            "actions": self.gather(self.rel_actions, win_start, win_len, S, PAD_REL_ACTIONS),
            "state_info": {"robot_obs": self.gather(self.robot_obs, win_start, win_len, S, PAD_REPEAT)},
            "idx": win_start,
        }
        if self.scene_obs is not None:
            d["state_info"]["scene_obs"] = self.gather(self.scene_obs, win_start, win_len, S, PAD_REPEAT)
        return d


    def batch_from_descriptors(self, desc: Dict[str, dict], S: int) -> Dict[str, dict]:
        """Device-side collate: ``desc[mod] = {"win_start" [B] int64, optional "win_len" [B] int32, optional
        "shift_<camera>" [B,S,2] int32, and for language modalities "lang" [B,384] + "use_for_aux_lang_loss" [B] bool}`` (all
        device tensors) -> the SURVEY 8b batch dict: cameras as ``ops.U8Frames`` windows of the resident store, per-step
        vectors gathered and padded by ``hulc2_window_gather_f32``.  Only library kernels run, so a trainer can capture it
        together with the step (``PolicyTrainer(collate=...)``) and replay it on refreshed descriptors."""
        out = {}
        for mod, d in desc.items():
            shifts = {k[len("shift_"):]: v for k, v in d.items() if k.startswith("shift_")}
            b = self.window_batch(d["win_start"], d.get("win_len"), S, shifts)
            for k in ("lang", "use_for_aux_lang_loss"):
                if k in d:
                    b[k] = d[k]
            out[mod] = b
        return out


# ----------------------------------------------------------------------------- loaders
class WindowLoader:
    """Iterable over one modality's batches (what ``DataLoader(NpzDataset)`` yields in the reference), with
    ``batch_size``, ``min_window_size``/``max_window_size``, ``pad=True`` semantics of ``BaseDataset``."""

    def __init__(self, store: DeviceEpisodeStore, episode_lookup: np.ndarray, batch_size: int, min_window_size: int = 16,
                 max_window_size: int = 32, train: bool = True, shift_pad: Optional[Dict[str, int]] = None, seed: int = 0,
                 lang_lookup: Optional[np.ndarray] = None, lang_emb: Optional[torch.Tensor] = None, aux_lang_loss_window: int = 1,
                 shuffle: Optional[bool] = None, drop_last: Optional[bool] = None, rank: Optional[int] = None, world: Optional[int] = None):
        self.store, self.lookup, self.B = store, np.asarray(episode_lookup, dtype=np.int64), int(batch_size)
        self.min_ws, self.max_ws, self.train = int(min_window_size), int(max_window_size), bool(train)
        self.shift_pad = dict(shift_pad or {}) if train else {}
        # Data parallel (the reference relies on Lightning injecting a DistributedSampler): every rank walks the SAME permutation
        # (order_rng, seeded identically) and takes entries rank, rank + world, ... of it, padded to an equal count; the window-size
        # and RandomShiftsAug draws come from a per-rank stream.
        if rank is None or world is None:
            import torch.distributed as dist

            on = dist.is_available() and dist.is_initialized()
            rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        self.rank, self.world = int(rank), int(world)
        self.order_rng = np.random.default_rng(seed)
        self.rng = np.random.default_rng([seed, self.rank])
        self.lang_lookup = None if lang_lookup is None else np.asarray(lang_lookup, dtype=np.int64)
        self.lang_emb = None if lang_emb is None else torch.as_tensor(lang_emb).to(store.device, torch.float32)
        self.aux_window = aux_lang_loss_window
        self.shuffle = bool(train) if shuffle is None else bool(shuffle)
        # training batches feed a captured step with static shapes: the ragged last batch is dropped by default
        self.drop_last = bool(train) if drop_last is None else bool(drop_last)
        if len(self.lookup) and int(self.lookup.max()) + self.max_ws > store.N + self.max_ws - self.min_ws:
            raise ValueError("episode_lookup points past the store: a window of min_window_size would read out of bounds")

    def _shard_len(self) -> int:
        return (len(self.lookup) + self.world - 1) // self.world

    def __len__(self) -> int:
        n = self._shard_len()
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def window_size(self, idx: int) -> int:
        if self.min_ws == self.max_ws:
            return self.max_ws
        mx = max_window_at(self.lookup, idx, self.min_ws, self.max_ws)
        if not self.train:       # the reference hashes idx (pyhash fnv1_32, absent here); any deterministic choice in range
            return self.min_ws + (idx * 2654435761 % 2**32) % (mx - self.min_ws + 1)
        return int(self.rng.integers(self.min_ws, mx + 1))

    def describe(self, ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray]]:
        """Index tensors of one batch: window starts, valid lengths, RandomShiftsAug offsets per camera [B,S,2] (padded
        steps repeat the last valid frame INCLUDING its shift, as the reference pads after the transform)."""
        starts = self.lookup[ids]
        lens = np.asarray([self.window_size(int(i)) for i in ids], dtype=np.int32)
        shifts = {}
        S = self.max_ws
        for cam, pad in self.shift_pad.items():
            draw = self.rng.integers(0, 2 * pad + 1, (len(ids), S, 2)) - pad
            last = np.take_along_axis(draw, (lens.astype(np.int64) - 1)[:, None, None].repeat(2, axis=2), axis=1)
            t = np.arange(S)[None, :, None]
            shifts[cam] = np.where(t < lens[:, None, None], draw, last).astype(np.int32)
        return starts, lens, shifts

    def __iter__(self) -> Iterator[dict]:
        order = self.order_rng.permutation(len(self.lookup)) if self.shuffle else np.arange(len(self.lookup))
        if self.world > 1:
            n = self._shard_len() * self.world                   # pad by wrapping so that every rank gets the same count
            order = np.resize(order, n)[self.rank :: self.world]
        for b in range(len(self)):
            ids = order[b * self.B : (b + 1) * self.B]
            starts, lens, shifts = self.describe(ids)
            dev = self.store.device
            batch = self.store.window_batch(torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev), self.max_ws,
                                            {k: torch.from_numpy(v).to(dev) for k, v in shifts.items()})
            batch["idx"] = torch.from_numpy(ids).to(dev)
            if self.lang_lookup is not None:
                batch["lang"] = self.lang_emb[torch.from_numpy(self.lang_lookup[ids]).to(dev)]
                batch["use_for_aux_lang_loss"] = torch.from_numpy(use_for_aux_lang_loss(self.lang_lookup, ids, self.aux_window)).to(dev)
            yield batch


class CombinedLoader:
    """``{"vis": loader, "lang": loader}`` -> ``{"vis": batch, "lang": batch}`` per step; the shorter loader cycles
    (Lightning's ``max_size_cycle``, which is what the trainer applies to the dict ``train_dataloader`` returns)."""

    def __init__(self, loaders: Dict[str, WindowLoader]):
        self.loaders = loaders

    def __len__(self) -> int:
        return max(len(v) for v in self.loaders.values())

    def __iter__(self):
        its = {k: iter(v) for k, v in self.loaders.items()}
        for _ in range(len(self)):
            out = {}
            for k in self.loaders:
                b = next(its[k], None)
                if b is None:
                    its[k] = iter(self.loaders[k])
                    b = next(its[k])
                out[k] = b
            yield out


class Hulc2DeviceDataModule:
    """Drop-in for ``Hulc2SimdDataModule`` (hulc2_sim_data_module.py:23-137) over device-resident stores: same
    ``setup`` / ``train_dataloader`` / ``val_dataloader`` / ``modalities`` surface.  ``datasets`` maps the modality key
    (``"vis"``, ``"lang"``) to ``dict(batch_size, min_window_size, max_window_size, [skip_frames, pretrain,
    aux_lang_loss_window])`` as in ``conf/datamodule/datasets/*.yaml``; ``train``/``val`` are dicts with a
    ``DeviceEpisodeStore`` under ``"store"``, ``"ep_start_end_ids"`` and, for language, ``"lang_start_end"`` +
    ``"lang_emb"`` (the ``auto_lang_ann.npy`` fields ``info.indx`` and ``language.emb``).  ``shift_pad`` is the
    RandomShiftsAug pad per camera from the transform config (rand_shift.yaml: static 10, gripper 4)."""

    def __init__(self, datasets: Dict[str, dict], train: dict, val: Optional[dict] = None, shift_pad: Optional[Dict[str, int]] = None,
                 seed: int = 0, **kwargs):
        self.datasets_cfg, self.train_split, self.val_split = datasets, train, val
        self.shift_pad = {"rgb_static": 10, "rgb_gripper": 4} if shift_pad is None else shift_pad
        self.seed = seed
        self.modalities: List[str] = []
        self.train_datasets: Dict[str, WindowLoader] = {}
        self.val_datasets: Dict[str, WindowLoader] = {}

    def prepare_data(self, *args, **kwargs):
        pass

    def _loader(self, key: str, cfg: dict, split: dict, train: bool) -> WindowLoader:
        mn, mx = cfg.get("min_window_size", 16), cfg.get("max_window_size", 32)
        pads = {k: v for k, v in self.shift_pad.items() if k in split["store"].rgb}
        if "lang" in key:
            look, ll = build_lang_lookup(split["lang_start_end"], mn, mx, cfg.get("skip_frames", 1), cfg.get("pretrain", False),
                                         cfg.get("aux_lang_loss_window", 1))
            return WindowLoader(split["store"], look, cfg["batch_size"], mn, mx, train, pads, self.seed + 1, ll, split["lang_emb"],
                                cfg.get("aux_lang_loss_window", 1))
        look = build_episode_lookup(split["ep_start_end_ids"], mn, mx)
        return WindowLoader(split["store"], look, cfg["batch_size"], mn, mx, train, pads, self.seed)

    def setup(self, stage=None):
        self.modalities = []
        for key, cfg in self.datasets_cfg.items():
            self.train_datasets[key] = self._loader(key, cfg, self.train_split, True)
            if self.val_split is not None:
                self.val_datasets[key] = self._loader(key, cfg, self.val_split, False)
            self.modalities.append(key)

    def train_dataloader(self):
        return CombinedLoader(self.train_datasets)

    def val_dataloader(self):
        return CombinedLoader(self.val_datasets)


def synthetic_store(n_frames: int, device="cuda", static_hw=(200, 200), gripper_hw=(84, 84), seed: int = 0) -> DeviceEpisodeStore:
    """CALVIN-shaped random episode frames drawn on the device (bench / tests; there is no dataset in the image)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rgb = {"rgb_static": torch.randint(0, 256, (n_frames, *static_hw, 3), generator=g, device=dev, dtype=torch.uint8),
           "rgb_gripper": torch.randint(0, 256, (n_frames, *gripper_hw, 3), generator=g, device=dev, dtype=torch.uint8)}

    def U(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, device=dev) * (hi - lo) + lo

    rel = U(n_frames, 7)
    e = torch.rand(n_frames, 6, generator=g, device=dev)
    rel[:, :6] = torch.where(e < 0.01, -1.0, torch.where(e > 0.99, 1.0, rel[:, :6]))
    rel[:, 6] = torch.where(torch.rand(n_frames, generator=g, device=dev) < 0.5, -1.0, 1.0)
    robot = U(n_frames, 15)
    robot[:, 3:6] = U(n_frames, 3, lo=-0.9 * np.pi / 2, hi=0.9 * np.pi / 2)
    return DeviceEpisodeStore(rgb, rel, robot, U(n_frames, 24), device=dev)
