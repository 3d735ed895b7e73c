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


def gimbal_pole_cases(n: int = 400000, seed: int = 1, keep: int = 64):
    """(action [1,K,7], robot_obs [1,K,15]) near the XYZ-Euler gimbal pole (pitch within 0.01 rad of +-pi/2), K = `keep` rows of
    which the first ones are chosen so that the reference's fp32 asin argument exceeds 1 by one ulp -- the inputs for which
    tcp_to_world_frame takes its quaternion NaN fallback (gripper_control.py:51-55) -- followed by ordinary near-pole rows."""
    import math

    from oracle import hulc2_oracle as O

    g = torch.Generator().manual_seed(seed)
    rob = torch.zeros(1, n, 15)
    rob[..., 3] = torch.rand(n, generator=g) * 2 - 1
    sgn = torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    rob[..., 4] = sgn * (math.pi / 2 - torch.rand(n, generator=g) * 0.01)
    rob[..., 5] = torch.rand(n, generator=g) * 2 - 1
    act = torch.rand(1, n, 7, generator=g) * 2 - 1
    w_T = O.euler_xyz_to_matrix(rob[..., 3:6]).float().view(-1, 3, 3)
    rel = O.euler_xyz_to_matrix(act[..., 3:6] * 0.01).float().view(-1, 3, 3)
    W = w_T @ torch.inverse(rel)
    bad = O.matrix_to_euler_xyz(W).isnan().any(-1)
    # the quaternion round trip rescues most of them; for a few the re-normalised matrix STILL has |m02| = 1 + ulp and the
    # reference dies on its `assert not isnan` (gripper_control.py:62) -- those are returned separately (`fatal`)
    again = O.matrix_to_euler_xyz(O.quaternion_to_matrix(O.matrix_to_quaternion(W))).isnan().any(-1)
    nan_rows = (bad & ~again).nonzero().flatten()[: keep // 2]
    rest = (~bad).nonzero().flatten()[: keep - len(nan_rows)]
    sel = torch.cat([nan_rows, rest])
    fatal = (bad & again).nonzero().flatten()[:8]
    return act[:, sel].clone(), rob[:, sel].clone(), int(len(nan_rows)), (act[:, fatal].clone(), rob[:, fatal].clone())


GRAD_SAMPLE = 4096


def grad_sample_index(name: str, numel: int) -> np.ndarray:
    """Seeded flat positions at which big gradients are compared with the fp64 fixture (tests/golden/make_grad_noise_floor.py)."""
    if numel <= GRAD_SAMPLE:
        return np.arange(numel, dtype=np.int64)
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2**31)
    return np.sort(np.random.default_rng(seed).choice(numel, GRAD_SAMPLE, replace=False)).astype(np.int64)


# ----------------------------------------------------------------------------- alternate blocks
ALT_GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hulc2_alt_golden.npz")
# tag -> config kwargs; B=2 windows per modality, hidden_size 256 keeps fixture generation and the oracle check fast
ALT_CASES = {
    "gru": dict(variant="calvin", static_hw=(200, 200), rnn_model="gru_decoder", distribution="discrete"),
    "lstm": dict(variant="calvin", static_hw=(200, 200), rnn_model="lstm_decoder", distribution="discrete"),
    "gauss_rw": dict(variant="real_world", static_hw=(150, 200), rnn_model="rnn_decoder", distribution="continuous"),
    "gauss_gru": dict(variant="calvin", static_hw=(200, 200), rnn_model="gru_decoder", distribution="continuous"),
    "rgbd_rw": dict(variant="real_world", static_hw=(150, 200), rnn_model="rnn_decoder", distribution="discrete", depth_static=True),
    "mlp": dict(variant="calvin", static_hw=(200, 200), rnn_model="mlp_decoder", distribution="discrete"),
}
ALT_B, ALT_HIDDEN, ALT_WEIGHT_SEED, ALT_BATCH_SEED, ALT_NOISE_SEED = 2, 256, 5, 17, 19
_alt_golden = None


def alt_gt(name) -> torch.Tensor:
    global _alt_golden
    if _alt_golden is None:
        _alt_golden = np.load(ALT_GOLDEN_PATH)
    return torch.from_numpy(np.asarray(_alt_golden[name]))


def alt_keys(prefix):
    alt_gt(next(iter(np.load(ALT_GOLDEN_PATH).files)))
    return [k for k in _alt_golden.files if k.startswith(prefix)]


def alt_case_inputs(tag):
    """(config kwargs, batch, per-modality plan draw): category indices [B,32] for the discrete plan, standard-normal
    eps [B,256] for the continuous one.  Shared by tests/golden/make_golden_alt.py and the tests."""
    from hulc2_b200.synthetic import synthetic_batch

    kw = dict(ALT_CASES[tag], dropout_p=0.0, hidden_size=ALT_HIDDEN)
    batch = synthetic_batch(ALT_B, seed=ALT_BATCH_SEED, static_hw=kw["static_hw"], aux="all", depth_static=kw.get("depth_static", False))
    g = torch.Generator().manual_seed(ALT_NOISE_SEED)
    if kw["distribution"] == "discrete":
        draw = {mod: torch.randint(0, 32, (ALT_B, 32), generator=g) for mod in batch}
    else:
        draw = {mod: torch.randn(ALT_B, 256, generator=g) for mod in batch}
    return kw, batch, draw


def build_alt_model(tag, device="cpu"):
    kw = dict(ALT_CASES[tag], dropout_p=0.0, hidden_size=ALT_HIDDEN)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = instantiate(hulc2_config(pkg="hulc2_b200", **kw))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    m.load_state_dict(synthetic_state_dict(shapes, seed=ALT_WEIGHT_SEED, skip=NON_LEARNED), strict=False)
    return m.to(device)


ALT_VAL_TAGS = ("lstm", "gauss_gru")


def alt_val_inputs(tag):
    """Validation batch + per-modality noise for the alternate-block fixtures: plan draws (category indices or normal eps)
    for the proposal and recognition plans, and the sampling uniforms of the two loss_and_act calls, in call order."""
    from hulc2_b200.synthetic import synthetic_batch

    kw = dict(ALT_CASES[tag], dropout_p=0.0, hidden_size=ALT_HIDDEN)
    batch = synthetic_batch(ALT_B, seed=ALT_BATCH_SEED + 1, static_hw=kw["static_hw"], aux="all", depth_static=kw.get("depth_static", False))
    g = torch.Generator().manual_seed(ALT_NOISE_SEED + 1)
    noise = {}
    for mod in batch:
        if kw["distribution"] == "discrete":
            d = [torch.randint(0, 32, (ALT_B, 32), generator=g) for _ in range(2)]
        else:
            d = [torch.randn(ALT_B, 256, generator=g) for _ in range(2)]
        u = [torch.rand(ALT_B, 32, 6, 10, generator=g), torch.rand(ALT_B, 32, 6, generator=g),
             torch.rand(ALT_B, 32, 6, 10, generator=g), torch.rand(ALT_B, 32, 6, generator=g)]
        noise[mod] = {"plan_idx_pp": d[0], "plan_idx_pr": d[1], "u1_pp": u[0], "u2_pp": u[1], "u1_pr": u[2], "u2_pr": u[3]}
    return kw, batch, noise
