"""Generates tests/golden/hulc2_golden.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference modules are imported through oracle/ref_import.py (SURVEY.md 8c shims) and
executed on deterministic synthetic inputs / weights / noise (hulc2_b200.synthetic, numpy
PCG64 seeds recorded below).  Big inputs (images) are NOT stored -- they are regenerated
from the seed; small op-level inputs are stored beside their outputs.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from hulc2_b200.config import hulc2_config  # noqa: E402
from hulc2_b200.synthetic import synthetic_batch, synthetic_obs, synthetic_state_dict  # noqa: E402
from oracle.ref_import import make_reference_model  # noqa: E402
from oracle.ref_noise import supplied_categories, supplied_uniforms  # noqa: E402

NON_LEARNED = (
    "x_map", "y_map", "temperature", "one_hot_embedding_eye", ".ones", "gripper_bounds",
    "action_max_bound", "action_min_bound",
)
G = {}


def put(name, t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    G[name] = np.asarray(t)


def build(variant, static_hw, dropout_p=0.0):
    cfg = hulc2_config(pkg="hulc2", variant=variant, static_hw=static_hw, dropout_p=dropout_p)
    m = make_reference_model(cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    m.load_state_dict(synthetic_state_dict(shapes, seed=0, skip=NON_LEARNED), strict=False)
    return m


def train_case(tag, variant, static_hw, B, aux):
    m = build(variant, static_hw)
    batch = synthetic_batch(B, seed=1, static_hw=static_hw, aux=aux)
    g = torch.Generator().manual_seed(5)
    idx = {mod: torch.randint(0, 32, (B, 32), generator=g) for mod in batch}
    with supplied_categories([idx[mod] for mod in batch]):
        loss = m.training_step(batch, 0)
    loss.backward()
    put(f"{tag}/loss", loss)
    for k, v in m.logged.items():
        put(f"{tag}/log/{k}", v)
    for mod in batch:
        put(f"{tag}/plan_idx/{mod}", idx[mod])
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        put(f"{tag}/gnorm/{n}", p.grad.norm())
        if p.numel() <= 2048:
            put(f"{tag}/grad/{n}", p.grad)
    with torch.no_grad():
        emb = m.perceptual_encoder(batch["vis"]["rgb_obs"], batch["vis"]["depth_obs"], batch["vis"]["robot_obs"])
    put(f"{tag}/perceptual_emb_vis", emb)
    return m


def rollout_case(m, tag, N, steps):
    m.eval()
    obs, goal = synthetic_obs(N, seed=2)
    g = torch.Generator().manual_seed(7)
    m.reset()
    m.replan_freq = 2  # exercise the replan branch inside a short trace
    for s in range(steps):
        pidx = torch.randint(0, 32, (N, 32), generator=g)
        u1 = torch.rand(N, 1, 6, 10, generator=g)
        u2 = torch.rand(N, 1, 6, generator=g)
        cats = [pidx] if s % m.replan_freq == 0 else []
        with supplied_categories(cats), supplied_uniforms([u1, u2]):
            a = m.step(obs, goal)
        put(f"{tag}/step{s}/plan_idx", pidx)
        put(f"{tag}/step{s}/u1", u1)
        put(f"{tag}/step{s}/u2", u2)
        put(f"{tag}/step{s}/action", a)
    m.replan_freq = 30


def val_case(m, tag, B):
    m.eval()

    class _DM:
        modalities = ["vis", "lang"]

    class _Tr:
        datamodule = _DM()

    m.trainer = _Tr()
    batch = synthetic_batch(B, seed=3, aux="all")
    g = torch.Generator().manual_seed(11)
    cats, unis = [], []
    for mod in batch:
        ipp = torch.randint(0, 32, (B, 32), generator=g)
        ipr = torch.randint(0, 32, (B, 32), generator=g)
        u = [torch.rand(B, 32, 6, 10, generator=g), torch.rand(B, 32, 6, generator=g),
             torch.rand(B, 32, 6, 10, generator=g), torch.rand(B, 32, 6, generator=g)]
        cats += [ipp, ipr]
        unis += u
        put(f"{tag}/{mod}/plan_idx_pp", ipp)
        put(f"{tag}/{mod}/plan_idx_pr", ipr)
        for i, n in enumerate(("u1_pp", "u2_pp", "u1_pr", "u2_pr")):
            put(f"{tag}/{mod}/{n}", u[i])
    with torch.no_grad(), supplied_categories(cats), supplied_uniforms(unis):
        out = m.validation_step(batch, 0)
    for k, v in m.logged.items():
        if k.startswith("val"):
            put(f"{tag}/log/{k}", v)
    for k, v in out.items():
        put(f"{tag}/out/{k}", v)


def op_cases(m):
    """Op-level fixtures with their (small) inputs stored."""
    g = torch.Generator().manual_seed(13)
    m.zero_grad()
    dec = m.action_decoder
    B, S = 3, 5
    lp = torch.randn(B, S, 6, 10, generator=g)
    ls = torch.randn(B, S, 6, 10, generator=g) * 3 - 2  # spans the -7 clamp
    mu = torch.randn(B, S, 6, 10, generator=g) * 0.5
    grip = torch.randn(B, S, 2, generator=g)
    act = torch.rand(B, S, 7, generator=g) * 2 - 1
    act[0, 0, 0], act[0, 0, 1], act[0, 1, 2], act[1, 0, 3] = -1.0, 1.0, 0.9995, -0.9995
    act[..., 6] = torch.where(torch.rand(B, S, generator=g) < 0.5, -1.0, 1.0)
    ls[2, 4] = -9.0  # below clamp: tiny scale -> cdf_delta < 1e-5 branch for far means
    mu[2, 4] = 5.0
    for n, t in (("lp", lp), ("ls", ls), ("mu", mu), ("grip", grip), ("act", act)):
        put(f"op/logistic/{n}", t)
    lpr, lsr, mur, gr = (t.clone().requires_grad_() for t in (lp, ls, mu, grip))
    loss = dec._loss(lpr, torch.clamp(lsr, min=-7.0), mur, gr, act)
    loss.backward()
    put("op/logistic/loss", loss)
    put("op/logistic/d_lp", lpr.grad)
    put("op/logistic/d_ls", lsr.grad)
    put("op/logistic/d_mu", mur.grad)
    put("op/logistic/d_grip", gr.grad)
    u1, u2 = torch.rand(B, S, 6, 10, generator=g), torch.rand(B, S, 6, generator=g)
    with supplied_uniforms([u1, u2]):
        smp = dec._sample(lp, torch.clamp(ls, min=-7.0), mu, grip)
    put("op/sample/u1", u1)
    put("op/sample/u2", u2)
    put("op/sample/out", smp)

    from hulc2.models.decoders.utils.gripper_control import tcp_to_world_frame, world_to_tcp_frame

    robot = torch.rand(B, S, 15, generator=g) * 2 - 1
    robot[..., 3:6] = (torch.rand(B, S, 3, generator=g) * 2 - 1) * 1.4
    put("op/frames/robot_obs", robot)
    put("op/frames/world_to_tcp", world_to_tcp_frame(act, robot))
    put("op/frames/tcp_to_world", tcp_to_world_frame(act, robot))

    pp = (torch.randn(B, 1024, generator=g) * 2).requires_grad_()
    pr = (torch.randn(B, 1024, generator=g) * 2).requires_grad_()
    from hulc2.utils.distributions import DiscState

    kl = m.compute_kl_loss(DiscState(pp), DiscState(pr))
    kl.backward()
    put("op/kl/pp", pp)
    put("op/kl/pr", pr)
    put("op/kl/loss", kl)
    put("op/kl/d_pp", pp.grad)
    put("op/kl/d_pr", pr.grad)

    Bc = 6
    sf = (torch.randn(Bc, 4096, generator=g)).requires_grad_()
    gl = (torch.randn(Bc, 32, generator=g)).requires_grad_()
    use = torch.tensor([True, False, True, True, False, True])
    cl = m.clip_auxiliary_loss(sf, gl, use)
    cl.backward()
    put("op/clip/seq_feat", sf)
    put("op/clip/goal", gl)
    put("op/clip/use", use)
    put("op/clip/loss", cl)
    put("op/clip/d_goal", gl.grad)
    put("op/clip/d_seq_feat_norm", sf.grad.norm(dim=1))
    put("op/clip/d_logit_scale", m.logit_scale.grad)
    m.zero_grad()

    x = torch.randn(2, 64, 21, 21, generator=g)
    put("op/ssm/x", x)
    put("op/ssm/out", m.perceptual_encoder.rgb_static_encoder.spatial_softmax(x))

    emb = torch.randn(2, 32, 128, generator=g)
    st, seq = m.plan_recognition(emb)
    put("op/pr/emb", emb)
    put("op/pr/logit", st.logit)
    put("op/pr/seq_feat_head", seq[:, :64])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    m = train_case("calvin_B2", "calvin", (200, 200), 2, "half")
    op_cases(m)
    rollout_case(m, "rollout_N4", 4, 4)
    val_case(m, "val_B2", 2)
    train_case("rw_B2", "real_world", (150, 200), 2, "all")
    put("meta/seeds", np.array([0, 1, 2, 3, 5, 7, 11, 13]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hulc2_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, len(G), "arrays", os.path.getsize(out) / 1e6, "MB")


if __name__ == "__main__":
    main()
