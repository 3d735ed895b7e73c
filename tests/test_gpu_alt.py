"""GPU parity of the alternate blocks reachable from the same Hydra groups, through the C-ABI:
GRU / LSTM decoder cells (decoders/utils/rnn.py:17-36), continuous latent plan (distributions.py:28-29,55-59,
hulc2.py:444-466) and the RGB-D static encoder (concat_encoders.py:74-80).  Checked against (a) fixtures produced by
the unmodified reference (tests/golden/make_golden_alt.py), (b) the CPU oracle on the same inputs, every parameter
gradient element-wise, and (c) plain torch formulas for the individual kernels."""
import pytest
import torch

from hulc2_b200 import noise, ops
from hulc2_b200._lib import call
from hulc2_b200.config import hulc2_config

from helpers import (ALT_CASES, alt_case_inputs, alt_gt, alt_keys, assert_close, build_alt_model, oracle_params, rel_err,
                     to_device)

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32():
    ops.set_precision("fp32")
    yield
    ops.set_precision("fp32")


def _rand(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


@pytest.mark.parametrize("B,H,with_prev", [(3, 8, True), (128, 2048, True), (5, 260, False)])
def test_gru_cell_fwd_bwd_vs_torch(B, H, with_prev):
    gi, gh = _rand(B, 3 * H, seed=1), _rand(B, 3 * H, seed=2)
    hp = _rand(B, H, seed=3) if with_prev else None
    dh_a, dh_b = _rand(B, H, seed=4), _rand(B, H, seed=5)
    h, save = torch.empty(B, H, device=DEV), torch.empty(B, 4 * H, device=DEV)
    call("hulc2_gru_cell_fwd", gi.data_ptr(), 3 * H, gh.data_ptr(), hp.data_ptr() if with_prev else None, h.data_ptr(), save.data_ptr(), B, H)
    dgi, dgh, dhp = torch.empty(B, 3 * H, device=DEV), torch.empty(B, 3 * H, device=DEV), torch.empty(B, H, device=DEV)
    call("hulc2_gru_cell_bwd", dh_a.data_ptr(), dh_b.data_ptr(), save.data_ptr(), hp.data_ptr() if with_prev else None,
         dgi.data_ptr(), 3 * H, dgh.data_ptr(), dhp.data_ptr(), B, H)
    gi_r, gh_r = gi.double().requires_grad_(), gh.double().requires_grad_()
    hp_r = (hp.double() if with_prev else torch.zeros(B, H, device=DEV, dtype=torch.float64)).requires_grad_()
    r = torch.sigmoid(gi_r[:, :H] + gh_r[:, :H])
    z = torch.sigmoid(gi_r[:, H:2 * H] + gh_r[:, H:2 * H])
    n = torch.tanh(gi_r[:, 2 * H:] + r * gh_r[:, 2 * H:])
    ref = (1 - z) * n + z * hp_r
    ref.backward((dh_a + dh_b).double())
    assert_close(h, ref, 2e-6, "h")
    assert_close(dgi, gi_r.grad, 1e-5, "dgi")
    assert_close(dgh, gh_r.grad, 1e-5, "dgh")
    assert_close(dhp, (dh_a + dh_b).double() * z.detach(), 1e-5, "dh_prev (direct path)")


@pytest.mark.parametrize("B,H,with_prev", [(3, 8, True), (128, 2048, True), (5, 260, False)])
def test_lstm_cell_fwd_bwd_vs_torch(B, H, with_prev):
    gi, gh = _rand(B, 4 * H, seed=1), _rand(B, 4 * H, seed=2)
    cp = _rand(B, H, seed=3) if with_prev else None
    dh_a, dh_b, dc = _rand(B, H, seed=4), _rand(B, H, seed=5), _rand(B, H, seed=6)
    h, c, save = torch.empty(B, H, device=DEV), torch.empty(B, H, device=DEV), torch.empty(B, 4 * H, device=DEV)
    call("hulc2_lstm_cell_fwd", gi.data_ptr(), 4 * H, gh.data_ptr(), cp.data_ptr() if with_prev else None, h.data_ptr(), c.data_ptr(),
         save.data_ptr(), B, H)
    dg, dcp = torch.empty(B, 4 * H, device=DEV), dc.clone()
    call("hulc2_lstm_cell_bwd", dh_a.data_ptr(), dh_b.data_ptr(), dcp.data_ptr(), save.data_ptr(), c.data_ptr(),
         cp.data_ptr() if with_prev else None, dg.data_ptr(), 4 * H, dcp.data_ptr(), B, H)   # dc_prev aliases dc
    g_r = (gi + gh).double().requires_grad_()
    cp_r = (cp.double() if with_prev else torch.zeros(B, H, device=DEV, dtype=torch.float64)).requires_grad_()
    i, f, gg, o = (g_r[:, k * H:(k + 1) * H] for k in range(4))
    c_ref = torch.sigmoid(f) * cp_r + torch.sigmoid(i) * torch.tanh(gg)
    h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
    torch.autograd.backward([h_ref, c_ref], [(dh_a + dh_b).double(), dc.double()])
    assert_close(h, h_ref, 2e-6, "h")
    assert_close(c, c_ref, 2e-6, "c")
    assert_close(dg, g_r.grad, 1e-5, "dgates")
    assert_close(dcp, cp_r.grad, 1e-5, "dc_prev")


def test_gauss_plan_kernels_vs_torch_distributions():
    import torch.distributions as D

    B, P, alpha, beta = 7, 256, 0.8, 0.01
    xq, xp, eps = _rand(B, 2 * P, seed=1, scale=2.0), _rand(B, 2 * P, seed=2, scale=2.0), _rand(B, P, seed=3)
    xq[0, P] = 25.0   # above torch's softplus threshold
    xq_, xp_ = xq.clone().requires_grad_(), xp.clone().requires_grad_()
    mq, sq = ops.GaussStateFunction.apply(xq_)
    mp, sp = ops.GaussStateFunction.apply(xp_)
    plan = ops.GaussRSampleFunction.apply(mp, sp, eps)
    kl = ops.GaussKLFunction.apply(mq, sq, mp, sp, alpha, beta)
    (kl * 3.0 + (plan * plan).sum() * 1e-3).backward()

    xq_r, xp_r = xq.clone().requires_grad_(), xp.clone().requires_grad_()

    def state(x):
        mean, var = torch.chunk(x, 2, dim=-1)
        return mean, torch.nn.functional.softplus(var) + 0.0001

    (mq_r, sq_r), (mp_r, sp_r) = state(xq_r), state(xp_r)
    pp, pr = D.Independent(D.Normal(mq_r, sq_r), 1), D.Independent(D.Normal(mp_r, sp_r), 1)
    pp_d, pr_d = D.Independent(D.Normal(mq_r.detach(), sq_r.detach()), 1), D.Independent(D.Normal(mp_r.detach(), sp_r.detach()), 1)
    kl_r = (alpha * D.kl_divergence(pr_d, pp).mean() + (1 - alpha) * D.kl_divergence(pr, pp_d).mean()) * beta
    plan_r = mp_r + sp_r * eps
    (kl_r * 3.0 + (plan_r * plan_r).sum() * 1e-3).backward()
    assert_close(sq, sq_r, 1e-6, "std")
    assert_close(plan, plan_r, 1e-6, "rsample")
    assert_close(kl, kl_r, 1e-5, "kl")
    assert_close(xq_.grad, xq_r.grad, 2e-5, "d prior state")
    assert_close(xp_.grad, xp_r.grad, 2e-5, "d posterior state")
    # segments: per-modality batch means
    seg = ops.GaussKLFunction.apply(mq.detach(), sq.detach(), mp.detach(), sp.detach(), alpha, beta, (3, 4))
    assert_close(seg[0], beta * D.kl_divergence(D.Independent(D.Normal(mp_r[:3], sp_r[:3]), 1), D.Independent(D.Normal(mq_r[:3], sq_r[:3]), 1)).mean(), 1e-5, "seg0")
    # default draw: Box-Muller over Philox uniforms is standard normal
    z = noise.normal((512, 1024), DEV)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3 and bool(torch.isfinite(z).all())


@pytest.mark.parametrize("Dh", [8, 16, 24, 32])
def test_attention_head_dims(Dh):
    """head_dim = latent / 8 heads: 16 for the RGB cameras, 24 with depth_static (192/8), 32 for RGBD_both (256/8)."""
    B, S, H = 3, 32, 8
    E = H * Dh
    qkv = _rand(B * S, 3 * E, seed=Dh).requires_grad_()
    keep = (torch.rand(B, H, S, S, generator=torch.Generator().manual_seed(1)) > 0.1).to(torch.uint8).to(DEV)
    scale = 1.0 / 0.9
    out = ops.AttentionFunction.apply(qkv, B, S, H, keep, scale)
    g = _rand(B * S, E, seed=2)
    out.backward(g)
    ref_in = qkv.detach().double().requires_grad_()
    q, k, v = (t.reshape(B, S, H, Dh).transpose(1, 2) for t in ref_in.chunk(3, dim=-1))
    p = torch.softmax(q @ k.transpose(-1, -2) / Dh ** 0.5, -1) * keep.double() * scale
    ref = (p @ v).transpose(1, 2).reshape(B * S, E)
    ref.backward(g.double())
    assert_close(out, ref, 1e-5, "attention out")
    assert_close(qkv.grad, ref_in.grad, 2e-5, "attention dqkv")


def _grad_tol(name):
    if name == "logit_scale":
        return 2e-3
    # conv weight gradient NORMS vs the reference fixture: 1e4-1e5 summands per element at B=2 windows
    return 1e-3 if "conv_model" in name else 5e-4


def _supplied(kw, batch, draw):
    key = "categories" if kw["distribution"] == "discrete" else "normals"
    return noise.supplied(**{key: [draw[mod] for mod in batch]})


@pytest.mark.parametrize("tag", list(ALT_CASES))
def test_alt_training_step_vs_reference_fixture_and_oracle(tag):
    from oracle import hulc2_oracle as O

    kw, batch, draw = alt_case_inputs(tag)
    m = build_alt_model(tag)
    P = oracle_params(m)
    out = O.training_step(batch, {mod: {"plan_idx": draw[mod]} for mod in batch}, P, hulc2_config(pkg="x", **kw))
    out["loss"].backward()
    m = m.to(DEV).train()
    with _supplied(kw, batch, draw):
        loss = m.training_step(to_device(batch, DEV), 0)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(loss, alt_gt(f"{tag}/loss"), 1e-5, "loss vs reference fixture")
    for k in alt_keys(f"{tag}/log/"):
        assert_close(m.logged[k[len(tag) + 5:]], alt_gt(k), 1e-5, k)
    grads = dict(m.named_parameters())
    n = 0
    for k in alt_keys(f"{tag}/grad_norm/"):
        name = k[len(tag) + 11:]
        assert_close(grads[name].grad.double().norm(), alt_gt(k), _grad_tol(name), k)
        n += 1
    assert n >= 95
    for name, prm in m.named_parameters():
        ref = P[name].grad
        if ref is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, name
            continue
        # element-wise (max-norm relative): conv biases / LayerNorm-adjacent sums are ill-conditioned (see
        # test_gpu_step.test_training_step_golden); conv weights measured up to 2.8e-3 (gauss_gru, conv2) while their norms
        # agree with the reference fixture to 1e-3 above
        tol = 1e-2 if (name.endswith("bias") and "conv_model" in name) else (5e-3 if "conv_model" in name else _grad_tol(name))
        assert_close(prm.grad, ref, tol, name)


@pytest.mark.parametrize("tag", list(ALT_CASES))
def test_alt_training_step_bf16(tag):
    kw, batch, draw = alt_case_inputs(tag)
    m = build_alt_model(tag).to(DEV).train()
    ops.set_precision("bf16")
    with _supplied(kw, batch, draw):
        loss = m.training_step(to_device(batch, DEV), 0)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(loss, alt_gt(f"{tag}/loss"), 2e-2, "bf16 loss vs reference fixture")
    for k in alt_keys(f"{tag}/log/"):
        assert_close(m.logged[k[len(tag) + 5:]], alt_gt(k), 3e-2, k)
    # bf16 operands at B = 2 windows / hidden 256: individual gradient norms move by up to ~15 % (bf16 rounding flips ReLU
    # masks of whole feature-map positions in the trunk, and the GRU/LSTM backward-through-time rounds 32 steps of gate
    # gradients; with 2 language rows the InfoNCE term is a 2 x 2 softmax at logit scale 14), so the gradient check here is
    # directional: cosine >= 0.8 against the fp32 oracle for every parameter with a non-negligible gradient and ~0.98 on
    # average (asserted >= 0.95; gauss_gru measures 0.965) (measured worst single parameter 0.89-0.94 on gauss_gru, depending on the conv kernels' summation order), norm
    # within 25 %.  The B = 8 bf16 test (test_gpu_bf16) holds 3e-2 / cosine 0.99.
    from oracle import hulc2_oracle as O

    P = oracle_params(build_alt_model(tag))
    out = O.training_step(batch, {mod: {"plan_idx": draw[mod]} for mod in batch}, P, hulc2_config(pkg="x", **kw))
    out["loss"].backward()
    worst, coss = 1.0, []
    for name, prm in m.named_parameters():
        ref = P[name].grad
        if ref is None or name == "logit_scale" or float(ref.norm()) < 1e-3:
            continue
        g = prm.grad.detach().cpu().double().flatten()
        r = ref.double().flatten()
        cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
        worst = min(worst, cos)
        coss.append(cos)
        assert cos >= 0.8, f"{name}: cosine {cos:.4f}"
        assert abs(float(g.norm() / r.norm()) - 1.0) < 0.25, f"{name}: norm ratio {float(g.norm() / r.norm()):.3f}"
    assert sum(coss) / len(coss) >= 0.95, f"mean cosine {sum(coss) / len(coss):.4f}"
    print(f"[{tag}] bf16 gradient cosine vs fp32 oracle: worst {worst:.4f}, mean {sum(coss) / len(coss):.4f}")


@pytest.mark.parametrize("rnn_model", ["gru_decoder", "lstm_decoder"])
def test_gated_decoder_act_carries_hidden_state(rnn_model):
    """LogisticDecoderRNN.act (logistic_decoder_rnn.py:101-116) keeps nn.GRU's h_n / nn.LSTM's (h_n, c_n) across calls."""
    from oracle import hulc2_oracle as O

    tag = "gru" if rnn_model == "gru_decoder" else "lstm"
    m = build_alt_model(tag)
    P = oracle_params(m, requires_grad=False)
    dec = m.action_decoder
    N = 5
    g = torch.Generator().manual_seed(3)
    plan, goal = torch.randn(N, 1024, generator=g), torch.randn(N, 32, generator=g)
    robot = torch.rand(N, 1, 15, generator=g) - 0.5
    hid = None
    m = m.to(DEV).eval()
    dec.clear_hidden_state()
    for step in range(3):
        emb = torch.randn(N, 1, 128, generator=g)
        u1 = torch.rand(N, 1, 6, 10, generator=g)
        u2 = torch.rand(N, 1, 6, generator=g).clamp(1e-4, 1 - 1e-4)
        lp, ls, mu, grip, hid = O.decoder_forward(plan, emb, goal, P, (64, 128), hid, 10, -7.0, rnn_model)
        ref = O.tcp_to_world_frame(O.decoder_sample(lp, ls, mu, grip, u1, u2, P), robot)
        with noise.supplied(uniforms=[u1, u2]):
            act = dec.act(plan.to(DEV), emb.to(DEV), goal.to(DEV), robot.to(DEV))
        assert_close(act, ref, 1e-4, f"step {step}")
        hs = dec.hidden_state if isinstance(dec.hidden_state, tuple) else (dec.hidden_state,)
        hr = hid if isinstance(hid, tuple) else (hid,)
        for a, b in zip(hs, hr):
            assert_close(a, b, 1e-5, f"hidden state after step {step}")


@pytest.mark.parametrize("tag", ["lstm", "gauss_gru"])
def test_alt_validation_step_vs_reference_fixture(tag):
    """Hulc2.validation_step (hulc2.py:510-598) with an LSTM decoder / a continuous plan + GRU decoder, same supplied noise as
    the unmodified reference: sampled plans, losses, MAE and gripper success-rate metrics."""
    from helpers import alt_val_inputs

    kw, batch, nz = alt_val_inputs(tag)
    m = build_alt_model(tag).to(DEV).eval()
    draws, unis = [], []
    for mod in batch:
        draws += [nz[mod]["plan_idx_pp"], nz[mod]["plan_idx_pr"]]
        unis += [nz[mod][k] for k in ("u1_pp", "u2_pp", "u1_pr", "u2_pr")]
    key = "categories" if kw["distribution"] == "discrete" else "normals"
    with torch.no_grad(), noise.supplied(uniforms=unis, **{key: draws}):
        out = m.validation_step(to_device(batch, DEV), 0)
    for k in alt_keys(f"{tag}/val/out/"):
        ref, mine = alt_gt(k), out[k[len(tag) + 9:]].cpu()
        if ref.dtype.is_floating_point:
            assert_close(mine, ref, 1e-5, k)
        else:
            assert torch.equal(mine, ref), k
    for k in alt_keys(f"{tag}/val/log/"):
        name = k[len(tag) + 9:]
        tol = 2e-4 if "mae" in name else 1e-5            # MAE of x100-scaled orientation deltas (see test_gpu_step)
        assert_close(m.logged[name], alt_gt(k), tol, name)
