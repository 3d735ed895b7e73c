"""GPU parity tests for the bf16 tcgen05 path (precision=1).  Kernel-level checks compare against fp32 torch
on bf16-ROUNDED operands (the tensor core multiplies bf16 exactly and accumulates in fp32, so only summation
order differs: 1e-4); path-level checks use north_star's bf16 tolerance, 2e-2 relative."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close, build_model, golden, gt, oracle_params, rel_err, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def _bf(x):
    return x.bfloat16().float()


@pytest.fixture(autouse=True)
def _precision():
    from hulc2_b200 import ops

    ops.set_precision("bf16")
    yield
    ops.set_precision("fp32")


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (70, 45, 33), (257, 130, 100), (3, 2048, 160), (64, 64, 2048), (300, 32, 192), (500, 700, 260)])
def test_gemm_bf16_nt_epilogues(M, N, K):
    from hulc2_b200 import ops

    A, B, bias, add = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    maskt = _rand(M, N, seed=5)
    keep = (torch.rand(M, N, generator=torch.Generator().manual_seed(6)) > 0.3).to(torch.uint8)
    C0 = _rand(M, N, seed=7)
    ref = _bf(A) @ _bf(B).t() + bias + add + C0
    ref = torch.relu(ref) * (maskt > 0) * keep * 1.25
    Cd = C0.to(DEV)
    ops.gemm(M, N, K, A.to(DEV), K, 1, B.to(DEV), K, 1, Cd, N, bias=bias.to(DEV), add=add.to(DEV), ld_add=N,
             mask=maskt.to(DEV), ld_mask=N, keep=keep.to(DEV), ld_keep=N, keep_scale=1.25, relu=True, accumulate=True, precision=1)
    torch.cuda.synchronize()
    assert_close(Cd, ref, 1e-4)


def test_gemm_bf16_plain_identity():
    """A = I: the output must reproduce bf16(B)^T exactly -- isolates smem layout / descriptor errors."""
    from hulc2_b200 import ops

    K = 128
    A = torch.eye(K)
    B = _rand(96, K, seed=9)
    out = torch.empty(K, 96, device=DEV)
    ops.gemm(K, 96, K, A.to(DEV), K, 1, B.to(DEV), K, 1, out, 96, precision=1)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), _bf(B).t().contiguous())


def test_gemm_bf16_nn_tn_splitk():
    from hulc2_b200 import ops

    M, N, K = 190, 170, 150
    dY, W, X = _rand(M, N, seed=1), _rand(N, K, seed=2), _rand(M, K, seed=3)
    dX = torch.empty(M, K, device=DEV)
    ops.gemm(M, K, N, dY.to(DEV), N, 1, W.to(DEV), 1, K, dX, K, precision=1)
    assert_close(dX, _bf(dY) @ _bf(W), 1e-4, "NN")
    dW = torch.empty(N, K, device=DEV)
    ops.gemm(N, K, M, dY.to(DEV), 1, N, X.to(DEV), 1, K, dW, K, precision=1)
    assert_close(dW, _bf(dY).t() @ _bf(X), 1e-4, "TN")
    rows, N2, K2 = 9000, 60, 300
    dY2, X2 = _rand(rows, N2, seed=4), _rand(rows, K2, seed=5)
    dW2 = torch.empty(N2, K2, device=DEV)
    ops.gemm(N2, K2, rows, dY2.to(DEV), 1, N2, X2.to(DEV), 1, K2, dW2, K2, precision=1)
    assert_close(dW2, _bf(dY2).double().t() @ _bf(X2).double(), 1e-4, "split-K TN")


def test_conv_bf16_primitives():
    from hulc2_b200 import ops
    from hulc2_b200._lib import call

    for (F_, Cin, H, W, Cout, k, s, nhwc) in [(2, 8, 13, 11, 16, 4, 2, 1), (2, 3, 40, 44, 32, 8, 4, 0), (3, 64, 9, 9, 64, 3, 1, 1)]:
        x = _bf(_rand(F_, Cin, H, W, seed=1)).requires_grad_()
        w = _bf(_rand(Cout, Cin, k, k, seed=2, scale=0.3)).requires_grad_()
        b = _rand(Cout, seed=3).requires_grad_()
        y = F.conv2d(x, w, b, stride=s)
        gout = _bf(_rand(*y.shape, seed=4))
        y.backward(gout)
        OH, OW = y.shape[2], y.shape[3]
        xin = (x.detach().permute(0, 2, 3, 1).contiguous() if nhwc else x.detach()).to(DEV)
        wd = w.detach().to(DEV)
        if nhwc:
            wk = torch.empty(Cout, k, k, Cin, device=DEV)
            call("hulc2_permute_conv_weight", wd.data_ptr(), wk.data_ptr(), Cout, Cin, k, k, 0, 0)
        else:
            wk = wd
        ws = ops.workspace(torch.device(DEV))
        yd = torch.empty(F_, OH, OW, Cout, device=DEV)
        bd = b.detach().to(DEV)
        a = ops._conv_args(F_, Cin, H, W, Cout, k, s, nhwc)
        a.precision = 1
        a.x, a.w, a.bias, a.y, a.relu = xin.data_ptr(), wk.data_ptr(), bd.data_ptr(), yd.data_ptr(), 0
        call("hulc2_conv2d_fwd", C.byref(a))
        assert_close(yd.permute(0, 3, 1, 2), y, 1e-4, f"conv fwd {Cin}->{Cout} k{k}s{s}")
        dyn = gout.permute(0, 2, 3, 1).contiguous().to(DEV)
        dw = torch.empty(Cout, k * k * Cin, device=DEV)
        a.dy, a.dw, a.accumulate, a.workspace, a.workspace_bytes = dyn.data_ptr(), dw.data_ptr(), 0, ws.data_ptr(), ws.numel()
        call("hulc2_conv2d_wgrad", C.byref(a))
        dwr = dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2) if nhwc else dw.view(Cout, Cin, k, k)
        assert_close(dwr, w.grad, 1e-4, "conv wgrad")
        if nhwc:
            whwoi = torch.empty(k, k, Cout, Cin, device=DEV)
            call("hulc2_permute_conv_weight", wd.data_ptr(), whwoi.data_ptr(), Cout, Cin, k, k, 2, 0)
            dx = torch.empty(F_, H, W, Cin, device=DEV)
            ones = torch.ones(F_, H, W, Cin, device=DEV)
            a.w, a.dx, a.xmask = whwoi.data_ptr(), dx.data_ptr(), ones.data_ptr()
            call("hulc2_conv2d_dgrad", C.byref(a))
            assert_close(dx.permute(0, 3, 1, 2), x.grad, 1e-4, "conv dgrad")


@pytest.mark.parametrize("tag,variant,hw,aux", [("calvin_B2", "calvin", (200, 200), "half"), ("rw_B2", "real_world", (150, 200), "all")])
def test_training_step_bf16_vs_reference_fixture(tag, variant, hw, aux):
    """Reference fixture (B=2): every logged loss within north_star's bf16 tolerance (2e-2 relative) and the
    gradients of the decoder / plan-proposal / goal-encoder parameters within 2e-2 on their norms.  The
    fixture's two windows are near-duplicates (uniform-noise frames), so the InfoNCE path -- which only sees
    the DIFFERENCE between rows -- is ill-conditioned under any operand rounding; full-gradient agreement
    is checked on a well-conditioned batch in test_training_step_bf16_all_gradients_vs_oracle."""
    from hulc2_b200 import noise
    from hulc2_b200.synthetic import synthetic_batch

    m = build_model(variant, hw).to(DEV).train()
    batch = to_device(synthetic_batch(2, seed=1, static_hw=hw, aux=aux), DEV)
    with noise.supplied(categories=[gt(f"{tag}/plan_idx/{mod}") for mod in batch]):
        loss = m.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(loss, gt(f"{tag}/loss"), 2e-2, "loss")
    for k in golden().files:
        if k.startswith(f"{tag}/log/"):
            assert_close(m.logged[k[len(tag) + 5 :]], gt(k), 2e-2, k)
    grads = dict(m.named_parameters())
    n = 0
    for k in golden().files:
        if k.startswith(f"{tag}/gnorm/"):
            name = k[len(tag) + 7 :]
            if name.startswith(("action_decoder.", "plan_proposal.", "visual_goal.", "plan_recognition.fc_state")):
                assert_close(grads[name].grad.norm(), gt(k), 2e-2, name)
                n += 1
    assert n >= 30
    with torch.no_grad():
        emb = m.perceptual_encoder(batch["vis"]["rgb_obs"], batch["vis"]["depth_obs"], batch["vis"]["robot_obs"])
    assert_close(emb, gt(f"{tag}/perceptual_emb_vis"), 2e-2, "perceptual_emb")


def test_training_step_bf16_all_gradients_vs_oracle():
    """torch default init (seed 0), B=8 per modality: loss within 2e-2 and EVERY parameter gradient within
    3e-2 on its norm with cosine >= 0.99 against the fp32 CPU oracle."""
    import json, os

    from hulc2_b200 import noise
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch
    from oracle import hulc2_oracle as O

    B = 8
    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.0))
    P = oracle_params(m)
    batch = synthetic_batch(B, seed=1, aux="all")
    idx = {mod: torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)) for mod in batch}
    out = O.training_step(batch, {mod: {"plan_idx": idx[mod]} for mod in batch}, P, hulc2_config(pkg="x", dropout_p=0.0))
    out["loss"].backward()
    m = m.to(DEV).train()
    with noise.supplied(categories=[idx[mod] for mod in batch]):
        loss = m.training_step(to_device(batch, DEV), 0)
    loss.backward()
    assert_close(loss, out["loss"], 2e-2, "loss")
    for k, v in m.logged.items():
        assert_close(v, out[k], 2e-2, k)
    report, bad = {}, {}
    for n, p in m.named_parameters():
        r = P[n].grad
        if r is None:
            continue
        g = p.grad.cpu()
        ratio = float(g.norm() / (r.norm() + 1e-30))
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        report[n] = (ratio, cos)
        if n != "logit_scale" and (abs(ratio - 1) > 3e-2 or cos < 0.99):
            bad[n] = (ratio, cos)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_bf16_oracle_B8.json", "w") as f:
        json.dump({"loss_rel_err": rel_err(loss, out["loss"]), "worst_ratio": max(abs(a - 1) for a, _ in report.values()),
                   "worst_cos": min(c for _, c in report.values()), "n_params": len(report)}, f, indent=1)
    assert not bad, bad


def test_training_step_bf16_bench_config_B64_vs_oracle():
    """The BASELINE configs[1] shape bench.py times -- B=64 windows per modality (M=128 decoder rows, F=4096 frames per
    encoder call: the halo conv kernels, split-K choices and the 4-CTA-cluster recurrence all depend on it), dropout off,
    supplied plan draw -- against the fp32 CPU oracle: loss and every logged scalar 2e-2, every parameter gradient 3e-2 on
    its norm with cosine >= 0.99 (north_star's bf16 tolerance)."""
    import json, os

    from hulc2_b200 import noise
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch
    from oracle import hulc2_oracle as O

    B = 64
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    m = instantiate(hulc2_config(dropout_p=0.0))
    P = oracle_params(m)
    batch = synthetic_batch(B, seed=1, aux="half")
    idx = {mod: torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)) for mod in batch}
    out = O.training_step(batch, {mod: {"plan_idx": idx[mod]} for mod in batch}, P, hulc2_config(pkg="x", dropout_p=0.0))
    out["loss"].backward()
    m = m.to(DEV).train()
    dev_batch = to_device(batch, DEV)
    with noise.supplied(categories=[idx[mod] for mod in batch]):
        loss = m.training_step(dev_batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(loss, out["loss"], 2e-2, "loss")
    for k, v in m.logged.items():
        assert_close(v, out[k], 2e-2, k)
    report, bad = {}, {}
    for n, p in m.named_parameters():
        r = P[n].grad
        if r is None:
            continue
        g = p.grad.cpu()
        ratio = float(g.norm() / (r.norm() + 1e-30))
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        report[n] = (ratio, cos)
        if n != "logit_scale" and (abs(ratio - 1) > 3e-2 or cos < 0.99):   # (ill-conditioned cancelling sum, see test_oracle_golden)
            bad[n] = (ratio, cos)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_bf16_oracle_B64.json", "w") as f:
        json.dump({"config": "configs[1]: B=64/modality, bf16, dropout 0, aux mask 50%", "loss": float(loss), "oracle_loss": float(out["loss"]),
                   "loss_rel_err": rel_err(loss, out["loss"]), "worst_ratio": max(abs(a - 1) for a, _ in report.values()),
                   "worst_cos": min(c for _, c in report.values()), "n_params": len(report),
                   "worst5_ratio": sorted(((n, a, c) for n, (a, c) in report.items()), key=lambda t: -abs(t[1] - 1))[:5],
                   "worst5_cos": sorted(((n, a, c) for n, (a, c) in report.items()), key=lambda t: t[2])[:5]}, f, indent=1)
    assert not bad, bad


def test_rollout_bf16_argmax_and_actions():
    from hulc2_b200 import noise
    from hulc2_b200.synthetic import synthetic_obs

    m = build_model("calvin").to(DEV).eval()
    obs, goal = synthetic_obs(4, seed=2)
    obs, goal = to_device(obs, DEV), to_device(goal, DEV)
    m.reset()
    m.replan_freq = 2
    for s in range(4):
        cats = [gt(f"rollout_N4/step{s}/plan_idx")] if s % 2 == 0 else []
        with noise.supplied(categories=cats, uniforms=[gt(f"rollout_N4/step{s}/u1"), gt(f"rollout_N4/step{s}/u2")]):
            a = m.step(obs, goal)
        assert a.shape == (4, 1, 7) and bool(torch.isfinite(a).all())
        assert set(a[..., -1].unique().tolist()) <= {-1.0, 1.0}


@pytest.fixture
def rnn_kernel_reset():
    yield
    from hulc2_b200 import _lib

    _lib.load_library().hulc2_rnn_select_kernel(0)


@pytest.mark.parametrize("kernel", [0, -1, 1], ids=["cluster_tma", "cluster", "persistent1d"])
@pytest.mark.parametrize("B,S,H,with_h0", [(64, 32, 2048, False), (128, 32, 2048, False), (5, 3, 2048, True), (128, 4, 1024, True),
                                           (33, 5, 512, True), (16, 5, 256, True)])
def test_persistent_rnn_matches_step_loop(B, S, H, with_h0, kernel, rnn_kernel_reset):
    """Persistent tcgen05 recurrences (one launch for all steps: cluster split-K kernel, or the 1-D kernel it falls
    back to) vs the fp32 per-step kernels."""
    from hulc2_b200 import ops, _lib
    from hulc2_b200._lib import call

    _lib.load_library().hulc2_rnn_select_kernel(kernel)

    dev = torch.device(DEV)
    pre = _rand(S, B, H, seed=1).to(dev)
    w = _rand(H, H, seed=2, scale=0.02).to(dev)
    h0 = _rand(B, H, seed=3).abs().to(dev) if with_h0 else None
    ws = ops.workspace(dev)
    outs = []
    for prec in (0, 1):
        h = torch.empty(S, B, H, device=dev)
        call("hulc2_rnn_relu_fwd", pre.data_ptr(), w.data_ptr(), None if h0 is None else h0.data_ptr(), h.data_ptr(), S, B, H, prec, ws.data_ptr(), ws.numel())
        outs.append(h)
    torch.cuda.synchronize()
    # no silent fallback: the kernel generation under test really served the call when the shape fits it
    path = _lib.load_library().hulc2_rnn_last_path()
    if H % 512 == 0 and kernel in (0, -1):
        assert path & 255 == (1 if kernel == 0 else 2), f"expected the cluster kernel, got path {path & 255} (reject {path >> 8})"
    elif kernel == 1 or H % 512:
        assert path & 255 == 3, path
    assert_close(outs[1], outs[0], 2e-2, "forward states")
    # ... and against plain PyTorch fp32 (nn.RNN relu recurrence, decoders/utils/rnn.py:5-14), not only this library's own kernel
    hp = h0 if h0 is not None else torch.zeros(B, H, device=dev)
    ref_h = []
    for t in range(S):
        hp = torch.relu(pre[t] + hp @ w.t())
        ref_h.append(hp)
    ref_h = torch.stack(ref_h)
    assert_close(outs[0], ref_h, 1e-5, "fp32 kernel vs torch")
    assert_close(outs[1], ref_h, 2e-2, "tcgen05 kernel vs torch")
    dh = _rand(S, B, H, seed=4).to(dev)
    res = []
    for prec in (0, 1):
        d = dh.clone()
        dh0 = torch.empty(B, H, device=dev)
        call("hulc2_rnn_relu_bwd", d.data_ptr(), w.data_ptr(), outs[0].data_ptr(), dh0.data_ptr(), S, B, H, prec, ws.data_ptr(), ws.numel())
        res.append((d, dh0))
    torch.cuda.synchronize()
    assert_close(res[1][0], res[0][0], 2e-2, "backward dz")
    assert_close(res[1][1], res[0][1], 2e-2, "dh0")
    # torch reference of the reverse recurrence: dz[t] = (dh[t] + dz[t+1] W_hh) * (h[t] > 0), dh0 = dz[0] W_hh
    carry = torch.zeros(B, H, device=dev)
    ref_dz = [None] * S
    for t in range(S - 1, -1, -1):
        ref_dz[t] = (dh[t] + carry) * (outs[0][t] > 0)
        carry = ref_dz[t] @ w
    assert_close(res[1][0], torch.stack(ref_dz), 2e-2, "tcgen05 backward dz vs torch")
    assert_close(res[1][1], carry, 2e-2, "tcgen05 dh0 vs torch")
    assert int(_lib.load_library().hulc2_rnn_device_error(1)) == 0
    # ReLU mask is exact: zeros where h == 0
    assert bool(((outs[0] <= 0) <= (res[1][0] == 0)).all())
