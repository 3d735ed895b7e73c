"""GPU parity tests of the persistent tcgen05 conv trunk (csrc/conv_sm100.cu, bf16 NHWC activations).

Every primitive is checked against torch's fp32 convolution (and its autograd) evaluated on the SAME bf16-rounded
operands, so only the fp32 summation order and the final bf16 rounding of the stored activation differ:
tolerance 1e-2 of the tensor's max-norm for bf16 outputs (bf16 has 8 mantissa bits: 2^-9 = 2e-3 per element),
2e-3 for fp32 weight/bias gradients.  Reference semantics: vision_network.py:38-48, vision_network_gripper.py:11-26.
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close

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


def _ste_bf16(t):
    """Round to bf16 in the forward pass, identity in the backward pass (the trunk stores activations as bf16)."""
    return t + (_bf(t) - t).detach()


def _ref_trunk(x, params):
    """fp32 torch model of the bf16 trunk with the SAME rounding points: bf16 weights, bf16 activations after every
    ReLU, bf16 gradients w.r.t. every pre-activation.  Only summation order differs from the kernels."""
    w1, b1, w2, b2, w3, b3 = params
    a = _bf(x)
    for w, b, s in ((w1, b1, 4), (w2, b2, 2), (w3, b3, 1)):
        z = F.conv2d(a, _ste_bf16(w), b, stride=s)
        z.register_hook(lambda g: _bf(g))
        a = _ste_bf16(F.relu(z))
    return a


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def _nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("Fr,C,H,W", [(3, 3, 200, 200), (2, 3, 84, 84), (2, 3, 150, 200), (2, 1, 150, 200)])
def test_pack_frames_exact(Fr, C, H, W):
    from hulc2_b200 import ops

    x = _rand(Fr, C, H, W, seed=1)
    xs = ops.pack_frames(x.to(DEV))
    torch.cuda.synchronize()
    H4, W4 = H // 4, W // 4
    ref = x[:, :, : 4 * H4, : 4 * W4].reshape(Fr, C, H4, 4, W4, 4).permute(0, 2, 4, 1, 3, 5).reshape(Fr, H4, W4, 16 * C).bfloat16()
    assert torch.equal(xs.cpu(), ref)


@pytest.mark.parametrize("Fr,H,W", [(5, 200, 200), (3, 84, 84), (2, 150, 200)])
def test_conv1_over_packed_frames(Fr, H, W):
    from hulc2_b200 import ops

    x, w, b = _rand(Fr, 3, H, W, seed=1), _rand(32, 3, 8, 8, seed=2, scale=0.1), _rand(32, seed=3, scale=0.1)
    xs = ops.pack_frames(x.to(DEV))
    wp = ops.pack_conv_weight(w.to(DEV), 1)
    y = ops.convb_fwd(xs, wp, b.to(DEV), Fr, 48, H // 4, W // 4, 32, 2, 1, name="t")
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_bf(x), _bf(w), b, stride=4))
    assert y.shape == (Fr, ref.shape[2], ref.shape[3], 32)
    got = _nchw(y.float().cpu())
    err = (got - ref).abs()
    bad = (err > 1e-2 * ref.abs().max()).nonzero()
    where = f"{len(bad)} elements off, first at (f, c, y, x) = {bad[:4].tolist()}, got {[float(got[tuple(i)]) for i in bad[:4]]} want {[float(ref[tuple(i)]) for i in bad[:4]]}"
    assert_close(got, ref, 1e-2, "conv1 fwd: " + where)


@pytest.mark.parametrize("Fr,C,H,W,Cout,k,s", [(7, 32, 49, 49, 64, 4, 2), (7, 64, 23, 23, 64, 3, 1), (3, 32, 20, 20, 64, 4, 2),
                                               (3, 64, 9, 9, 64, 3, 1), (2, 32, 36, 49, 64, 4, 2), (1, 64, 17, 23, 64, 3, 1)])
def test_conv_fwd_dgrad_wgrad_nhwc(Fr, C, H, W, Cout, k, s):
    from hulc2_b200 import ops

    x = F.relu(_rand(Fr, C, H, W, seed=1))                       # a post-ReLU activation (has exact zeros for the mask)
    w, b = _rand(Cout, C, k, k, seed=2, scale=0.05), _rand(Cout, seed=3, scale=0.1)
    xb = _nhwc(x).bfloat16().to(DEV)
    y = ops.convb_fwd(xb, ops.pack_conv_weight(w.to(DEV), 0), b.to(DEV), Fr, C, H, W, Cout, k, s, name="t")
    xr = _bf(x).requires_grad_(True)
    wr = _bf(w).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    pre = F.conv2d(xr, wr, br, stride=s)
    ref = F.relu(pre)
    torch.cuda.synchronize()
    assert_close(_nchw(y.float().cpu()), ref, 1e-2, "fwd")

    dy = _rand(*ref.shape, seed=4) * (ref > 0)                   # gradient wrt the pre-activation
    dyb = _nhwc(dy).bfloat16()
    pre.backward(dyb.float().permute(0, 3, 1, 2))
    dx = ops.convb_dgrad(dyb.to(DEV), w.to(DEV), xb, Fr, C, H, W, Cout, k, s, name="t")
    dw, db = ops.convb_wgrad(xb, dyb.to(DEV), Fr, C, H, W, Cout, k, s, (Cout, C, k, k), name="t")
    torch.cuda.synchronize()
    assert_close(_nchw(dx.float().cpu()), xr.grad * (x > 0), 1e-2, "dgrad")
    assert_close(dw.cpu(), wr.grad, 2e-3, "wgrad")
    assert_close(db.cpu(), br.grad, 2e-3, "bgrad")


def test_conv1_wgrad_packed_frames():
    from hulc2_b200 import ops

    Fr, H, W = 3, 84, 84
    x, w = _rand(Fr, 3, H, W, seed=1), _rand(32, 3, 8, 8, seed=2, scale=0.1)
    xr, wr = _bf(x), _bf(w).requires_grad_(True)
    br = torch.zeros(32, requires_grad=True)
    pre = F.conv2d(xr, wr, br, stride=4)
    dy = _rand(*pre.shape, seed=4)
    dyb = _nhwc(dy).bfloat16()
    pre.backward(dyb.float().permute(0, 3, 1, 2))
    xs = ops.pack_frames(x.to(DEV))
    dw, db = ops.convb_wgrad(xs, dyb.to(DEV), Fr, 48, H // 4, W // 4, 32, 2, 1, (32, 3, 8, 8), dw_layout=1, name="t")
    torch.cuda.synchronize()
    assert_close(dw.cpu(), wr.grad, 2e-3, "wgrad1")
    assert_close(db.cpu(), br.grad, 2e-3, "bgrad1")


@pytest.mark.parametrize("Fr,H,W", [(150, 200, 200), (40, 150, 200)])
def test_static_trunk_many_tiles(Fr, H, W):
    """More tiles than SMs x stages: exercises the persistent tile loop, ring wrap-around and TMEM double buffering."""
    from hulc2_b200 import ops

    x = _rand(Fr, 3, H, W, seed=1)
    ws = [_rand(32, 3, 8, 8, seed=2, scale=0.08), _rand(64, 32, 4, 4, seed=3, scale=0.05), _rand(64, 64, 3, 3, seed=4, scale=0.05)]
    bs = [_rand(32, seed=5, scale=0.1), _rand(64, seed=6, scale=0.1), _rand(64, seed=7, scale=0.1)]
    d = [t.to(DEV) for t in (x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    xs, y1, y2, y3, bits = ops._convb_trunk_fwd(*d)
    torch.cuda.synchronize()
    # ReLU sign bits written by the forward kernels: bit i of the packed tensor == (element i of the activation > 0)
    assert bits[0] is None                      # (requested for conv2's output only, see ops._convb_trunk_fwd)
    for y, m in ((y2, bits[1]),):
        assert m is not None, "the halo forward kernels write the sign bits"
        ref_bits = (y.reshape(-1) > 0).cpu()
        got = ((m.cpu().view(-1, 1) >> torch.arange(8, dtype=torch.uint8)) & 1).bool().reshape(-1)[: ref_bits.numel()]
        assert torch.equal(got, ref_bits)
    r1 = F.relu(F.conv2d(_bf(x), _bf(ws[0]), bs[0], stride=4))
    assert_close(_nchw(y1.float().cpu()), r1, 1e-2, "y1")
    r2 = F.relu(F.conv2d(_nchw(y1.float().cpu()), _bf(ws[1]), bs[1], stride=2))
    assert_close(_nchw(y2.float().cpu()), r2, 1e-2, "y2")
    r3 = F.relu(F.conv2d(_nchw(y2.float().cpu()), _bf(ws[2]), bs[2], stride=1))
    assert_close(_nchw(y3.float().cpu()), r3, 1e-2, "y3")


@pytest.mark.parametrize("hw", [(200, 200), (150, 200)])
def test_static_encoder_bf16_trunk_vs_torch(hw):
    """StaticConvSSM (conv trunk + SpatialSoftmax) forward and all six parameter gradients vs the fp32 torch model of
    the trunk with identical bf16 rounding points (step-level parity vs the un-rounded reference is in test_gpu_bf16.py)."""
    from hulc2_b200 import ops
    from hulc2_b200.models.perceptual_encoders.vision_network import SpatialSoftmax

    Fr = 16
    H, W = hw
    x = _rand(Fr, 3, H, W, seed=1)
    ws = [_rand(32, 3, 8, 8, seed=2, scale=0.08), _rand(64, 32, 4, 4, seed=3, scale=0.05), _rand(64, 64, 3, 3, seed=4, scale=0.05)]
    bs = [_rand(32, seed=5, scale=0.1), _rand(64, seed=6, scale=0.1), _rand(64, seed=7, scale=0.1)]
    oh = lambda h, k, s: (h - k) // s + 1
    h3, w3 = oh(oh(oh(H, 8, 4), 4, 2), 3, 1), oh(oh(oh(W, 8, 4), 4, 2), 3, 1)
    ssm = SpatialSoftmax(num_rows=w3, num_cols=h3, temperature=1.0)
    dev = [t.to(DEV).requires_grad_(True) for pair in zip(ws, bs) for t in pair]
    out = ops.StaticConvSSM.apply(x.to(DEV), *dev, ssm.x_map.to(DEV), ssm.y_map.to(DEV), ssm.temperature.to(DEV))
    g = _rand(Fr, 128, seed=9)
    out.backward(g.to(DEV))
    torch.cuda.synchronize()

    ref_p = [t.clone().requires_grad_(True) for pair in zip(ws, bs) for t in pair]
    a = _ref_trunk(x, ref_p)
    n, c, h, w = a.shape
    sm = torch.softmax(a.reshape(-1, h * w), dim=1)
    ex = (sm * ssm.x_map).sum(1, keepdim=True)
    ey = (sm * ssm.y_map).sum(1, keepdim=True)
    ref = torch.cat((ex, ey), 1).view(n, 2 * c)
    ref.backward(g)
    assert_close(out.cpu(), ref, 5e-3, "keypoints")
    for d, r, nm in zip(dev, ref_p, ("w1", "b1", "w2", "b2", "w3", "b3")):
        assert_close(d.grad.cpu(), r.grad, 2e-2, nm)


def test_gripper_trunk_bf16_vs_torch():
    from hulc2_b200 import ops

    Fr = 5
    x = _rand(Fr, 3, 84, 84, seed=1)
    ws = [_rand(32, 3, 8, 8, seed=2, scale=0.08), _rand(64, 32, 4, 4, seed=3, scale=0.05), _rand(64, 64, 3, 3, seed=4, scale=0.05)]
    bs = [_rand(32, seed=5, scale=0.1), _rand(64, seed=6, scale=0.1), _rand(64, seed=7, scale=0.1)]
    dev = [t.to(DEV).requires_grad_(True) for pair in zip(ws, bs) for t in pair]
    flat = ops.GripperConvFlatten.apply(x.to(DEV), *dev)
    g = _rand(Fr, 3136, seed=9)
    flat.backward(g.to(DEV))
    torch.cuda.synchronize()
    ref_p = [t.clone().requires_grad_(True) for pair in zip(ws, bs) for t in pair]
    a = _ref_trunk(x, ref_p).flatten(1)
    a.backward(g)
    assert_close(flat.cpu(), a, 1e-2, "flatten")
    for d, r, nm in zip(dev, ref_p, ("w1", "b1", "w2", "b2", "w3", "b3")):
        assert_close(d.grad.cpu(), r.grad, 2e-2, nm)


@pytest.mark.parametrize("Fr,H,W", [(40, 200, 200), (24, 150, 200), (64, 84, 84)])
def test_trunk_backward_sign_bits_equal_activation_mask(Fr, H, W):
    """Input gradients masked with the forward's ReLU sign bits (2 bytes per 16 channels) are BIT-IDENTICAL to those masked
    with the bf16 activations themselves (32 bytes per 16 channels), for both stride classes of the trunk."""
    from hulc2_b200 import ops

    x = _rand(Fr, 3, H, W, seed=1)
    ws = [_rand(32, 3, 8, 8, seed=2, scale=0.08), _rand(64, 32, 4, 4, seed=3, scale=0.05), _rand(64, 64, 3, 3, seed=4, scale=0.05)]
    bs = [_rand(32, seed=5, scale=0.1), _rand(64, seed=6, scale=0.1), _rand(64, seed=7, scale=0.1)]
    d = [t.to(DEV) for t in (x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])]
    xs, y1, y2, y3, bits = ops._convb_trunk_fwd(*d)
    assert bits[1] is not None
    # conv1's bits are not produced by the trunk (not worth it there): ask the kernel for them directly
    H4, W4 = H // 4, W // 4
    y1b, m1 = ops.convb_fwd(xs, ops.pack_conv_weight(d[1], 1), d[2], Fr, 48, H4, W4, 32, 2, 1, name="c1", sign_bits=True)
    assert m1 is not None and torch.equal(y1b.view(torch.int16), y1.view(torch.int16))
    bits = (m1, bits[1])
    dz3 = (_rand(*y3.shape, seed=8).to(DEV) * (y3 > 0)).bfloat16()
    H1, W1, H2, W2 = y1.shape[1], y1.shape[2], y2.shape[1], y2.shape[2]
    dz2_a = ops.convb_dgrad(dz3, d[5], y2, Fr, 64, H2, W2, 64, 3, 1, name="c3")
    dz2_b = ops.convb_dgrad(dz3, d[5], y2, Fr, 64, H2, W2, 64, 3, 1, name="c3", sign_bits=bits[1])
    dz1_a = ops.convb_dgrad(dz2_a, d[3], y1, Fr, 32, H1, W1, 64, 4, 2, name="c2")
    dz1_b = ops.convb_dgrad(dz2_a, d[3], y1, Fr, 32, H1, W1, 64, 4, 2, name="c2", sign_bits=bits[0])
    torch.cuda.synchronize()
    assert torch.equal(dz2_a.view(torch.int16), dz2_b.view(torch.int16))
    assert torch.equal(dz1_a.view(torch.int16), dz1_b.view(torch.int16))
    assert float(dz1_a.float().abs().sum()) > 0
