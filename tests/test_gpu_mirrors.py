"""Producer-written bf16 operand mirrors (the "_m" C-ABI entry points + ops' activation-mirror registry).

A kernel that produces the fp32 input of a tcgen05 contraction (LayerNorm, attention, positional add, the decoder
recurrence) also writes the bf16 copy the contraction multiplies, so no separate fp32 -> bf16 pass runs.  Bar: every
mirror is BIT-IDENTICAL to ``hulc2_f32_to_bf16`` of the fp32 result, the consumers really pick the mirrors up (launch
counts drop), and a training step gives bit-identical loss and gradients with the mirrors on and off."""
import pytest
import torch

from helpers import to_device

DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * scale).to(DEV)


@pytest.fixture(autouse=True)
def _bf16_precision():
    from hulc2_b200 import ops

    ops.set_precision("bf16")
    ops.clear_mirrors()
    yield
    ops.set_precision("fp32")
    ops.producer_mirrors = True
    ops.clear_mirrors()


def _mirror_of(t):
    from hulc2_b200 import ops

    hit = ops._find_mirror(t.view(-1, t.shape[-1]))
    assert hit is not None, "producer did not register a mirror"
    m16, ld = hit
    return m16.reshape(-1, ld)[:, : t.shape[-1]]


def _same_bits(m16, t):
    ref = t.reshape(-1, t.shape[-1]).bfloat16()
    return torch.equal(m16.view(torch.int16), ref.view(torch.int16))


@pytest.mark.gpu
def test_layernorm_mirrors_bit_exact():
    from hulc2_b200 import ops

    rows, D = 4096, 128
    x, res = _rand(rows, D, seed=1).requires_grad_(), _rand(rows, D, seed=2).requires_grad_()
    gamma, beta = _rand(D, seed=3).requires_grad_(), _rand(D, seed=4).requires_grad_()
    keep = (torch.rand(rows, D, generator=torch.Generator().manual_seed(5)) > 0.1).to(torch.uint8).to(DEV)
    y = ops.layer_norm(x, gamma, beta, res=res, keep=keep, keep_scale=1.0 / 0.9)
    assert _same_bits(_mirror_of(y), y)
    (dx, dres) = torch.autograd.grad(y, (x, res), _rand(rows, D, seed=6))
    assert _same_bits(_mirror_of(dres), dres)          # the residual branch's gradient feeds the linear's backward
    # without a residual the mirrored gradient is dx
    y2 = ops.layer_norm(x, gamma, beta)
    (dx2,) = torch.autograd.grad(y2, (x,), _rand(rows, D, seed=7))
    assert _same_bits(_mirror_of(dx2), dx2)
    # D that is not a multiple of 8: mirror rows padded to 8 columns
    x3, g3, b3 = _rand(64, 36, seed=8), _rand(36, seed=9), _rand(36, seed=10)
    y3 = ops.layer_norm(x3, g3, b3)
    assert _same_bits(_mirror_of(y3), y3)


@pytest.mark.gpu
@pytest.mark.parametrize("Dh", [16, 24, 32])
def test_attention_and_add_pos_mirrors_bit_exact(Dh):
    from hulc2_b200 import ops

    B, S, H = 16, 32, 8
    E = H * Dh
    emb, pos = _rand(B, S, E, seed=1).requires_grad_(), _rand(S, E, seed=2)
    x = ops.AddPosFunction.apply(emb, pos, None, 1.0)
    assert _same_bits(_mirror_of(x), x)
    qkv = _rand(B * S, 3 * E, seed=3).requires_grad_()
    keep = (torch.rand(B, H, S, S, generator=torch.Generator().manual_seed(4)) > 0.1).to(torch.uint8).to(DEV)
    out = ops.AttentionFunction.apply(qkv, B, S, H, keep, 1.0 / 0.9)
    assert _same_bits(_mirror_of(out), out)
    (dqkv,) = torch.autograd.grad(out, (qkv,), _rand(B * S, E, seed=5))
    assert _same_bits(_mirror_of(dqkv), dqkv)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [0, -1, 1, 2], ids=["cluster_tma", "cluster", "persistent1d", "per_step_gemm"])
@pytest.mark.parametrize("B,S,H,with_h0", [(128, 32, 2048, False), (5, 3, 2048, True), (64, 4, 1024, True)])
def test_recurrence_state_mirrors_bit_exact(B, S, H, with_h0, kernel):
    """h16 / dz16 of hulc2_rnn_relu_{fwd,bwd}_m: slot t + 1 == bf16(state t) for every kernel of the fallback chain."""
    from hulc2_b200 import _lib, ops

    lib = _lib.load_library()
    prev = lib.hulc2_rnn_select_kernel(kernel)
    try:
        ws = ops.workspace(torch.device(DEV))
        pre = _rand(S, B, H, seed=1, scale=0.5)
        w = _rand(H, H, seed=2, scale=1.0 / H ** 0.5)
        h0 = _rand(B, H, seed=3).abs() if with_h0 else None
        h = torch.empty(S, B, H, device=DEV)
        h16 = torch.full((S + 1, B, H), float("nan"), device=DEV, dtype=torch.bfloat16)
        _lib.call("hulc2_rnn_relu_fwd_m", pre.data_ptr(), w.data_ptr(), ops._p(h0), h.data_ptr(), h16.data_ptr(), S, B, H, 1,
                  ws.data_ptr(), ws.numel())
        # same call without the mirror: identical fp32 states
        h_ref = torch.empty(S, B, H, device=DEV)
        _lib.call("hulc2_rnn_relu_fwd", pre.data_ptr(), w.data_ptr(), ops._p(h0), h_ref.data_ptr(), S, B, H, 1, ws.data_ptr(), ws.numel())
        torch.cuda.synchronize()
        assert torch.equal(h, h_ref)
        assert torch.equal(h16[1:].view(torch.int16), h.bfloat16().view(torch.int16))
        dh = _rand(S, B, H, seed=4)
        dz, dz_ref = dh.clone(), dh.clone()
        dz16 = torch.full((S + 1, B, H), float("nan"), device=DEV, dtype=torch.bfloat16)
        _lib.call("hulc2_rnn_relu_bwd_m", dz.data_ptr(), w.data_ptr(), h.data_ptr(), None, dz16.data_ptr(), S, B, H, 1, ws.data_ptr(), ws.numel())
        _lib.call("hulc2_rnn_relu_bwd", dz_ref.data_ptr(), w.data_ptr(), h.data_ptr(), None, S, B, H, 1, ws.data_ptr(), ws.numel())
        torch.cuda.synchronize()
        assert torch.equal(dz, dz_ref)
        assert torch.equal(dz16[1:].view(torch.int16), dz.bfloat16().view(torch.int16))
        assert lib.hulc2_rnn_device_error(1) == 0
    finally:
        lib.hulc2_rnn_select_kernel(prev)


def _one_step(mirrors: bool, B=4):
    from hulc2_b200 import _lib, noise, ops
    from hulc2_b200._compat import instantiate
    from hulc2_b200.config import hulc2_config
    from hulc2_b200.synthetic import synthetic_batch

    ops.producer_mirrors = mirrors
    ops.clear_mirrors()
    torch.manual_seed(0)
    m = instantiate(hulc2_config(dropout_p=0.1)).to(DEV).train()
    batch = to_device(synthetic_batch(B, seed=1, aux="all"), DEV)
    idx = [torch.randint(0, 32, (B, 32), generator=torch.Generator().manual_seed(5)).to(DEV) for _ in batch]
    noise.manual_seed(123)
    n0 = _lib.launch_count
    with noise.supplied(categories=idx):
        loss = m.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    calls = _lib.launch_count - n0
    return float(loss), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, calls


@pytest.mark.gpu
def test_training_step_identical_with_and_without_producer_mirrors():
    """Same step, mirrors written by the producers vs converted by the consumers: loss and EVERY gradient bit-identical, and
    the step with producer mirrors makes >= 15 fewer C-ABI calls (the fp32 -> bf16 passes that are gone)."""
    loss_on, g_on, calls_on = _one_step(True)
    loss_off, g_off, calls_off = _one_step(False)
    assert loss_on == loss_off
    assert g_on.keys() == g_off.keys()
    diff = [n for n in g_on if not torch.equal(g_on[n], g_off[n])]
    assert not diff, diff
    assert calls_off - calls_on >= 15, (calls_on, calls_off)


def test_mirror_registry_guards_cpu():
    """Host logic of the registry (no GPU): shape / contiguity / version guards, clearing at step boundaries."""
    from hulc2_b200 import ops

    ops.clear_mirrors()
    src = torch.zeros(6, 8)
    m16 = torch.zeros(6, 8, dtype=torch.bfloat16)
    ops.register_mirror(src, m16, 8, 6, 8)
    assert ops._find_mirror(src) is not None
    assert ops._find_mirror(src.view(6, 8)) is not None            # an alias of the same memory
    assert ops._find_mirror(src.view(12, 4)) is None               # other shape
    assert ops._find_mirror(src[:, :4]) is None                    # not dense
    assert ops._find_mirror(torch.zeros(6, 8)) is None             # other memory
    src.add_(1.0)                                                  # in-place update (autograd accumulation): mirror is stale
    assert ops._find_mirror(src) is None
    ops.register_mirror(src, m16, 8, 6, 8)
    assert ops._find_mirror(src) is not None
    ops.begin_grad_step()
    assert ops._find_mirror(src) is None
    ops.register_mirror(src, m16, 8, 6, 8)
    ops.invalidate_weight_mirrors()
    assert ops._find_mirror(src) is None
    # a loop that never reaches a step boundary cannot pin more than 512 tensors
    keep = [torch.zeros(2, 8) for _ in range(600)]
    for t in keep:
        ops.register_mirror(t, m16, 8, 2, 8)
    assert len(ops._act16) <= 512
    ops.clear_mirrors()
    # the consumer-facing lookups refuse CPU tensors like every other op (no CPU fallback), registry hit or not
    h = torch.zeros(3, 2, 8)
    ops.register_mirror(h, torch.ones(3, 2, 8, dtype=torch.bfloat16), 8, 6, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mirror_like(h)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mirror2d(h.view(6, 8))
    ops.clear_mirrors()
    assert ops._bits_saved((None, h))[0].numel() == 0 and ops._bits_saved((None, h))[1] is h
