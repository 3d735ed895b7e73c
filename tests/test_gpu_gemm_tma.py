"""GPU parity tests of the TMA-fed tcgen05 GEMM (csrc/gemm_tma_sm100.cu): bf16 operand mirrors, all four operand
orientations, the fused epilogue, the bf16 output mirror, split-K and ragged edges -- against fp32 torch on the same
bf16-rounded operands (products are exact in fp32, only the summation order differs: 1e-4)."""
import pytest
import torch

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


def _launches():
    from hulc2_b200 import _lib

    return _lib.load_library().hulc2_tma_gemm_count()


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (70, 48, 40), (257, 136, 104), (3, 2048, 160), (128, 2048, 2048), (300, 32, 192),
                                   (500, 704, 264), (4096, 256, 128), (1000, 184, 2048)])
def test_tma_nt_epilogues_and_mirror(M, N, K):
    from hulc2_b200 import ops

    A, B, bias, add = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    maskt = _rand(M, N, seed=5)
    keep = (torch.rand(M, N, generator=torch.Generator().manual_seed(6)) > 0.3).to(torch.uint8)
    C0 = _rand(M, N, seed=7)
    ref = _bf(A) @ _bf(B).t() + bias + add + C0
    ref = torch.relu(ref) * (maskt > 0) * keep * 1.25
    Cd = C0.to(DEV)
    Ad, Bd = A.to(DEV), B.to(DEV)
    C16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(M, N, K, Ad, K, 1, Bd, K, 1, Cd, N, bias=bias.to(DEV), add=add.to(DEV), ld_add=N,
             mask=maskt.to(DEV), ld_mask=N, keep=keep.to(DEV), ld_keep=N, keep_scale=1.25, relu=True, accumulate=True,
             A16=ops.to_bf16(Ad), B16=ops.to_bf16(Bd), C16=C16, ld16=N)
    torch.cuda.synchronize()
    assert_close(Cd, ref, 1e-4)
    assert torch.equal(C16.float().cpu(), Cd.cpu().bfloat16().float()), "bf16 mirror must be the rounded fp32 result"


def test_tma_identity_exact():
    """A = I reproduces bf16(B)^T exactly: isolates TMA box / swizzle / descriptor errors (K- and MN-major)."""
    from hulc2_b200 import ops

    K = 256
    n0 = _launches()
    A = torch.eye(K).to(DEV)
    B = _rand(192, K, seed=9).to(DEV)
    out = torch.empty(K, 192, device=DEV)
    ops.gemm(K, 192, K, A, K, 1, B, K, 1, out, 192, A16=ops.to_bf16(A), B16=ops.to_bf16(B))
    assert torch.equal(out.cpu(), _bf(B.cpu()).t().contiguous())
    # MN-major B: out2[m, n] = sum_k I[m, k] Bt[k, n] with Bt stored [K, N]
    Bt = _rand(K, 192, seed=10).to(DEV)
    out2 = torch.empty(K, 192, device=DEV)
    ops.gemm(K, 192, K, A, K, 1, Bt, 1, 192, out2, 192, A16=ops.to_bf16(A), B16=ops.to_bf16(Bt))
    assert torch.equal(out2.cpu(), _bf(Bt.cpu()))
    # MN-major A: out3[m, n] = sum_k At[k, m] I[n, k] with At stored [K, M]
    At = _rand(K, 320, seed=11).to(DEV)
    out3 = torch.empty(320, K, device=DEV)
    ops.gemm(320, K, K, At, 1, 320, A, K, 1, out3, K, A16=ops.to_bf16(At), B16=ops.to_bf16(A))
    assert torch.equal(out3.cpu(), _bf(At.cpu()).t().contiguous())
    assert _launches() - n0 == 3


@pytest.mark.parametrize("M,N,K", [(190, 176, 152), (4096, 2048, 184), (128, 1024, 2048)])
def test_tma_dgrad_wgrad_orientations(M, N, K):
    """dX = dY W (B MN-major) and dW = dY^T X (both MN-major, split-K on the long row axis)."""
    from hulc2_b200 import ops

    dY, W, X = _rand(M, N, seed=1).to(DEV), _rand(N, K, seed=2).to(DEV), _rand(M, K, seed=3).to(DEV)
    dY16, W16, X16 = ops.to_bf16(dY), ops.to_bf16(W), ops.to_bf16(X)
    dX = torch.empty(M, K, device=DEV)
    n0 = _launches()
    ops.gemm(M, K, N, dY, N, 1, W, 1, K, dX, K, A16=dY16, B16=W16)
    assert_close(dX, (_bf(dY.cpu()).double() @ _bf(W.cpu()).double()).float(), 1e-4, "dgrad")
    dW = torch.empty(N, K, device=DEV)
    ops.gemm(N, K, M, dY, 1, N, X, 1, K, dW, K, A16=dY16, B16=X16)
    assert_close(dW, (_bf(dY.cpu()).double().t() @ _bf(X.cpu()).double()).float(), 1e-4, "wgrad")
    assert _launches() - n0 == 2, "both contractions must take the TMA path"


def test_tma_subblocks_offsets_and_accumulate():
    """Column sub-blocks of wider matrices (W_ih[:, P:P+E], heads at a column offset) through element offsets."""
    from hulc2_b200 import ops

    M, H, In, P, E = 256, 512, 1120, 1024, 64
    x, W = _rand(M, E, seed=1).to(DEV), _rand(H, In, seed=2).to(DEV)
    out = _rand(M, H, seed=3).to(DEV)
    ref = out.cpu() + _bf(x.cpu()) @ _bf(W.cpu()[:, P : P + E]).t()
    ops.gemm(M, H, E, x, E, 1, W, In, 1, out, H, b_off=P, accumulate=True, A16=ops.to_bf16(x), B16=ops.to_bf16(W))
    assert_close(out, ref, 1e-4, "sub-block B")
    heads = torch.zeros(M, 184, device=DEV)
    Wh = _rand(182, H, seed=4).to(DEV)
    hid = _rand(M, H, seed=5).to(DEV)
    ops.gemm(M, 182, H, hid, H, 1, Wh, H, 1, heads, 184, A16=ops.to_bf16(hid), B16=ops.to_bf16(Wh))
    assert_close(heads[:, :182], _bf(hid.cpu()) @ _bf(Wh.cpu()).t(), 1e-4, "N=182 into ld 184")
    assert float(heads[:, 182:].abs().max()) == 0.0


def test_tma_falls_back_when_unaligned():
    """Row strides that are not multiples of 8 elements cannot be TMA boxes: the gather kernel serves them."""
    from hulc2_b200 import ops

    M, N, K = 70, 45, 33
    A, B = _rand(M, K, seed=1).to(DEV), _rand(N, K, seed=2).to(DEV)
    out = torch.empty(M, N, device=DEV)
    n0 = _launches()
    ops.gemm(M, N, K, A, K, 1, B, K, 1, out, N, A16=ops.to_bf16(A), B16=ops.to_bf16(B))
    assert_close(out, _bf(A.cpu()) @ _bf(B.cpu()).t(), 1e-4)
    assert _launches() == n0


@pytest.mark.parametrize("Mw,Kw,R", [(2048, 2048, 4096), (128, 2048, 4096), (2048, 128, 4096), (184, 2048, 4096), (2048, 160, 128),
                                     (64, 512, 4096), (384, 128, 4096), (32, 128, 200), (1024, 4096, 128), (72, 40, 1000)])
def test_tma_rowsum_bias_gradient(Mw, Kw, R):
    """Weight-gradient contraction dW[Mw,Kw] = g^T x over R rows (both operands MN-major) with the bias gradient
    sum_rows g riding along as an extra ones-operand MMA (hulc2_gemm_args.rowsum): every tiling / split-K choice."""
    from hulc2_b200 import ops

    g, x = _rand(R, Mw, seed=21), _rand(R, Kw, seed=22)
    g16, x16 = ops.to_bf16(g.to(DEV)), ops.to_bf16(x.to(DEV))
    dW = torch.empty(Mw, Kw, device=DEV)
    db = torch.full((Mw,), 7.0, device=DEV)           # overwritten, not accumulated
    n0 = _launches()
    ops.gemm16(Mw, Kw, R, g16, 1, Mw, x16, 1, Kw, dW, Kw, rowsum=db)
    torch.cuda.synchronize()
    assert _launches() - n0 == 1
    assert_close(dW, _bf(g).t() @ _bf(x), 1e-4)
    assert_close(db, _bf(g).double().sum(0).float(), 1e-5)
    # K-major A as well: rowsum[m] = sum_k A[m, k]
    A = _rand(Mw, R, seed=23)
    out = torch.empty(Mw, Kw, device=DEV)
    rs = torch.empty(Mw, device=DEV)
    ops.gemm16(Mw, Kw, R, ops.to_bf16(A.to(DEV)), R, 1, x16, 1, Kw, out, Kw, rowsum=rs)
    torch.cuda.synchronize()
    assert_close(out, _bf(A) @ _bf(x), 1e-4)
    assert_close(rs, _bf(A).double().sum(1).float(), 1e-5)


def test_rowsum_needs_the_tma_path():
    """No silent fallback: the fp32 CUDA-core path and the bf16 gather path do not serve rowsum -> error."""
    from hulc2_b200 import ops

    A, B = _rand(64, 64, seed=1).to(DEV), _rand(64, 64, seed=2).to(DEV)
    out, rs = torch.empty(64, 64, device=DEV), torch.empty(64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.gemm(64, 64, 64, A, 64, 1, B, 64, 1, out, 64, precision=0, rowsum=rs)
    with pytest.raises(RuntimeError):
        ops.gemm(64, 64, 64, A, 64, 1, B, 64, 1, out, 64, precision=1, rowsum=rs)       # no bf16 mirrors -> gather kernel
