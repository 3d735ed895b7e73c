"""Autograd operators of the HULC++ policy step, each backed by the C-ABI CUDA kernels.

Every ``torch.autograd.Function`` here launches only kernels of ``libhulc2_b200.so`` (through
``_lib.call``); torch provides device memory (``torch.empty``), views and the autograd tape.
There is no eager/PyTorch fallback: calling any op without the library or without a CUDA device
raises ``RuntimeError``.

Layout notes
  * activations inside the conv encoders are NHWC (the implicit-GEMM output layout);
  * the decoder runs time-major ``[S,B,*]`` so every recurrence step is one contiguous matrix;
  * 2-D operands may be row-strided views (``stride(1) == 1``), e.g. ``emb[:, 0]``.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import byref as C_byref
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvArgs, ConvbArgs, GemmArgs, call

_PRECISION = {"fp32": 0, "bf16": 1}
_precision = 0
_WS_BYTES = 96 << 20
_ws = {}


def set_precision(p: str) -> None:
    """'fp32' = CUDA-core path (1e-5 parity); 'bf16' = tcgen05 tensor-core contractions."""
    global _precision
    _precision = _PRECISION[p]


def get_precision() -> str:
    return "bf16" if _precision else "fp32"


def workspace(device) -> torch.Tensor:
    key = (device.type, device.index, _on_side)      # launches on the weight-gradient side stream get a scratch buffer of their own
    if key not in _ws:
        _ws[key] = torch.empty(_WS_BYTES, dtype=torch.uint8, device=device)
    return _ws[key]


# ----------------------------------------------------------------------------- weight gradients off the critical path
# The gradient of a layer's weights is read by nobody before the optimizer (or its bucket's all-reduce), while the gradient of
# its input is the next link of backward's dependency chain.  Most of those launches are latency-bound (M = 64..128 rows, or
# 4096 x 128 outputs), so running the weight-gradient contractions on a SIDE stream lets them fill the SMs the chain leaves
# idle; in the captured step the fork / join become graph edges.  Rules that keep this race-free:
#   * a section starts after everything enqueued so far on the caller's stream (side.wait_stream(main));
#   * its launches use their own workspace; the tensors they read / write are kept alive until the join, so the allocator cannot
#     hand their memory to a later launch of the main stream;
#   * the join (main waits for side) happens at the end of the backward pass (autograd final callback), before a bucket of the
#     gradient arena is all-reduced (ddp.GradBucketReducer), and before a parameter receives a SECOND gradient in the same step
#     (grad_buffer), because autograd adds that one on the main stream.
# MEASURED (r02, B200, bench step): 6.556 ms with the side stream, 6.575 ms without -- the weight-gradient contractions are
# wide enough (>= 128 CTAs after split-K) that they do not fit beside the chain's kernels, so the fork / join buys 0.3 %.  The
# mechanism therefore stays OFF by default (one stream, nothing to reason about); HULC2_SIDE_WGRAD=1 switches it on (A/B).
import os as _os0

side_wgrads = _os0.environ.get("HULC2_SIDE_WGRAD", "0") == "1"
_on_side = False
_side_streams: dict = {}
_side_active = False
_side_keepalive: list = []


def _in_grad_arena(t: torch.Tensor) -> bool:
    ptr = t.data_ptr()
    seen = set()
    for g, _o, _n in _grad_views.values():
        if id(g) in seen:
            continue
        seen.add(id(g))
        base = g.data_ptr()
        if base <= ptr < base + 4 * g.numel() and g.device == t.device:
            return True
    return False


def wgrad_section(fn, reads=(), outs=()) -> None:
    """Runs ``fn()`` -- weight-gradient launches only -- on the side stream of the current device (see above).  ``reads``: the
    operand tensors (kept alive until the join); ``outs``: the gradients written.  Only gradients that live in the optimizer's
    gradient arena qualify: autograd adopts such a slice as ``param.grad`` without launching anything, whereas any other tensor
    may be cloned or accumulated on the main stream right after the backward function returns -- then ``fn`` runs inline."""
    global _on_side, _side_active
    if not side_wgrads or _on_side or not outs or not all(o is None or _in_grad_arena(o) for o in outs):
        fn()
        return
    main = torch.cuda.current_stream()
    key = main.device.index
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device=main.device)
    side.wait_stream(main)
    _on_side = True
    try:
        with torch.cuda.stream(side):
            fn()
    finally:
        _on_side = False
    _side_keepalive.extend(t for t in reads if t is not None)
    if not _side_active:
        _side_active = True
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_side)
        except RuntimeError:          # not inside a backward pass: nothing will join later
            join_side()


def join_side() -> None:
    """The caller's stream waits for the weight-gradient side stream (no-op when nothing is outstanding)."""
    global _side_active
    if not _side_active:
        return
    for idx, side in _side_streams.items():
        torch.cuda.current_stream(side.device).wait_stream(side)
    _side_active = False
    _side_keepalive.clear()


def _p(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"hulc2_b200 ops take float32 tensors, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("hulc2_b200 ops need CUDA tensors (no CPU fallback)")
    return t


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """Returns a 2-D tensor with unit inner stride (copying only if the view is not expressible)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


# ----------------------------------------------------------------------------- raw kernel wrappers
def gemm(M, N, K, A, a_rs, a_ks, B, b_rs, b_ks, Cout, ldc, *, a_off=0, b_off=0, c_off=0, bias=None, add=None, ld_add=0,
         mask=None, ld_mask=0, keep=None, ld_keep=0, keep_scale=1.0, relu=False, accumulate=False, alpha=1.0,
         a_inner=0, a_rs_outer=0, a_rs_inner=0, c_inner=0, c_rs_outer=0, c_rs_inner=0, precision=None,
         A16=None, B16=None, C16=None, ld16=0, c16_off=0, rowsum=None):
    """C[m,n] = epi(alpha * sum_k A(m,k) B(n,k)); offsets are in elements.
    A16/B16: bf16 mirrors of the tensors A/B (same flat layout, see :func:`to_bf16`) -> TMA-fed tcgen05 kernel;
    C16: bf16 tensor that additionally receives the result (row stride ld16)."""
    ws = workspace(Cout.device)
    g = GemmArgs()
    g.M, g.N, g.K = int(M), int(N), int(K)
    g.A = A.data_ptr() + 4 * a_off
    g.a_rs, g.a_ks, g.a_inner, g.a_rs_outer, g.a_rs_inner = a_rs, a_ks, a_inner, a_rs_outer, a_rs_inner
    g.B = B.data_ptr() + 4 * b_off
    g.b_rs, g.b_ks = b_rs, b_ks
    g.C = Cout.data_ptr() + 4 * c_off
    g.ldc, g.c_inner, g.c_rs_outer, g.c_rs_inner = ldc, c_inner, c_rs_outer, c_rs_inner
    g.bias = _p(bias)
    g.add, g.ld_add = _p(add), ld_add
    g.mask, g.ld_mask = _p(mask), ld_mask
    g.keep, g.ld_keep, g.keep_scale = _p(keep), ld_keep, keep_scale
    g.relu, g.accumulate, g.alpha = int(relu), int(accumulate), alpha
    g.precision = _precision if precision is None else precision
    g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    if g.precision == 1:
        g.A16 = None if A16 is None else A16.data_ptr() + 2 * a_off
        g.B16 = None if B16 is None else B16.data_ptr() + 2 * b_off
        g.C16 = None if C16 is None else C16.data_ptr() + 2 * c16_off
        g.ld16 = ld16
    g.rowsum = _p(rowsum)
    _lib.tag(f"gemm[M={int(M)},N={int(N)},K={int(K)}]", 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N))
    call("hulc2_gemm", C.byref(g))


def to_bf16(t: torch.Tensor) -> torch.Tensor:
    """bf16 mirror of a dense fp32 tensor (same shape, element i <-> element i): the GEMM operand copies."""
    t = _f32(t)
    assert t.is_contiguous(), "to_bf16 mirrors dense tensors"
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    call("hulc2_f32_to_bf16", t.data_ptr(), out.data_ptr(), t.numel())
    return out


def mirror_like(t: torch.Tensor) -> torch.Tensor:
    """bf16 mirror with the shape of the dense fp32 tensor ``t``: the one its producer wrote, else a conversion."""
    t = _f32(t)
    if t.is_contiguous() and t.dim() >= 2:
        hit = _find_mirror(t.view(-1, t.shape[-1]))
        if hit is not None and hit[1] == t.shape[-1]:
            return hit[0].view(t.shape)
    return to_bf16(t)


def _pad8(n: int) -> int:
    return (int(n) + 7) & ~7


# Activation mirrors written by their PRODUCER (the "_m" entry points: LayerNorm, attention, positional add, the recurrence):
# data_ptr of the fp32 tensor -> (fp32 tensor, its version, bf16 mirror, mirror row stride, rows, cols).  The entry holds the
# fp32 tensor, so its memory cannot be recycled under a live key, and the version guards against in-place torch updates
# (autograd accumulating into a gradient buffer).  Cleared at every step boundary (zero_grad, optimizer.step, start of a
# training / validation / rollout step).  A consumer that finds no entry converts as before.
_act16: dict = {}
producer_mirrors = os.environ.get("HULC2_PRODUCER_MIRRORS", "1") != "0"     # A/B switch


def emit_mirrors() -> bool:
    return producer_mirrors and _precision == 1


def register_mirror(src: torch.Tensor, m16: torch.Tensor, ld: int, rows: int, cols: int) -> None:
    if len(_act16) >= 512:          # a caller that never reaches a step boundary (a long no-grad loop) must not pin memory
        _act16.clear()
    _act16[src.data_ptr()] = (src, src._version, m16, int(ld), int(rows), int(cols))


def clear_mirrors() -> None:
    _act16.clear()


def _find_mirror(x2: torch.Tensor):
    ent = _act16.get(x2.data_ptr())
    if ent is None:
        return None
    src, ver, m16, ld, rows, cols = ent
    if x2.dim() != 2 or x2.shape[0] != rows or x2.shape[1] != cols or not x2.is_contiguous() or x2._version != ver or src._version != ver:
        return None
    return m16, ld


def mirror2d(x2: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Compact bf16 mirror [rows, pad8(cols)] of a 2-D fp32 tensor with unit inner stride (rows may be strided)."""
    x2 = _f32(x2)
    hit = _find_mirror(x2)
    if hit is not None:
        return hit
    rows, cols = x2.shape
    ld = _pad8(cols)
    out = torch.empty(rows, ld, device=x2.device, dtype=torch.bfloat16)
    if x2.is_contiguous() and ld == cols:
        call("hulc2_f32_to_bf16", x2.data_ptr(), out.data_ptr(), x2.numel())
    elif rows > 0:
        assert x2.stride(1) == 1
        call("hulc2_f32_to_bf16_2d", x2.data_ptr(), _ld(x2), out.data_ptr(), ld, rows, cols)
    return out, ld


# bf16 mirrors of the parameters: one flat conversion of the optimizer's parameter arena (registered by FusedAdam) or
# one per tensor, reused by every contraction of the step; invalidated whenever the fp32 masters may have changed
# (start of every training / validation step and rollout re-plan, optimizer.step, load_state_dict).
_w16: dict = {}
_arenas: list = []      # [fp32 arena tensor, bf16 mirror or None, arena._version at conversion]


def register_param_arena(arena: torch.Tensor) -> None:
    _arenas[:] = [a for a in _arenas if a[0].data_ptr() != arena.data_ptr()]
    _arenas.append([arena, None, -1])


def invalidate_weight_mirrors() -> None:
    _w16.clear()
    _act16.clear()
    for a in _arenas:
        a[1] = None


# Gradient sinks: when the optimizer keeps all gradients in one arena (optim.FusedAdam), the FIRST gradient a backward
# kernel produces for a parameter in a step is written straight into that parameter's arena slice and returned as the
# gradient; with ``param.grad is None`` autograd's AccumulateGrad then adopts the tensor instead of launching
# ``grad += new`` (108 parameters = ~120 tiny kernels per step).  Later contributions in the same step (a parameter used
# twice) get fresh buffers and are accumulated by autograd as usual.
_grad_views: dict = {}      # parameter data_ptr -> (gradient arena, offset, numel)
_grad_claimed: set = set()
direct_grads = True


def register_grad_views(arena_g: torch.Tensor, params, offs) -> list:
    keys = []
    for prm, o in zip(params, offs):
        _grad_views[prm.data_ptr()] = (arena_g, int(o), prm.numel())
        keys.append(prm.data_ptr())
    return keys


def unregister_grad_views(keys) -> None:
    for k in keys:
        _grad_views.pop(k, None)


def begin_grad_step() -> None:
    """Start of a step (the gradient arena has just been zero-filled): every parameter's slice can be claimed again."""
    _grad_claimed.clear()
    _act16.clear()


def grad_buffer(W: torch.Tensor, zero: bool = False) -> torch.Tensor:
    """Buffer for a gradient of parameter ``W`` (same shape): its gradient-arena slice on the first request of a step,
    else a fresh tensor (zero-filled when ``zero``; an arena slice is zero at the start of a step)."""
    if direct_grads:
        key = W.data_ptr()
        ent = _grad_views.get(key)
        if ent is not None and key not in _grad_claimed and ent[2] == W.numel() and ent[0].device == W.device:
            _grad_claimed.add(key)
            g, o, n = ent
            return g[o : o + n].view(W.shape)
        if ent is not None and key in _grad_claimed:
            join_side()       # autograd will ADD this second gradient to the first one on the main stream: that one must be complete
    return torch.zeros_like(W, dtype=torch.float32) if zero else torch.empty_like(W, dtype=torch.float32)


def weight16(W: torch.Tensor) -> torch.Tensor:
    """bf16 mirror of a (contiguous) parameter tensor, same shape."""
    ptr, n = W.data_ptr(), W.numel()
    for a in _arenas:
        base = a[0].data_ptr()
        if base <= ptr and ptr + 4 * n <= base + 4 * a[0].numel() and a[0].device == W.device:
            if a[1] is None or a[2] != a[0]._version:     # torch in-place writes through any view bump the version
                a[1], a[2] = to_bf16(a[0]), a[0]._version
            off = (ptr - base) // 4
            return a[1][off : off + n].view(W.shape)
    key = (ptr, n, W._version)
    hit = _w16.get(key)
    if hit is None:
        hit = _w16[key] = to_bf16(W.detach().contiguous())
    return hit


def gemm16(M, N, K, A16, a_rs, a_ks, B16, b_rs, b_ks, Cout, ldc, *, a_off=0, b_off=0, c_off=0, bias=None, add=None, ld_add=0,
           mask=None, ld_mask=0, keep=None, ld_keep=0, keep_scale=1.0, relu=False, accumulate=False, alpha=1.0,
           C16=None, ld16=0, rowsum=None):
    """C[m,n] = epi(alpha * sum_k A(m,k) B(n,k)) with bf16-only operands (TMA-fed tcgen05 kernel): A(m,k) =
    A16[a_off + m*a_rs + k*a_ks], one of (a_rs, a_ks) being 1; same for B.  C fp32 (+ optional bf16 copy C16).
    rowsum: optional fp32 [M] that receives sum_k A(m,k) -- the bias gradient when this is a weight-gradient contraction."""
    ws = workspace(Cout.device)
    g = GemmArgs()
    g.M, g.N, g.K = int(M), int(N), int(K)
    g.A, g.B = None, None
    g.A16, g.a_rs, g.a_ks = A16.data_ptr() + 2 * a_off, a_rs, a_ks
    g.B16, g.b_rs, g.b_ks = B16.data_ptr() + 2 * b_off, b_rs, b_ks
    g.C, g.ldc = Cout.data_ptr() + 4 * c_off, ldc
    g.bias = _p(bias)
    g.add, g.ld_add = _p(add), ld_add
    g.mask, g.ld_mask = _p(mask), ld_mask
    g.keep, g.ld_keep, g.keep_scale = _p(keep), ld_keep, keep_scale
    g.relu, g.accumulate, g.alpha = int(relu), int(accumulate), alpha
    g.precision = 1
    g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    g.C16, g.ld16 = _p(C16), ld16
    g.rowsum = _p(rowsum)
    _lib.tag(f"gemm16[M={int(M)},N={int(N)},K={int(K)}]", 2.0 * M * N * K,
             2.0 * (M * K + N * K) + M * N * (4.0 + (2.0 if C16 is not None else 0.0) + (4.0 if accumulate or add is not None else 0.0)))
    call("hulc2_gemm", C.byref(g))


def _conv_args(F, Cin, H, W, Cout, k, stride, in_nhwc) -> ConvArgs:
    a = ConvArgs()
    a.F, a.C, a.H, a.W, a.Cout, a.KH, a.KW, a.stride, a.in_nhwc = F, Cin, H, W, Cout, k, k, stride, int(in_nhwc)
    a.precision = _precision
    return a


def colsum(x: torch.Tensor, ld: int, rows: int, cols: int, out: torch.Tensor, accumulate=False, x_off=0):
    ws = workspace(x.device)
    call("hulc2_colsum", x.data_ptr() + 4 * x_off, ld, rows, cols, out.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel())


# ----------------------------------------------------------------------------- MLP (stack of nn.Linear [+ReLU] [+dropout])
class MLPFunction(torch.autograd.Function):
    """y = L_n(...act(L_1(x))) with L_i = nn.Linear.  Covers goal_encoders.py:19-27/52-60, plan_proposal_net.py:25-40,
    proj_vis_lang.py:10-21, vision_network.py:49-52, the transformer FFN and every single nn.Linear on the path.
    args: x, n_layers, relu flags tuple, keeps tuple (u8 mask or None per layer), keep_scale, then W1, b1, W2, b2, ..."""

    @staticmethod
    def forward(ctx, x, relus, keeps, keep_scale, *wb):
        x2 = _rows2d(_f32(x))
        M = x2.shape[0]
        n = len(wb) // 2
        fp32_only = isinstance(relus, _Fp32Flags)      # see mlp(..., fp32=True)
        ctx.b16 = _precision == 1 and not fp32_only and M > 0 and all(wb[2 * i].shape[1] % 8 == 0 for i in range(n))
        if ctx.b16:
            # bf16 operand mirrors: every layer's epilogue also emits the bf16 copy the next layer (and the weight
            # gradient) multiplies; the fp32 activations stay for the ReLU masks and non-GEMM consumers
            cur16, ld = mirror2d(x2)
            acts, mirrors = [], [cur16]
            for i in range(n):
                W, b = wb[2 * i], wb[2 * i + 1]
                N, K = W.shape
                y = torch.empty(M, N, device=x.device, dtype=torch.float32)
                last = i == n - 1
                y16 = None if last else torch.empty(M, _pad8(N), device=x.device, dtype=torch.bfloat16)
                gemm16(M, N, K, cur16, ld, 1, weight16(W), K, 1, y, N, bias=b, relu=relus[i], keep=keeps[i], ld_keep=N,
                       keep_scale=keep_scale, C16=y16, ld16=_pad8(N))
                acts.append(y)
                if not last:
                    mirrors.append(y16)
                    cur16, ld = y16, _pad8(N)
            ctx.relus, ctx.keeps, ctx.keep_scale, ctx.n = relus, keeps, keep_scale, n
            ctx.x_shape = x.shape
            ctx.save_for_backward(x2, *acts, *wb, *mirrors)
            return acts[-1].view(*x.shape[:-1], acts[-1].shape[1])
        acts = []
        cur, ld = x2, _ld(x2)
        for i in range(n):
            W, b = wb[2 * i], wb[2 * i + 1]
            N, K = W.shape
            y = torch.empty(M, N, device=x.device, dtype=torch.float32)
            if M > 0:
                gemm(M, N, K, cur, ld, 1, W, K, 1, y, N, bias=b, relu=relus[i], keep=keeps[i], ld_keep=N, keep_scale=keep_scale,
                     precision=0 if fp32_only else None)
            acts.append(y)
            cur, ld = y, N
        ctx.fp32_only = fp32_only
        ctx.relus, ctx.keeps, ctx.keep_scale, ctx.n = relus, keeps, keep_scale, n
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, *acts, *wb)
        return acts[-1].view(*x.shape[:-1], acts[-1].shape[1])

    @staticmethod
    def backward(ctx, dout):
        n = ctx.n
        saved = ctx.saved_tensors
        x2, acts, wb = saved[0], saved[1 : 1 + n], saved[1 + n : 1 + 3 * n]
        mirrors = saved[1 + 3 * n :]
        M = x2.shape[0]
        g = _rows2d(dout.contiguous())
        if g.data_ptr() == dout.data_ptr() and (ctx.relus[n - 1] or ctx.keeps[n - 1] is not None):
            g = g.clone()
        if ctx.relus[n - 1]:
            assert ctx.keeps[n - 1] is None, "dropout after the last layer of an MLP stack is not supported"
            call("hulc2_relu_mask", g.data_ptr(), acts[n - 1].data_ptr(), g.data_ptr(), g.numel())
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        dx = None
        if ctx.b16:
            g16, ldg = mirror2d(g)
            for i in range(n - 1, -1, -1):
                W = wb[2 * i]
                N, K = W.shape
                in16 = mirrors[i]
                ld_in = in16.shape[1]
                db = grad_buffer(wb[2 * i + 1]) if ctx.needs_input_grad[5 + 2 * i] else None
                if ctx.needs_input_grad[4 + 2 * i]:
                    dW = grad_buffer(W)
                    # dW = g^T inp (both MN-major); the bias gradient = row sums of g^T rides along (one more narrow MMA).  Off the
                    # critical path: the input gradient below is what the rest of backward waits for
                    wgrad_section(lambda g16=g16, ldg=ldg, in16=in16, ld_in=ld_in, dW=dW, db=db, N=N, K=K:
                                  gemm16(N, K, M, g16, 1, ldg, in16, 1, ld_in, dW, K, rowsum=db), (g16, in16), (dW, db))
                    grads[2 * i] = dW
                elif db is not None:
                    colsum(g, N, M, N, db)
                grads[2 * i + 1] = db
                if i > 0 or ctx.needs_input_grad[0]:
                    gi = torch.empty(M, K, device=W.device, dtype=torch.float32)
                    gi16 = torch.empty(M, _pad8(K), device=W.device, dtype=torch.bfloat16) if i > 0 else None
                    if i > 0:
                        gemm16(M, K, N, g16, ldg, 1, weight16(W), 1, K, gi, K, mask=acts[i - 1] if ctx.relus[i - 1] else None,
                               ld_mask=K, keep=ctx.keeps[i - 1], ld_keep=K, keep_scale=ctx.keep_scale, C16=gi16, ld16=_pad8(K))
                    else:
                        gemm16(M, K, N, g16, ldg, 1, weight16(W), 1, K, gi, K)
                    g, g16, ldg = gi, gi16, _pad8(K)
                    if i == 0:
                        dx = gi.view(ctx.x_shape)
            return (dx, None, None, None, *grads)
        prec = 0 if ctx.fp32_only else None
        for i in range(n - 1, -1, -1):
            W = wb[2 * i]
            N, K = W.shape
            inp = x2 if i == 0 else acts[i - 1]
            ld_in = _ld(x2) if i == 0 else K
            if ctx.needs_input_grad[4 + 2 * i]:
                dW = grad_buffer(W)
                gemm(N, K, M, g, 1, N, inp, 1, ld_in, dW, K, precision=prec)           # dW = g^T inp
                grads[2 * i] = dW
            if ctx.needs_input_grad[5 + 2 * i]:
                db = grad_buffer(wb[2 * i + 1])
                colsum(g, N, M, N, db)
                grads[2 * i + 1] = db
            if i > 0 or ctx.needs_input_grad[0]:
                gi = torch.empty(M, K, device=W.device, dtype=torch.float32)
                if M > 0:
                    if i > 0:
                        gemm(M, K, N, g, N, 1, W, 1, K, gi, K, mask=acts[i - 1] if ctx.relus[i - 1] else None, ld_mask=K,
                             keep=ctx.keeps[i - 1], ld_keep=K, keep_scale=ctx.keep_scale, precision=prec)
                    else:
                        gemm(M, K, N, g, N, 1, W, 1, K, gi, K, precision=prec)
                g = gi
                if i == 0:
                    dx = gi.view(ctx.x_shape)
        return (dx, None, None, None, *grads)


class _Fp32Flags(tuple):
    """ReLU flags of an MLP stack that must run on the fp32 CUDA-core GEMM even under precision 'bf16'."""


# Under precision 'bf16' the visuo-lingual contrastive path (hulc2.py:472-508) -- plan_recognition.fc (128 -> 4096, M = B rows),
# ProjVisLang (4096 -> 128 -> 32, 32 -> 128 -> 32) -- runs on the fp32 GEMM: InfoNCE only sees the DIFFERENCE between the rows
# of a batch whose common mode is ~100x larger, so bf16 operand rounding (2^-9 of the common mode) reaches 4-6 % of every
# gradient that flows back through it (measured at B = 64: static-encoder tail 5.7 % off, cosine 0.994; with these ~400 MFLOP
# of M <= 128 contractions in fp32: see profiles/parity_r02_bf16_oracle_B64.json).  HULC2_CLIP_FP32=0 restores bf16 (A/B switch).
import os as _os

clip_fp32 = _os.environ.get("HULC2_CLIP_FP32", "1") != "0"
clip_fp32_level = int(_os.environ.get("HULC2_CLIP_FP32", "2") or 0)     # 1 = ProjVisLang only, 2 = + plan_recognition.fc


def mlp(x, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], relus: Sequence[bool], keeps=None, keep_scale=1.0, fp32: bool = False):
    keeps = tuple(keeps) if keeps is not None else (None,) * len(layers)
    flat = [t for wb in layers for t in wb]
    flags = tuple(bool(r) for r in relus)
    return MLPFunction.apply(x, _Fp32Flags(flags) if fp32 else flags, keeps, float(keep_scale), *flat)


def linear(x, W, b, relu=False, fp32: bool = False):
    return mlp(x, [(W, b)], [relu], fp32=fp32)


# ----------------------------------------------------------------------------- conv trunks
def _conv_trunk_fwd(x, w1, b1, w2, b2, w3, b3):
    """conv(k8,s4)+ReLU -> conv(k4,s2)+ReLU -> conv(k3,s1)+ReLU; x NCHW, outputs NHWC."""
    F_, Cin, H, W = x.shape
    dev = x.device
    ws = workspace(dev)

    def osz(h, k, s):
        return (h - k) // s + 1

    H1, W1 = osz(H, 8, 4), osz(W, 8, 4)
    H2, W2 = osz(H1, 4, 2), osz(W1, 4, 2)
    H3, W3 = osz(H2, 3, 1), osz(W2, 3, 1)
    y1 = torch.empty(F_, H1, W1, 32, device=dev, dtype=torch.float32)
    y2 = torch.empty(F_, H2, W2, 64, device=dev, dtype=torch.float32)
    y3 = torch.empty(F_, H3, W3, 64, device=dev, dtype=torch.float32)
    w2p = torch.empty(64, 4, 4, 32, device=dev, dtype=torch.float32)
    w3p = torch.empty(64, 3, 3, 64, device=dev, dtype=torch.float32)
    call("hulc2_permute_conv_weight", w2.data_ptr(), w2p.data_ptr(), 64, 32, 4, 4, 0, 0)
    call("hulc2_permute_conv_weight", w3.data_ptr(), w3p.data_ptr(), 64, 64, 3, 3, 0, 0)
    for (xin, w, b, y, cin, h, wd, cout, k, s, nhwc) in (
        (x, w1, b1, y1, Cin, H, W, 32, 8, 4, 0),
        (y1, w2p, b2, y2, 32, H1, W1, 64, 4, 2, 1),
        (y2, w3p, b3, y3, 64, H2, W2, 64, 3, 1, 1),
    ):
        a = _conv_args(F_, cin, h, wd, cout, k, s, nhwc)
        a.x, a.w, a.bias, a.y, a.relu = xin.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 1
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.tag(f"conv_fwd[F={F_},{cin}x{h}x{wd}->{cout},k{k}s{s}]", 2.0 * y.numel() * cin * k * k)
        call("hulc2_conv2d_fwd", C.byref(a))
    return y1, y2, y3


def _conv_trunk_bwd(x, y1, y2, dz3, w2, w3, need):
    """dz3 = gradient wrt the PRE-activation of conv3 (already ReLU-masked), NHWC. Returns param grads."""
    F_, Cin, H, W = x.shape
    dev = x.device
    ws = workspace(dev)
    H1, W1, H2, W2 = y1.shape[1], y1.shape[2], y2.shape[1], y2.shape[2]
    H3, W3 = dz3.shape[1], dz3.shape[2]

    def wgrad(xin, dz, cin, h, wd, cout, k, s, nhwc, shape_oihw):
        a = _conv_args(F_, cin, h, wd, cout, k, s, nhwc)
        dw = torch.empty(cout, cin * k * k, device=dev, dtype=torch.float32)
        a.x, a.dy, a.dw, a.accumulate = xin.data_ptr(), dz.data_ptr(), dw.data_ptr(), 0
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.tag(f"conv_wgrad[F={F_},{cin}x{h}x{wd}->{cout},k{k}s{s}]", 2.0 * dz.numel() * cin * k * k)
        call("hulc2_conv2d_wgrad", C.byref(a))
        if nhwc:
            out = torch.empty(shape_oihw, device=dev, dtype=torch.float32)
            call("hulc2_permute_conv_weight", dw.data_ptr(), out.data_ptr(), cout, cin, k, k, 1, 0)
            return out
        return dw.view(shape_oihw)

    def bgrad(dz, cout):
        db = torch.empty(cout, device=dev, dtype=torch.float32)
        colsum(dz, cout, dz.numel() // cout, cout, db)
        return db

    def dgrad(dz, w_oihw, xmask, cin, h, wd, cout, k, s):
        whwoi = torch.empty(k, k, cout, cin, device=dev, dtype=torch.float32)
        call("hulc2_permute_conv_weight", w_oihw.data_ptr(), whwoi.data_ptr(), cout, cin, k, k, 2, 0)
        dx = torch.empty(F_, h, wd, cin, device=dev, dtype=torch.float32)
        a = _conv_args(F_, cin, h, wd, cout, k, s, 1)
        a.dy, a.w, a.dx, a.xmask = dz.data_ptr(), whwoi.data_ptr(), dx.data_ptr(), xmask.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.tag(f"conv_dgrad[F={F_},{cin}x{h}x{wd}<-{cout},k{k}s{s}]", 2.0 * dz.numel() * cin * k * k)
        call("hulc2_conv2d_dgrad", C.byref(a))
        return dx

    g = {}
    if need[2]:
        g["w3"] = wgrad(y2, dz3, 64, H2, W2, 64, 3, 1, 1, (64, 64, 3, 3))
        g["b3"] = bgrad(dz3, 64)
    dz2 = dgrad(dz3, w3, y2, 64, H2, W2, 64, 3, 1)
    if need[1]:
        g["w2"] = wgrad(y1, dz2, 32, H1, W1, 64, 4, 2, 1, (64, 32, 4, 4))
        g["b2"] = bgrad(dz2, 64)
    dz1 = dgrad(dz2, w2, y1, 32, H1, W1, 64, 4, 2)
    if need[0]:
        g["w1"] = wgrad(x, dz1, Cin, H, W, 32, 8, 4, 0, (32, Cin, 8, 8))
        g["b1"] = bgrad(dz1, 32)
    return g


# ----------------------------------------------------------------------------- bf16 conv trunk (csrc/conv_sm100.cu)
def _osz(h, k, s):
    return (h - k) // s + 1


def convb_trunk_supported(x) -> bool:
    """The persistent tcgen05 trunk serves precision 'bf16' on the 3-conv stacks of the reference (k8s4 -> k4s2 -> k3s1)."""
    x = x[0] if isinstance(x, (tuple, list)) else x
    if _precision != 1 or x.dim() != 4:
        return False
    lib = _lib.load_library()
    Cin = x.shape[1]
    return bool(lib.hulc2_convb_supported(16 * Cin, 32, 2, 2, 1) and lib.hulc2_convb_supported(32, 64, 4, 4, 2)
                and lib.hulc2_convb_supported(64, 64, 3, 3, 1) and x.shape[2] >= 8 and x.shape[3] >= 8)


def _cb(F, C, H, W, Cout, k, stride) -> ConvbArgs:
    a = ConvbArgs()
    a.F, a.C, a.H, a.W, a.Cout, a.KH, a.KW, a.stride = F, C, H, W, Cout, k, k, stride
    return a


def _bf16(*shape, device):
    """bf16 activation buffer with 64 elements of slack behind it: the halo kernels read 32/48-channel pixels as 64-element TMA
    rows (the overlap meets zero weights / ignored accumulator columns), so the last pixel's row ends past the tensor."""
    n = 1
    for d in shape:
        n *= int(d)
    return torch.empty(n + 64, device=device, dtype=torch.bfloat16)[:n].view(*shape)


def pack_conv_weight(w: torch.Tensor, mode: int, stride: int = 1) -> torch.Tensor:
    Cout, Cin, KH, KW = w.shape
    wp = torch.empty(w.numel(), device=w.device, dtype=torch.bfloat16)
    call("hulc2_convb_pack_weight", w.data_ptr(), wp.data_ptr(), Cout, Cin, KH, KW, mode, stride)
    return wp


relu_sign_bits = os.environ.get("HULC2_RELU_BITS", "1") != "0"      # A/B switch


def convb_fwd(x, wp, bias, F, C, H, W, Cout, k, stride, relu=True, name="conv", sign_bits=False):
    """x bf16 NHWC [F,H,W,C] -> bf16 NHWC [F,OH,OW,Cout].  sign_bits: also return (y > 0) packed 1 bit per element (uint8 tensor,
    NHWC element order) when the kernel that ran can write it, else None -- what the next layer's input gradient masks with."""
    OH, OW = _osz(H, k, stride), _osz(W, k, stride)
    y = _bf16(F, OH, OW, Cout, device=x.device)
    a = _cb(F, C, H, W, Cout, k, stride)
    a.x, a.w, a.bias, a.y, a.relu = x.data_ptr(), wp.data_ptr(), _p(bias), y.data_ptr(), int(relu)
    ws = workspace(x.device)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()        # scratch for re-tiled weights (conv1 halo path)
    bits = None
    if sign_bits and relu and relu_sign_bits:
        bits = torch.empty(2 * ((y.numel() + 15) // 16), device=x.device, dtype=torch.uint8)
        a.mask_bits = bits.data_ptr()
    _lib.tag(f"convb_fwd[{name},F={F},{C}x{H}x{W}->{Cout},k{k}s{stride}]", 2.0 * y.numel() * C * k * k,
             2.0 * (x.numel() + y.numel()) + (bits.numel() if bits is not None else 0))
    call("hulc2_convb_fwd", C_byref(a))
    if not sign_bits:
        return y
    return y, (bits if bits is not None and a.mask_bits_written else None)


def convb_dgrad(dy, w_oihw, xmask, F, C, H, W, Cout, k, stride, name="conv", sign_bits=None) -> torch.Tensor:
    """dy bf16 [F,OH,OW,Cout] -> dx bf16 [F,H,W,C], zeroed where xmask <= 0 (sign_bits: the same mask, 1 bit per element, from
    ``convb_fwd(..., sign_bits=True)`` -- read instead of the bf16 activations by the halo kernels)."""
    wp = pack_conv_weight(w_oihw, 2, stride)
    dx = _bf16(F, H, W, C, device=dy.device)
    a = _cb(F, C, H, W, Cout, k, stride)
    a.dy, a.w, a.dx, a.xmask = dy.data_ptr(), wp.data_ptr(), dx.data_ptr(), _p(xmask)
    a.mask_bits = _p(sign_bits)
    # algorithmic bytes: with sign bits the mask costs 1/16 of the activation bytes
    mask_bytes = 0.0 if xmask is None else (dx.numel() / 8.0 if sign_bits is not None else 2.0 * dx.numel())
    _lib.tag(f"convb_dgrad[{name},F={F},{C}x{H}x{W}<-{Cout},k{k}s{stride}]", 2.0 * dy.numel() * C * k * k,
             2.0 * (dy.numel() + dx.numel()) + mask_bytes)
    call("hulc2_convb_dgrad", C_byref(a))
    return dx


def convb_wgrad(x, dy, F, C, H, W, Cout, k, stride, dw_shape, dw_layout=0, name="conv", param=None, bias_param=None):
    """(dw fp32 in dw_shape (OIHW), db fp32 [Cout]) from the bf16 NHWC input x and output gradient dy.  ``param``: the weight
    tensor the gradient belongs to (lets the result land directly in the optimizer's gradient arena, see grad_buffer)."""
    ws = workspace(x.device)
    dw = grad_buffer(param) if param is not None and tuple(param.shape) == tuple(dw_shape) else torch.empty(dw_shape, device=x.device, dtype=torch.float32)
    db = grad_buffer(bias_param) if bias_param is not None and bias_param.numel() == Cout else torch.empty(Cout, device=x.device, dtype=torch.float32)
    a = _cb(F, C, H, W, Cout, k, stride)
    a.x, a.dy, a.dw, a.db, a.dw_layout = x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), dw_layout
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.tag(f"convb_wgrad[{name},F={F},{C}x{H}x{W}->{Cout},k{k}s{stride}]", 2.0 * dy.numel() * C * k * k,
             2.0 * (F * H * W * C + dy.numel()))
    call("hulc2_convb_wgrad", C_byref(a))
    return dw, db


class U8Frames:
    """Camera frames as the dataset stores them: uint8 HWC ``[N,H,W,C]`` (episode_utils.py:61-86), plus what the
    reference's dataloader workers would apply before the encoder sees them -- the window slice with
    pad_with_repetition (``win_start [B]`` int64 / ``win_len [B]`` int32 into a resident frame store;
    base_dataset.py:121-163) and the RandomShiftsAug draw (``shift [F,2]`` int32 = (dx,dy) in [-pad,pad];
    transforms.py:85-106).  Scale + Normalize(0.5,0.5) always apply.  Quacks like an NCHW tensor for shape queries;
    the conversion happens inside the trunk's pack kernel (bf16) or ``to_f32`` (fp32 precision)."""

    def __init__(self, u8: torch.Tensor, shift: Optional[torch.Tensor] = None, win_start: Optional[torch.Tensor] = None,
                 win_len: Optional[torch.Tensor] = None, S: int = 1):
        if u8.dtype != torch.uint8 or u8.dim() != 4:
            raise ValueError("U8Frames expects a uint8 [N,H,W,C] tensor")
        if not u8.is_cuda:
            raise RuntimeError("hulc2_b200: frames must be CUDA tensors (no CPU fallback)")
        self.u8 = u8.contiguous()
        self.S = int(S)
        self.win_start = None if win_start is None else win_start.to(torch.int64).contiguous()
        self.win_len = None if win_len is None else win_len.to(torch.int32).contiguous()
        self.F = int(self.win_start.numel()) * self.S if self.win_start is not None else int(u8.shape[0])
        self.shift = None if shift is None else shift.to(torch.int32).reshape(-1, 2).contiguous()
        if self.shift is not None and self.shift.shape[0] != self.F:
            raise ValueError(f"shift must hold one (dx,dy) per output frame: {tuple(self.shift.shape)} vs F={self.F}")
        self.device = u8.device

    @property
    def shape(self):
        _, H, W, C = self.u8.shape
        return (self.F, C, H, W)

    def dim(self) -> int:
        return 4

    def _args(self):
        F_, C, H, W = self.shape
        return (self.u8.data_ptr(), _p(self.win_start), _p(self.win_len), _p(self.shift)), (F_, self.S, C, H, W)

    def to_f32(self) -> torch.Tensor:
        """fp32 NCHW frames in [-1,1]: exactly the tensor the reference's dataloader would have produced."""
        F_, C, H, W = self.shape
        out = torch.empty(F_, C, H, W, device=self.device, dtype=torch.float32)
        ptrs, dims = self._args()
        _lib.tag(f"frames_u8_to_f32[F={F_},{C}x{H}x{W}]", 0.0, 5.0 * F_ * C * H * W)
        call("hulc2_frames_u8_to_f32", *ptrs, out.data_ptr(), *dims)
        return out

    def pack_into(self, xs_ptr: int) -> None:
        F_, C, H, W = self.shape
        ptrs, dims = self._args()
        _lib.tag(f"frames_u8_pack[F={F_},{C}x{H}x{W}]", 0.0, 3.0 * F_ * C * H * W)            # 1 B in, 2 B out per pixel-channel
        call("hulc2_frames_u8_pack_bf16", *ptrs, xs_ptr, *dims)


def _frame_groups(x):
    """A frame operand is one NCHW tensor, a ``U8Frames``, or a tuple of them (the same camera of several modalities,
    encoded by one trunk call).  -> (list of contiguous fp32 [F_i,C,H,W] tensors / U8Frames, (F_total, C, H, W))."""
    groups = [t if isinstance(t, U8Frames) else _f32(t).contiguous() for t in (x if isinstance(x, (tuple, list)) else (x,))]
    shp = tuple(groups[0].shape[1:])
    assert all(t.dim() == 4 and tuple(t.shape[1:]) == shp for t in groups), "frame groups must share C,H,W"
    return groups, (sum(t.shape[0] for t in groups), *shp)


def _frames_tensor(x) -> torch.Tensor:
    groups, _ = _frame_groups(x)
    groups = [t.to_f32() if isinstance(t, U8Frames) else t for t in groups]
    return groups[0] if len(groups) == 1 else torch.cat(groups, 0)   # fp32 reference-precision path only


def pack_frames(x) -> torch.Tensor:
    """Frames (fp32 NCHW tensors or uint8 ``U8Frames``; one or a tuple of frame groups) -> bf16 [F, H/4, W/4, 16*C]
    (space-to-depth by the first conv's stride); groups are packed back to back, so no concatenated fp32 copy is ever
    made -- and for uint8 frames no fp32 frame exists at all."""
    groups, (F_, Cin, H, W) = _frame_groups(x)
    per_frame = (H // 4) * (W // 4) * 16 * Cin
    # (_bf16 leaves slack behind the last pixel: conv1's halo path reads packed pixels as 64-element rows at a 48-element pitch)
    xs = _bf16(F_, H // 4, W // 4, 16 * Cin, device=groups[0].device)
    f0 = 0
    for t in groups:
        if isinstance(t, U8Frames):
            t.pack_into(xs.data_ptr() + 2 * f0 * per_frame)
        else:
            _lib.tag(f"pack_frames[F={t.shape[0]},{Cin}x{H}x{W}]", 0.0, 6.0 * t.numel())
            call("hulc2_pack_frames_bf16", t.data_ptr(), xs.data_ptr() + 2 * f0 * per_frame, t.shape[0], Cin, H, W)
        f0 += t.shape[0]
    return xs


def _convb_trunk_fwd(x, w1, b1, w2, b2, w3, b3):
    _, (F_, Cin, H, W) = _frame_groups(x)
    xs = pack_frames(x)
    H4, W4 = H // 4, W // 4
    H1, W1 = H4 - 1, W4 - 1
    H2, W2 = _osz(H1, 4, 2), _osz(W1, 4, 2)
    # sign bits for conv2's output only: on conv1 the forward pays more for writing them than conv2's input gradient gains
    y1, m1 = convb_fwd(xs, pack_conv_weight(w1, 1), b1, F_, 16 * Cin, H4, W4, 32, 2, 1, name="c1"), None
    y2, m2 = convb_fwd(y1, pack_conv_weight(w2, 0), b2, F_, 32, H1, W1, 64, 4, 2, name="c2", sign_bits=True)
    y3 = convb_fwd(y2, pack_conv_weight(w3, 0), b3, F_, 64, H2, W2, 64, 3, 1, name="c3")
    return xs, y1, y2, y3, (m1, m2)


def _bits_saved(bits):
    """(m1, m2) -> tensors for save_for_backward (an empty tensor stands for "no sign bits")."""
    return tuple(m if m is not None else torch.empty(0, dtype=torch.uint8) for m in bits)


def _convb_trunk_bwd(xs, y1, y2, dz3, w1, w2, w3, biases=(None, None, None), bits=(None, None)):
    """dz3 = bf16 gradient wrt conv3's pre-activation (already ReLU-masked).  Returns the six parameter gradients.
    bits: ReLU sign bits of (y1, y2) from the forward (convb_fwd(sign_bits=True)); empty / None = mask with the activations."""
    m1, m2 = (m if m is not None and m.numel() > 0 else None for m in bits)
    F_, H4, W4, C16 = xs.shape
    H1, W1, H2, W2 = y1.shape[1], y1.shape[2], y2.shape[1], y2.shape[2]
    dw3, db3 = convb_wgrad(y2, dz3, F_, 64, H2, W2, 64, 3, 1, tuple(w3.shape), name="c3", param=w3, bias_param=biases[2])
    dz2 = convb_dgrad(dz3, w3, y2, F_, 64, H2, W2, 64, 3, 1, name="c3", sign_bits=m2)
    dw2, db2 = convb_wgrad(y1, dz2, F_, 32, H1, W1, 64, 4, 2, tuple(w2.shape), name="c2", param=w2, bias_param=biases[1])
    dz1 = convb_dgrad(dz2, w2, y1, F_, 32, H1, W1, 64, 4, 2, name="c2", sign_bits=m1)
    dw1, db1 = convb_wgrad(xs, dz1, F_, C16, H4, W4, 32, 2, 1, tuple(w1.shape), dw_layout=1, name="c1", param=w1, bias_param=biases[0])
    return {"w1": dw1, "b1": db1, "w2": dw2, "b2": db2, "w3": dw3, "b3": db3}


class StaticConvSSM(torch.autograd.Function):
    """Static-camera trunk: 3 convs + SpatialSoftmax (vision_network.py:55-58, 100-108) -> [F, 2*64]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, x_map, y_map, temperature):
        _, (F_, _c, _h, _w) = _frame_groups(x)
        out = torch.empty(F_, 128, device=w1.device, dtype=torch.float32)
        ctx.bf16 = convb_trunk_supported(x)
        ctx.biases = (b1, b2, b3)              # only their identity is used: the bias gradients land in their arena slices
        if not ctx.bf16:
            x = _frames_tensor(x)
        if ctx.bf16:
            xs, y1, y2, y3, bits = _convb_trunk_fwd(x, w1, b1, w2, b2, w3, b3)
            HW = y3.shape[1] * y3.shape[2]
            _lib.tag(f"spatial_softmax_fwd_bf16[F={F_},HW={HW}]", 0.0, 2.0 * y3.numel() + 4.0 * out.numel())
            # the softmax statistics are kept for the backward when a gradient will be asked for (one pass there instead of three)
            stats = None
            if any(ctx.needs_input_grad) and _lib.load_library().hulc2_spatial_softmax_stats_supported(HW, 64):
                stats = torch.empty(F_, 128, device=w1.device, dtype=torch.float32)
                call("hulc2_spatial_softmax_fwd_bf16_stats", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
                     out.data_ptr(), stats.data_ptr(), F_, HW, 64)
            else:
                call("hulc2_spatial_softmax_fwd_bf16", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
                     out.data_ptr(), F_, HW, 64)
            ctx.has_stats = stats is not None
            ctx.save_for_backward(xs, y1, y2, y3, w1, w2, w3, x_map, y_map, temperature, *_bits_saved(bits),
                                  *((out, stats) if stats is not None else ()))
            return out
        y1, y2, y3 = _conv_trunk_fwd(x, w1, b1, w2, b2, w3, b3)
        HW = y3.shape[1] * y3.shape[2]
        call("hulc2_spatial_softmax_fwd", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
             out.data_ptr(), F_, HW, 64)
        ctx.save_for_backward(x, y1, y2, y3, w2, w3, x_map, y_map, temperature, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        if ctx.bf16:
            xs, y1, y2, y3, w1, w2, w3, x_map, y_map, temperature, m1, m2, *extra = ctx.saved_tensors
            F_ = xs.shape[0]
            HW = y3.shape[1] * y3.shape[2]
            dz3 = torch.empty_like(y3)
            dtemp = torch.zeros(1, device=xs.device, dtype=torch.float32) if ctx.needs_input_grad[9] else None
            _lib.tag(f"spatial_softmax_bwd_bf16[F={F_},HW={HW}]", 0.0, 4.0 * y3.numel() + 4.0 * dout.numel())  # y3 in, dz3 out
            if ctx.has_stats:
                fwd_out, stats = extra
                call("hulc2_spatial_softmax_bwd_bf16_stats", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
                     fwd_out.data_ptr(), stats.data_ptr(), dout.data_ptr(), dz3.data_ptr(), _p(dtemp), F_, HW, 64, 1)
            else:
                call("hulc2_spatial_softmax_bwd_bf16", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
                     dout.data_ptr(), dz3.data_ptr(), _p(dtemp), F_, HW, 64, 1)
            g = _convb_trunk_bwd(xs, y1, y2, dz3, w1, w2, w3, ctx.biases, bits=(m1, m2))
            return (None, g["w1"], g["b1"], g["w2"], g["b2"], g["w3"], g["b3"], None, None, dtemp)
        x, y1, y2, y3, w2, w3, x_map, y_map, temperature, out = ctx.saved_tensors
        F_ = x.shape[0]
        HW = y3.shape[1] * y3.shape[2]
        dz3 = torch.empty_like(y3)
        dtemp = None
        if ctx.needs_input_grad[9]:
            dtemp = torch.zeros(1, device=x.device, dtype=torch.float32)
        call("hulc2_spatial_softmax_bwd", y3.data_ptr(), x_map.data_ptr(), y_map.data_ptr(), temperature.data_ptr(),
             out.data_ptr(), dout.data_ptr(), dz3.data_ptr(), _p(dtemp), F_, HW, 64, 1)
        g = _conv_trunk_bwd(x, y1, y2, dz3, w2, w3, (True, True, True))
        return (None, g["w1"], g["b1"], g["w2"], g["b2"], g["w3"], g["b3"], None, None, dtemp)


class GripperConvFlatten(torch.autograd.Function):
    """Gripper trunk: 3 convs + nn.Flatten in (C,H,W) order (vision_network_gripper.py:11-22) -> [F, 64*7*7]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        _, (F_, _c, _h, _w) = _frame_groups(x)
        ctx.bf16 = convb_trunk_supported(x)
        ctx.biases = (b1, b2, b3)              # only their identity is used: the bias gradients land in their arena slices
        if not ctx.bf16:
            x = _frames_tensor(x)
        if ctx.bf16:
            xs, y1, y2, y3, bits = _convb_trunk_fwd(x, w1, b1, w2, b2, w3, b3)
            HW = y3.shape[1] * y3.shape[2]
            flat = torch.empty(F_, 64 * HW, device=w1.device, dtype=torch.float32)
            call("hulc2_nhwc_bf16_to_nchw", y3.data_ptr(), flat.data_ptr(), F_, HW, 64)
            ctx.save_for_backward(xs, y1, y2, y3, w1, w2, w3, *_bits_saved(bits))
            return flat
        y1, y2, y3 = _conv_trunk_fwd(x, w1, b1, w2, b2, w3, b3)
        HW = y3.shape[1] * y3.shape[2]
        flat = torch.empty(F_, 64 * HW, device=x.device, dtype=torch.float32)
        call("hulc2_nhwc_to_nchw", y3.data_ptr(), flat.data_ptr(), F_, HW, 64)
        ctx.save_for_backward(x, y1, y2, y3, w2, w3)
        return flat

    @staticmethod
    def backward(ctx, dflat):
        dflat = dflat.contiguous()
        if ctx.bf16:
            xs, y1, y2, y3, w1, w2, w3, m1, m2 = ctx.saved_tensors
            F_ = xs.shape[0]
            HW = y3.shape[1] * y3.shape[2]
            dz3 = torch.empty_like(y3)
            call("hulc2_nchw_to_nhwc_bf16", dflat.data_ptr(), dz3.data_ptr(), F_, HW, 64, y3.data_ptr())
            g = _convb_trunk_bwd(xs, y1, y2, dz3, w1, w2, w3, ctx.biases, bits=(m1, m2))
            return (None, g["w1"], g["b1"], g["w2"], g["b2"], g["w3"], g["b3"])
        x, y1, y2, y3, w2, w3 = ctx.saved_tensors
        F_ = x.shape[0]
        HW = y3.shape[1] * y3.shape[2]
        dz3 = torch.empty_like(y3)
        call("hulc2_nchw_to_nhwc", dflat.data_ptr(), dz3.data_ptr(), F_, HW, 64, y3.data_ptr())
        g = _conv_trunk_bwd(x, y1, y2, dz3, w2, w3, (True, True, True))
        return (None, g["w1"], g["b1"], g["w2"], g["b2"], g["w3"], g["b3"])


# ----------------------------------------------------------------------------- LayerNorm
class LayerNormFunction(torch.autograd.Function):
    """y = LN(x + dropout(res)) * gamma + beta (eps 1e-5); res/keep optional.  Used for the encoder / goal
    LayerNorms and the post-LN residual blocks of nn.TransformerEncoderLayer."""

    @staticmethod
    def forward(ctx, x, res, keep, keep_scale, gamma, beta, eps):
        x2 = _rows2d(_f32(x))
        rows, D = x2.shape
        y = torch.empty(rows, D, device=x.device, dtype=torch.float32)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        r2 = _rows2d(res) if res is not None else None
        t = torch.empty(rows, D, device=x.device, dtype=torch.float32) if res is not None else None
        y16 = torch.empty(rows, _pad8(D), device=x.device, dtype=torch.bfloat16) if emit_mirrors() and rows > 0 else None
        call("hulc2_layernorm_fwd_m", x2.data_ptr(), _ld(x2), _p(r2), _ld(r2) if r2 is not None else 0, _p(keep), keep_scale,
             gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), D, _p(t), mean.data_ptr(), rstd.data_ptr(), rows, D, eps,
             _p(y16), _pad8(D))
        if y16 is not None:
            register_mirror(y, y16, _pad8(D), rows, D)
        ctx.save_for_backward(t if t is not None else x2, gamma, mean, rstd)
        ctx.beta = beta                        # identity only (grad_buffer)
        ctx.keep, ctx.keep_scale, ctx.has_res, ctx.shape = keep, keep_scale, res is not None, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        t, gamma, mean, rstd = ctx.saved_tensors
        beta = ctx.beta
        rows, D = t.shape[0], gamma.shape[0]
        dy2 = _rows2d(dy.contiguous())
        dx = torch.empty(rows, D, device=dy.device, dtype=torch.float32)
        dres = torch.empty(rows, D, device=dy.device, dtype=torch.float32) if ctx.has_res else None
        dgamma = grad_buffer(gamma, zero=True)
        dbeta = grad_buffer(beta, zero=True)
        # the gradient that goes straight into a contraction's backward: the residual branch's when there is one, else dx
        g16 = torch.empty(rows, _pad8(D), device=dy.device, dtype=torch.bfloat16) if emit_mirrors() and rows > 0 else None
        call("hulc2_layernorm_bwd_m", dy2.data_ptr(), D, t.data_ptr(), _ld(t), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
             dx.data_ptr(), D, _p(dres), _p(ctx.keep), ctx.keep_scale, dgamma.data_ptr(), dbeta.data_ptr(), rows, D,
             _p(g16) if dres is None else None, _p(g16) if dres is not None else None, _pad8(D))
        if g16 is not None:
            register_mirror(dres if dres is not None else dx, g16, _pad8(D), rows, D)
        return (dx.view(ctx.shape), dres.view(ctx.shape) if dres is not None else None, None, None, dgamma, dbeta, None)


def layer_norm(x, gamma, beta, res=None, keep=None, keep_scale=1.0, eps=1e-5):
    return LayerNormFunction.apply(x, res, keep, float(keep_scale), gamma, beta, float(eps))


# ----------------------------------------------------------------------------- transformer pieces
class AddPosFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, pos, keep, keep_scale):
        emb = _f32(emb).contiguous()
        B, S, E = emb.shape
        out = torch.empty_like(emb)
        out16 = torch.empty(B * S, E, device=emb.device, dtype=torch.bfloat16) if emit_mirrors() and E % 8 == 0 and B * S > 0 else None
        call("hulc2_add_pos_fwd_m", emb.data_ptr(), pos.data_ptr(), _p(keep), keep_scale, out.data_ptr(), _p(out16), B, S, E)
        if out16 is not None:
            register_mirror(out, out16, E, B * S, E)
        ctx.keep, ctx.keep_scale, ctx.pos_shape = keep, keep_scale, pos.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        B, S, E = dout.shape
        demb = torch.empty_like(dout)
        dpos = torch.zeros(ctx.pos_shape, device=dout.device, dtype=torch.float32)
        call("hulc2_add_pos_bwd", dout.data_ptr(), _p(ctx.keep), ctx.keep_scale, demb.data_ptr(), dpos.data_ptr(), B, S, E)
        return demb, dpos, None, None


class AttentionFunction(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) v per (window, head) with dropout on the probabilities; qkv [B*S, 3E]."""

    @staticmethod
    def forward(ctx, qkv, B, S, H, keep, keep_scale):
        qkv = _f32(qkv).contiguous()
        E = qkv.shape[-1] // 3
        out = torch.empty(B * S, E, device=qkv.device, dtype=torch.float32)
        probs = torch.empty(B, H, S, S, device=qkv.device, dtype=torch.float32)
        _lib.tag(f"attention_fwd[B={B},S={S},H={H},dh={E // H}]", 4.0 * B * H * S * S * (E // H),
                 4.0 * B * S * 4 * E + B * H * S * S * (4.0 + (1.0 if keep is not None else 0.0)))
        out16 = torch.empty(B * S, _pad8(E), device=qkv.device, dtype=torch.bfloat16) if emit_mirrors() and B * S > 0 else None
        call("hulc2_attention_fwd_m", qkv.data_ptr(), _p(keep), keep_scale, out.data_ptr(), probs.data_ptr(), _p(out16), _pad8(E),
             B, S, H, E // H)
        if out16 is not None:
            register_mirror(out, out16, _pad8(E), B * S, E)
        ctx.save_for_backward(qkv, probs)
        ctx.dims, ctx.keep, ctx.keep_scale = (B, S, H, E // H), keep, keep_scale
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, probs = ctx.saved_tensors
        B, S, H, Dh = ctx.dims
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        _lib.tag(f"attention_bwd[B={B},S={S},H={H},dh={Dh}]", 8.0 * B * H * S * S * Dh,
                 4.0 * B * S * 7 * H * Dh + B * H * S * S * (4.0 + (1.0 if ctx.keep is not None else 0.0)))
        E3 = 3 * H * Dh
        d16 = torch.empty(B * S, _pad8(E3), device=qkv.device, dtype=torch.bfloat16) if emit_mirrors() and B * S > 0 else None
        call("hulc2_attention_bwd_m", qkv.data_ptr(), probs.data_ptr(), _p(ctx.keep), ctx.keep_scale, dout.data_ptr(),
             dqkv.data_ptr(), _p(d16), _pad8(E3), B, S, H, Dh)
        if d16 is not None:
            register_mirror(dqkv, d16, _pad8(E3), B * S, E3)
        return dqkv, None, None, None, None, None


class MeanSeqFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32(x).contiguous()
        B, S, E = x.shape
        out = torch.empty(B, E, device=x.device, dtype=torch.float32)
        call("hulc2_mean_seq_fwd", x.data_ptr(), out.data_ptr(), B, S, E)
        ctx.dims = (B, S, E)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, S, E = ctx.dims
        dout = dout.contiguous()
        dx = torch.empty(B, S, E, device=dout.device, dtype=torch.float32)
        call("hulc2_mean_seq_bwd", dout.data_ptr(), dx.data_ptr(), B, S, E)
        return dx


# ----------------------------------------------------------------------------- latent plan
class KLFunction(torch.autograd.Function):
    """hulc2.py:444-466 for the discrete plan: balanced categorical KL, scaled by kl_beta -> scalar.
    With ``segments`` (row counts per modality, summing to B) the result is the vector of per-modality batch means."""

    @staticmethod
    def forward(ctx, pp, pr, cats, classes, alpha, beta, segments=None):
        pp, pr = _f32(pp).contiguous(), _f32(pr).contiguous()
        B = pp.shape[0]
        segs = tuple(int(n) for n in segments) if segments is not None else (B,)
        assert sum(segs) == B
        loss = torch.empty(len(segs), device=pp.device, dtype=torch.float32)
        row = pp.shape[1] * 4
        b0 = 0
        for i, n in enumerate(segs):
            _lib.tag(f"kl_fwd[B={n},{cats}x{classes}]", 0.0, 8.0 * n * cats * classes)          # 8 KB / window forward
            call("hulc2_kl_fwd", pp.data_ptr() + b0 * row, pr.data_ptr() + b0 * row, loss.data_ptr() + 4 * i, n, cats, classes, alpha, beta)
            b0 += n
        ctx.save_for_backward(pp, pr)
        ctx.cfg = (segs, cats, classes, alpha, beta)
        return loss if segments is not None else loss.view(())

    @staticmethod
    def backward(ctx, g):
        pp, pr = ctx.saved_tensors
        segs, cats, classes, alpha, beta = ctx.cfg
        g = g.contiguous().view(-1)
        dpp, dpr = torch.empty_like(pp), torch.empty_like(pr)
        row = pp.shape[1] * 4
        b0 = 0
        for i, n in enumerate(segs):
            _lib.tag(f"kl_bwd[B={n},{cats}x{classes}]", 0.0, 16.0 * n * cats * classes)
            call("hulc2_kl_bwd", pp.data_ptr() + b0 * row, pr.data_ptr() + b0 * row, g.data_ptr() + 4 * i, dpp.data_ptr() + b0 * row,
                 dpr.data_ptr() + b0 * row, n, cats, classes, alpha, beta)
            b0 += n
        return dpp, dpr, None, None, None, None, None


class GaussStateFunction(torch.autograd.Function):
    """distributions.py:55-59: x [B,2P] -> (mean, std = softplus(var) + 1e-4)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x).contiguous()
        B, P2 = x.shape
        P = P2 // 2
        mean = torch.empty(B, P, device=x.device, dtype=torch.float32)
        std = torch.empty(B, P, device=x.device, dtype=torch.float32)
        call("hulc2_gauss_state_fwd", x.data_ptr(), mean.data_ptr(), std.data_ptr(), B, P)
        ctx.save_for_backward(x)
        return mean, std

    @staticmethod
    def backward(ctx, dmean, dstd):
        (x,) = ctx.saved_tensors
        B, P2 = x.shape
        dx = torch.empty_like(x)
        call("hulc2_gauss_state_bwd", x.data_ptr(), _p(dmean.contiguous() if dmean is not None else None),
             _p(dstd.contiguous() if dstd is not None else None), dx.data_ptr(), B, P2 // 2)
        return dx


class GaussRSampleFunction(torch.autograd.Function):
    """Independent(Normal(mean, std), 1).rsample() with caller-visible noise: plan = mean + std * eps (hulc2.py:235)."""

    @staticmethod
    def forward(ctx, mean, std, eps):
        mean, std, eps = _f32(mean).contiguous(), _f32(std).contiguous(), _f32(eps).contiguous()
        plan = torch.empty_like(mean)
        call("hulc2_gauss_rsample", mean.data_ptr(), std.data_ptr(), eps.data_ptr(), plan.data_ptr(), mean.numel())
        ctx.save_for_backward(eps)
        return plan

    @staticmethod
    def backward(ctx, g):
        (eps,) = ctx.saved_tensors
        g = g.contiguous()
        dstd = torch.empty_like(g)
        call("hulc2_gauss_rsample", None, g.data_ptr(), eps.data_ptr(), dstd.data_ptr(), g.numel())
        return g, dstd, None


def box_muller(u1: torch.Tensor, u2: torch.Tensor) -> torch.Tensor:
    eps = torch.empty_like(u1)
    call("hulc2_box_muller", u1.contiguous().data_ptr(), u2.contiguous().data_ptr(), eps.data_ptr(), eps.numel())
    return eps


class GaussKLFunction(torch.autograd.Function):
    """hulc2.py:444-466 for the continuous plan: balanced KL between diagonal normals, scaled by kl_beta -> scalar
    (or the vector of per-modality batch means with ``segments``)."""

    @staticmethod
    def forward(ctx, pp_mean, pp_std, pr_mean, pr_std, alpha, beta, segments=None):
        ts = [_f32(t).contiguous() for t in (pp_mean, pp_std, pr_mean, pr_std)]
        B, P = ts[0].shape
        segs = tuple(int(n) for n in segments) if segments is not None else (B,)
        assert sum(segs) == B
        loss = torch.empty(len(segs), device=ts[0].device, dtype=torch.float32)
        b0 = 0
        for i, n in enumerate(segs):
            off = 4 * b0 * P
            call("hulc2_gauss_kl_fwd", *(t.data_ptr() + off for t in ts), loss.data_ptr() + 4 * i, n, P, alpha, beta)
            b0 += n
        ctx.save_for_backward(*ts)
        ctx.cfg = (segs, alpha, beta)
        return loss if segments is not None else loss.view(())

    @staticmethod
    def backward(ctx, g):
        ts = ctx.saved_tensors
        segs, alpha, beta = ctx.cfg
        B, P = ts[0].shape
        g = g.contiguous().view(-1)
        ds = [torch.empty_like(t) for t in ts]
        b0 = 0
        for i, n in enumerate(segs):
            off = 4 * b0 * P
            call("hulc2_gauss_kl_bwd", *(t.data_ptr() + off for t in ts), g.data_ptr() + 4 * i, *(d.data_ptr() + off for d in ds),
                 n, P, alpha, beta)
            b0 += n
        return (*ds, None, None, None)


class PlanRSampleFunction(torch.autograd.Function):
    """OneHotCategoricalStraightThrough.rsample given the drawn indices (hulc2.py:235, distributions.py:23-26)."""

    @staticmethod
    def forward(ctx, logits, idx, cats, classes):
        logits = _f32(logits).contiguous()
        B = logits.shape[0]
        plan = torch.empty(B, cats * classes, device=logits.device, dtype=torch.float32)
        call("hulc2_onehot_fwd", idx.contiguous().data_ptr(), plan.data_ptr(), B, cats, classes)
        ctx.save_for_backward(logits)
        ctx.cfg = (B, cats, classes)
        return plan

    @staticmethod
    def backward(ctx, dplan):
        (logits,) = ctx.saved_tensors
        B, cats, classes = ctx.cfg
        dplan = dplan.contiguous()
        dl = torch.empty_like(logits)
        call("hulc2_st_onehot_bwd", logits.data_ptr(), dplan.data_ptr(), dl.data_ptr(), B, cats, classes)
        return dl, None, None, None


def onehot(idx: torch.Tensor, cats: int, classes: int) -> torch.Tensor:
    B = idx.shape[0]
    plan = torch.empty(B, cats * classes, device=idx.device, dtype=torch.float32)
    call("hulc2_onehot_fwd", idx.contiguous().data_ptr(), plan.data_ptr(), B, cats, classes)
    return plan


def categorical_sample(logits: torch.Tensor, u: torch.Tensor, cats: int, classes: int) -> torch.Tensor:
    B = logits.shape[0]
    idx = torch.empty(B, cats, device=logits.device, dtype=torch.int64)
    call("hulc2_categorical_sample", logits.contiguous().data_ptr(), u.contiguous().data_ptr(), idx.data_ptr(), B, cats, classes)
    return idx


def uniform(shape, device, seed: int, offset: int = 0, epoch: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty(shape, device=device, dtype=torch.float32)
    call("hulc2_philox_uniform_ep", out.data_ptr(), out.numel(), seed, offset, _p(epoch))
    return out


def dropout_mask(shape, p: float, device, seed: int, offset: int = 0, epoch: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = torch.empty(shape, device=device, dtype=torch.uint8)
    call("hulc2_dropout_mask_ep", out.data_ptr(), out.numel(), p, seed, offset, _p(epoch))
    return out


# ----------------------------------------------------------------------------- decoder recurrence
class RNNDecoderFunction(torch.autograd.Function):
    """2-layer Elman ReLU RNN over x_t = [plan | emb_t | goal] (logistic_decoder_rnn.py:262-268, rnn.py:5-14).
    plan and goal are constant over time, so W_ih x_t splits into a per-window base term and a per-step
    embedding term (same algebra, 32x fewer FLOPs than materialising [B,S,1120]).  Returns time-major
    hidden states of the last layer [S,B,H] and h_n [2,B,H]."""

    @staticmethod
    def forward(ctx, plan, emb, goal, h0, wi0, wh0, bi0, bh0, wi1, wh1, bi1, bh1):
        plan, goal = _f32(plan).contiguous(), _f32(goal).contiguous()
        B, S, Es = emb.shape
        P, G = plan.shape[1], goal.shape[1]
        H = wh0.shape[0]
        In = wi0.shape[1]
        assert In == P + Es + G
        dev = plan.device
        ws = workspace(dev)
        embT = torch.empty(S, B, Es, device=dev, dtype=torch.float32)
        assert emb.stride(2) == 1
        call("hulc2_transpose01", emb.data_ptr(), emb.stride(0), emb.stride(1), embT.data_ptr(), Es, B, S, Es, 0)
        bsum0 = torch.empty(H, device=dev, dtype=torch.float32)
        bsum1 = torch.empty(H, device=dev, dtype=torch.float32)
        for bs, bi, bh in ((bsum0, bi0, bh0), (bsum1, bi1, bh1)):
            call("hulc2_copy2d", bi.data_ptr(), H, bs.data_ptr(), H, 1, H, 0)
            call("hulc2_axpy", bh.data_ptr(), bs.data_ptr(), H, 1.0)
        b16 = _precision == 1 and In % 8 == 0 and P % 8 == 0 and Es % 8 == 0 and H % 8 == 0
        base = torch.empty(B, H, device=dev, dtype=torch.float32)
        pre = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        if b16:
            wi0h, wi1h = weight16(wi0), weight16(wi1)
            plan16, _ = mirror2d(plan)
            goal16, ldg = mirror2d(goal)
            embT16, _ = mirror2d(embT.view(S * B, Es))
            gemm16(B, H, P, plan16, P, 1, wi0h, In, 1, base, H, bias=bsum0)
            gemm16(B, H, G, goal16, ldg, 1, wi0h, In, 1, base, H, b_off=P + Es, accumulate=True)
            call("hulc2_copy2d", base.data_ptr(), 0, pre.data_ptr(), B * H, S, B * H, 0)
            gemm16(S * B, H, Es, embT16, Es, 1, wi0h, In, 1, pre, H, b_off=P, accumulate=True)
        else:
            gemm(B, H, P, plan, P, 1, wi0, In, 1, base, H, bias=bsum0)
            gemm(B, H, G, goal, G, 1, wi0, In, 1, base, H, b_off=P + Es, accumulate=True)
            call("hulc2_copy2d", base.data_ptr(), 0, pre.data_ptr(), B * H, S, B * H, 0)
            gemm(S * B, H, Es, embT, Es, 1, wi0, In, 1, pre, H, b_off=P, accumulate=True)
        H0 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        h00 = h0[0].contiguous() if h0 is not None else None
        h01 = h0[1].contiguous() if h0 is not None else None
        _lib.tag(f"rnn_relu_fwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 12.0 * B * H))
        # bf16 mirrors of the states come out of the recurrence kernel itself ([S+1,B,H]: slot t + 1 = state t is the kernel's own
        # step-to-step operand store): no fp32 -> bf16 pass over 32 MB per layer and direction
        S16 = (lambda: torch.empty(S + 1, B, H, device=dev, dtype=torch.bfloat16)) if b16 and emit_mirrors() else (lambda: None)
        H0s = S16()
        call("hulc2_rnn_relu_fwd_m", pre.data_ptr(), wh0.data_ptr(), _p(h00), H0.data_ptr(), _p(H0s), S, B, H, _lib_precision(),
             ws.data_ptr(), ws.numel())
        H0h = None
        if b16:
            H0h = H0s[1:] if H0s is not None else to_bf16(H0)
            gemm16(S * B, H, H, H0h, H, 1, wi1h, H, 1, pre, H, bias=bsum1)
        else:
            gemm(S * B, H, H, H0, H, 1, wi1, H, 1, pre, H, bias=bsum1)
        H1 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        H1s = S16()
        _lib.tag(f"rnn_relu_fwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 12.0 * B * H))
        call("hulc2_rnn_relu_fwd_m", pre.data_ptr(), wh1.data_ptr(), _p(h01), H1.data_ptr(), _p(H1s), S, B, H, _lib_precision(),
             ws.data_ptr(), ws.numel())
        H1h = H1s[1:] if H1s is not None else None
        if H1h is not None:
            register_mirror(H1, H1h, H, S * B, H)          # the decoder heads' operand (heads_forward / DecoderLossFunction)
        hn = torch.empty(2, B, H, device=dev, dtype=torch.float32)
        call("hulc2_copy2d", H0.data_ptr() + 4 * (S - 1) * B * H, B * H, hn.data_ptr(), B * H, 1, B * H, 0)
        call("hulc2_copy2d", H1.data_ptr() + 4 * (S - 1) * B * H, B * H, hn.data_ptr() + 4 * B * H, B * H, 1, B * H, 0)
        ctx.b16 = b16
        ctx.bias_params = (bi0, bh0, bi1, bh1)     # identity only: their gradients land in the arena slices (grad_buffer)
        extra = (plan16, goal16, embT16, H0h) + ((H1h,) if H1h is not None else ()) if b16 else ()
        ctx.save_for_backward(plan, embT, goal, H0, H1, wi0, wh0, wi1, wh1, h00 if h00 is not None else plan.new_empty(0),
                              h01 if h01 is not None else plan.new_empty(0), *extra)
        ctx.dims = (B, S, Es, P, G, H, In)
        ctx.has_h0 = h0 is not None
        ctx.mark_non_differentiable(hn)
        return H1, hn

    @staticmethod
    def backward(ctx, dH1, _dhn):
        if ctx.b16:
            return RNNDecoderFunction._backward16(ctx, dH1)
        plan, embT, goal, H0, H1, wi0, wh0, wi1, wh1, h00, h01 = ctx.saved_tensors
        B, S, Es, P, G, H, In = ctx.dims
        dev = plan.device
        ws = workspace(dev)
        prec = _lib_precision()
        step = B * H
        dH1 = dH1.contiguous()
        dz1 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        call("hulc2_copy2d", dH1.data_ptr(), step, dz1.data_ptr(), step, S, step, 0)
        _lib.tag(f"rnn_relu_bwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 16.0 * B * H))
        call("hulc2_rnn_relu_bwd", dz1.data_ptr(), wh1.data_ptr(), H1.data_ptr(), None, S, B, H, prec, ws.data_ptr(), ws.numel())
        dwh1 = grad_buffer(wh1)
        if S > 1:
            gemm(H, H, (S - 1) * B, dz1, 1, H, H1, 1, H, dwh1, H, a_off=step)
        else:
            call("hulc2_fill", dwh1.data_ptr(), dwh1.numel(), 0.0)
        if ctx.has_h0:
            gemm(H, H, B, dz1, 1, H, h01, 1, H, dwh1, H, accumulate=True)
        dwi1 = grad_buffer(wi1)
        gemm(H, H, S * B, dz1, 1, H, H0, 1, H, dwi1, H)
        db1 = torch.empty(H, device=dev, dtype=torch.float32)
        colsum(dz1, H, S * B, H, db1)
        dz0 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        gemm(S * B, H, H, dz1, H, 1, wi1, 1, H, dz0, H)                      # dH0 = dz1 W_ih1
        _lib.tag(f"rnn_relu_bwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 16.0 * B * H))
        call("hulc2_rnn_relu_bwd", dz0.data_ptr(), wh0.data_ptr(), H0.data_ptr(), None, S, B, H, prec, ws.data_ptr(), ws.numel())
        dwh0 = grad_buffer(wh0)
        if S > 1:
            gemm(H, H, (S - 1) * B, dz0, 1, H, H0, 1, H, dwh0, H, a_off=step)
        else:
            call("hulc2_fill", dwh0.data_ptr(), dwh0.numel(), 0.0)
        if ctx.has_h0:
            gemm(H, H, B, dz0, 1, H, h00, 1, H, dwh0, H, accumulate=True)
        db0 = torch.empty(H, device=dev, dtype=torch.float32)
        colsum(dz0, H, S * B, H, db0)
        dzsum = torch.empty(B, H, device=dev, dtype=torch.float32)
        colsum(dz0, B * H, S, B * H, dzsum)                                  # sum over time
        dwi0 = grad_buffer(wi0)
        gemm(H, P, B, dzsum, 1, H, plan, 1, P, dwi0, In)
        gemm(H, Es, S * B, dz0, 1, H, embT, 1, Es, dwi0, In, c_off=P)
        gemm(H, G, B, dzsum, 1, H, goal, 1, G, dwi0, In, c_off=P + Es)
        dplan = dgoal = demb = None
        if ctx.needs_input_grad[0]:
            dplan = torch.empty(B, P, device=dev, dtype=torch.float32)
            gemm(B, P, H, dzsum, H, 1, wi0, 1, In, dplan, P)
        if ctx.needs_input_grad[2]:
            dgoal = torch.empty(B, G, device=dev, dtype=torch.float32)
            gemm(B, G, H, dzsum, H, 1, wi0, 1, In, dgoal, G, b_off=P + Es)
        if ctx.needs_input_grad[1]:
            dembT = torch.empty(S, B, Es, device=dev, dtype=torch.float32)
            gemm(S * B, Es, H, dz0, H, 1, wi0, 1, In, dembT, Es, b_off=P)
            demb = torch.empty(B, S, Es, device=dev, dtype=torch.float32)
            call("hulc2_transpose01", dembT.data_ptr(), B * Es, Es, demb.data_ptr(), Es, S, B, Es, 0)
        return (dplan, demb, dgoal, None, dwi0, dwh0, db0, db0.clone(), dwi1, dwh1, db1, db1.clone())

    @staticmethod
    def _backward16(ctx, dH1):
        """Same gradient algebra as ``backward`` on bf16 operand mirrors (TMA-fed contractions)."""
        plan, embT, goal, H0, H1, wi0, wh0, wi1, wh1, h00, h01, plan16, goal16, embT16, H0h, *rest = ctx.saved_tensors
        H1h_saved = rest[0] if rest else None
        B, S, Es, P, G, H, In = ctx.dims
        dev = plan.device
        ws = workspace(dev)
        step = B * H
        ldg = goal16.shape[1]
        wi0h, wi1h = weight16(wi0), weight16(wi1)
        dH1 = dH1.contiguous()
        dz1 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        call("hulc2_copy2d", dH1.data_ptr(), step, dz1.data_ptr(), step, S, step, 0)
        _lib.tag(f"rnn_relu_bwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 16.0 * B * H))
        S16 = (lambda: torch.empty(S + 1, B, H, device=dev, dtype=torch.bfloat16)) if emit_mirrors() else (lambda: None)
        dz1s = S16()
        call("hulc2_rnn_relu_bwd_m", dz1.data_ptr(), wh1.data_ptr(), H1.data_ptr(), None, _p(dz1s), S, B, H, 1, ws.data_ptr(), ws.numel())
        dz1h = dz1s[1:] if dz1s is not None else to_bf16(dz1)
        H1h = H1h_saved if H1h_saved is not None else torch.empty(H1.shape, device=dev, dtype=torch.bfloat16)
        dwh1, dwi1 = grad_buffer(wh1), grad_buffer(wi1)
        bi0, bh0, bi1, bh1 = ctx.bias_params
        db1, db1h = grad_buffer(bi1), grad_buffer(bh1)                           # b_hh receives the same gradient as b_ih

        def layer1_wgrads():
            if H1h_saved is None:
                call("hulc2_f32_to_bf16", H1.data_ptr(), H1h.data_ptr(), H1.numel())
            if S > 1:
                gemm16(H, H, (S - 1) * B, dz1h, 1, H, H1h, 1, H, dwh1, H, a_off=step)
            else:
                call("hulc2_fill", dwh1.data_ptr(), dwh1.numel(), 0.0)
            if ctx.has_h0:
                gemm(H, H, B, dz1, 1, H, h01, 1, H, dwh1, H, accumulate=True)
            gemm16(H, H, S * B, dz1h, 1, H, H0h, 1, H, dwi1, H, rowsum=db1)     # + bias gradient = row sums of dz1^T
            call("hulc2_copy2d", db1.data_ptr(), H, db1h.data_ptr(), H, 1, H, 0)

        # the four weight-gradient contractions of a layer (2 x 34 GFLOP) run beside the next layer's recurrence, which leaves
        # the tensor pipes 94 % idle
        wgrad_section(layer1_wgrads, (dz1, dz1h, H1, H1h, H0h, h01), (dwh1, dwi1, db1, db1h))
        dz0 = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        gemm16(S * B, H, H, dz1h, H, 1, wi1h, 1, H, dz0, H)                      # dH0 = dz1 W_ih1
        _lib.tag(f"rnn_relu_bwd[S={S},B={B},H={H}]", 2.0 * S * B * H * H, S * (2.0 * H * H + 16.0 * B * H))
        dz0s = S16()
        call("hulc2_rnn_relu_bwd_m", dz0.data_ptr(), wh0.data_ptr(), H0.data_ptr(), None, _p(dz0s), S, B, H, 1, ws.data_ptr(), ws.numel())
        dz0h = dz0s[1:] if dz0s is not None else to_bf16(dz0)
        dzsum = torch.empty(B, H, device=dev, dtype=torch.float32)
        colsum(dz0, B * H, S, B * H, dzsum)                                      # sum over time
        dzsumh = to_bf16(dzsum)
        dwh0, dwi0 = grad_buffer(wh0), grad_buffer(wi0)
        db0, db0h = grad_buffer(bi0), grad_buffer(bh0)

        def layer0_wgrads():
            if S > 1:
                gemm16(H, H, (S - 1) * B, dz0h, 1, H, H0h, 1, H, dwh0, H, a_off=step)
            else:
                call("hulc2_fill", dwh0.data_ptr(), dwh0.numel(), 0.0)
            if ctx.has_h0:
                gemm(H, H, B, dz0, 1, H, h00, 1, H, dwh0, H, accumulate=True)
            gemm16(H, P, B, dzsumh, 1, H, plan16, 1, P, dwi0, In)
            gemm16(H, Es, S * B, dz0h, 1, H, embT16, 1, Es, dwi0, In, c_off=P, rowsum=db0)   # + bias gradient = row sums of dz0^T
            gemm16(H, G, B, dzsumh, 1, H, goal16, 1, ldg, dwi0, In, c_off=P + Es)
            call("hulc2_copy2d", db0.data_ptr(), H, db0h.data_ptr(), H, 1, H, 0)

        wgrad_section(layer0_wgrads, (dz0, dz0h, dzsumh, H0h, h00, plan16, embT16, goal16), (dwh0, dwi0, db0, db0h))
        dplan = dgoal = demb = None
        if ctx.needs_input_grad[0]:
            dplan = torch.empty(B, P, device=dev, dtype=torch.float32)
            gemm16(B, P, H, dzsumh, H, 1, wi0h, 1, In, dplan, P)
        if ctx.needs_input_grad[2]:
            dgoal = torch.empty(B, G, device=dev, dtype=torch.float32)
            gemm16(B, G, H, dzsumh, H, 1, wi0h, 1, In, dgoal, G, b_off=P + Es)
        if ctx.needs_input_grad[1]:
            dembT = torch.empty(S, B, Es, device=dev, dtype=torch.float32)
            gemm16(S * B, Es, H, dz0h, H, 1, wi0h, 1, In, dembT, Es, b_off=P)
            demb = torch.empty(B, S, Es, device=dev, dtype=torch.float32)
            call("hulc2_transpose01", dembT.data_ptr(), B * Es, Es, demb.data_ptr(), Es, S, B, Es, 0)
        return (dplan, demb, dgoal, None, dwi0, dwh0, db0, db0h, dwi1, dwh1, db1, db1h)


class GatedRNNDecoderFunction(torch.autograd.Function):
    """2-layer nn.GRU / nn.LSTM decoder (decoders/utils/rnn.py:17-36) over x_t = [plan | emb_t | goal]
    (logistic_decoder_rnn.py:262-268).  Same split of W_ih x_t into a per-window base term and a per-step embedding
    term as the ReLU RNN.  Per step: one recurrent contraction h_{t-1} W_hh^T (+ b_hh) and one fused cell kernel
    (csrc/cells.cu).  Returns time-major hidden states of the last layer [S,B,H], h_n [2,B,H] and c_n [2,B,H]
    (LSTM; empty for GRU)."""

    @staticmethod
    def forward(ctx, kind, plan, emb, goal, h0, c0, wi0, wh0, bi0, bh0, wi1, wh1, bi1, bh1):
        assert kind in ("gru", "lstm")
        Gn = 3 if kind == "gru" else 4
        plan, goal = _f32(plan).contiguous(), _f32(goal).contiguous()
        B, S, Es = emb.shape
        P, G = plan.shape[1], goal.shape[1]
        H = wh0.shape[1]
        GH = Gn * H
        In = wi0.shape[1]
        assert In == P + Es + G and wi0.shape[0] == GH and wh0.shape[0] == GH
        dev = plan.device
        embT = torch.empty(S, B, Es, device=dev, dtype=torch.float32)
        assert emb.stride(2) == 1
        call("hulc2_transpose01", emb.data_ptr(), emb.stride(0), emb.stride(1), embT.data_ptr(), Es, B, S, Es, 0)
        # layer-0 input gates: per-window base (plan, goal, b_ih) broadcast over time + per-step embedding term
        base = torch.empty(B, GH, device=dev, dtype=torch.float32)
        gi = torch.empty(S, B, GH, device=dev, dtype=torch.float32)
        gemm(B, GH, P, plan, P, 1, wi0, In, 1, base, GH, bias=bi0)
        gemm(B, GH, G, goal, G, 1, wi0, In, 1, base, GH, b_off=P + Es, accumulate=True)
        call("hulc2_copy2d", base.data_ptr(), 0, gi.data_ptr(), B * GH, S, B * GH, 0)
        gemm(S * B, GH, Es, embT, Es, 1, wi0, In, 1, gi, GH, b_off=P, accumulate=True)
        gh = torch.empty(B, GH, device=dev, dtype=torch.float32)
        Hs, Cs, saves = [], [], []
        hn = torch.empty(2, B, H, device=dev, dtype=torch.float32)
        cn = torch.empty(2 if kind == "lstm" else 0, B, H, device=dev, dtype=torch.float32)
        h0s = [h0[l].contiguous() if h0 is not None else None for l in range(2)]
        c0s = [c0[l].contiguous() if c0 is not None else None for l in range(2)]
        for l, (wi, wh, bi, bh) in enumerate(((wi0, wh0, bi0, bh0), (wi1, wh1, bi1, bh1))):
            if l == 1:
                gemm(S * B, GH, H, Hs[0], H, 1, wi, H, 1, gi, GH, bias=bi)
            Hl = torch.empty(S, B, H, device=dev, dtype=torch.float32)
            Cl = torch.empty(S, B, H, device=dev, dtype=torch.float32) if kind == "lstm" else None
            sv = torch.empty(S, B, 4 * H, device=dev, dtype=torch.float32)
            for t in range(S):
                hp = Hl[t - 1] if t > 0 else h0s[l]
                if hp is None:
                    call("hulc2_copy2d", bh.data_ptr(), 0, gh.data_ptr(), GH, B, GH, 0)
                else:
                    gemm(B, GH, H, hp, H, 1, wh, H, 1, gh, GH, bias=bh)
                _lib.tag(f"{kind}_cell_fwd[B={B},H={H}]", 0.0, (48.0 if kind == "gru" else 60.0) * B * H)
                if kind == "gru":
                    call("hulc2_gru_cell_fwd", gi[t].data_ptr(), GH, gh.data_ptr(), _p(hp), Hl[t].data_ptr(), sv[t].data_ptr(), B, H)
                else:
                    cp = Cl[t - 1] if t > 0 else c0s[l]
                    call("hulc2_lstm_cell_fwd", gi[t].data_ptr(), GH, gh.data_ptr(), _p(cp), Hl[t].data_ptr(), Cl[t].data_ptr(),
                         sv[t].data_ptr(), B, H)
            Hs.append(Hl)
            Cs.append(Cl)
            saves.append(sv)
            call("hulc2_copy2d", Hl[S - 1].data_ptr(), B * H, hn[l].data_ptr(), B * H, 1, B * H, 0)
            if kind == "lstm":
                call("hulc2_copy2d", Cl[S - 1].data_ptr(), B * H, cn[l].data_ptr(), B * H, 1, B * H, 0)
        empty = plan.new_empty(0)
        ctx.save_for_backward(plan, embT, goal, wi0, wh0, wi1, wh1, Hs[0], Hs[1], saves[0], saves[1],
                              *(Cs if kind == "lstm" else (empty, empty)),
                              *(h if h is not None else empty for h in h0s), *(c if c is not None else empty for c in c0s))
        ctx.kind, ctx.dims = kind, (B, S, Es, P, G, H, In)
        ctx.has_h0, ctx.has_c0 = h0 is not None, c0 is not None
        ctx.mark_non_differentiable(hn, cn)
        return Hs[1], hn, cn

    @staticmethod
    def backward(ctx, dH1, _dhn, _dcn):
        (plan, embT, goal, wi0, wh0, wi1, wh1, H0, H1, sv0, sv1, C0, C1, h00, h01, c00, c01) = ctx.saved_tensors
        kind = ctx.kind
        B, S, Es, P, G, H, In = ctx.dims
        Gn = 3 if kind == "gru" else 4
        GH = Gn * H
        dev = plan.device
        f32 = dict(device=dev, dtype=torch.float32)

        def layer_bwd(dHs, Hl, Cl, sv, wh, h0l, c0l):
            """-> dgi [S,B,GH] (gradient of the input gates), dW_hh, db_hh."""
            dgi = torch.empty(S, B, GH, **f32)
            dgh = torch.empty(S, B, GH, **f32) if kind == "gru" else dgi
            rec = [torch.empty(B, H, **f32), torch.empty(B, H, **f32)]
            dc = torch.empty(B, H, **f32) if kind == "lstm" else None
            for t in range(S - 1, -1, -1):
                hp = Hl[t - 1] if t > 0 else h0l
                dhb = rec[(t + 1) & 1] if t < S - 1 else None
                _lib.tag(f"{kind}_cell_bwd[B={B},H={H}]", 0.0, (52.0 if kind == "gru" else 56.0) * B * H)
                if kind == "gru":
                    call("hulc2_gru_cell_bwd", dHs[t].data_ptr(), _p(dhb), sv[t].data_ptr(), _p(hp), dgi[t].data_ptr(), GH,
                         dgh[t].data_ptr(), rec[t & 1].data_ptr(), B, H)
                else:
                    cp = Cl[t - 1] if t > 0 else c0l
                    call("hulc2_lstm_cell_bwd", dHs[t].data_ptr(), _p(dhb), dc.data_ptr() if t < S - 1 else None, sv[t].data_ptr(),
                         Cl[t].data_ptr(), _p(cp), dgi[t].data_ptr(), GH, dc.data_ptr(), B, H)
                if t > 0:   # gradient into h_{t-1} through the recurrent contraction (h_0 itself needs no gradient)
                    gemm(B, H, GH, dgh[t], GH, 1, wh, 1, H, rec[t & 1], H, accumulate=(kind == "gru"))
            dwh = torch.empty_like(wh)
            if S > 1:
                gemm(GH, H, (S - 1) * B, dgh, 1, GH, Hl, 1, H, dwh, H, a_off=B * GH)
            else:
                call("hulc2_fill", dwh.data_ptr(), dwh.numel(), 0.0)
            if h0l is not None:
                gemm(GH, H, B, dgh, 1, GH, h0l, 1, H, dwh, H, accumulate=True)
            dbh = torch.empty(GH, **f32)
            colsum(dgh, GH, S * B, GH, dbh)
            return dgi, dwh, dbh

        dH1 = dH1.contiguous()
        dg1, dwh1, dbh1 = layer_bwd(dH1, H1, C1 if kind == "lstm" else None, sv1, wh1,
                                    h01 if ctx.has_h0 else None, c01 if ctx.has_c0 else None)
        dwi1 = grad_buffer(wi1)
        gemm(GH, H, S * B, dg1, 1, GH, H0, 1, H, dwi1, H)
        if kind == "gru":
            dbi1 = torch.empty(GH, **f32)
            colsum(dg1, GH, S * B, GH, dbi1)
        else:
            dbi1 = dbh1.clone()
        dH0 = torch.empty(S, B, H, **f32)
        gemm(S * B, H, GH, dg1, GH, 1, wi1, 1, H, dH0, H)
        dg0, dwh0, dbh0 = layer_bwd(dH0, H0, C0 if kind == "lstm" else None, sv0, wh0,
                                    h00 if ctx.has_h0 else None, c00 if ctx.has_c0 else None)
        if kind == "gru":
            dbi0 = torch.empty(GH, **f32)
            colsum(dg0, GH, S * B, GH, dbi0)
        else:
            dbi0 = dbh0.clone()
        dgsum = torch.empty(B, GH, **f32)
        colsum(dg0, B * GH, S, B * GH, dgsum)                                   # sum over time
        dwi0 = grad_buffer(wi0)
        gemm(GH, P, B, dgsum, 1, GH, plan, 1, P, dwi0, In)
        gemm(GH, Es, S * B, dg0, 1, GH, embT, 1, Es, dwi0, In, c_off=P)
        gemm(GH, G, B, dgsum, 1, GH, goal, 1, G, dwi0, In, c_off=P + Es)
        dplan = dgoal = demb = None
        if ctx.needs_input_grad[1]:
            dplan = torch.empty(B, P, **f32)
            gemm(B, P, GH, dgsum, GH, 1, wi0, 1, In, dplan, P)
        if ctx.needs_input_grad[3]:
            dgoal = torch.empty(B, G, **f32)
            gemm(B, G, GH, dgsum, GH, 1, wi0, 1, In, dgoal, G, b_off=P + Es)
        if ctx.needs_input_grad[2]:
            dembT = torch.empty(S, B, Es, **f32)
            gemm(S * B, Es, GH, dg0, GH, 1, wi0, 1, In, dembT, Es, b_off=P)
            demb = torch.empty(B, S, Es, **f32)
            call("hulc2_transpose01", dembT.data_ptr(), B * Es, Es, demb.data_ptr(), Es, S, B, Es, 0)
        return (None, dplan, demb, dgoal, None, None, dwi0, dwh0, dbi0, dbh0, dwi1, dwh1, dbi1, dbh1)


def _lib_precision() -> int:
    return _precision


# ----------------------------------------------------------------------------- heads + logistic-mixture loss
HEAD_LD = 184  # 3*A*M + 2 = 182 columns, row-padded to 16 bytes


def _heads16_ok(Hs, wp, wg) -> bool:
    return _precision == 1 and wg is not None and Hs.shape[-1] % 8 == 0 and 3 * wp.shape[0] + 2 <= HEAD_LD and Hs.numel() > 0


def heads_forward(Hs, wp, bp, wm, bm, wsc, bsc, wg, bg, Hs16=None) -> torch.Tensor:
    """Four nn.Linear heads (logistic_decoder_rnn.py:269-274) into one [rows, 184] buffer
    [logit_probs | means | log_scales | gripper]."""
    rows, H = Hs.shape[0] * Hs.shape[1], Hs.shape[2]
    AM = wp.shape[0]
    heads = torch.empty(rows, HEAD_LD, device=Hs.device, dtype=torch.float32)
    if _heads16_ok(Hs, wp, wg):
        # one contraction over the stacked head weights [3*A*M + 2, H] (bf16 mirrors)
        Wcat = torch.cat([weight16(wp), weight16(wm), weight16(wsc), weight16(wg)], 0)
        bcat = torch.cat([bp, bm, bsc, bg], 0)
        Hs16 = Hs16 if Hs16 is not None else mirror_like(Hs.contiguous())
        gemm16(rows, 3 * AM + 2, H, Hs16, H, 1, Wcat, H, 1, heads, HEAD_LD, bias=bcat)
        return heads
    for W, b, off in ((wp, bp, 0), (wm, bm, AM), (wsc, bsc, 2 * AM)):
        gemm(rows, AM, H, Hs, H, 1, W, H, 1, heads, HEAD_LD, c_off=off, bias=b)
    if wg is not None:
        gemm(rows, 2, H, Hs, H, 1, wg, H, 1, heads, HEAD_LD, c_off=3 * AM, bias=bg)
    return heads


class DecoderLossFunction(torch.autograd.Function):
    """heads + discretised logistic-mixture NLL + gripper cross-entropy (logistic_decoder_rnn.py:133-152,181-228).
    Hs is time-major [S,B,H]; actions [B,S,A+1] (already in the tcp frame) -> scalar loss.
    ``actions`` may also be a tuple of per-modality tensors [B_i,S,A+1] with sum(B_i) = B (windows of several
    modalities decoded by one recurrence call): the result is then the vector of per-modality mean losses [n]."""

    @staticmethod
    def forward(ctx, Hs, actions, amin, amax, cfg, wp, bp, wm, bm, wsc, bsc, wg, bg):
        Hs = _f32(Hs).contiguous()
        S, B, H = Hs.shape
        A, M, num_classes, ls_min, alpha = cfg
        ctx.b16 = _heads16_ok(Hs, wp, wg)
        Hs16 = mirror_like(Hs) if ctx.b16 else None
        heads = heads_forward(Hs, wp, bp, wm, bm, wsc, bsc, wg, bg, Hs16=Hs16)
        ws = workspace(Hs.device)
        segmented = isinstance(actions, (tuple, list))
        segs = [_f32(a).contiguous() for a in (actions if segmented else (actions,))]
        assert sum(a.shape[0] for a in segs) == B, "per-modality actions must cover the decoded batch"
        out = torch.empty(len(segs), 3, device=Hs.device, dtype=torch.float32)
        b0 = 0
        for i, a in enumerate(segs):
            _lib.tag(f"logistic_loss_fwd[B={a.shape[0]},S={S}]", 0.0, 124.0 * a.shape[0] * S * A)  # 124 B per (b,s,dim) row
            call("hulc2_logistic_loss_seg_fwd", heads.data_ptr() + 4 * b0 * HEAD_LD, HEAD_LD, a.data_ptr(), amin.data_ptr(),
                 amax.data_ptr(), out.data_ptr() + 12 * i, a.shape[0], S, A, M, num_classes, ls_min, alpha, 1, B, ws.data_ptr(), ws.numel())
            b0 += a.shape[0]
        ctx.save_for_backward(Hs if Hs16 is None else Hs16, heads, amin, amax, wp, wm, wsc, wg, *segs)
        ctx.cfg = cfg
        return out[:, 0] if segmented else out[0, 0]

    @staticmethod
    def backward(ctx, g):
        Hs, heads, amin, amax, wp, wm, wsc, wg, *segs = ctx.saved_tensors
        S, B, H = Hs.shape
        A, M, num_classes, ls_min, alpha = ctx.cfg
        AM = A * M
        rows = S * B
        dev = Hs.device
        g = g.contiguous().view(-1)
        dheads = torch.empty_like(heads)
        b0 = 0
        for i, a in enumerate(segs):
            _lib.tag(f"logistic_loss_bwd[B={a.shape[0]},S={S}]", 0.0, 244.0 * a.shape[0] * S * A)
            call("hulc2_logistic_loss_seg_bwd", heads.data_ptr() + 4 * b0 * HEAD_LD, HEAD_LD, a.data_ptr(), amin.data_ptr(),
                 amax.data_ptr(), g.data_ptr() + 4 * i, dheads.data_ptr() + 4 * b0 * HEAD_LD, a.shape[0], S, A, M, num_classes,
                 ls_min, alpha, 1, B)
            b0 += a.shape[0]
        dbias = torch.empty(3 * AM + 2, device=dev, dtype=torch.float32)
        dH = torch.empty(S, B, H, device=dev, dtype=torch.float32)
        if not ctx.b16:
            colsum(dheads, HEAD_LD, rows, 3 * AM + 2, dbias)
        if ctx.b16:
            NH = 3 * AM + 2
            Hs16 = Hs                                                     # the bf16 mirror was saved instead of the fp32 states
            dh16, ldh = mirror2d(dheads[:, :NH])
            Wcat = torch.cat([weight16(wp), weight16(wm), weight16(wsc), weight16(wg)], 0)
            gemm16(rows, H, NH, dh16, ldh, 1, Wcat, 1, H, dH, H)          # dH = dheads Wcat
            dW = torch.empty(NH, H, device=dev, dtype=torch.float32)
            gemm16(NH, H, rows, dh16, 1, ldh, Hs16, 1, H, dW, H, rowsum=dbias)   # dWcat = dheads^T Hs, dbias = its row sums
            return (dH, None, None, None, None, dW[0:AM], dbias[0:AM], dW[AM : 2 * AM], dbias[AM : 2 * AM], dW[2 * AM : 3 * AM],
                    dbias[2 * AM : 3 * AM], dW[3 * AM :], dbias[3 * AM :])
        wgrads = []
        for i, (W, n) in enumerate(((wp, AM), (wm, AM), (wsc, AM), (wg, 2))):
            off = i * AM
            dW = torch.empty_like(W)
            gemm(n, H, rows, dheads, 1, HEAD_LD, Hs, 1, H, dW, H, a_off=off)
            wgrads.append(dW)
            gemm(rows, H, n, dheads, HEAD_LD, 1, W, 1, H, dH, H, a_off=off, accumulate=(i > 0))
        return (dH, None, None, None, None, wgrads[0], dbias[0:AM], wgrads[1], dbias[AM : 2 * AM], wgrads[2],
                dbias[2 * AM : 3 * AM], wgrads[3], dbias[3 * AM :])


def heads_unpack(heads, B, S, A, M, ls_min, time_major=True):
    dev = heads.device
    lp = torch.empty(B, S, A, M, device=dev, dtype=torch.float32)
    ls = torch.empty(B, S, A, M, device=dev, dtype=torch.float32)
    mu = torch.empty(B, S, A, M, device=dev, dtype=torch.float32)
    gr = torch.empty(B, S, 2, device=dev, dtype=torch.float32)
    call("hulc2_heads_unpack", heads.data_ptr(), HEAD_LD, lp.data_ptr(), ls.data_ptr(), mu.data_ptr(), gr.data_ptr(), B, S, A, M,
         ls_min, int(time_major))
    return lp, ls, mu, gr


def heads_pack(lp, ls, mu, gr) -> torch.Tensor:
    """[B,S,A,M] tensors -> batch-major fused heads buffer (for the public _loss/_sample API)."""
    B, S, A, M = lp.shape
    AM = A * M
    heads = torch.empty(B * S, HEAD_LD, device=lp.device, dtype=torch.float32)
    for t, off in ((lp, 0), (mu, AM), (ls, 2 * AM)):
        t = _f32(t).contiguous()
        call("hulc2_copy2d", t.data_ptr(), AM, heads.data_ptr() + 4 * off, HEAD_LD, B * S, AM, 0)
    if gr is not None:
        gr = _f32(gr).contiguous()
        call("hulc2_copy2d", gr.data_ptr(), 2, heads.data_ptr() + 4 * 3 * AM, HEAD_LD, B * S, 2, 0)
    return heads


class LogisticLossFunction(torch.autograd.Function):
    """Public ``_loss`` on separate [B,S,A,M] tensors (logistic_decoder_rnn.py:133-152)."""

    @staticmethod
    def forward(ctx, lp, ls, mu, gr, actions, amin, amax, cfg):
        B, S, A, M = lp.shape
        _, _, num_classes, ls_min, alpha = cfg
        heads = heads_pack(lp, ls, mu, gr)
        out = torch.empty(3, device=lp.device, dtype=torch.float32)
        ws = workspace(lp.device)
        actions = _f32(actions).contiguous()
        call("hulc2_logistic_loss_fwd", heads.data_ptr(), HEAD_LD, actions.data_ptr(), amin.data_ptr(), amax.data_ptr(),
             out.data_ptr(), B, S, A, M, num_classes, ls_min, alpha, 0, ws.data_ptr(), ws.numel())
        ctx.save_for_backward(heads, actions, amin, amax)
        ctx.cfg, ctx.dims = cfg, (B, S, A, M)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        heads, actions, amin, amax = ctx.saved_tensors
        B, S, A, M = ctx.dims
        _, _, num_classes, ls_min, alpha = ctx.cfg
        dheads = torch.empty_like(heads)
        call("hulc2_logistic_loss_bwd", heads.data_ptr(), HEAD_LD, actions.data_ptr(), amin.data_ptr(), amax.data_ptr(),
             g.contiguous().data_ptr(), dheads.data_ptr(), B, S, A, M, num_classes, ls_min, alpha, 0)
        AM = A * M
        outs = []
        for off, n, shape in ((0, AM, (B, S, A, M)), (2 * AM, AM, (B, S, A, M)), (AM, AM, (B, S, A, M)), (3 * AM, 2, (B, S, 2))):
            t = torch.empty(B * S, n, device=heads.device, dtype=torch.float32)
            call("hulc2_copy2d", dheads.data_ptr() + 4 * off, HEAD_LD, t.data_ptr(), n, B * S, n, 0)
            outs.append(t.view(shape))
        return (outs[0], outs[1], outs[2], outs[3], None, None, None, None)


def logistic_sample(heads, u1, u2, gripper_bounds, B, S, A, M, ls_min, time_major) -> torch.Tensor:
    act = torch.empty(B, S, A + 1, device=heads.device, dtype=torch.float32)
    call("hulc2_logistic_sample", heads.data_ptr(), HEAD_LD, _f32(u1).contiguous().data_ptr(), _f32(u2).contiguous().data_ptr(),
         gripper_bounds.data_ptr(), act.data_ptr(), B, S, A, M, ls_min, int(time_major))
    return act


def val_metrics(pred: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """hulc2.py:292-302 + the reductions logged in validation_step (:559-575) in one launch:
    -> float32[4] = (total MAE, position MAE, orientation MAE, discrete-gripper success rate) for pred / actions [B,S,A+1]."""
    p, a = _f32(pred).contiguous(), _f32(actions).contiguous()
    B, S, A1 = a.shape
    out = torch.empty(4, device=p.device, dtype=torch.float32)
    call("hulc2_val_metrics", p.data_ptr(), a.data_ptr(), out.data_ptr(), B, S, A1 - 1)
    return out


def world_to_tcp(actions: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    a, r = _f32(actions).contiguous(), _f32(robot_obs).contiguous()
    out = torch.empty_like(a)
    call("hulc2_world_to_tcp", a.data_ptr(), r.data_ptr(), r.shape[-1], out.data_ptr(), a.numel() // 7)
    return out


def tcp_to_world(actions: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    a, r = _f32(actions).contiguous(), _f32(robot_obs).contiguous()
    out = torch.empty_like(a)
    call("hulc2_tcp_to_world", a.data_ptr(), r.data_ptr(), r.shape[-1], out.data_ptr(), a.numel() // 7)
    return out


# ----------------------------------------------------------------------------- InfoNCE
class InfoNCEFunction(torch.autograd.Function):
    """hulc2.py:494-508 on projected features; ``use`` = uint8 row mask (static shapes, no host sync)."""

    @staticmethod
    def forward(ctx, img, txt, logit_scale, use):
        img, txt = _f32(img).contiguous(), _f32(txt).contiguous()
        B, D = img.shape
        loss = torch.empty(1, device=img.device, dtype=torch.float32)
        ws = workspace(img.device)
        _lib.tag(f"infonce_fwd[B={B},D={D}]", 4.0 * B * B * D, 8.0 * B * D)
        call("hulc2_infonce_fwd", img.data_ptr(), txt.data_ptr(), _p(use), logit_scale.data_ptr(), loss.data_ptr(), B, D,
             ws.data_ptr(), ws.numel())
        ctx.save_for_backward(img, txt, logit_scale)
        ctx.use = use
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        img, txt, logit_scale = ctx.saved_tensors
        B, D = img.shape
        dimg, dtxt = torch.empty_like(img), torch.empty_like(txt)
        dls = torch.zeros((), device=img.device, dtype=torch.float32)
        ws = workspace(img.device)
        _lib.tag(f"infonce_bwd[B={B},D={D}]", 8.0 * B * B * D, 16.0 * B * D)
        call("hulc2_infonce_bwd", img.data_ptr(), txt.data_ptr(), _p(ctx.use), logit_scale.data_ptr(), g.contiguous().data_ptr(),
             dimg.data_ptr(), dtxt.data_ptr(), dls.data_ptr(), B, D, ws.data_ptr(), ws.numel())
        return dimg, dtxt, dls, None


# ----------------------------------------------------------------------------- scalar combine
class WeightedSumFunction(torch.autograd.Function):
    """total = sum_i w_i * x_i for 0-d tensors (hulc2.py:243,426-430) without leaving the library."""

    @staticmethod
    def forward(ctx, weights, *xs):
        n = len(xs)
        assert 1 <= n <= 8 and len(weights) == n
        xs = [_f32(x).contiguous() for x in xs]
        out = torch.empty((), device=xs[0].device, dtype=torch.float32)
        call("hulc2_weighted_sum", (C.c_void_p * n)(*[x.data_ptr() for x in xs]), (C.c_float * n)(*weights), n, out.data_ptr())
        ctx.weights = weights
        return out

    @staticmethod
    def backward(ctx, g):
        n = len(ctx.weights)
        outs = torch.empty(n, device=g.device, dtype=torch.float32)
        call("hulc2_weighted_fanout", g.contiguous().data_ptr(), (C.c_float * n)(*ctx.weights), n, outs.data_ptr())
        return (None, *outs.unbind(0))


def weighted_sum(weights, xs):
    return WeightedSumFunction.apply(tuple(float(w) for w in weights), *xs)


class ConcatColsFunction(torch.autograd.Function):
    """[a | b] along the feature axis for 2-D row-strided operands (plan_proposal_net.py:43)."""

    @staticmethod
    def forward(ctx, a, b):
        a2, b2 = _rows2d(_f32(a)), _rows2d(_f32(b))
        M, Ka, Kb = a2.shape[0], a2.shape[1], b2.shape[1]
        out = torch.empty(M, Ka + Kb, device=a.device, dtype=torch.float32)
        call("hulc2_copy2d", a2.data_ptr(), _ld(a2), out.data_ptr(), Ka + Kb, M, Ka, 0)
        call("hulc2_copy2d", b2.data_ptr(), _ld(b2), out.data_ptr() + 4 * Ka, Ka + Kb, M, Kb, 0)
        ctx.dims = (M, Ka, Kb)
        return out

    @staticmethod
    def backward(ctx, dout):
        M, Ka, Kb = ctx.dims
        dout = dout.contiguous()
        da = torch.empty(M, Ka, device=dout.device, dtype=torch.float32)
        db = torch.empty(M, Kb, device=dout.device, dtype=torch.float32)
        call("hulc2_copy2d", dout.data_ptr(), Ka + Kb, da.data_ptr(), Ka, M, Ka, 0)
        call("hulc2_copy2d", dout.data_ptr() + 4 * Ka, Ka + Kb, db.data_ptr(), Kb, M, Kb, 0)
        return da, db


def concat_cols(a, b):
    return ConcatColsFunction.apply(a, b)


class DecoderInputFunction(torch.autograd.Function):
    """x_t = [plan | emb_t | goal] for every step (logistic_decoder_rnn.py:262-268), TIME-major rows [S*B, P+Es+G].  Used by the
    MLP decoder (decoders/utils/rnn.py:39-46), which has no recurrence to fold the constant terms into."""

    @staticmethod
    def forward(ctx, plan, emb, goal):
        plan, goal = _f32(plan).contiguous(), _f32(goal).contiguous()
        B, S, Es = emb.shape
        P, G = plan.shape[1], goal.shape[1]
        In = P + Es + G
        assert emb.stride(2) == 1
        x = torch.empty(S * B, In, device=plan.device, dtype=torch.float32)
        # dst[d1 = s, d0 = b, :D2] = src[b * s0 + s * s1 + :D2]: s1 = 0 broadcasts plan / goal over time
        call("hulc2_transpose01", plan.data_ptr(), P, 0, x.data_ptr(), In, B, S, P, 0)
        call("hulc2_transpose01", emb.data_ptr(), emb.stride(0), emb.stride(1), x.data_ptr() + 4 * P, In, B, S, Es, 0)
        call("hulc2_transpose01", goal.data_ptr(), G, 0, x.data_ptr() + 4 * (P + Es), In, B, S, G, 0)
        ctx.dims = (B, S, Es, P, G)
        return x

    @staticmethod
    def backward(ctx, dx):
        B, S, Es, P, G = ctx.dims
        In = P + Es + G
        dx = dx.contiguous()
        dev = dx.device
        dplan = demb = dgoal = None
        tsum = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
            tsum = torch.empty(B, In, device=dev, dtype=torch.float32)
            colsum(dx, B * In, S, B * In, tsum)                                  # sum over time, per window
        if ctx.needs_input_grad[0]:
            dplan = torch.empty(B, P, device=dev, dtype=torch.float32)
            call("hulc2_copy2d", tsum.data_ptr(), In, dplan.data_ptr(), P, B, P, 0)
        if ctx.needs_input_grad[2]:
            dgoal = torch.empty(B, G, device=dev, dtype=torch.float32)
            call("hulc2_copy2d", tsum.data_ptr() + 4 * (P + Es), In, dgoal.data_ptr(), G, B, G, 0)
        if ctx.needs_input_grad[1]:
            demb = torch.empty(B, S, Es, device=dev, dtype=torch.float32)
            # dst[d1 = b, d0 = s, :Es] = dx[s * (B*In) + b * In + P + :Es]
            call("hulc2_transpose01", dx.data_ptr() + 4 * P, B * In, In, demb.data_ptr(), Es, S, B, Es, 0)
        return dplan, demb, dgoal


class ConcatRowsFunction(torch.autograd.Function):
    """Stacks 2-D row blocks [B_i, D] into [sum B_i, D] (latent goals of several modalities -> one decoder batch)."""

    @staticmethod
    def forward(ctx, *xs):
        xs2 = [_rows2d(_f32(x)) for x in xs]
        D = xs2[0].shape[1]
        out = torch.empty(sum(x.shape[0] for x in xs2), D, device=xs2[0].device, dtype=torch.float32)
        r0 = 0
        for x in xs2:
            call("hulc2_copy2d", x.data_ptr(), _ld(x), out.data_ptr() + 4 * r0 * D, D, x.shape[0], D, 0)
            r0 += x.shape[0]
        ctx.rows = [x.shape[0] for x in xs2]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        outs, r0 = [], 0
        for n in ctx.rows:
            outs.append(dout[r0 : r0 + n])
            r0 += n
        return tuple(outs)


def concat_rows(xs):
    return xs[0] if len(xs) == 1 else ConcatRowsFunction.apply(*xs)
