"""ctypes binding of the C-ABI in ``include/hulc2_b200.h``.

The library is loaded lazily on the first op call.  There is NO fallback: a missing
``libhulc2_b200.so`` or a missing CUDA device raises ``RuntimeError`` (the product path must
fail loudly when the CUDA extension is absent).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhulc2_b200.so")

c_float_p = C.c_void_p  # device pointers travel as integers
LL = C.c_longlong
I = C.c_int
F = C.c_float
P = C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", I), ("N", I), ("K", I),
        ("A", P), ("a_rs", LL), ("a_ks", LL), ("a_inner", I), ("a_rs_outer", LL), ("a_rs_inner", LL),
        ("B", P), ("b_rs", LL), ("b_ks", LL),
        ("C", P), ("ldc", LL), ("c_inner", I), ("c_rs_outer", LL), ("c_rs_inner", LL),
        ("bias", P),
        ("add", P), ("ld_add", LL),
        ("mask", P), ("ld_mask", LL),
        ("keep", P), ("ld_keep", LL), ("keep_scale", F),
        ("relu", I), ("accumulate", I),
        ("alpha", F),
        ("precision", I),
        ("workspace", P), ("workspace_bytes", LL),
        ("A16", P), ("B16", P), ("C16", P), ("ld16", LL),
        ("rowsum", P),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("F", I), ("C", I), ("H", I), ("W", I), ("Cout", I), ("KH", I), ("KW", I), ("stride", I), ("in_nhwc", I),
        ("x", P), ("w", P), ("bias", P), ("y", P), ("relu", I),
        ("dy", P), ("dw", P), ("dx", P), ("xmask", P), ("accumulate", I),
        ("precision", I),
        ("workspace", P), ("workspace_bytes", LL),
    ]


class ConvbArgs(C.Structure):
    _fields_ = [
        ("F", I), ("C", I), ("H", I), ("W", I), ("Cout", I), ("KH", I), ("KW", I), ("stride", I),
        ("x", P), ("w", P), ("bias", P), ("y", P), ("relu", I),
        ("dy", P), ("dx", P), ("xmask", P),
        ("dw", P), ("db", P), ("dw_layout", I),
        ("workspace", P), ("workspace_bytes", LL),
        ("mask_bits", P), ("mask_bits_written", I),
    ]


# name -> argtypes (without the trailing stream); every function returns int and takes a stream last
_SIGS = {
    "hulc2_gemm": [C.POINTER(GemmArgs)],
    "hulc2_f32_to_bf16": [P, P, LL],
    "hulc2_f32_to_bf16_2d": [P, LL, P, LL, LL, I],
    "hulc2_conv2d_fwd": [C.POINTER(ConvArgs)],
    "hulc2_conv2d_wgrad": [C.POINTER(ConvArgs)],
    "hulc2_conv2d_dgrad": [C.POINTER(ConvArgs)],
    "hulc2_permute_conv_weight": [P, P, I, I, I, I, I, I],
    "hulc2_pack_frames_bf16": [P, P, I, I, I, I],
    "hulc2_frames_u8_pack_bf16": [P, P, P, P, P, I, I, I, I, I],
    "hulc2_frames_u8_to_f32": [P, P, P, P, P, I, I, I, I, I],
    "hulc2_window_gather_f32": [P, P, P, P, I, I, I, I],
    "hulc2_convb_pack_weight": [P, P, I, I, I, I, I, I],
    "hulc2_convb_fwd": [C.POINTER(ConvbArgs)],
    "hulc2_convb_dgrad": [C.POINTER(ConvbArgs)],
    "hulc2_convb_wgrad": [C.POINTER(ConvbArgs)],
    "hulc2_spatial_softmax_fwd_bf16": [P, P, P, P, P, I, I, I],
    "hulc2_spatial_softmax_bwd_bf16": [P, P, P, P, P, P, P, I, I, I, I],
    "hulc2_spatial_softmax_fwd_bf16_stats": [P, P, P, P, P, P, I, I, I],
    "hulc2_spatial_softmax_bwd_bf16_stats": [P, P, P, P, P, P, P, P, P, I, I, I, I],
    "hulc2_nhwc_bf16_to_nchw": [P, P, I, I, I],
    "hulc2_nchw_to_nhwc_bf16": [P, P, I, I, I, P],
    "hulc2_copy2d": [P, LL, P, LL, LL, I, I],
    "hulc2_transpose01": [P, LL, LL, P, LL, I, I, I, I],
    "hulc2_fill": [P, LL, F],
    "hulc2_axpy": [P, P, LL, F],
    "hulc2_weighted_sum": [P, P, I, P],
    "hulc2_weighted_fanout": [P, P, I, P],
    "hulc2_colsum": [P, LL, LL, I, P, I, P, LL],
    "hulc2_relu_mask": [P, P, P, LL],
    "hulc2_nhwc_to_nchw": [P, P, I, I, I],
    "hulc2_nchw_to_nhwc": [P, P, I, I, I, P],
    "hulc2_spatial_softmax_fwd": [P, P, P, P, P, I, I, I],
    "hulc2_spatial_softmax_bwd": [P, P, P, P, P, P, P, P, I, I, I, I],
    "hulc2_layernorm_fwd": [P, LL, P, LL, P, F, P, P, P, LL, P, P, P, LL, I, F],
    "hulc2_layernorm_bwd": [P, LL, P, LL, P, P, P, P, LL, P, P, F, P, P, LL, I],
    "hulc2_layernorm_fwd_m": [P, LL, P, LL, P, F, P, P, P, LL, P, P, P, LL, I, F, P, LL],
    "hulc2_layernorm_bwd_m": [P, LL, P, LL, P, P, P, P, LL, P, P, F, P, P, LL, I, P, P, LL],
    "hulc2_add_pos_fwd": [P, P, P, F, P, I, I, I],
    "hulc2_add_pos_fwd_m": [P, P, P, F, P, P, I, I, I],
    "hulc2_attention_fwd_m": [P, P, F, P, P, P, LL, I, I, I, I],
    "hulc2_attention_bwd_m": [P, P, P, F, P, P, P, LL, I, I, I, I],
    "hulc2_add_pos_bwd": [P, P, F, P, P, I, I, I],
    "hulc2_attention_fwd": [P, P, F, P, P, I, I, I, I],
    "hulc2_attention_bwd": [P, P, P, F, P, P, I, I, I, I],
    "hulc2_mean_seq_fwd": [P, P, I, I, I],
    "hulc2_mean_seq_bwd": [P, P, I, I, I],
    "hulc2_kl_fwd": [P, P, P, I, I, I, F, F],
    "hulc2_kl_bwd": [P, P, P, P, P, I, I, I, F, F],
    "hulc2_onehot_fwd": [P, P, I, I, I],
    "hulc2_st_onehot_bwd": [P, P, P, I, I, I],
    "hulc2_categorical_sample": [P, P, P, I, I, I],
    "hulc2_logistic_loss_fwd": [P, LL, P, P, P, P, I, I, I, I, I, F, F, I, P, LL],
    "hulc2_logistic_loss_bwd": [P, LL, P, P, P, P, P, I, I, I, I, I, F, F, I],
    "hulc2_logistic_loss_seg_fwd": [P, LL, P, P, P, P, I, I, I, I, I, F, F, I, I, P, LL],
    "hulc2_logistic_loss_seg_bwd": [P, LL, P, P, P, P, P, I, I, I, I, I, F, F, I, I],
    "hulc2_logistic_sample": [P, LL, P, P, P, P, I, I, I, I, F, I],
    "hulc2_val_metrics": [P, P, P, I, I, I],
    "hulc2_heads_unpack": [P, LL, P, P, P, P, I, I, I, I, F, I],
    "hulc2_world_to_tcp": [P, P, I, P, LL],
    "hulc2_tcp_to_world": [P, P, I, P, LL],
    "hulc2_infonce_fwd": [P, P, P, P, P, I, I, P, LL],
    "hulc2_infonce_bwd": [P, P, P, P, P, P, P, P, I, I, P, LL],
    "hulc2_rnn_relu_fwd": [P, P, P, P, I, I, I, I, P, LL],
    "hulc2_rnn_relu_bwd": [P, P, P, P, I, I, I, I, P, LL],
    "hulc2_rnn_relu_fwd_m": [P, P, P, P, P, I, I, I, I, P, LL],
    "hulc2_rnn_relu_bwd_m": [P, P, P, P, P, I, I, I, I, P, LL],
    "hulc2_gru_cell_fwd": [P, LL, P, P, P, P, I, I],
    "hulc2_gru_cell_bwd": [P, P, P, P, P, LL, P, P, I, I],
    "hulc2_lstm_cell_fwd": [P, LL, P, P, P, P, P, I, I],
    "hulc2_lstm_cell_bwd": [P, P, P, P, P, P, P, LL, P, I, I],
    "hulc2_gauss_state_fwd": [P, P, P, I, I],
    "hulc2_gauss_state_bwd": [P, P, P, P, I, I],
    "hulc2_gauss_rsample": [P, P, P, P, LL],
    "hulc2_box_muller": [P, P, P, LL],
    "hulc2_gauss_kl_fwd": [P, P, P, P, P, I, I, F, F],
    "hulc2_gauss_kl_bwd": [P, P, P, P, P, P, P, P, P, I, I, F, F],
    "hulc2_adam_step": [P, P, P, P, LL, F, F, F, F, F, I, F],
    "hulc2_philox_uniform": [P, LL, C.c_ulonglong, C.c_ulonglong],
    "hulc2_dropout_mask": [P, LL, F, C.c_ulonglong, C.c_ulonglong],
    "hulc2_adam_step_dev": [P, P, P, P, LL, F, F, F, F, F, P, I, F],
    "hulc2_adam_step_graph": [P, P, P, P, LL, P, F, F, F, F, P, I, F],
    "hulc2_philox_uniform_ep": [P, LL, C.c_ulonglong, C.c_ulonglong, P],
    "hulc2_dropout_mask_ep": [P, LL, F, C.c_ulonglong, C.c_ulonglong, P],
    "hulc2_counter_add": [P, C.c_ulonglong],
}
_NO_STREAM = {"hulc2_last_error": (C.c_char_p, []), "hulc2_version": (I, []), "hulc2_device_supports_tcgen05": (I, []),
              "hulc2_launch_count": (C.c_ulonglong, []), "hulc2_tma_gemm_count": (C.c_ulonglong, []), "hulc2_convb_supported": (I, [I, I, I, I, I]),
              "hulc2_rnn_select_kernel": (I, [I]), "hulc2_spatial_softmax_stats_supported": (I, [I, I]), "hulc2_rnn_cluster_capacity": (I, [I]), "hulc2_rnn_device_error": (I, [I]), "hulc2_rnn_last_path": (I, []),
              "hulc2_rnn_set_trace": (I, [P])}

EXPORTED_SYMBOLS = sorted(list(_SIGS) + list(_NO_STREAM))

_lib: Optional[C.CDLL] = None
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it as gpu_launches evidence)


def load_library(require_cuda: bool = True) -> C.CDLL:
    """Loads libhulc2_b200.so and declares prototypes.  ``require_cuda=False`` is for the CPU-side
    symbol-export test only; compute entry points still need a device."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"hulc2_b200: CUDA library not found at {LIB_PATH}. Build it with `python -m hulc2_b200.build` "
                "(there is no CPU or PyTorch fallback)."
            )
        lib = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = I
            fn.argtypes = list(args) + [P]
        for name, (res, args) in _NO_STREAM.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_cuda and not torch.cuda.is_available():
        raise RuntimeError("hulc2_b200: no CUDA device available; this package has no CPU fallback")
    return _lib


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_prof = None   # list of (key, flops, start_event, end_event) while profiling
_tag = None    # (key, flops) annotation for the next call
_flush = None  # (ptr, n floats) of a buffer larger than L2: written before every profiled call
_prof_external = False   # events become event-record NODES of the CUDA graph being captured (timeline of a replay)


def tag(key: str, flops: float = 0.0, nbytes: float = 0.0) -> None:
    """Annotates the next C-ABI call (shape key + algorithmic FLOPs and HBM bytes) for bench.py's per-kernel timing."""
    global _tag
    if _prof is not None:
        _tag = (key, flops, nbytes)


def profile_begin(flush: Optional[torch.Tensor] = None) -> None:
    """Starts per-call CUDA-event timing of every C-ABI call.  ``flush``: an fp32 device buffer larger than L2; it is
    re-written (``hulc2_fill``) before EVERY profiled call.  That does two things: each kernel starts with a cold L2 (as under
    ncu), and the ~40 us the GPU spends on the fill keep it behind the host, so the event pair brackets the kernel alone
    instead of the host's launch latency (an eager step is host-bound: without the fill every small call reads 15-20 us)."""
    global _prof, _flush, _prof_external
    _prof = []
    _flush = None if flush is None else (flush.data_ptr(), flush.numel())
    _prof_external = False


def profile_begin_graph() -> None:
    """Per-call timing INSIDE a CUDA graph: call while the step is being captured.  Every C-ABI call is bracketed by two
    external timing events (cudaEventRecordExternal -> event-record nodes of the graph); after a replay ``profile_read_graph``
    returns what each call took in the replayed step -- no host latency, warm caches, the order and overlap of the real step."""
    global _prof, _flush, _prof_external
    _prof, _flush, _prof_external = [], None, True


def profile_read_gaps(top: int = 12) -> list:
    """After a replay of the graph captured under ``profile_begin_graph``: the ``top`` largest intervals BETWEEN two consecutive
    C-ABI calls, each with the calls on either side -- where the replayed step waits (the gradient all-reduce before the optimizer
    step, torch kernels, stream joins): [{"ms", "after", "before", "index"}]."""
    torch.cuda.synchronize()
    out = []
    for i in range(1, len(_prof)):
        out.append({"ms": max(_prof[i - 1][4].elapsed_time(_prof[i][3]), 0.0), "after": _prof[i - 1][0], "before": _prof[i][0], "index": i})
    out.sort(key=lambda g: -g["ms"])
    return out[:top]


def profile_read_graph(other_key: str = "(between calls: torch kernels, memsets, launch gaps)") -> dict:
    """After a replay of the graph captured under ``profile_begin_graph``: {key: {key, ms, calls, flops, bytes}}; the time
    between the end of one C-ABI call and the start of the next is booked under ``other_key``."""
    torch.cuda.synchronize()
    out = {}
    prev = None
    for key, flops, nbytes, e0, e1 in _prof:
        r = out.setdefault(key, {"key": key, "ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        r["ms"] += e0.elapsed_time(e1)
        r["calls"] += 1
        r["flops"] += flops
        r["bytes"] += nbytes
        if prev is not None:
            g = out.setdefault(other_key, {"key": other_key, "ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
            g["ms"] += max(prev.elapsed_time(e0), 0.0)
            g["calls"] += 1
        prev = e1
    return out


def profile_end_graph() -> None:
    global _prof, _tag, _flush, _prof_external
    _prof, _tag, _flush, _prof_external = None, None, None, False


def profile_end() -> dict:
    """Returns {key: {key, ms, calls, flops}} aggregated over the profiled region (CUDA events on the launch stream)."""
    global _prof, _tag, _flush, _prof_external
    _prof_external = False
    torch.cuda.synchronize()
    out = {}
    for key, flops, nbytes, e0, e1 in _prof:
        r = out.setdefault(key, {"key": key, "ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        r["ms"] += e0.elapsed_time(e1)
        r["calls"] += 1
        r["flops"] += flops
        r["bytes"] += nbytes
    _prof, _tag, _flush = None, None, None
    return out


_AUTO_KEY = {"hulc2_f32_to_bf16": (2,), "hulc2_f32_to_bf16_2d": (4, 5), "hulc2_axpy": (2,), "hulc2_copy2d": (4, 5), "hulc2_colsum": (2, 3),
             "hulc2_fill": (1,), "hulc2_layernorm_fwd": (14, 15), "hulc2_layernorm_bwd": (14, 15), "hulc2_dropout_mask_ep": (1,),
             "hulc2_layernorm_fwd_m": (14, 15), "hulc2_layernorm_bwd_m": (14, 15)}


# algorithmic HBM bytes of the untagged memory-bound helpers (same argument positions as above): product of the size
# arguments times bytes moved per element
_AUTO_BYTES = {"hulc2_f32_to_bf16": 6, "hulc2_f32_to_bf16_2d": 6, "hulc2_axpy": 12, "hulc2_copy2d": 8, "hulc2_colsum": 4, "hulc2_fill": 4,
               "hulc2_layernorm_fwd": 16, "hulc2_layernorm_bwd": 16, "hulc2_dropout_mask_ep": 1,
               "hulc2_layernorm_fwd_m": 18, "hulc2_layernorm_bwd_m": 18}


def _auto_bytes(name: str, args) -> float:
    idx = _AUTO_KEY.get(name)
    if idx is None:
        return 0.0
    n = 1.0
    for i in idx:
        n *= float(args[i])
    return n * _AUTO_BYTES[name]


def _auto_key(name: str, args) -> str:
    """Profile key of an untagged call: the entry point plus its size arguments (profiling only)."""
    idx = _AUTO_KEY.get(name)
    base = name[:-2] if name.endswith("_m") else name       # "_m" = same operation + bf16 mirror output: one family
    return base if idx is None else f"{base}[{','.join(str(args[i]) for i in idx)}]"


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on torch's current stream; raise RuntimeError on failure."""
    global launch_count, _tag
    lib = load_library()
    if _prof is not None:
        key, flops, nbytes = _tag if _tag is not None else (_auto_key(name, args), 0.0, _auto_bytes(name, args))
        _tag = None
        e0, e1 = torch.cuda.Event(enable_timing=True, external=_prof_external), torch.cuda.Event(enable_timing=True, external=_prof_external)
        if _flush is not None:
            lib.hulc2_fill(_flush[0], _flush[1], 0.0, stream())
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        _prof.append((key, flops, nbytes, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream())
    launch_count += 1
    if rc != 0:
        msg = lib.hulc2_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed with code {rc}: {msg}")
