"""ctypes binding of libpaid_attn.so (include/paid_attn.h).

PyTorch is used for device memory and streams only: every call passes raw device
pointers (``tensor.data_ptr()``) and the current ``cudaStream_t``.  There is no
CPU or PyTorch fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PAID_LIB_PATH") or os.path.join(_HERE, "lib", "libpaid_attn.so")

PAID_OK, PAID_EINVAL, PAID_EUNSUPPORTED, PAID_ECUDA, PAID_EWORKSPACE = 0, -1, -2, -3, -4
PAID_F16, PAID_BF16 = 0, 1
PAID_PLAIN, PAID_OUTER, PAID_INNER = 0, 1, 2
FLAG_GENERIC_KERNELS = 1
FLAG_ONE_WARPGROUP = 2
MODES = {"plain": PAID_PLAIN, "outer": PAID_OUTER, "inner": PAID_INNER}

EXPORTS = [
    "paid_attn_abi_version", "paid_attn_workspace_bytes", "paid_attn_core_workspace_bytes", "paid_attn_forward",
    "paid_attn_core", "paid_attn_project_endpoints", "paid_attn_project_kv", "paid_linear", "paid_attn_last_error",
    "paid_attn_launch_count", "paid_attn_last_kernel", "paid_attn_profile_enable", "paid_attn_profile_read",
    "paid_attn_profile_rows",
    "paid_linear_geglu", "paid_geglu", "paid_add_layer_norm", "paid_group_norm_nhwc", "paid_group_norm_workspace_bytes",
    "paid_residual_bias_add",
]


class PaidAttnParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32),
        ("dtype", C.c_int32), ("mode", C.c_int32), ("fused", C.c_int32),
        ("N", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("Cc", C.c_int32),
        ("heads", C.c_int32), ("scale", C.c_float),
        ("begin_frame", C.c_int32), ("end_frame", C.c_int32),
        ("x", C.c_void_p), ("ctx", C.c_void_p), ("wq", C.c_void_p), ("wk", C.c_void_p), ("wv", C.c_void_p),
        ("wo", C.c_void_p), ("bo", C.c_void_p), ("coef", C.c_void_p), ("kv_ext", C.c_void_p),
        ("y", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
        ("k_pre", C.c_void_p), ("v_pre", C.c_void_p), ("kv_pre_broadcast", C.c_int32), ("plain_tail", C.c_int32),
        ("kv_ext_ready_event", C.c_void_p),
    ]


class PaidProfileRow(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("d", C.c_int64 * 4), ("launches", C.c_uint64),
                ("total_ms", C.c_double), ("flops", C.c_double)]


PROFILE_KINDS = {0: "attention", 1: "linear", 2: "linear_geglu"}


class PaidCoreParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32),
        ("dtype", C.c_int32), ("mode", C.c_int32), ("fused", C.c_int32),
        ("N", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("begin_frame", C.c_int32), ("end_frame", C.c_int32),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("kv_ext", C.c_void_p), ("coef", C.c_void_p),
        ("out", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
        ("accumulate", C.c_int32), ("out_scale", C.c_float), ("out_frame_scale", C.c_void_p),
        ("kv_broadcast", C.c_int32),
    ]


_lib = None


def load_library() -> C.CDLL:
    """Load the in-tree shared library (built by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.paid_attn_abi_version.restype = C.c_int
    lib.paid_attn_workspace_bytes.restype = C.c_uint64
    lib.paid_attn_workspace_bytes.argtypes = [C.POINTER(PaidAttnParams)]
    lib.paid_attn_core_workspace_bytes.restype = C.c_uint64
    lib.paid_attn_core_workspace_bytes.argtypes = [C.POINTER(PaidCoreParams)]
    lib.paid_attn_forward.restype = C.c_int
    lib.paid_attn_forward.argtypes = [C.POINTER(PaidAttnParams), C.c_void_p]
    lib.paid_attn_core.restype = C.c_int
    lib.paid_attn_core.argtypes = [C.POINTER(PaidCoreParams), C.c_void_p]
    lib.paid_attn_project_endpoints.restype = C.c_int
    lib.paid_attn_project_endpoints.argtypes = [C.POINTER(PaidAttnParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.paid_attn_project_kv.restype = C.c_int
    lib.paid_attn_project_kv.argtypes = [C.POINTER(PaidAttnParams), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.paid_linear.restype = C.c_int
    lib.paid_linear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                C.c_int32, C.c_uint32, C.c_void_p]
    lib.paid_linear_geglu.restype = C.c_int
    lib.paid_linear_geglu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_uint32, C.c_void_p]
    lib.paid_geglu.restype = C.c_int
    lib.paid_geglu.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.paid_add_layer_norm.restype = C.c_int
    lib.paid_add_layer_norm.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p]
    lib.paid_residual_bias_add.restype = C.c_int
    lib.paid_residual_bias_add.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.paid_group_norm_workspace_bytes.restype = C.c_uint64
    lib.paid_group_norm_workspace_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    lib.paid_group_norm_nhwc.restype = C.c_int
    lib.paid_group_norm_nhwc.argtypes = [C.c_void_p] * 6 + [C.c_uint64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                         C.c_int32, C.c_int32, C.c_void_p]
    lib.paid_attn_last_error.restype = C.c_char_p
    lib.paid_attn_launch_count.restype = C.c_uint64
    lib.paid_attn_last_kernel.restype = C.c_char_p
    lib.paid_attn_profile_enable.restype = C.c_int
    lib.paid_attn_profile_enable.argtypes = [C.c_int]
    lib.paid_attn_profile_read.restype = C.c_int
    lib.paid_attn_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int]
    lib.paid_attn_profile_rows.restype = C.c_int
    lib.paid_attn_profile_rows.argtypes = [C.POINTER(PaidProfileRow), C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    if lib.paid_attn_abi_version() != 3:
        raise RuntimeError("libpaid_attn.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().paid_attn_last_error().decode()


def last_kernel() -> str:
    return load_library().paid_attn_last_kernel().decode()


def launch_count() -> int:
    return int(load_library().paid_attn_launch_count())


def profile_enable(on: bool):
    _check(load_library().paid_attn_profile_enable(int(on)), "paid_attn_profile_enable")


def profile_read(reset: bool = True):
    """(total kernel ms, launches, algorithmic flops) of the attention-core launches since the last reset."""
    ms, n, fl = C.c_double(0), C.c_uint64(0), C.c_double(0)
    _check(load_library().paid_attn_profile_read(C.byref(ms), C.byref(n), C.byref(fl), int(reset)),
           "paid_attn_profile_read")
    return ms.value, int(n.value), fl.value


def profile_rows(reset: bool = True):
    """Per (kernel kind, shape) rows of the launches since the last reset: dicts with kind, d (shape key, see
    include/paid_attn.h), launches, ms, flops."""
    lib = load_library()
    n = C.c_uint64(0)
    _check(lib.paid_attn_profile_rows(None, 0, C.byref(n), 0), "paid_attn_profile_rows")
    rows = (PaidProfileRow * max(1, n.value))()
    _check(lib.paid_attn_profile_rows(rows, n.value, C.byref(n), int(reset)), "paid_attn_profile_rows")
    return [dict(kind=PROFILE_KINDS.get(r.kind, str(r.kind)), d=[int(v) for v in r.d], launches=int(r.launches), ms=r.total_ms,
                 flops=r.flops) for r in rows[:n.value]]


def _check(status: int, what: str):
    if status != PAID_OK:
        raise RuntimeError(f"{what} failed with status {status}: {last_error()}")


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return PAID_F16
    if t.dtype == torch.bfloat16:
        return PAID_BF16
    raise NotImplementedError(f"libpaid_attn computes in fp16/bf16 (fp32 accumulate); got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _dev_check(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("libpaid_attn needs CUDA tensors; there is no CPU path")
        if not t.is_contiguous():
            raise RuntimeError("libpaid_attn needs contiguous tensors")


def _stream(t: torch.Tensor):
    # the library launches on the CURRENT device (tensor maps, kernel attributes and the launch itself are per device)
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "wrap the call in torch.cuda.device(tensor.device)")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


_workspaces: dict = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One growing scratch buffer per device; calls are stream-ordered so layers can share it."""
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def make_params(x, ctx, wq, wk, wv, wo, bo, coef, heads: int, mode: int, fused: bool, scale: Optional[float] = None,
                begin_frame: Optional[int] = None, end_frame: Optional[int] = None, kv_ext=None, flags: int = 0,
                y=None, plain_tail: int = 0) -> PaidAttnParams:
    N, S, Cdim = x.shape
    N -= plain_tail                  # trailing classifier-free-guidance rows (PaidAttnParams.plain_tail)
    if plain_tail < 0 or N <= 0:
        raise ValueError(f"plain_tail={plain_tail} does not leave an interpolation sequence in a batch of {x.shape[0]}")
    L, Cc = (S, Cdim) if ctx is None else (ctx.shape[1], ctx.shape[2])
    p = PaidAttnParams()
    p.struct_size = C.sizeof(PaidAttnParams)
    p.flags = flags
    p.dtype, p.mode, p.fused = _dtype_code(x), mode, int(bool(fused))
    p.N, p.S, p.L, p.C, p.Cc, p.heads = N, S, L, Cdim, Cc, heads
    p.scale = float((Cdim // heads) ** -0.5 if scale is None else scale)
    p.begin_frame = 0 if begin_frame is None else begin_frame
    p.end_frame = N - 1 if end_frame is None else end_frame
    p.x, p.ctx, p.wq, p.wk, p.wv, p.wo, p.bo = map(_ptr, (x, ctx, wq, wk, wv, wo, bo))
    p.coef, p.kv_ext, p.y = _ptr(coef), _ptr(kv_ext), _ptr(y)
    p.plain_tail = plain_tail
    return p


def attn_forward(x, ctx, wq, wk, wv, wo, bo, coef, heads: int, mode: int, fused: bool, scale=None,
                 begin_frame=None, end_frame=None, kv_ext=None, flags: int = 0, out=None, k_pre=None, v_pre=None,
                 kv_pre_broadcast: bool = False, kv_ext_ready=None, plain_tail: int = 0) -> torch.Tensor:
    """One processor call through ``paid_attn_forward``.  Tensors: x (N,S,C), ctx None|(N,L,Cc), weights as in
    nn.Linear, coef fp32 (N,) on the device (None for plain mode).  ``plain_tail``: the last ``plain_tail`` frames of x /
    ctx / k_pre are the unconditional rows of a classifier-free-guidance batch and get stock attention (the first
    ``N - plain_tail`` frames are the interpolation sequence; coef has that many entries).  ``k_pre`` / ``v_pre``: K / V of the context projected
    earlier with ``project_kv`` ((N,L,C), or (1,L,C) with ``kv_pre_broadcast``); ``kv_ext_ready``: a ``torch.cuda.Event``
    the stream waits for before the attention core (the endpoint K/V in ``kv_ext`` arrive on another stream)."""
    lib = load_library()
    _dev_check(x, ctx, wq, wk, wv, wo, bo, coef, kv_ext, k_pre, v_pre)
    for t in (ctx, wq, wk, wv, wo, bo, kv_ext, k_pre, v_pre):
        if t is not None and t.dtype != x.dtype:
            raise RuntimeError("all tensors of a call must share one dtype")
    if coef is not None and coef.dtype != torch.float32:
        raise RuntimeError("coef must be fp32")
    y = torch.empty_like(x) if out is None else out
    p = make_params(x, ctx, wq, wk, wv, wo, bo, coef, heads, mode, fused, scale, begin_frame, end_frame, kv_ext, flags, y,
                    plain_tail)
    if k_pre is not None:
        if ctx is None and k_pre.shape[1] != x.shape[1]:
            raise RuntimeError("k_pre of a self-attention call must have S tokens")
        p.L = k_pre.shape[1]
        p.k_pre, p.v_pre, p.kv_pre_broadcast = k_pre.data_ptr(), v_pre.data_ptr(), int(bool(kv_pre_broadcast))
    if kv_ext_ready is not None:
        p.kv_ext_ready_event = kv_ext_ready.cuda_event
    need = lib.paid_attn_workspace_bytes(C.byref(p))
    if need == 0:
        raise RuntimeError(f"paid_attn_workspace_bytes rejected the parameters: {last_error()}")
    ws = _workspace(x.device, int(need))
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    _check(lib.paid_attn_forward(C.byref(p), _stream(x)), "paid_attn_forward")
    return y


def project_endpoints(x, ctx, wk, wv, heads: int, local_frame: int, k_out: torch.Tensor, v_out: torch.Tensor,
                      flags: int = 0):
    lib = load_library()
    _dev_check(x, ctx, wk, wv, k_out, v_out)
    p = make_params(x, ctx, None, wk, wv, None, None, None, heads, PAID_PLAIN, False, flags=flags)
    _check(lib.paid_attn_project_endpoints(C.byref(p), local_frame, k_out.data_ptr(), v_out.data_ptr(), _stream(x)),
           "paid_attn_project_endpoints")


def project_kv(x, ctx, wk, wv, heads: int, k_out: torch.Tensor, v_out: torch.Tensor, flags: int = 0):
    """K / V of every frame of the context (``paid_attn_project_kv``) into k_out / v_out (N,L,C)."""
    lib = load_library()
    _dev_check(x, ctx, wk, wv, k_out, v_out)
    p = make_params(x, ctx, None, wk, wv, None, None, None, heads, PAID_PLAIN, False, flags=flags)
    _check(lib.paid_attn_project_kv(C.byref(p), k_out.data_ptr(), v_out.data_ptr(), _stream(x)), "paid_attn_project_kv")


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, flags: int = 0,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = load_library()
    _dev_check(x, w, bias, out)
    K = x.shape[-1]
    M = x.numel() // K
    y = torch.empty(*x.shape[:-1], w.shape[0], dtype=x.dtype, device=x.device) if out is None else out
    _check(lib.paid_linear(x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), M, w.shape[0], K, _dtype_code(x),
                           flags, _stream(x)), "paid_linear")
    return y


def linear_geglu(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """``paid_linear_geglu``: (x Wa^T + ba) * gelu(x Wg^T + bg) for w = [Wa ; Wg] (2D, K): the feed-forward's first
    Linear with GEGLU in the GEMM epilogue; returns (..., D)."""
    lib = load_library()
    _dev_check(x, w, bias)
    K = x.shape[-1]
    M = x.numel() // K
    D = w.shape[0] // 2
    y = torch.empty(*x.shape[:-1], D, dtype=x.dtype, device=x.device)
    _check(lib.paid_linear_geglu(x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), M, D, K, _dtype_code(x), flags,
                                 _stream(x)), "paid_linear_geglu")
    return y


def geglu(h: torch.Tensor) -> torch.Tensor:
    """a * gelu(g) for h = [a | g] along the last dim (``paid_geglu``)."""
    lib = load_library()
    _dev_check(h)
    D = h.shape[-1] // 2
    out = torch.empty(*h.shape[:-1], D, dtype=h.dtype, device=h.device)
    _check(lib.paid_geglu(h.data_ptr(), out.data_ptr(), h.numel() // (2 * D), D, _dtype_code(h), _stream(h)), "paid_geglu")
    return out


def add_layer_norm(x: torch.Tensor, delta: Optional[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor,
                   eps: float = 1e-5):
    """``paid_add_layer_norm``: returns (x + delta, LayerNorm(x + delta) * weight + bias) over the last dim; with
    ``delta=None`` the first element is ``x`` itself.  x, delta: (..., C) contiguous."""
    lib = load_library()
    _dev_check(x, delta, weight, bias)
    Cdim = x.shape[-1]
    if not x.is_contiguous() or (delta is not None and (not delta.is_contiguous() or delta.shape != x.shape)):
        raise ValueError("add_layer_norm: x and delta must be contiguous and of the same shape")
    h = torch.empty_like(x)
    x_out = torch.empty_like(x) if delta is not None else x
    _check(lib.paid_add_layer_norm(x.data_ptr(), None if delta is None else delta.data_ptr(), weight.data_ptr(),
                                   bias.data_ptr(), None if delta is None else x_out.data_ptr(), h.data_ptr(),
                                   x.numel() // Cdim, Cdim, float(eps), _dtype_code(x), _stream(x)), "paid_add_layer_norm")
    return x_out, h


def group_norm_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, eps: float = 1e-5,
                    silu: bool = False, pre_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``paid_group_norm_nhwc`` on a channels-last (N, C, H, W) feature map; returns a channels-last tensor of the same
    logical shape.  ``pre_bias`` (N, C) is added to x before the statistics (the ResNet time-embedding add)."""
    lib = load_library()
    _dev_check(weight, bias, pre_bias)
    if not x.is_cuda:
        raise RuntimeError("libpaid_attn needs CUDA tensors; there is no CPU path")
    if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("group_norm_nhwc: x must be a 4-D channels_last tensor")
    N, Cdim, H, W = x.shape
    if pre_bias is not None and (pre_bias.shape != (N, Cdim) or not pre_bias.is_contiguous()):
        raise ValueError("group_norm_nhwc: pre_bias must be a contiguous (N, C) tensor")
    need = int(lib.paid_group_norm_workspace_bytes(N, H * W, Cdim, groups))
    if need == 0:
        raise RuntimeError(f"paid_group_norm_nhwc: unsupported geometry C={Cdim} groups={groups}")
    ws = _workspace(x.device, need)          # the per-device scratch shared (stream-ordered) with the attention calls
    y = torch.empty_like(x, memory_format=torch.channels_last)
    _check(lib.paid_group_norm_nhwc(x.data_ptr(), None if pre_bias is None else pre_bias.data_ptr(), weight.data_ptr(),
                                    bias.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel(), N, H * W, Cdim, groups,
                                    float(eps), int(bool(silu)), _dtype_code(x), _stream(x)), "paid_group_norm_nhwc")
    return y


def residual_bias_add(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``paid_residual_bias_add`` on two channels-last (N, C, H, W) feature maps: a + b + bias[None, :, None, None]."""
    lib = load_library()
    _dev_check(bias)
    if a.shape != b.shape or a.dim() != 4 or not a.is_cuda or not b.is_cuda or not (
            a.is_contiguous(memory_format=torch.channels_last) and b.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("residual_bias_add: a and b must be channels_last CUDA tensors of the same 4-D shape")
    N, Cdim, H, W = a.shape
    out = torch.empty_like(a, memory_format=torch.channels_last)
    _check(lib.paid_residual_bias_add(a.data_ptr(), b.data_ptr(), bias.data_ptr(), out.data_ptr(), N * H * W, Cdim,
                                      _dtype_code(a), _stream(a)), "paid_residual_bias_add")
    return out


def attn_core(q, k, v, coef, heads: int, mode: int, fused: bool, scale=None, begin_frame=None, end_frame=None,
              kv_ext=None, flags: int = 0, out=None, accumulate: bool = False, out_scale: float = 1.0,
              out_frame_scale=None, kv_broadcast: bool = False) -> torch.Tensor:
    """Attention on projected tensors through ``paid_attn_core``: q (N,S,C), k/v (N,L,C) (or (1,L,C) with
    kv_broadcast).  ``out`` + ``accumulate`` add ``out_scale * out_frame_scale[n] * attention`` to an existing result."""
    lib = load_library()
    _dev_check(q, k, v, coef, kv_ext, out, out_frame_scale)
    N, S, Cdim = q.shape
    p = PaidCoreParams()
    p.struct_size = C.sizeof(PaidCoreParams)
    p.flags = flags
    p.dtype, p.mode, p.fused = _dtype_code(q), mode, int(bool(fused))
    p.N, p.S, p.L, p.heads, p.head_dim = N, S, k.shape[1], heads, Cdim // heads
    p.scale = float((Cdim // heads) ** -0.5 if scale is None else scale)
    p.begin_frame = 0 if begin_frame is None else begin_frame
    p.end_frame = N - 1 if end_frame is None else end_frame
    if accumulate and out is None:
        raise RuntimeError("accumulate needs an existing out tensor")
    out = torch.empty_like(q) if out is None else out
    p.accumulate, p.out_scale, p.out_frame_scale, p.kv_broadcast = int(accumulate), float(out_scale), _ptr(out_frame_scale), int(kv_broadcast)
    p.q, p.k, p.v, p.kv_ext, p.coef, p.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), _ptr(kv_ext), _ptr(coef), out.data_ptr()
    need = lib.paid_attn_core_workspace_bytes(C.byref(p))
    if need:
        ws = _workspace(q.device, int(need))
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    _check(lib.paid_attn_core(C.byref(p), _stream(q)), "paid_attn_core")
    return out
