"""torch.library custom ops (namespace ``hol``) over the C-ABI CUDA library.

These replace, for CUDA tensors, the chain of ATen calls the reference issues in
``Piecewise.forward`` (reference PolynomialLayers.py:294-336, :639-676 -> Basis.py:492-512)
and ``PiecewisePolynomialConvolution.forward`` (FunctionalConvolution.py:441-447), and the
backward autograd derives for them.  CUDA dispatch only: there is no CPU kernel.

PyTorch is plumbing here (device memory from the caching allocator, the current stream);
the arithmetic happens in csrc/*.cu.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

_NAN = float("nan")
_node_cache: dict = {}


def lobatto_nodes(n: int, length: float = 2.0) -> Tensor:
    """Interpolation nodes as the reference builds them (LagrangePolynomial.py:9-22, :191-195):
    ``(length/2) * -cos(pi*arange(n)/(n-1))`` evaluated by torch in fp32 on the CPU (n == 1 -> [0]).
    Cached host tensor; the kernels receive it as a plain ``const float*``."""
    key = (int(n), float(length))
    t = _node_cache.get(key)
    if t is None:
        if n == 1:
            base = torch.tensor([0.0], dtype=torch.float32, device="cpu")
        else:
            base = -torch.cos(torch.pi * torch.arange(n, device="cpu") / (n - 1))
        t = ((length / 2.0) * base).to(torch.float32).contiguous()
        _node_cache[key] = t
    return t


def _check_inputs(x: Tensor, w: Tensor, what: str):
    if not x.is_cuda or not w.is_cuda:
        raise RuntimeError(f"{what}: tensors must live on a CUDA device (this build has no CPU path)")
    if x.dtype != torch.float32 or w.dtype != torch.float32:
        raise TypeError(f"{what}: float32 tensors required, got {x.dtype} / {w.dtype}")
    if x.device != w.device:
        raise RuntimeError(f"{what}: x on {x.device} but weights on {w.device}")


def _stream(t: Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _per(periodicity: Optional[float]) -> float:
    return _NAN if periodicity is None else float(periodicity)


def _nb(n: int, segments: int, discontinuous: bool) -> int:
    return n * segments if discontinuous else (n - 1) * segments + 1


def set_plan_overrides(mask: int) -> int:
    """Diagnostic (tests, A/B measurements): force kernel families for the calls that follow
    (``_lib.HOL_PLAN_*`` bits, 0 = the planner's own choice); returns the previous mask.  Must not
    change between a forward and its backward."""
    return int(_lib.load().hol_set_plan_overrides(int(mask)))


# ------------------------------------------------------------------------------------------
# segment index (Piecewise.which_segment)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("hol::pw_segment_index", mutates_args=(), device_types="cuda")
def pw_segment_index(x: Tensor, segments: int, length: float, periodicity: Optional[float]) -> Tensor:
    if x.dtype != torch.float32:
        raise TypeError("pw_segment_index: float32 input required")
    lib = _lib.load()
    xc = x.contiguous()
    out = torch.empty(xc.shape, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_segment_index_i64(_ptr(xc), _ptr(out), xc.numel(), segments, length,
                                                _per(periodicity), _stream(x)), "hol_pw_segment_index_i64")
    return out


@pw_segment_index.register_fake
def _(x, segments, length, periodicity):
    return torch.empty(x.shape, dtype=torch.int64, device=x.device)


# ------------------------------------------------------------------------------------------
# fully connected layer
# ------------------------------------------------------------------------------------------
def fc_workspace(x: Tensor, w: Tensor, n: int, segments: int, discontinuous: bool) -> Tensor:
    """Scratch for one layer call (hol_pw_workspace_bytes), from torch's caching allocator."""
    nbytes = _fc_workspace_bytes(x.shape[0], x.shape[1], w.shape[0], n, segments, discontinuous)
    return torch.empty(nbytes, dtype=torch.uint8, device=x.device)


def _fc_workspace_bytes(B: int, I: int, O: int, n: int, segments: int, discontinuous: bool) -> int:
    nbytes = _lib.load().hol_pw_workspace_bytes(B, I, O, n, segments, int(discontinuous))
    if nbytes == 0:
        _lib.check(-2 if (n > _lib.HOL_MAX_N or segments > _lib.HOL_MAX_SEGMENTS) else -1, "hol_pw_workspace_bytes")
    return int(nbytes)


@torch.library.custom_op("hol::pw_forward", mutates_args=(), device_types="cuda")
def pw_forward(x: Tensor, w: Tensor, n: int, segments: int, length: float, discontinuous: bool,
               periodicity: Optional[float]) -> Tuple[Tensor, Tensor]:
    """-> (y[B, O], ws).  `ws` is the scratch workspace of the call (hol_pw_workspace_bytes, from torch's
    caching allocator): it keeps the prepared state (transposed local coordinates, packed weights, Phi
    tiles) that ``hol::pw_backward`` reuses.  A uint8 tensor, so autograd never tracks it."""
    _check_inputs(x, w, "pw_forward")
    if x.dim() != 2:
        raise ValueError(f"pw_forward: x must be [batch, in_features], got {tuple(x.shape)}")
    B, I = x.shape
    O = w.shape[0]
    if tuple(w.shape) != (O, I, _nb(n, segments, discontinuous)):
        raise ValueError(f"pw_forward: weight shape {tuple(w.shape)} does not match "
                         f"[out, {I}, {_nb(n, segments, discontinuous)}]")
    lib = _lib.load()
    xc, wc = x.contiguous(), w.contiguous()
    y = torch.empty((B, O), dtype=torch.float32, device=x.device)
    ws = torch.empty(_fc_workspace_bytes(B, I, O, n, segments, discontinuous), dtype=torch.uint8, device=x.device)
    nodes = lobatto_nodes(n, 2.0)  # FC layers always build LagrangePoly(n) with length 2 (PolynomialLayers.py:234)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_forward_f32(_ptr(xc), _ptr(wc), _ptr(nodes), _ptr(y), B, I, O, n, segments, length,
                                          int(discontinuous), _per(periodicity), _ptr(ws), ws.numel(), _stream(x)),
                   "hol_pw_forward_f32")
    return y, ws


@pw_forward.register_fake
def _(x, w, n, segments, length, discontinuous, periodicity):
    B, I = x.shape
    nbytes = _fc_workspace_bytes(int(B), int(I), int(w.shape[0]), n, segments, discontinuous)
    return (torch.empty((B, w.shape[0]), dtype=torch.float32, device=x.device),
            torch.empty((nbytes,), dtype=torch.uint8, device=x.device))


@torch.library.custom_op("hol::pw_backward", mutates_args=("ws",), device_types="cuda")
def pw_backward(gy: Tensor, x: Tensor, w: Tensor, ws: Tensor, n: int, segments: int, length: float,
                discontinuous: bool, periodicity: Optional[float], need_dx: bool, need_dw: bool,
                ws_prepared: bool, gy_stats_valid: bool = False) -> Tuple[Tensor, Tensor]:
    """-> (dx, dw) (an empty tensor stands for a gradient that was not asked for).  `ws_prepared`: the
    workspace is the forward's; `gy_stats_valid`: it also holds max|gy| of this gy already (second half of
    a backward that was split into a dW call and a dX call)."""
    _check_inputs(x, w, "pw_backward")
    B, I = x.shape
    O = w.shape[0]
    lib = _lib.load()
    gyc, xc, wc = gy.contiguous(), x.contiguous(), w.contiguous()
    dx = torch.empty_like(xc) if need_dx else torch.empty((0,), dtype=torch.float32, device=x.device)
    dw = torch.empty_like(wc) if need_dw else torch.empty((0,), dtype=torch.float32, device=x.device)
    nodes = lobatto_nodes(n, 2.0)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_backward_f32(_ptr(gyc), _ptr(xc), _ptr(wc), _ptr(nodes), _ptr(dx if need_dx else None),
                                           _ptr(dw if need_dw else None), B, I, O, n, segments, length,
                                           int(discontinuous), _per(periodicity), _ptr(ws), ws.numel(),
                                           (_lib.HOL_WS_PREPARED if ws_prepared else 0)
                                           | (_lib.HOL_GY_STATS_VALID if gy_stats_valid else 0), _stream(x)),
                   "hol_pw_backward_f32")
    return dx, dw


@pw_backward.register_fake
def _(gy, x, w, ws, n, segments, length, discontinuous, periodicity, need_dx, need_dw, ws_prepared,
      gy_stats_valid=False):
    e = torch.empty((0,), dtype=torch.float32, device=x.device)
    return (torch.empty_like(x) if need_dx else e, torch.empty_like(w) if need_dw else e)


@torch.library.custom_op("hol::pw_backward_dw_into", mutates_args=("ws", "dw_out"), device_types="cuda")
def pw_backward_dw_into(gy: Tensor, x: Tensor, w: Tensor, ws: Tensor, dw_out: Tensor, n: int, segments: int,
                        length: float, discontinuous: bool, periodicity: Optional[float], ws_prepared: bool,
                        gy_stats_valid: bool = False) -> None:
    """dW only, written (not accumulated) into `dw_out` -- a caller-owned contiguous fp32 tensor of
    w's shape, e.g. the parameter's slice of a flat data-parallel gradient buffer (parallel.py)."""
    _check_inputs(x, w, "pw_backward_dw_into")
    if dw_out.dtype != torch.float32 or not dw_out.is_contiguous() or dw_out.shape != w.shape:
        raise ValueError("pw_backward_dw_into: dw_out must be a contiguous float32 tensor of the weight's shape")
    B, I = x.shape
    O = w.shape[0]
    lib = _lib.load()
    gyc, xc, wc = gy.contiguous(), x.contiguous(), w.contiguous()
    nodes = lobatto_nodes(n, 2.0)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_backward_f32(_ptr(gyc), _ptr(xc), _ptr(wc), _ptr(nodes), _ptr(None), _ptr(dw_out), B, I, O,
                                           n, segments, length, int(discontinuous), _per(periodicity), _ptr(ws),
                                           ws.numel(), (_lib.HOL_WS_PREPARED if ws_prepared else 0)
                                           | (_lib.HOL_GY_STATS_VALID if gy_stats_valid else 0), _stream(x)),
                   "hol_pw_backward_f32")


@pw_backward_dw_into.register_fake
def _(gy, x, w, ws, dw_out, n, segments, length, discontinuous, periodicity, ws_prepared, gy_stats_valid=False):
    return None


def _grad_sink(w: Tensor):
    """A data-parallel gradient buffer that wants dW of this parameter written straight into it
    (parallel.FlatGradients sets ``w._hol_grad_sink``), with the tensor to write into -- or (None, None)."""
    sink = getattr(w, "_hol_grad_sink", None)
    if sink is None:
        return None, None
    out = sink.target(w)
    return (sink, out) if out is not None else (None, None)


def _pw_setup(ctx, inputs, output):
    x, w, n, segments, length, discontinuous, periodicity = inputs
    ctx.save_for_backward(x, w)
    ctx.ws = output[1]       # plain attribute: the backward ops mutate the workspace (no version check wanted)
    ctx.set_materialize_grads(False)
    ctx.w_ref = w            # the caller's tensor object (carries the gradient-sink attribute)
    ctx.args = (n, segments, length, discontinuous, periodicity)


def _pw_bwd(ctx, gy, _gws=None):
    """Backward of hol::pw_forward.  dW first (csrc/abi.cu): when the weight has a gradient sink the
    slice's all-reduce is issued between the dW and the dX launches and overlaps the latter."""
    x, w = ctx.saved_tensors
    if gy is None:  # y did not take part in the loss
        return (None,) * 7
    dx, dw = _layer_backward(ctx, gy, x, w, False)
    return dx, dw, None, None, None, None, None


def _layer_backward(ctx, gy, x, w, gy_stats_valid: bool):
    """dW, then dX, of one layer from the forward's workspace; `gy_stats_valid`: max|gy| is already there."""
    ws = ctx.ws
    need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    dx = dw = None
    sink, out = _grad_sink(ctx.w_ref) if need_dw else (None, None)
    if sink is not None and need_dx and sink.wants_deferred(ctx.w_ref):
        # a larger gradient is still to come (parallel.py, "largest message first"): dX now, dW from finish()
        dx = pw_backward(gy, x, w, ws, *ctx.args, True, False, True, gy_stats_valid)[0]
        args, w_ref = ctx.args, ctx.w_ref
        sink.defer(w_ref, lambda: pw_backward_dw_into(gy, x, w, ws, out, *args, True, True))
    elif sink is not None:
        pw_backward_dw_into(gy, x, w, ws, out, *ctx.args, True, gy_stats_valid)
        sink.ready(ctx.w_ref)
        if need_dx:
            dx = pw_backward(gy, x, w, ws, *ctx.args, True, False, True, True)[0]
    elif need_dx or need_dw:
        rdx, rdw = pw_backward(gy, x, w, ws, *ctx.args, need_dx, need_dw, True, gy_stats_valid)
        dx = rdx if need_dx else None
        dw = rdw if need_dw else None
    return dx, dw


pw_forward.register_autograd(_pw_bwd, setup_context=_pw_setup)


def piecewise_polynomial(x: Tensor, w: Tensor, n: int, segments: int, length: float = 2.0,
                         discontinuous: bool = False, periodicity: Optional[float] = None) -> Tensor:
    """Functional form of the layer: y = Piecewise(n, segments)(x) with weights w[O, I, nb].
    ``hol::pw_forward`` is a functional, differentiable op (torch.library.register_autograd); its second
    output is the scratch workspace the backward reuses."""
    if not x.is_meta:
        _check_inputs(x, w, "piecewise_polynomial")
    return pw_forward(x, w, n, segments, float(length), bool(discontinuous), periodicity)[0]


# ------------------------------------------------------------------------------------------
# a layer with the per-sample normalisation that follows it in the network
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("hol::pw_forward_norm", mutates_args=(), device_types="cuda")
def pw_forward_norm(x: Tensor, w: Tensor, n: int, segments: int, length: float, discontinuous: bool,
                    periodicity: Optional[float], norm_kind: int, eps: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (y_norm[B, O], y_raw[B, O], stats[B, 4], ws): the layer and the normalisation HighOrderMLP puts
    behind it (reference networks.py:235-270) in one call -- on the tensor-core path the forward's last
    kernel writes both rows (hol_pw_forward_norm_f32)."""
    _check_inputs(x, w, "pw_forward_norm")
    if x.dim() != 2:
        raise ValueError(f"pw_forward_norm: x must be [batch, in_features], got {tuple(x.shape)}")
    B, I = x.shape
    O = w.shape[0]
    if tuple(w.shape) != (O, I, _nb(n, segments, discontinuous)):
        raise ValueError(f"pw_forward_norm: weight shape {tuple(w.shape)} does not match "
                         f"[out, {I}, {_nb(n, segments, discontinuous)}]")
    lib = _lib.load()
    xc, wc = x.contiguous(), w.contiguous()
    y_raw = torch.empty((B, O), dtype=torch.float32, device=x.device)
    y_norm = torch.empty((B, O), dtype=torch.float32, device=x.device)
    stats = torch.empty((B, 4), dtype=torch.float32, device=x.device)
    ws = torch.empty(_fc_workspace_bytes(B, I, O, n, segments, discontinuous), dtype=torch.uint8, device=x.device)
    nodes = lobatto_nodes(n, 2.0)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_forward_norm_f32(_ptr(xc), _ptr(wc), _ptr(nodes), _ptr(y_raw), _ptr(y_norm), _ptr(stats),
                                               norm_kind, eps, B, I, O, n, segments, length, int(discontinuous),
                                               _per(periodicity), _ptr(ws), ws.numel(), _stream(x)),
                   "hol_pw_forward_norm_f32")
    return y_norm, y_raw, stats, ws


@pw_forward_norm.register_fake
def _(x, w, n, segments, length, discontinuous, periodicity, norm_kind, eps):
    B, I = x.shape
    O = w.shape[0]
    nbytes = _fc_workspace_bytes(int(B), int(I), int(O), n, segments, discontinuous)
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=x.device)
    return f(B, O), f(B, O), f(B, 4), torch.empty((nbytes,), dtype=torch.uint8, device=x.device)


@torch.library.custom_op("hol::pw_norm_backward", mutates_args=("ws",), device_types="cuda")
def pw_norm_backward(gy_norm: Tensor, y_raw: Tensor, stats: Tensor, ws: Tensor, norm_kind: int, in_features: int,
                     n: int, segments: int, discontinuous: bool) -> Tensor:
    """Gradient w.r.t. the layer's raw output from the gradient w.r.t. the normalised one; the same sweep
    leaves max|gy_raw| in the layer's workspace for its tensor-core backward passes (hol_pw_norm_backward_f32)."""
    lib = _lib.load()
    g = gy_norm.contiguous()
    B, O = y_raw.shape
    out = torch.empty_like(y_raw)
    with torch.cuda.device(y_raw.device):
        _lib.check(lib.hol_pw_norm_backward_f32(norm_kind, _ptr(g), _ptr(y_raw), _ptr(stats), _ptr(out), B, in_features, O,
                                                n, segments, int(discontinuous), _ptr(ws), ws.numel(), _stream(y_raw)),
                   "hol_pw_norm_backward_f32")
    return out


@pw_norm_backward.register_fake
def _(gy_norm, y_raw, stats, ws, norm_kind, in_features, n, segments, discontinuous):
    return torch.empty_like(y_raw)


def _pwn_setup(ctx, inputs, output):
    x, w, n, segments, length, discontinuous, periodicity, norm_kind, _eps = inputs
    ctx.save_for_backward(x, w, output[1], output[2])
    ctx.ws = output[3]
    ctx.set_materialize_grads(False)
    ctx.w_ref = w
    ctx.args = (n, segments, length, discontinuous, periodicity)
    ctx.norm_kind = norm_kind


def _pwn_bwd(ctx, g_norm, g_raw=None, _gstats=None, _gws=None):
    x, w, y_raw, stats = ctx.saved_tensors
    if g_norm is None and g_raw is None:
        return (None,) * 9
    n, segments, _length, discontinuous, _per_ = ctx.args
    stats_valid = False
    gy = g_raw
    if g_norm is not None:
        gy = pw_norm_backward(g_norm, y_raw, stats, ctx.ws, ctx.norm_kind, x.shape[1], n, segments, discontinuous)
        stats_valid = g_raw is None
        if g_raw is not None:  # the raw output was used as well: its gradient adds, the statistic is stale
            gy = gy + g_raw
    dx, dw = _layer_backward(ctx, gy, x, w, stats_valid)
    return dx, dw, None, None, None, None, None, None, None


pw_forward_norm.register_autograd(_pwn_bwd, setup_context=_pwn_setup)

NORM_KINDS = {"max_abs": 0, "max_center": 1, "l2": 2}


def piecewise_polynomial_normalized(x: Tensor, w: Tensor, n: int, segments: int, norm: str, eps: float = 1e-6,
                                    length: float = 2.0, discontinuous: bool = False,
                                    periodicity: Optional[float] = None) -> Tensor:
    """normalisation(Piecewise(n, segments)(x)) for norm in {"max_abs", "max_center", "l2"} as ONE layer call
    each way: bit-identical to ``normalize_rows(piecewise_polynomial(...), norm, eps)``."""
    if not x.is_meta:
        _check_inputs(x, w, "piecewise_polynomial_normalized")
    return pw_forward_norm(x, w, n, segments, float(length), bool(discontinuous), periodicity, NORM_KINDS[norm],
                           float(eps))[0]


# ------------------------------------------------------------------------------------------
# unfolded convolution (rows = B*Ho*Wo, inputs = C*Kh*Kw); 1-d convolutions are H = Kh = 1
# ------------------------------------------------------------------------------------------
def conv_out_size(h: int, k: int, stride: int, padding: int, dilation: int) -> int:
    return (h + 2 * padding - dilation * (k - 1) - 1) // stride + 1


def _pair(v, name: str) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        if len(v) == 1:
            return int(v[0]), int(v[0])
        if len(v) != 2:
            raise ValueError(f"{name}={v}: one value or a pair expected")
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _geom(C: int, H: int, W: int, kernel, stride, padding, dilation):
    """{C, H, W, Kh, Kw, sh, sw, ph, pw, dh, dw} for the hol_pw_conv_* entry points + the output size."""
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = (_pair(kernel, "kernel_size"), _pair(stride, "stride"),
                                              _pair(padding, "padding"), _pair(dilation, "dilation"))
    arr = (ctypes.c_int32 * 11)(C, H, W, kh, kw, sh, sw, ph, pw, dh, dw)
    return arr, conv_out_size(H, kh, sh, ph, dh), conv_out_size(W, kw, sw, pw, dw)


def _conv_workspace_bytes(B, C, H, W, O, kernel, stride, padding, dilation, n, segments, discontinuous) -> int:
    g, _, _ = _geom(C, H, W, kernel, stride, padding, dilation)
    nbytes = _lib.load().hol_pw_conv_workspace_bytes(B, ctypes.cast(g, ctypes.c_void_p), O, n, segments, int(discontinuous))
    if nbytes == 0:
        _lib.check(-2, "hol_pw_conv_workspace_bytes")
    return int(nbytes)


def conv_workspace(x: Tensor, w_fc: Tensor, n: int, segments: int, discontinuous: bool, kernel_size, stride, padding,
                   dilation) -> Tensor:
    B, C, H, W = x.shape
    nbytes = _conv_workspace_bytes(B, C, H, W, w_fc.shape[0], kernel_size, stride, padding, dilation, n, segments,
                                   discontinuous)
    return torch.empty(nbytes, dtype=torch.uint8, device=x.device)


@torch.library.custom_op("hol::pw_conv2d_rows", mutates_args=(), device_types="cuda")
def pw_conv2d_rows(x: Tensor, w_fc: Tensor, n: int, segments: int, length: float, discontinuous: bool,
                   periodicity: Optional[float], kernel_size: Sequence[int], stride: Sequence[int],
                   padding: Sequence[int], dilation: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """x[B,C,H,W], w_fc[O, C*Kh*Kw, nb] -> (y_rows[B*Ho*Wo, O], ws): the workspace of the call, reused by
    ``hol::pw_conv2d_rows_backward``.  kernel_size / stride / padding / dilation are (h, w) pairs."""
    _check_inputs(x, w_fc, "pw_conv2d_rows")
    if x.dim() != 4:
        raise ValueError(f"pw_conv2d_rows: x must be [B, C, H, W], got {tuple(x.shape)}")
    B, C, H, W = x.shape
    O = w_fc.shape[0]
    kh, kw = _pair(kernel_size, "kernel_size")
    if tuple(w_fc.shape) != (O, C * kh * kw, _nb(n, segments, discontinuous)):
        raise ValueError(f"pw_conv2d_rows: weight shape {tuple(w_fc.shape)} does not match the layer")
    g, Ho, Wo = _geom(C, H, W, kernel_size, stride, padding, dilation)
    if Ho <= 0 or Wo <= 0:
        raise ValueError("pw_conv2d_rows: kernel does not fit the input")
    lib = _lib.load()
    xc, wc = x.contiguous(), w_fc.contiguous()
    ws = torch.empty(_conv_workspace_bytes(B, C, H, W, O, kernel_size, stride, padding, dilation, n, segments,
                                           discontinuous), dtype=torch.uint8, device=x.device)
    y = torch.empty((B * Ho * Wo, O), dtype=torch.float32, device=x.device)
    nodes = lobatto_nodes(n, length)  # conv builds its basis with the layer's length (LagrangePolynomial.py:242-246)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_conv_forward_f32(_ptr(xc), _ptr(wc), _ptr(nodes), _ptr(y), B, ctypes.cast(g, ctypes.c_void_p),
                                               O, n, segments, length, int(discontinuous), _per(periodicity), _ptr(ws),
                                               ws.numel(), _stream(x)), "hol_pw_conv_forward_f32")
    return y, ws


@pw_conv2d_rows.register_fake
def _(x, w_fc, n, segments, length, discontinuous, periodicity, kernel_size, stride, padding, dilation):
    B, C, H, W = x.shape
    _, Ho, Wo = _geom(int(C), int(H), int(W), kernel_size, stride, padding, dilation)
    nbytes = _conv_workspace_bytes(int(B), int(C), int(H), int(W), int(w_fc.shape[0]), kernel_size, stride, padding,
                                   dilation, n, segments, discontinuous)
    return (torch.empty((B * Ho * Wo, w_fc.shape[0]), dtype=torch.float32, device=x.device),
            torch.empty((nbytes,), dtype=torch.uint8, device=x.device))


@torch.library.custom_op("hol::pw_conv2d_rows_backward", mutates_args=("ws",), device_types="cuda")
def pw_conv2d_rows_backward(gy: Tensor, x: Tensor, w_fc: Tensor, ws: Tensor, n: int, segments: int, length: float,
                            discontinuous: bool, periodicity: Optional[float], kernel_size: Sequence[int],
                            stride: Sequence[int], padding: Sequence[int], dilation: Sequence[int], need_dx: bool,
                            need_dw: bool, ws_prepared: bool) -> Tuple[Tensor, Tensor]:
    _check_inputs(x, w_fc, "pw_conv2d_rows_backward")
    B, C, H, W = x.shape
    O = w_fc.shape[0]
    lib = _lib.load()
    gyc, xc, wc = gy.contiguous(), x.contiguous(), w_fc.contiguous()
    dx = torch.empty_like(xc) if need_dx else torch.empty((0,), dtype=torch.float32, device=x.device)
    dw = torch.empty_like(wc) if need_dw else torch.empty((0,), dtype=torch.float32, device=x.device)
    nodes = lobatto_nodes(n, length)
    g, _, _ = _geom(C, H, W, kernel_size, stride, padding, dilation)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_conv_backward_f32(_ptr(gyc), _ptr(xc), _ptr(wc), _ptr(nodes),
                                                _ptr(dx if need_dx else None), _ptr(dw if need_dw else None), B,
                                                ctypes.cast(g, ctypes.c_void_p), O, n, segments, length,
                                                int(discontinuous), _per(periodicity), _ptr(ws), ws.numel(),
                                                _lib.HOL_WS_PREPARED if ws_prepared else 0, _stream(x)),
                   "hol_pw_conv_backward_f32")
    return dx, dw


@pw_conv2d_rows_backward.register_fake
def _(gy, x, w_fc, ws, n, segments, length, discontinuous, periodicity, kernel_size, stride, padding, dilation,
      need_dx, need_dw, ws_prepared):
    e = torch.empty((0,), dtype=torch.float32, device=x.device)
    return (torch.empty_like(x) if need_dx else e, torch.empty_like(w_fc) if need_dw else e)


def _conv_setup(ctx, inputs, output):
    x, w_fc = inputs[:2]
    ctx.save_for_backward(x, w_fc)
    ctx.ws = output[1]
    ctx.args = tuple(inputs[2:])
    ctx.set_materialize_grads(False)


def _conv_bwd(ctx, gy, _gws=None):
    x, w_fc = ctx.saved_tensors
    ws = ctx.ws
    need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    dx = dw = None
    if gy is not None and (need_dx or need_dw):
        rdx, rdw = pw_conv2d_rows_backward(gy, x, w_fc, ws, *ctx.args, need_dx, need_dw, True)
        dx = rdx if need_dx else None
        dw = rdw if need_dw else None
    return (dx, dw) + (None,) * 9


pw_conv2d_rows.register_autograd(_conv_bwd, setup_context=_conv_setup)


def piecewise_polynomial_conv2d(x: Tensor, weight: Tensor, n: int, segments: int, kernel_size,
                                length: float = 2.0, discontinuous: bool = False,
                                periodicity: Optional[float] = None, stride=1, padding=0,
                                dilation=1, rescale: float = 1.0) -> Tensor:
    """conv.weight[O, C*nb, Kh, Kw] layout in, NCHW out: the unfolded path of SURVEY 3.3.  kernel_size,
    stride, padding and dilation are ints or (h, w) pairs."""
    B, C, H, W = x.shape
    O = weight.shape[0]
    kh, kw = _pair(kernel_size, "kernel_size")
    nb = _nb(n, segments, discontinuous)
    w_fc = weight.view(O, C, nb, kh, kw).permute(0, 1, 3, 4, 2).reshape(O, C * kh * kw, nb)
    _check_inputs(x, w_fc, "piecewise_polynomial_conv2d")
    ks, st, pd, dl = (list(_pair(v, k)) for v, k in ((kernel_size, "kernel_size"), (stride, "stride"),
                                                     (padding, "padding"), (dilation, "dilation")))
    y_rows = pw_conv2d_rows(x, w_fc, n, segments, float(length), bool(discontinuous), periodicity, ks, st, pd, dl)[0]
    Ho, Wo = conv_out_size(H, kh, st[0], pd[0], dl[0]), conv_out_size(W, kw, st[1], pd[1], dl[1])
    y = y_rows.view(B, Ho, Wo, O).permute(0, 3, 1, 2)
    if rescale != 1.0:
        y = y * rescale
    return y


def piecewise_polynomial_conv1d(x: Tensor, weight: Tensor, n: int, segments: int, kernel_size: int,
                                length: float = 2.0, discontinuous: bool = False,
                                periodicity: Optional[float] = None, stride: int = 1, padding: int = 0,
                                dilation: int = 1, rescale: float = 1.0) -> Tensor:
    """PiecewisePolynomialConvolution1d (reference FunctionalConvolution.py:493-519, :622-650): the 2-d path
    with H = Kh = 1 -- the prep kernel unfolds the taps itself and flags the zero-padded ones invalid, so
    padding, stride and dilation are all handled by the CUDA library.  x[B, C, L], conv.weight[O, C*nb, K]."""
    if x.dim() != 3:
        raise ValueError(f"piecewise_polynomial_conv1d: x must be [B, C, L], got {tuple(x.shape)}")
    y = piecewise_polynomial_conv2d(x.unsqueeze(2), weight.unsqueeze(2), n, segments, (1, int(kernel_size)), length,
                                    discontinuous, periodicity, (1, int(stride)), (0, int(padding)), (1, int(dilation)),
                                    rescale)
    return y.squeeze(2)


# ------------------------------------------------------------------------------------------
# conv-transpose: the piecewise layer per input pixel, then a fold of the taps onto the output grid
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("hol::fold_rows", mutates_args=(), device_types="cuda")
def fold_rows(z_rows: Tensor, B: int, O: int, H: int, W: int, Ho: int, Wo: int, K: int, stride: int, padding: int,
              dilation: int) -> Tensor:
    """z_rows[B*H*W, O*K*K] -> y[B, O, Ho, Wo] (hol_fold_rows_f32)."""
    if z_rows.dtype != torch.float32 or tuple(z_rows.shape) != (B * H * W, O * K * K):
        raise ValueError(f"fold_rows: float32 [{B * H * W}, {O * K * K}] expected, got {z_rows.dtype} {tuple(z_rows.shape)}")
    zc = z_rows.contiguous()
    y = torch.empty((B, O, Ho, Wo), dtype=torch.float32, device=z_rows.device)
    with torch.cuda.device(z_rows.device):
        _lib.check(_lib.load().hol_fold_rows_f32(_ptr(zc), _ptr(y), B, O, H, W, Ho, Wo, K, stride, padding, dilation,
                                                 _stream(z_rows)), "hol_fold_rows_f32")
    return y


@fold_rows.register_fake
def _(z_rows, B, O, H, W, Ho, Wo, K, stride, padding, dilation):
    return torch.empty((B, O, Ho, Wo), dtype=torch.float32, device=z_rows.device)


@torch.library.custom_op("hol::unfold_rows", mutates_args=(), device_types="cuda")
def unfold_rows(gy: Tensor, B: int, O: int, H: int, W: int, Ho: int, Wo: int, K: int, stride: int, padding: int,
                dilation: int) -> Tensor:
    """Adjoint of fold_rows: gy[B, O, Ho, Wo] -> gz_rows[B*H*W, O*K*K] (hol_unfold_rows_f32)."""
    gc = gy.contiguous()
    gz = torch.empty((B * H * W, O * K * K), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        _lib.check(_lib.load().hol_unfold_rows_f32(_ptr(gc), _ptr(gz), B, O, H, W, Ho, Wo, K, stride, padding, dilation,
                                                   _stream(gy)), "hol_unfold_rows_f32")
    return gz


@unfold_rows.register_fake
def _(gy, B, O, H, W, Ho, Wo, K, stride, padding, dilation):
    return torch.empty((B * H * W, O * K * K), dtype=torch.float32, device=gy.device)


def _fold_setup(ctx, inputs, output):
    ctx.args = tuple(inputs[1:])


def _fold_bwd(ctx, gy):
    return (unfold_rows(gy, *ctx.args),) + (None,) * 10


def _unfold_setup(ctx, inputs, output):
    ctx.args = tuple(inputs[1:])


def _unfold_bwd(ctx, gz):
    return (fold_rows(gz, *ctx.args),) + (None,) * 10


fold_rows.register_autograd(_fold_bwd, setup_context=_fold_setup)
unfold_rows.register_autograd(_unfold_bwd, setup_context=_unfold_setup)


def piecewise_polynomial_conv_transpose2d(x: Tensor, weight: Tensor, n: int, segments: int, kernel_size: int,
                                          length: float = 2.0, discontinuous: bool = False,
                                          periodicity: Optional[float] = None, stride: int = 1, padding: int = 0,
                                          output_padding: int = 0, dilation: int = 1, rescale: float = 1.0) -> Tensor:
    """PiecewisePolynomialConvolutionTranspose2d (reference FunctionalConvolutionTranspose.py:75-160):
    x[B, C, H, W], conv.weight[C*nb, O, K, K] (nn.ConvTranspose2d layout) -> y[B, O, Ho, Wo]."""
    B, C, H, W = x.shape
    K = int(kernel_size)
    nb = _nb(n, segments, discontinuous)
    O = weight.shape[1]
    if tuple(weight.shape) != (C * nb, O, K, K):
        raise ValueError(f"piecewise_polynomial_conv_transpose2d: weight shape {tuple(weight.shape)} does not match "
                         f"[{C * nb}, out_channels, {K}, {K}]")
    # per input pixel: inputs = the C channels, outputs = (o, kh, kw)
    w_fc = weight.view(C, nb, O, K, K).permute(2, 3, 4, 0, 1).reshape(O * K * K, C, nb)
    _check_inputs(x, w_fc, "piecewise_polynomial_conv_transpose2d")
    z = pw_conv2d_rows(x, w_fc, n, segments, float(length), bool(discontinuous), periodicity, [1, 1], [1, 1], [0, 0],
                       [1, 1])[0]
    Ho = (H - 1) * stride - 2 * padding + dilation * (K - 1) + output_padding + 1
    Wo = (W - 1) * stride - 2 * padding + dilation * (K - 1) + output_padding + 1
    y = fold_rows(z, B, O, H, W, Ho, Wo, K, stride, padding, dilation)
    return y * rescale if rescale != 1.0 else y


# ------------------------------------------------------------------------------------------
# per-sample normalisations between layers: one fused kernel each way
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("hol::row_norm", mutates_args=(), device_types="cuda")
def row_norm(x: Tensor, kind: int, eps: float) -> Tuple[Tensor, Tensor]:
    """x[B, W] -> (y[B, W], stats[B, 4]); kind 0 max_abs, 1 max_center, 2 l2 (reference utils.py:8-58)."""
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError(f"row_norm: float32 [batch, width] input required, got {x.dtype} {tuple(x.shape)}")
    lib = _lib.load()
    xc = x.contiguous()
    y = torch.empty_like(xc)
    stats = torch.empty((xc.shape[0], 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_row_norm_forward_f32(kind, _ptr(xc), _ptr(y), _ptr(stats), xc.shape[0], xc.shape[1], eps,
                                                _stream(x)), "hol_row_norm_forward_f32")
    return y, stats


@row_norm.register_fake
def _(x, kind, eps):
    return torch.empty_like(x), torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)


@torch.library.custom_op("hol::row_norm_backward", mutates_args=(), device_types="cuda")
def row_norm_backward(gy: Tensor, x: Tensor, stats: Tensor, kind: int) -> Tensor:
    lib = _lib.load()
    gyc, xc = gy.contiguous(), x.contiguous()
    dx = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_row_norm_backward_f32(kind, _ptr(gyc), _ptr(xc), _ptr(stats), _ptr(dx), xc.shape[0],
                                                 xc.shape[1], _stream(x)), "hol_row_norm_backward_f32")
    return dx


@row_norm_backward.register_fake
def _(gy, x, stats, kind):
    return torch.empty_like(x)


def _row_norm_setup(ctx, inputs, output):
    x, kind, _ = inputs
    ctx.save_for_backward(x, output[1])
    ctx.kind = kind


def _row_norm_bwd(ctx, gy, _gstats):
    x, stats = ctx.saved_tensors
    return row_norm_backward(gy, x, stats, ctx.kind), None, None


row_norm.register_autograd(_row_norm_bwd, setup_context=_row_norm_setup)


def normalize_rows(x: Tensor, kind: str, eps: float = 1e-6) -> Tensor:
    """Fused per-sample normalisation over dim 1 of a CUDA fp32 [B, W] tensor."""
    return row_norm(x, NORM_KINDS[kind], float(eps))[0]


def launch_count(kind: int, B: int, I: int, O: int, n: int, segments: int, discontinuous: bool, want_dx: bool = True,
                 want_dw: bool = True, prepared: bool = True, fused_norm: bool = False) -> int:
    """Kernels of this library launched by one forward (kind 0) / backward (kind 1) call of a layer;
    `fused_norm`: the layer runs with the normalisation folded in (pw_forward_norm and its backward)."""
    flags = (_lib.HOL_WS_PREPARED if prepared else 0) | (_lib.HOL_COUNT_FUSED_NORM | _lib.HOL_GY_STATS_VALID
                                                        if fused_norm else 0)
    return _lib.load().hol_pw_launch_count(kind, B, I, O, n, segments, int(discontinuous), int(want_dx), int(want_dw),
                                           flags)


def plan_flags(B: int, I: int, O: int, n: int, segments: int, discontinuous: bool) -> dict:
    """Which kernels the planner picks for a layer shape (hol_pw_plan_flags)."""
    f = _lib.load().hol_pw_plan_flags(B, I, O, n, segments, int(discontinuous))
    if f < 0:
        raise ValueError("plan_flags: shape outside the envelope")
    return {"tc_forward": bool(f & 1), "tc_dx": bool(f & 2), "tc_dw": bool(f & 4), "dense_dw": bool(f & 8),
            "shared_phi": bool(f & 16), "rotated_gather": bool(f & 32), "tc_dw_ksplit": (f >> 8) & 255,
            "tc_fwd_ksplit": (f >> 16) & 255}


def uses_tensor_cores(B: int, I: int, O: int, n: int, segments: int, discontinuous: bool) -> bool:
    """True when the layer shape is planned onto the tcgen05 path (csrc/tc_gemm.cu: tc_plan)."""
    return bool(_lib.load().hol_pw_plan_flags(B, I, O, n, segments, int(discontinuous)) & 7)


def profile_layer(x: Tensor, w: Tensor, gy: Tensor, n: int, segments: int, length: float = 2.0,
                  discontinuous: bool = False, periodicity: Optional[float] = None) -> dict:
    """Diagnostic: CUDA-event durations (ms) of the stages of one forward+backward of a layer
    (hol_pw_profile_f32; synchronises).  Keys: prep, pack, forward, dx, bucket_sort, dw and, for a
    tensor-core layer, gemm_fwd / gemm_dx / gemm_dw (the tcgen05 GEMM launches alone)."""
    _check_inputs(x, w, "profile_layer")
    B, I = x.shape
    O = w.shape[0]
    lib = _lib.load()
    ws = fc_workspace(x, w, n, segments, discontinuous)
    y = torch.empty((B, O), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    ms = (ctypes.c_float * 9)()
    nodes = lobatto_nodes(n, 2.0)
    with torch.cuda.device(x.device):
        _lib.check(lib.hol_pw_profile_f32(_ptr(gy.contiguous()), _ptr(x.contiguous()), _ptr(w.contiguous()),
                                          _ptr(nodes), _ptr(y), _ptr(dx), _ptr(dw), B, I, O, n, segments, length,
                                          int(discontinuous), _per(periodicity), _ptr(ws), ws.numel(),
                                          ctypes.cast(ms, ctypes.c_void_p), _stream(x)), "hol_pw_profile_f32")
    return dict(zip(("prep", "pack", "forward", "dx", "bucket_sort", "dw", "gemm_fwd", "gemm_dx", "gemm_dw"),
                    (float(v) for v in ms)))
