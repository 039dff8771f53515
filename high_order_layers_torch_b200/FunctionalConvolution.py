"""PiecewisePolynomialConvolution2d on the unfolded fully-connected kernels.

Reference FunctionalConvolution.py: PiecewisePolynomialConvolution (:392-447),
PiecewisePolynomialConvolution2d (:450-490), conv_wrapper (:53-103),
constant_random_initialization_conv2d (:17-50).  The reference expands the input to C*V dense
channels and calls nn.Conv2d; here the conv IS the piecewise layer applied to im2col rows
(SURVEY 3.3), evaluated by the same CUDA kernels without materialising either tensor.  The
parameter keeps the reference's name and layout: ``conv.weight [O, C*V, K, K]``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


def _as_int(v, name: str) -> int:
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise ValueError(f"{name}={v}: only square (equal) values are supported")
        v = v[0]
    return int(v)


class _PiecewiseConvolution2dBase(nn.Module):
    _discontinuous = False

    def __init__(self, n: int, segments: int, in_channels: int, out_channels: int, kernel_size: int,
                 length: float = 2.0, rescale_output: bool = False, periodicity: Optional[float] = None,
                 weight_magnitude: float = 1.0, device: str = "cpu", *args, stride: int = 1, padding: int = 0,
                 dilation: int = 1, groups: int = 1, padding_mode: str = "zeros", **kwargs):
        super().__init__()
        name = type(self).__name__
        if groups != 1:
            raise ValueError(f"{name}: groups != 1 is not supported by the CUDA path")
        if padding_mode != "zeros":
            raise ValueError(f"{name}: only padding_mode='zeros' is supported")
        if isinstance(padding, str):
            raise ValueError(f"{name}: string padding modes are not supported")
        self._n = n
        self._segments = segments
        self._length = length
        # the reference takes an int kernel_size (its initialiser multiplies it); stride / padding / dilation
        # go to nn.Conv2d unchanged, so (h, w) pairs are accepted for them
        self._kernel_size = _as_int(kernel_size, "kernel_size")
        self._stride = ops._pair(stride, "stride")
        self._padding = ops._pair(padding, "padding")
        self._dilation = ops._pair(dilation, "dilation")
        self._in_channels = in_channels
        self._nodes = n * segments if self._discontinuous else (n - 1) * segments + 1
        self._channels = self._nodes * in_channels
        self._out_channels = out_channels
        self.periodicity = periodicity
        k = self._kernel_size
        # parameter container with the reference's key (conv.weight); its own forward is never
        # used.  The three steps below consume the RNG exactly like the reference (conv_wrapper
        # then constant_random_initialization_conv2d), so equal seeds give equal weights.
        self.conv = nn.Conv2d(in_channels=self._channels, out_channels=out_channels, kernel_size=k,
                              stride=self._stride, padding=self._padding, dilation=self._dilation, groups=1,
                              bias=False, padding_mode="zeros")
        in_features = self._channels * k * k
        if rescale_output is False:
            self.conv.weight.data.uniform_(-weight_magnitude / in_features, weight_magnitude / in_features)
        elif rescale_output is True:
            self.conv.weight.data.uniform_(-weight_magnitude, weight_magnitude)
        self._total_in = in_channels * k * k
        self._rescale = 1.0 / self._total_in if rescale_output is True else 1.0
        # one random constant per (out, in, kh, kw) link, identical over the nodes
        link_features = k * k * self._channels // self._nodes
        values = torch.rand(out_channels, in_channels, k, k, device=device) * weight_magnitude / link_features
        with torch.no_grad():
            self.conv.weight.copy_(values.unsqueeze(2).expand(out_channels, in_channels, self._nodes, k, k)
                                   .reshape(out_channels, self._channels, k, k))
        if str(device) != "cpu":
            self.conv.to(device)

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError(
                f"{type(self).__name__}: input is on {x.device}; this build runs on CUDA (sm_100a) only "
                "and has no CPU fallback")
        return ops.piecewise_polynomial_conv2d(x, self.conv.weight, self._n, self._segments, self._kernel_size,
                                               length=float(self._length), discontinuous=self._discontinuous,
                                               periodicity=self.periodicity, stride=self._stride,
                                               padding=self._padding, dilation=self._dilation,
                                               rescale=self._rescale)


class PiecewisePolynomialConvolution2d(_PiecewiseConvolution2dBase):
    """Continuous piecewise polynomial 2-d convolution (reference FunctionalConvolution.py:450-490)."""
    _discontinuous = False


class PiecewiseDiscontinuousPolynomialConvolution2d(_PiecewiseConvolution2dBase):
    """Discontinuous variant: every segment owns its n nodes, C*n*segments expanded channels
    (reference FunctionalConvolution.py:522-619; expansion Basis.py:220-296)."""
    _discontinuous = True


class _PiecewiseConvolution1dBase(nn.Module):
    """1-d variants (reference FunctionalConvolution.py:493-519, :622-650): Expansion1d + nn.Conv1d in
    the reference, here the convolution kernels with H = Kh = 1 (stride, zero padding and dilation handled
    by the CUDA library's prep stage).  Parameter ``conv.weight
    [O, C*V, K]`` initialised by the reference's conv_wrapper recipe (uniform in +-wm/(C*V*K*K); the
    reference's 1-d classes do not apply the constant-per-link initialisation of the 2-d ones)."""
    _discontinuous = False

    def __init__(self, n: int, segments: int, in_channels: int, kernel_size: int, length: float = 2.0,
                 rescale_output: bool = False, periodicity: Optional[float] = None, *args, out_channels: int,
                 stride: int = 1, padding: int = 0, dilation: int = 1, groups: int = 1,
                 padding_mode: str = "zeros", weight_magnitude: float = 1.0, device: str = "cpu", **kwargs):
        super().__init__()
        name = type(self).__name__
        if groups != 1 or padding_mode != "zeros" or isinstance(padding, str):
            raise ValueError(f"{name}: groups != 1 / non-zero padding modes are not supported by the CUDA path")
        self._n, self._segments, self._length = n, segments, length
        self._kernel_size = _as_int(kernel_size, "kernel_size")
        self._stride, self._dilation = _as_int(stride, "stride"), _as_int(dilation, "dilation")
        self._padding = _as_int(padding, "padding")
        self._nodes = n * segments if self._discontinuous else (n - 1) * segments + 1
        self._channels = self._nodes * in_channels
        self._out_channels = out_channels
        self.periodicity = periodicity
        k = self._kernel_size
        self.conv = nn.Conv1d(in_channels=self._channels, out_channels=out_channels, kernel_size=k,
                              stride=self._stride, padding=self._padding, dilation=self._dilation, groups=1,
                              bias=False)
        in_features = self._channels * k * k          # the reference's conv_wrapper uses k*k for every rank
        if rescale_output is False:
            self.conv.weight.data.uniform_(-weight_magnitude / in_features, weight_magnitude / in_features)
        elif rescale_output is True:
            self.conv.weight.data.uniform_(-weight_magnitude, weight_magnitude)
        self._total_in = in_channels * k * k
        self._rescale = 1.0 / self._total_in if rescale_output is True else 1.0
        if str(device) != "cpu":
            self.conv.to(device)

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: input is on {x.device}; this build runs on CUDA (sm_100a) "
                               "only and has no CPU fallback")
        return ops.piecewise_polynomial_conv1d(x, self.conv.weight, self._n, self._segments, self._kernel_size,
                                               length=float(self._length), discontinuous=self._discontinuous,
                                               periodicity=self.periodicity, stride=self._stride,
                                               padding=self._padding, dilation=self._dilation, rescale=self._rescale)


class PiecewisePolynomialConvolution1d(_PiecewiseConvolution1dBase):
    _discontinuous = False


class PiecewiseDiscontinuousPolynomialConvolution1d(_PiecewiseConvolution1dBase):
    _discontinuous = True
