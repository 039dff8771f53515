"""PiecewisePolynomialConvolutionTranspose2d on the sm_100a kernels.

Reference FunctionalConvolutionTranspose.py: conv_transpose_wrapper (:17-72),
PiecewisePolynomialConvolutionTranspose (:75-134), PiecewisePolynomialConvolutionTranspose2d
(:137-160).  The reference expands the input to C*V one-hot-like channels and calls
nn.ConvTranspose2d; here the layer is the piecewise polynomial evaluated per input pixel
(rows = B*H*W, inputs = the C channels, outputs = (o, kh, kw): the same CUDA kernels as the fully
connected layer, through the convolution entry point with a 1x1 kernel) followed by a fold of the
K*K taps onto the output grid (csrc/prep.cu: pw_fold_rows_kernel).  The parameter keeps the
reference's name and layout: ``conv.weight [C*V, O, K, K]``.
"""
from __future__ import annotations

from typing import Optional

import torch.nn as nn
from torch import Tensor

from . import ops
from .FunctionalConvolution import _as_int


class PiecewisePolynomialConvolutionTranspose2d(nn.Module):
    def __init__(self, n: int, segments: int, in_channels: int, kernel_size: int, length: float = 2.0,
                 rescale_output: bool = False, periodicity: Optional[float] = None, *args, out_channels: int,
                 stride: int = 1, padding: int = 0, output_padding: int = 0, dilation: int = 1, groups: int = 1,
                 padding_mode: str = "zeros", weight_magnitude: float = 1.0, device: str = "cpu", **kwargs):
        super().__init__()
        if groups != 1 or padding_mode != "zeros":
            raise ValueError("PiecewisePolynomialConvolutionTranspose2d: groups != 1 / non-zero padding modes are "
                             "not supported by the CUDA path")
        self._n, self._segments, self._length = n, segments, length
        self._kernel_size = _as_int(kernel_size, "kernel_size")
        self._stride = _as_int(stride, "stride")
        self._padding = _as_int(padding, "padding")
        self._output_padding = _as_int(output_padding, "output_padding")
        self._dilation = _as_int(dilation, "dilation")
        self._channels = ((n - 1) * segments + 1) * in_channels
        self._out_channels = out_channels
        self.periodicity = periodicity
        k = self._kernel_size
        # parameter container with the reference's key; same RNG consumption as conv_transpose_wrapper
        self.conv = nn.ConvTranspose2d(in_channels=self._channels, out_channels=out_channels, kernel_size=k,
                                       stride=self._stride, padding=self._padding,
                                       output_padding=self._output_padding, dilation=self._dilation, groups=1,
                                       bias=False, padding_mode="zeros")
        in_features = self._channels * k * k
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
            raise RuntimeError("PiecewisePolynomialConvolutionTranspose2d: input is on "
                               f"{x.device}; this build runs on CUDA (sm_100a) only and has no CPU fallback")
        return ops.piecewise_polynomial_conv_transpose2d(
            x, self.conv.weight, self._n, self._segments, self._kernel_size, length=float(self._length),
            discontinuous=False, periodicity=self.periodicity, stride=self._stride, padding=self._padding,
            output_padding=self._output_padding, dilation=self._dilation, rescale=self._rescale)
