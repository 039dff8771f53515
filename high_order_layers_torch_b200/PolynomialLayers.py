"""Piecewise Lagrange-polynomial fully connected layers on the sm_100a kernels.

Same constructor arguments, attributes, parameter name/shape (``w: [out, in, nb]``), initial
values for a given RNG seed and helper methods as reference PolynomialLayers.py
(Piecewise :196-357, PiecewisePolynomial :532-555, PiecewiseDiscontinuous :584-703,
PiecewiseDiscontinuousPolynomial :706-729); ``forward`` dispatches to the ``hol::pw_forward``
custom op instead of building the gathered [B, I, O, n] weight tensor.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .LagrangePolynomial import LagrangePoly
from .utils import make_periodic


def constant_random_initialization(w: Tensor, out_features: int, in_features: int, weight_magnitude: float,
                                   device) -> None:
    """One uniform value in [0, weight_magnitude/in_features) per link, constant over its nodes
    (reference PolynomialLayers.py:12-23; consumes the RNG identically: one rand(out, in))."""
    values = torch.rand(out_features, in_features, device=device) * weight_magnitude / in_features
    with torch.no_grad():
        w.copy_(values.unsqueeze(2).expand(-1, -1, w.size(2)))


class _PiecewiseBase(nn.Module):
    _discontinuous = False

    def __init__(self, n: int, in_features: int, out_features: int, segments: int, length: float = 2.0,
                 weight_magnitude: float = 1.0, poly=None, periodicity: Optional[float] = None, device: str = "cpu",
                 initialize: str = "constant_random", **kwargs):
        super().__init__()
        if poly is None:
            poly = LagrangePoly
        self._poly = poly(n)
        self._n = n
        self._segments = segments
        self.in_features = in_features
        self.out_features = out_features
        self.periodicity = periodicity
        self.device = device
        nb = n * segments if self._discontinuous else (n - 1) * segments + 1
        self.w = nn.Parameter(torch.empty(out_features, in_features, nb, device=device), requires_grad=True)
        if initialize == "constant_random":
            constant_random_initialization(self.w, out_features, in_features, weight_magnitude, device)
        else:
            bound = (1.0 if self._discontinuous else weight_magnitude) / in_features
            self.w.data.uniform_(-bound, bound)
        self.wrange = None
        self._length = length
        self._half = 0.5 * length

    # -- helpers used by refinement / inspection code ------------------------------------
    @property
    def n(self) -> int:
        return self._n

    def _eta(self, index):
        """Left edge of segment `index` in [-1, 1] (reference :338-344)."""
        eta = index / float(self._segments)
        return eta * 2 - 1

    def which_segment(self, x: Tensor) -> Tensor:
        """Segment index of every element (int64), bit-identical to the reference's CPU result."""
        if x.is_cuda:
            return ops.pw_segment_index(x.to(torch.float32), self._segments, float(self._length), self.periodicity)
        if self.periodicity is not None:
            x = make_periodic(x, self.periodicity)
        idx = (((x + self._half) / self._length) * self._segments).long()
        return idx.clamp(0, self._segments - 1)

    def x_local(self, x_global: Tensor, index: Tensor) -> Tensor:
        lo, hi = self._eta(index), self._eta(index + 1)
        return self._length * ((x_global - lo) / (hi - lo)) - self._half

    def x_global(self, x_local: Tensor, index: Tensor) -> Tensor:
        lo, hi = self._eta(index), self._eta(index + 1)
        return ((x_local + self._half) / self._length) * (hi - lo) + lo

    # -- the hot path ---------------------------------------------------------------------
    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError(
                f"{type(self).__name__}: input is on {x.device}; this build runs on CUDA (sm_100a) only and has "
                "no CPU fallback -- move the module and its input to a CUDA device")
        if x.dim() != 2:
            raise ValueError(f"{type(self).__name__} expects [batch, in_features], got {tuple(x.shape)}")
        return ops.piecewise_polynomial(x, self.w, self._n, self._segments, float(self._length),
                                        self._discontinuous, self.periodicity)

    def forward_normalized(self, x: Tensor, norm: str, eps: float = 1e-6) -> Tensor:
        """norm(self(x)) for the per-sample normalisations (max_abs / max_center / l2) as one layer call:
        HighOrderMLP uses it for a layer followed by its normalisation module."""
        if not x.is_cuda or x.dim() != 2:
            return ops.normalize_rows(self.forward(x), norm, eps)  # forward raises the layer's own errors
        return ops.piecewise_polynomial_normalized(x, self.w, self._n, self._segments, norm, eps, float(self._length),
                                                   self._discontinuous, self.periodicity)

    def extra_repr(self) -> str:
        return (f"n={self._n}, in_features={self.in_features}, out_features={self.out_features}, "
                f"segments={self._segments}, length={self._length}, periodicity={self.periodicity}")


class Piecewise(_PiecewiseBase):
    """Continuous piecewise polynomial: neighbouring segments share their end node."""
    _discontinuous = False

    def interpolate(self, layer_out: "Piecewise") -> None:
        """p-refinement into `layer_out` (reference :346-350)."""
        from .refinement import interpolate_polynomial_layer
        interpolate_polynomial_layer(self, layer_out)

    def refine(self, layer_out: "Piecewise") -> None:
        """h-refinement into `layer_out` (reference :352-356)."""
        from .refinement import refine_polynomial_layer
        refine_polynomial_layer(layer_in=self, layer_out=layer_out)


class PiecewiseDiscontinuous(_PiecewiseBase):
    """Discontinuous piecewise polynomial: every segment owns its n nodes."""
    _discontinuous = True

    def interpolate(self, layer_out: "PiecewiseDiscontinuous") -> None:
        from .refinement import interpolate_polynomial_layer
        interpolate_polynomial_layer(layer_in=self, layer_out=layer_out)

    def refine(self, layer_out: "PiecewiseDiscontinuous") -> None:
        from .refinement import refine_discontinuous_polynomial_layer
        refine_discontinuous_polynomial_layer(layer_in=self, layer_out=layer_out)


class PiecewisePolynomial(Piecewise):
    def __init__(self, n: int, in_features: int, out_features: int, segments: int, length=2.0, weight_magnitude=1.0,
                 periodicity: Optional[float] = None, device: str = "cpu", **kwargs):
        # like the reference (:545-555), extra kwargs (rescale_output, scale, initialization, ...) are
        # accepted and ignored, so the initialisation is always "constant_random"
        super().__init__(n, in_features, out_features, segments, length, weight_magnitude, poly=LagrangePoly,
                         periodicity=periodicity, device=device)


class PiecewiseDiscontinuousPolynomial(PiecewiseDiscontinuous):
    def __init__(self, n: int, in_features: int, out_features: int, segments: int, length=2.0, weight_magnitude=1.0,
                 periodicity: Optional[float] = None, device: str = "cpu", **kwargs):
        super().__init__(n, in_features, out_features, segments, length, weight_magnitude, poly=LagrangePoly,
                         periodicity=periodicity, device=device)


class Polynomial(nn.Module):
    """One Lagrange polynomial per link on [-1, 1] (reference PolynomialLayers.py:26-76: Function with
    LagrangePolyFlat): y[b,o] = sum_i sum_j phi_j(x[b,i]) w[o,i,j], no segment clamp.  It is the
    single-segment case of the piecewise kernels; only the default length 2.0 maps onto
    them (x_in == x), other lengths raise.  Keeps the reference's unused ``result`` parameter so
    that state_dict keys match."""

    def __init__(self, n: int, in_features: int, out_features: int, length: float = 2.0, weight_magnitude: float = 1.0,
                 periodicity: Optional[float] = None, initialize: str = "constant_random", device: str = "cpu",
                 **kwargs):
        super().__init__()
        if float(length) != 2.0:
            raise ValueError("Polynomial: only length=2.0 is supported by the CUDA path")
        self.poly = LagrangePoly(n, length=length)
        self.n = n
        self.in_features = in_features
        self.out_features = out_features
        self.periodicity = periodicity
        self.w = nn.Parameter(torch.empty(out_features, in_features, n, device=device), requires_grad=True)
        if initialize == "constant_random":
            constant_random_initialization(self.w, out_features, in_features, weight_magnitude, device)
        else:
            self.w.data.uniform_(-weight_magnitude / in_features, weight_magnitude / in_features)
        self.result = nn.Parameter(torch.empty(out_features), requires_grad=True)

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"Polynomial: input is on {x.device}; this build runs on CUDA (sm_100a) only and has no "
                               "CPU fallback")
        return ops.piecewise_polynomial(x, self.w, self.n, 1, 2.0, True, self.periodicity)
