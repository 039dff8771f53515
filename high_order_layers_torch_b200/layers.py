"""Layer registries and factories with the reference's names and error behaviour
(reference layers.py:258-295, :357-372) restricted to the hot path: the keys ``continuous`` /
``discontinuous`` / ``polynomial`` (fully connected) and ``continuous2d`` / ``discontinuous2d``
(convolution) and ``continuous2d`` (conv-transpose).  Asking for any other key
raises the same ValueError the reference raises for an unknown layer type."""
from __future__ import annotations

import torch.nn as nn

from .FunctionalConvolution import (PiecewiseDiscontinuousPolynomialConvolution1d,
                                    PiecewiseDiscontinuousPolynomialConvolution2d, PiecewisePolynomialConvolution1d,
                                    PiecewisePolynomialConvolution2d)
from .FunctionalConvolutionTranspose import PiecewisePolynomialConvolutionTranspose2d
from .PolynomialLayers import PiecewiseDiscontinuousPolynomial, PiecewisePolynomial, Polynomial
from .utils import (l2_normalization, max_abs_normalization, max_abs_normalization_last, max_abs_normalization_nd,
                    max_center_normalization, max_center_normalization_nd)


class MaxAbsNormalization(nn.Module):
    """x / (max|x| + eps) per sample, the normalisation HighOrderMLP puts between layers
    (reference layers.py:70-81)."""

    def __init__(self, eps: float = 1e-6, dim: int = 1):
        super().__init__()
        self._eps = eps
        self._dim = dim

    def forward(self, x):
        return max_abs_normalization(x, eps=self._eps, dim=self._dim)


class MaxAbsNormalizationLast(nn.Module):
    """reference layers.py:44-53"""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self._eps = eps

    def forward(self, x):
        return max_abs_normalization_last(x, eps=self._eps)


class MaxCenterNormalization(nn.Module):
    """(x - midrange) / (max - midrange + eps) per sample (reference layers.py:84-96)."""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self._eps = eps

    def forward(self, x):
        return max_center_normalization(x, eps=self._eps)


class MaxCenterNormalizationND(nn.Module):
    """reference layers.py:99-112"""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self._eps = eps

    def forward(self, x):
        return max_center_normalization_nd(x, eps=self._eps)


class MaxAbsNormalizationND(nn.Module):
    """reference layers.py:115-127"""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self._eps = eps

    def forward(self, x):
        return max_abs_normalization_nd(x, eps=self._eps)


class L2Normalization(nn.Module):
    """x / (||x||_2 + eps) per sample (reference layers.py:130-141)."""

    def __init__(self, eps: float = 1e-6):
        super().__init__()
        self._eps = eps

    def forward(self, x):
        return l2_normalization(x, eps=self._eps)


class SumLayer(nn.Module):
    """Sum of the outputs of several layers on the same input (resnet option of HighOrderMLP)."""

    def __init__(self, layer_list):
        super().__init__()
        self.layer_list = nn.ModuleList(layer_list)

    def forward(self, x):
        out = self.layer_list[0](x)
        for layer in self.layer_list[1:]:
            out = out + layer(x)
        return out


normalization_layers = {"max_abs": MaxAbsNormalization, "max_center": MaxCenterNormalization, "l2": L2Normalization}

fc_layers = {
    "continuous": PiecewisePolynomial,
    "discontinuous": PiecewiseDiscontinuousPolynomial,
    "polynomial": Polynomial,
}

convolutional_layers = {
    "continuous2d": PiecewisePolynomialConvolution2d,
    "continuous1d": PiecewisePolynomialConvolution1d,
    "discontinuous2d": PiecewiseDiscontinuousPolynomialConvolution2d,
    "discontinuous1d": PiecewiseDiscontinuousPolynomialConvolution1d,
}


convolutional_transpose_layers = {
    "continuous2d": PiecewisePolynomialConvolutionTranspose2d,
}


def high_order_fc_layers(layer_type: str, **kwargs):
    if layer_type in fc_layers.keys():
        return fc_layers[layer_type](**kwargs)
    raise ValueError(
        f"Fully connected layer type {layer_type} not recognized.  Must be one of {list(fc_layers.keys())}")


def high_order_convolution_layers(layer_type: str, **kwargs):
    if layer_type in convolutional_layers.keys():
        return convolutional_layers[layer_type](**kwargs)
    raise ValueError(
        f"Convolutional layer type {layer_type} not recognized.  Must be one of {list(convolutional_layers.keys())}")


def high_order_convolution_transpose_layers(layer_type: str, **kwargs):
    if layer_type in convolutional_transpose_layers.keys():
        return convolutional_transpose_layers[layer_type](**kwargs)
    raise ValueError(
        f"ConvolutionalTranspose layer type {layer_type} not recognized.  Must be one of "
        f"{list(convolutional_transpose_layers.keys())}")
