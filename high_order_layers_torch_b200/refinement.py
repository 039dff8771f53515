"""p- / h-refinement, linear initialisation and smoothing of piecewise-polynomial layers
(SURVEY 8f-4).  Same functions, arguments and results as reference PolynomialLayers.py:757-1023 and
networks.py:1002-1131, but each one is a handful of batched tensor operations on the device the
weights live on -- a small interpolation matrix applied to all O*I links at once -- instead of the
reference's Python loops over outputs x inputs x segments (minutes for a 784x128 layer).

All of them act on weights only (no activations), so they work on CPU and CUDA tensors alike.
"""
from __future__ import annotations

import random
from typing import Callable, Union

import torch
import torch.nn as nn
from torch import Tensor

from .PolynomialLayers import (Piecewise, PiecewiseDiscontinuous, PiecewiseDiscontinuousPolynomial,
                               PiecewisePolynomial)


def _basis_matrix(poly, x: Tensor) -> Tensor:
    """[len(x), n] matrix phi_j(x_k) of a layer's basis (fp32, evaluated like LagrangeBasis.__call__)."""
    x = x.reshape(-1).to(torch.float32).cpu()
    return torch.stack([poly.basis(x, j) for j in range(poly.basis.n)], dim=-1)


def _segment_view(w: Tensor, n: int, segments: int, discontinuous: bool) -> Tensor:
    """w[O, I, nb] -> [O, I, S, n]: the n node weights of every segment (neighbouring segments of a
    continuous layer overlap in their end node; a continuous n == 1 layer has one weight in all)."""
    if discontinuous:
        return w.reshape(w.shape[0], w.shape[1], segments, n)
    if n == 1:
        return w[:, :, :1].unsqueeze(2).expand(-1, -1, segments, 1)
    return w.unfold(2, n, n - 1)


def _store_segments(w_out: Tensor, values: Tensor, n: int, discontinuous: bool) -> None:
    """values[O, I, S, n] -> w_out[O, I, nb] with the reference's write order: segments ascending,
    so a shared end node keeps the value computed from the segment to its RIGHT."""
    O, I, S, _ = values.shape
    if discontinuous:
        w_out.copy_(values.reshape(O, I, S * n))
    elif n == 1:
        w_out[:, :, 0] = values[:, :, S - 1, 0]
    else:
        w_out[:, :, : S * (n - 1)] = values[:, :, :, : n - 1].reshape(O, I, S * (n - 1))
        w_out[:, :, S * (n - 1)] = values[:, :, S - 1, n - 1]


def interpolate_polynomial_layer(layer_in: Union[PiecewisePolynomial, PiecewiseDiscontinuousPolynomial],
                                 layer_out: Union[PiecewisePolynomial, PiecewiseDiscontinuousPolynomial]) -> None:
    """p-refinement: weights of `layer_out` (same segments, other polynomial order) from `layer_in`
    (reference PolynomialLayers.py:757-812).  Exact when n_out >= n_in."""
    if layer_in._segments != layer_out._segments:
        raise ValueError(f"Number of input and output segments must be the same, got {layer_in._segments} and "
                         f"{layer_out._segments}")
    disc = isinstance(layer_in, PiecewiseDiscontinuous)
    n_in, n_out, S = layer_in._poly.basis.n, layer_out._poly.basis.n, layer_in._segments
    with torch.no_grad():
        w_in = layer_in.w
        M = _basis_matrix(layer_in._poly, layer_out._poly.basis.X).to(w_in.device)     # [n_out, n_in]
        values = torch.matmul(_segment_view(w_in, n_in, S, disc), M.t())               # [O, I, S, n_out]
        _store_segments(layer_out.w, values.to(layer_out.w.device), n_out, disc)


def refine_polynomial_layer(layer_in: PiecewisePolynomial, layer_out: PiecewisePolynomial) -> None:
    """h-refinement of a continuous layer: the nodes of every segment of `layer_out` are evaluated in
    `layer_in` (reference PolynomialLayers.py:884-948)."""
    n_in, n_out = layer_in._poly.basis.n, layer_out._poly.basis.n
    S_out = layer_out._segments
    with torch.no_grad():
        w_in = layer_in.w
        X_out = layer_out._poly.basis.X.reshape(1, -1).to(torch.float32).cpu()          # [1, n_out]
        # x_global(x_out, j) with the reference's scalar arithmetic: eta in Python floats, applied in fp32
        lo = [j / float(S_out) * 2 - 1 for j in range(S_out)]
        hi = [(j + 1) / float(S_out) * 2 - 1 for j in range(S_out)]
        scale = torch.tensor([h - l for h, l in zip(hi, lo)], dtype=torch.float64).to(torch.float32).reshape(-1, 1)
        shift = torch.tensor(lo, dtype=torch.float64).to(torch.float32).reshape(-1, 1)
        xg = ((X_out + layer_out._half) / layer_out._length) * scale + shift            # [S_out, n_out]
        idx = layer_in.which_segment(xg)                                                # CPU path (tiny)
        xl = layer_in.x_local(xg, idx)
        phi = _basis_matrix(layer_in._poly, xl).to(w_in.device)                         # [S_out*n_out, n_in]
        cols = (idx.reshape(-1, 1) * (n_in - 1) + torch.arange(n_in).reshape(1, -1)).to(w_in.device)
        gathered = w_in[:, :, cols]                                                     # [O, I, S_out*n_out, n_in]
        values = (gathered * phi).sum(-1).reshape(w_in.shape[0], w_in.shape[1], S_out, n_out)
        _store_segments(layer_out.w, values.to(layer_out.w.device), n_out, False)


def refine_discontinuous_polynomial_layer(layer_in: PiecewiseDiscontinuousPolynomial,
                                          layer_out: PiecewiseDiscontinuousPolynomial):
    """The reference raises here (PolynomialLayers.py:815-881: node positions on a discontinuity
    belong to two segments); kept so that callers see the same error."""
    raise ValueError("Since boundaries are doubled up at discontinuities there is ambiguity and this does not work.  "
                     "This function needs to be fixed.")


def smooth_discontinuous_layer(layer: PiecewiseDiscontinuous, factor: float):
    """Move the two values at every interior discontinuity towards each other by `factor`/2 of
    their difference each (reference PolynomialLayers.py:951-969)."""
    n = layer.n
    with torch.no_grad():
        w = layer.w
        upper = w[:, :, (n - 1):-1:n]      # last node of segment s
        lower = w[:, :, n:-1:n]            # first node of segment s + 1
        step = 0.5 * factor * (upper - lower)
        new_upper, new_lower = upper - step, lower + step
        w[:, :, (n - 1):-1:n] = new_upper
        w[:, :, n:-1:n] = new_lower


def initialize_polynomial_layer(layer_in: Union[PiecewisePolynomial, PiecewiseDiscontinuous], max_slope: float = 1.0,
                                max_offset: float = 0.0) -> None:
    """Every link starts as the straight line a*x + b over the whole domain, a and b drawn from
    Python's `random` in the reference's order (outputs, inputs; slope then offset)
    (reference PolynomialLayers.py:972-1023)."""
    if not isinstance(layer_in, (PiecewisePolynomial, PiecewiseDiscontinuous)):
        return
    n, S = layer_in._poly.basis.n, layer_in._segments
    disc = isinstance(layer_in, PiecewiseDiscontinuous)
    O, I = layer_in.w.shape[0], layer_in.w.shape[1]
    draws = [random.random() for _ in range(2 * O * I)]
    with torch.no_grad():
        a = max_slope * (torch.tensor(draws[0::2], dtype=torch.float64) * 2 - 1.0)
        b = max_offset * (torch.tensor(draws[1::2], dtype=torch.float64) * 2 - 1.0)
        X = layer_in._poly.basis.X.reshape(1, -1).to(torch.float32).cpu()
        lo = [j / float(S) * 2 - 1 for j in range(S)]
        hi = [(j + 1) / float(S) * 2 - 1 for j in range(S)]
        scale = torch.tensor([h - l for h, l in zip(hi, lo)], dtype=torch.float64).to(torch.float32).reshape(-1, 1)
        shift = torch.tensor(lo, dtype=torch.float64).to(torch.float32).reshape(-1, 1)
        xg = ((X + layer_in._half) / layer_in._length) * scale + shift                  # [S, n] fp32
        a32 = a.to(torch.float32).reshape(O, I, 1, 1)
        b32 = b.to(torch.float32).reshape(O, I, 1, 1)
        values = a32 * xg.reshape(1, 1, S, n) + b32
        _store_segments(layer_in.w, values.to(layer_in.w.device), n, disc)


# ------------------------------------------------------------------------------------------
# network level (reference networks.py:1002-1131)
# ------------------------------------------------------------------------------------------
def _leaf_modules(module: nn.Module):
    return [m for m in module.modules() if not isinstance(m, nn.Sequential)]


def initialize_network_polynomial_layers(network: nn.Module, max_slope: float, max_offset: float,
                                         scale_slope: Callable[[float], float] = lambda input_size: 1) -> None:
    for layer in _leaf_modules(network):
        if isinstance(layer, (PiecewisePolynomial, PiecewiseDiscontinuous)):
            initialize_polynomial_layer(layer, max_slope=max_slope * scale_slope(layer.in_features),
                                        max_offset=max_offset)


def interpolate_high_order_mlp(network_in, network_out) -> None:
    """p-refinement of a whole HighOrderMLP."""
    for l_in, l_out in zip(_leaf_modules(network_in.model), _leaf_modules(network_out.model)):
        if hasattr(l_in, "interpolate"):
            l_in.interpolate(l_out)


def hp_refine_high_order_mlp(network_in, network_out) -> None:
    """h- (and p-) refinement of a whole HighOrderMLP."""
    for l_in, l_out in zip(_leaf_modules(network_in.model), _leaf_modules(network_out.model)):
        if hasattr(l_in, "refine"):
            l_in.refine(l_out)


def smooth_discontinuous_network(network_in: nn.Module, factor: float) -> None:
    for layer in network_in.model.modules():
        if isinstance(layer, PiecewiseDiscontinuous):
            smooth_discontinuous_layer(layer=layer, factor=factor)
