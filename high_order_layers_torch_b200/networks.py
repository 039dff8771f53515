"""HighOrderMLP with the reference's constructor (reference networks.py:160-288)."""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch.nn as nn
from torch import Tensor

from .layers import L2Normalization, MaxAbsNormalization, MaxCenterNormalization, SumLayer, high_order_fc_layers
from .PolynomialLayers import _PiecewiseBase


def _fusable_norm(layer, norm):
    """(kind, eps) when `norm` is one of the per-sample normalisations the layer kernels can apply themselves
    and neither module carries hooks that would have to see the intermediate tensor; else None."""
    if not isinstance(layer, _PiecewiseBase) or not hasattr(layer, "forward_normalized"):
        return None
    for m in (layer, norm):
        if m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks:
            return None
    if type(norm) is MaxAbsNormalization and norm._dim in (1, -1):
        return "max_abs", norm._eps
    if type(norm) is MaxCenterNormalization:
        return "max_center", norm._eps
    if type(norm) is L2Normalization:
        return "l2", norm._eps
    return None


class HighOrderMLP(nn.Module):
    """input layer -> [normalization][non_linearity][dropout] hidden layer ... -> output layer.

    As in the reference, ``scale`` / ``rescale_output`` / ``initialization`` are forwarded to the
    layers and swallowed there, and the first and last layers are registered twice
    (``input_layer`` / ``output_layer`` and inside ``model``), which keeps state_dict keys identical.
    """

    def __init__(self, layer_type: str, n: int, in_width: int, out_width: int, hidden_layers: int,
                 hidden_width: int, scale: float = 2.0, n_in: Optional[int] = None, n_out: Optional[int] = None,
                 n_hidden: Optional[int] = None, rescale_output: bool = False, periodicity: Optional[float] = None,
                 non_linearity: Optional[Callable[[Tensor], Tensor]] = None, in_segments: Optional[int] = None,
                 out_segments: Optional[int] = None, hidden_segments: Optional[int] = None,
                 normalization: Optional[Callable[[], Any]] = None, resnet: bool = False, device: str = "cpu",
                 layer_type_in: Optional[str] = None, initialization: str = "constant_random",
                 dropout: float = 0.0) -> None:
        super().__init__()
        n_in = n_in or n
        n_hidden = n_hidden or n
        n_out = n_out or n
        common = dict(rescale_output=rescale_output, scale=scale, periodicity=periodicity, device=device)
        self.dropout_layer = nn.Dropout(p=dropout)

        def between(stack):
            if normalization is not None:
                stack.append(normalization())
            if non_linearity is not None:
                stack.append(non_linearity)
            if dropout > 0:
                stack.append(self.dropout_layer)

        self.input_layer = high_order_fc_layers(layer_type=layer_type_in or layer_type, n=n_in,
                                                in_features=in_width, out_features=hidden_width,
                                                segments=in_segments, intialization=initialization, **common)
        stack = [self.input_layer]
        for k in range(hidden_layers):
            between(stack)
            hidden = high_order_fc_layers(layer_type=layer_type, n=n_hidden, in_features=hidden_width,
                                          out_features=hidden_width, segments=hidden_segments,
                                          initialization=initialization, **common)
            if resnet is True and k > 0:
                hidden = SumLayer(layer_list=[hidden, stack[-1]])
            stack.append(hidden)
        between(stack)
        self.output_layer = high_order_fc_layers(layer_type=layer_type, n=n_out, in_features=hidden_width,
                                                 out_features=out_width, segments=out_segments,
                                                 initialization=initialization, **common)
        stack.append(self.output_layer)
        self.model = nn.Sequential(*stack)
        self.fuse_normalization = True  # False: run every module of `model` on its own (same results)

    def forward(self, x: Tensor) -> Tensor:
        """``self.model(x)``, except that a layer directly followed by its per-sample normalisation runs as
        one fused call (``forward_normalized``: the forward's last kernel also normalises, the
        normalisation's backward also takes the statistic the layer's backward needs) -- same values."""
        mods = list(self.model)
        k = 0
        while k < len(mods):
            fused = _fusable_norm(mods[k], mods[k + 1]) if (self.fuse_normalization and k + 1 < len(mods)
                                                            and x.is_cuda and x.dim() == 2) else None
            if fused is not None:
                x = mods[k].forward_normalized(x, *fused)
                k += 2
            else:
                x = mods[k](x)
                k += 1
        return x
