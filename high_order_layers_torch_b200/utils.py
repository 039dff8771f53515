"""Small elementwise helpers kept for API compatibility (reference utils.py:8-79).

They are ordinary torch ops (memory-bound glue between layers), not part of the CUDA hot path;
inside the layers the periodic fold is fused into the prep kernel (csrc/prep.cu).
"""
from __future__ import annotations

import torch
from torch import Tensor


def max_abs(x: Tensor, dim: int = 1) -> Tensor:
    return x.abs().max(dim=dim, keepdim=True)[0]


def max_abs_normalization(x: Tensor, eps: float = 1e-6, dim: int = 1) -> Tensor:
    """x / (max|x| + eps) per sample -- reference utils.py:12-13."""
    return x / (max_abs(x, dim=dim) + eps)


def max_abs_normalization_nd(x: Tensor, eps: float = 1e-6) -> Tensor:
    flat = x.reshape(x.shape[0], -1)
    return (flat / (max_abs(flat) + eps)).reshape(x.shape)


def make_periodic(x: Tensor, periodicity: float) -> Tensor:
    """Mirror-periodic fold onto [-p/2, p/2] -- reference utils.py:74-79."""
    shifted = torch.remainder(x + 0.5 * periodicity, 2 * periodicity)
    folded = torch.where(shifted > periodicity, 2 * periodicity - shifted, shifted)
    return folded - 0.5 * periodicity
