"""Small elementwise helpers kept for API compatibility (reference utils.py:8-79).

The per-sample normalisations HighOrderMLP puts between layers run as one fused CUDA kernel each
way for CUDA fp32 [B, W] inputs (csrc/norm.cu); inside the layers the periodic fold
is fused into the prep kernel (csrc/prep.cu).
"""
from __future__ import annotations

import torch
from torch import Tensor


def _fused(x: Tensor, dim: int) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and dim in (1, -1)


def max_abs(x: Tensor, dim: int = 1) -> Tensor:
    return x.abs().max(dim=dim, keepdim=True)[0]


def max_abs_normalization(x: Tensor, eps: float = 1e-6, dim: int = 1) -> Tensor:
    """x / (max|x| + eps) per sample -- reference utils.py:12-13.  CUDA fp32 [B, W] inputs take the
    fused kernel pair (csrc/norm.cu); anything else the reference's torch expression."""
    if _fused(x, dim):
        from .ops import normalize_rows
        return normalize_rows(x, "max_abs", eps)
    return x / (max_abs(x, dim=dim) + eps)


def max_abs_normalization_last(x: Tensor, eps: float = 1e-6) -> Tensor:
    """reference utils.py:16-17"""
    if x.is_cuda and x.dtype == torch.float32 and x.dim() >= 2:
        from .ops import normalize_rows
        return normalize_rows(x.reshape(-1, x.shape[-1]), "max_abs", eps).reshape(x.shape)
    return x / (max_abs(x, dim=x.dim() - 1) + eps)


def max_center_normalization(x: Tensor, eps: float = 1e-6, dim: int = 1) -> Tensor:
    """(x - midrange) / (max - midrange + eps) per sample -- reference utils.py:20-28."""
    if _fused(x, dim):
        from .ops import normalize_rows
        return normalize_rows(x, "max_center", eps)
    max_x = torch.max(x, dim=dim, keepdim=True)[0]
    min_x = torch.min(x, dim=dim, keepdim=True)[0]
    midrange = 0.5 * (max_x + min_x)
    return (x - midrange) / (max_x - midrange + eps)


def max_center_normalization_nd(x: Tensor, eps: float = 1e-6) -> Tensor:
    """reference utils.py:42-54"""
    return max_center_normalization(x.reshape(x.shape[0], -1), eps=eps).reshape(x.shape)


def l2_normalization(x: Tensor, eps: float = 1e-6) -> Tensor:
    """x / (||x||_2 + eps) per sample -- reference utils.py:57-58."""
    if _fused(x, 1):
        from .ops import normalize_rows
        return normalize_rows(x, "l2", eps)
    return x / (x.norm(2, 1, keepdim=True) + eps)


def max_abs_normalization_nd(x: Tensor, eps: float = 1e-6) -> Tensor:
    """reference utils.py:61-65"""
    flat = x.reshape(x.shape[0], -1)
    return max_abs_normalization(flat, eps=eps).reshape(x.shape)


norm_type = {"max_abs": max_abs_normalization, "l2": l2_normalization}


def make_periodic(x: Tensor, periodicity: float) -> Tensor:
    """Mirror-periodic fold onto [-p/2, p/2] -- reference utils.py:74-79."""
    shifted = torch.remainder(x + 0.5 * periodicity, 2 * periodicity)
    folded = torch.where(shifted > periodicity, 2 * periodicity - shifted, shifted)
    return folded - 0.5 * periodicity
