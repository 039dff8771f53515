"""Small elementwise helpers kept for API compatibility (reference utils.py:8-79).

The per-sample normalisations HighOrderMLP puts between layers run as one fused CUDA kernel each
way (csrc/norm.cu) -- CUDA fp32 only, no torch fallback; inside the layers the periodic
fold is fused into the prep kernel (csrc/prep.cu).  ``make_periodic`` is the host-side helper the
inspection methods (``which_segment`` on CPU tensors) use.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _rows(x: Tensor, dim: int, what: str) -> Tensor:
    """The fused kernels take CUDA fp32 [B, W] tensors normalised over dim 1; like the layers there is no
    CPU / torch fallback (north_star): anything else raises."""
    if not x.is_cuda:
        raise RuntimeError(f"{what}: input is on {x.device}; this build runs on CUDA (sm_100a) only and has no CPU fallback")
    if x.dtype != torch.float32:
        raise TypeError(f"{what}: float32 input required, got {x.dtype}")
    if x.dim() != 2 or dim not in (1, -1):
        raise ValueError(f"{what}: a [batch, width] tensor normalised over dim 1 is required, got {tuple(x.shape)}, dim={dim} "
                         "(use the *_nd / *_last variants for other ranks)")
    return x


def max_abs(x: Tensor, dim: int = 1) -> Tensor:
    return x.abs().max(dim=dim, keepdim=True)[0]


def max_abs_normalization(x: Tensor, eps: float = 1e-6, dim: int = 1) -> Tensor:
    """x / (max|x| + eps) per sample -- reference utils.py:12-13, as one fused kernel each way (csrc/norm.cu)."""
    from .ops import normalize_rows
    return normalize_rows(_rows(x, dim, "max_abs_normalization"), "max_abs", eps)


def max_abs_normalization_last(x: Tensor, eps: float = 1e-6) -> Tensor:
    """reference utils.py:16-17: normalised over the last dimension"""
    from .ops import normalize_rows
    flat = x.reshape(-1, x.shape[-1])
    return normalize_rows(_rows(flat, 1, "max_abs_normalization_last"), "max_abs", eps).reshape(x.shape)


def max_center_normalization(x: Tensor, eps: float = 1e-6, dim: int = 1) -> Tensor:
    """(x - midrange) / (max - midrange + eps) per sample -- reference utils.py:20-28."""
    from .ops import normalize_rows
    return normalize_rows(_rows(x, dim, "max_center_normalization"), "max_center", eps)


def max_center_normalization_nd(x: Tensor, eps: float = 1e-6) -> Tensor:
    """reference utils.py:42-54"""
    return max_center_normalization(x.reshape(x.shape[0], -1), eps=eps).reshape(x.shape)


def l2_normalization(x: Tensor, eps: float = 1e-6) -> Tensor:
    """x / (||x||_2 + eps) per sample -- reference utils.py:57-58."""
    from .ops import normalize_rows
    return normalize_rows(_rows(x, 1, "l2_normalization"), "l2", eps)


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
