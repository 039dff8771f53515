"""Host-side description of the Lagrange basis at Chebyshev-Lobatto nodes.

Mirrors the names of reference LagrangePolynomial.py (chebyshevLobatto :9-22, LagrangeBasis
:190-211, LagrangeBasis1 :214-227, get_lagrange_basis :230-234, LagrangePoly :305-307) so that
code poking at ``layer._poly.basis.X`` keeps working.  The hot evaluation lives in the CUDA
kernels; the torch implementations here are helpers for small host-side uses (refinement,
inspection) and run on whatever device their inputs are on.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .ops import lobatto_nodes


def chebyshevLobatto(n: int) -> Tensor:
    """n points -cos(pi k/(n-1)) in [-1, 1], fp32, evaluated exactly like the reference."""
    return lobatto_nodes(n, 2.0).clone()


class LagrangeBasis:
    def __init__(self, n: int, length: float = 2.0):
        self.n = n
        self.num_basis = n
        self.X = lobatto_nodes(n, length).clone()
        diff = self.X.unsqueeze(1) - self.X.unsqueeze(0)          # [j, m] = X_j - X_m (fp32)
        self.denominators = torch.where(torch.eye(n, dtype=torch.bool), torch.ones_like(diff), diff)

    def __call__(self, x: Tensor, j: int) -> Tensor:
        X = self.X.to(x.device)
        den = self.denominators[j].to(x.device)
        factors = (x.unsqueeze(-1) - X) / den
        keep = torch.arange(self.n, device=x.device) != j
        return torch.where(keep, factors, torch.ones((), device=x.device)).prod(dim=-1)


class LagrangeBasis1:
    """Degenerate constant basis (n == 1)."""

    def __init__(self, length: float = 2.0):
        self.n = 1
        self.num_basis = 1
        self.X = torch.tensor([0.0])

    def __call__(self, x: Tensor, j: int) -> Tensor:
        return torch.ones_like(x)


def get_lagrange_basis(n: int, length: float = 2.0):
    return LagrangeBasis1(length=length) if n == 1 else LagrangeBasis(n, length=length)


class LagrangePoly:
    """``basis`` + ``interpolate(x[B,I], w[B,I,O,n]) -> [B,O]`` (reference Basis.py:486-512)."""

    def __init__(self, n: int, length: float = 2.0, **kwargs):
        self.n = n
        self.basis = get_lagrange_basis(n, length=length)

    def interpolate(self, x: Tensor, w: Tensor) -> Tensor:
        phi = torch.stack([self.basis(x, j) for j in range(self.n)])       # [n, B, I]
        return torch.einsum("nbi,bion->bo", phi, w)
