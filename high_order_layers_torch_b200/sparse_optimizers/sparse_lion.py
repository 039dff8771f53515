"""SparseLion (reference sparse_optimizers/sparse_lion.py:13-92): Lion that only touches the
weights whose gradient is non-zero -- for the piecewise layers that is the n weights of the one
segment each sample activated per link; every other entry of dW is exactly zero.

Every parameter takes one fused kernel (csrc/optim.cu) through the
``hol::sparse_lion_step`` op; like the layers there is no CPU / torch fallback.  As in the
reference, ``weight_decay`` has no effect: its decay line multiplies the copy that advanced
indexing returns (sparse_lion.py:19), so no weight is ever decayed.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import Tensor
from torch.optim.optimizer import Optimizer

from .. import _lib
from ..ops import _ptr, _stream


@torch.library.custom_op("hol::sparse_lion_step", mutates_args=("p", "exp_avg"), device_types="cuda")
def sparse_lion_step(p: Tensor, grad: Tensor, exp_avg: Tensor, lr: float, wd: float, beta1: float,
                     beta2: float) -> None:
    if p.dtype != torch.float32 or not (p.is_contiguous() and exp_avg.is_contiguous()):
        raise TypeError("sparse_lion_step: contiguous float32 tensors required")
    lib = _lib.load()
    g = grad.contiguous()
    with torch.cuda.device(p.device):
        _lib.check(lib.hol_sparse_lion_step_f32(_ptr(p), _ptr(g), _ptr(exp_avg), p.numel(), lr, wd, beta1, beta2,
                                                _stream(p)), "hol_sparse_lion_step_f32")


class SparseLion(Optimizer):
    def __init__(self, params, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.99), weight_decay: float = 0.0,
                 use_triton: bool = False):
        assert lr > 0.0
        assert all(0.0 <= beta <= 1.0 for beta in betas)
        # use_triton: accepted for signature compatibility; there is no Triton path in this build
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure: Optional[Callable] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, wd, (beta1, beta2) = group["lr"], group["weight_decay"], group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state["exp_avg"] = torch.zeros_like(p)
                exp_avg = state["exp_avg"]
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("SparseLion: parameters must be contiguous float32 CUDA tensors (this build runs "
                                       "on CUDA (sm_100a) only and has no CPU fallback)")
                sparse_lion_step(p, p.grad, exp_avg, lr, wd, beta1, beta2)
        return loss
