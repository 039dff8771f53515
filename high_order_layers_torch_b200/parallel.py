"""Plain data parallelism over the batch (north_star / SURVEY 8e): weights replicated, one
process per GPU, ONE exchange step per training step -- a sum-allreduce of the weight gradients.

``FlatGradients`` re-homes every parameter gradient as a view into one flat fp32 buffer, so the
dW kernels' results are copied once into it and the step needs a single NCCL all-reduce over
NVLink (6.1 MB for the 784->128->128->10 MLP, 168 MB for the 1024x1024 n=5 S=8 layer) instead
of one collective per tensor.  Works with any torch.distributed backend (``gloo`` in the CPU
tests, ``nccl`` on the GPUs).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Rows [lo, hi) of a batch of `total` samples owned by `rank` (contiguous, balanced)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FlatGradients:
    def __init__(self, params: Iterable[torch.nn.Parameter]):
        seen = set()
        self.params: List[torch.nn.Parameter] = []
        for p in params:  # HighOrderMLP registers its first/last layer twice: deduplicate
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)  # autograd accumulates in place
            off += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = True, async_op: bool = False):
        """Sum (or mean) the flat gradient over the ranks; a no-op for a single process."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            if async_op:
                work.wait()
            self.flat.div_(dist.get_world_size(group))
            return None
        return work
