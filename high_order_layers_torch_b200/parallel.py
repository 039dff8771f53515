"""Plain data parallelism over the batch (north_star / SURVEY 8e): weights replicated, one
process per GPU, ONE exchange step per layer and training step -- a sum-allreduce of the weight
gradient, started as soon as that layer's dW kernel has been queued.

``FlatGradients`` re-homes every parameter gradient as a view into one flat fp32 buffer.

* The piecewise layers see the buffer as a *gradient sink* (``w._hol_grad_sink``): their backward
  writes dW straight into the parameter's slice (no temporary, no autograd accumulation, no zero fill)
  and calls ``ready(w)``.  dW is computed BEFORE dX (csrc/abi.cu), so the all-reduce of the slice is
  issued on a side stream while the dX kernels of the same layer -- and the whole backward of the
  layers below -- still run: the collective is hidden behind compute.
* Every other parameter (conv weights reached through a permuted view, ``nn.Linear``) accumulates
  through autograd as usual and is reduced by ``finish()`` in as few collectives as its slices allow.

A step is ``begin()`` -> forward -> backward -> ``finish()``.  Works with any torch.distributed
backend (``gloo`` in the CPU tests, ``nccl`` on the GPUs); with a single process the collectives are
skipped and only the direct dW write remains.  The whole step, collectives included, can be
captured in a CUDA graph: the side stream forks from and joins the capturing stream through events.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Rows [lo, hi) of a batch of `total` samples owned by `rank` (contiguous, balanced)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FlatGradients:
    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 average: bool = True, direct: bool = True, overlap: bool = True):
        seen = set()
        self.params: List[torch.nn.Parameter] = []
        for p in params:  # HighOrderMLP registers its first/last layer twice: deduplicate
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.group = group
        self.average = average
        self.direct = direct
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        self._range: Dict[int, Tuple[int, int]] = {}
        off = 0
        for p in self.params:
            self._range[id(p)] = (off, off + p.numel())
            off += p.numel()
        self._alias()
        self._side = torch.cuda.Stream(device=dev) if (overlap and dev.type == "cuda") else None
        self._works: list = []
        self._side_used = False
        self._done: set = set()
        self._direct_ids: set = set()  # parameters whose layers wrote into the sink in the last step

    # -- plumbing -------------------------------------------------------------------------
    def _world(self) -> int:
        if not dist.is_available() or not dist.is_initialized():
            return 1
        return dist.get_world_size(self.group)

    def _alias(self) -> None:
        """(Re-)establish p.grad as a view of the flat buffer.  ``zero_grad(set_to_none=True)`` (the
        default of optimizers and modules) breaks the aliasing: a gradient that autograd then
        allocates afresh is copied into its slice here, so nothing is ever reduced from stale zeros."""
        for p in self.params:
            lo, hi = self._range[id(p)]
            view = self.flat[lo:hi].view_as(p)
            g = p.grad
            if g is None or g.data_ptr() != view.data_ptr():
                if g is not None:
                    view.copy_(g)
                p.grad = view
            if self.direct:
                p._hol_grad_sink = self

    def _ranges(self, ids) -> List[Tuple[int, int]]:
        """Coalesced [lo, hi) ranges of the given parameters in buffer order."""
        out: List[Tuple[int, int]] = []
        for lo, hi in sorted(self._range[i] for i in ids):
            if out and out[-1][1] == lo:
                out[-1] = (out[-1][0], hi)
            else:
                out.append((lo, hi))
        return out

    def _reduce(self, lo: int, hi: int) -> None:
        world = self._world()
        if world == 1:
            return
        buf = self.flat[lo:hi]
        backend = dist.get_backend(self.group)
        if self.average and backend == "nccl":
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            if self.average:
                work.wait()
                buf.div_(world)
            else:
                self._works.append(work)

    # -- the step -------------------------------------------------------------------------
    def begin(self) -> None:
        """Start of a step: zero the slices that accumulate through autograd (the slices the layers
        overwrite through the sink need no fill) and forget the previous step's state."""
        self._alias()
        keep = self._direct_ids if self.direct else set()
        for lo, hi in self._ranges(i for i in self._range if i not in keep):
            self.flat[lo:hi].zero_()
        self._done.clear()
        self._works.clear()
        self._side_used = False

    zero_ = begin  # the older name

    # sink protocol, called by the layers' backward (ops.py) -------------------------------
    def target(self, p: torch.nn.Parameter) -> Optional[torch.Tensor]:
        """The slice dW of `p` is to be written into, or None if `p` is not managed / not aliased."""
        if not self.direct or id(p) not in self._range:
            return None
        g = p.grad
        lo, hi = self._range[id(p)]
        if g is None or g.data_ptr() != self.flat[lo:hi].data_ptr():
            return None
        if id(p) in self._done:
            raise RuntimeError("FlatGradients: a parameter received two gradients in one backward; build the "
                               "FlatGradients with direct=False for modules that are applied more than once")
        return g

    def ready(self, p: torch.nn.Parameter) -> None:
        """dW of `p` has been queued on the current stream: reduce its slice on the side stream."""
        self._done.add(id(p))
        self._direct_ids.add(id(p))
        if self._world() == 1:
            return
        lo, hi = self._range[id(p)]
        if self._side is None:
            self._reduce(lo, hi)
            return
        ev = torch.cuda.Event()
        ev.record()
        self._side_used = True
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._reduce(lo, hi)

    def finish(self) -> None:
        """End of the backward: reduce what autograd accumulated and join the side stream."""
        self._alias()
        rest = [i for i in self._range if i not in self._done]
        if self._world() > 1:
            for lo, hi in self._ranges(rest):
                self._reduce(lo, hi)          # on the current stream: nothing is left to overlap with
        if self._side is not None and self._side_used:
            # join the side stream (it forked from this one through an event, so the join is also valid
            # inside a CUDA graph capture; an unused side stream must NOT be touched during a capture)
            with torch.cuda.stream(self._side):
                for w in self._works:
                    w.wait()
            torch.cuda.current_stream().wait_stream(self._side)
        else:
            for w in self._works:
                w.wait()
        self._works.clear()
        self._side_used = False
        # a parameter that was direct once but took the autograd route this time must be zero-filled again
        self._direct_ids &= self._done

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = True, async_op: bool = False):
        """Older interface: everything at the end of the backward (equivalent to finish())."""
        if group is not None:
            self.group = group
        self.average = average
        self.finish()
        return None
