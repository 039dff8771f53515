"""Plain data parallelism over the batch (north_star / SURVEY 8e): weights replicated, one
process per GPU, ONE exchange step per layer and training step -- a sum-allreduce of the weight
gradient, started as soon as that layer's dW kernel has been queued.

``FlatGradients`` re-homes every parameter gradient as a view into one flat fp32 buffer.

* The piecewise layers see the buffer as a *gradient sink* (``w._hol_grad_sink``): their backward
  writes dW straight into the parameter's slice (no temporary, no autograd accumulation, no zero fill)
  and calls ``ready(w)``.  dW is computed BEFORE dX (csrc/abi.cu), so the all-reduce of the slice is
  issued on a side stream while the dX kernels of the same layer -- and the whole backward of the
  layers below -- still run: the collective is hidden behind compute.
* Largest message first: the backward reaches the layers last-to-first, so the FIRST layer's gradient
  -- in an MLP by far the largest one (784 -> 128: 5.2 of 6.1 MB) -- is the last kernel of the backward
  and its exchange would sit exposed behind it.  A layer whose dX is still needed, and that is followed by
  a parameter with a larger gradient, therefore only computes dX in the backward and hands its dW launch to
  ``defer()``; ``finish()`` runs the deferred dW launches (largest first), so they -- not idle time --
  cover the exchange of the large slice.  With one process, or one layer, nothing is deferred.
* Every other parameter (conv weights reached through a permuted view, ``nn.Linear``) accumulates
  through autograd as usual and is reduced by ``finish()`` in as few collectives as its slices allow.

The exchange itself runs on our own kernels over NVLink peer memory where it can (csrc/p2p.cu): on
CUDA with the ``nccl`` backend the flat buffer is allocated in symmetric memory (torch's
``_symmetric_memory`` hands out the mapping: plumbing), every rank reads its peers' gradients with
plain loads over NVLink / NVSwitch, adds them in rank order and stores the result into all
buffers -- deterministic, graph-capturable, and latency-wise ~3x cheaper than NCCL's all-reduce at
the 5 MB size that sits exposed at the end of an MLP's backward.  ``collective="nccl"`` (or a backend
without peer memory: ``gloo`` in the CPU tests) uses ``dist.all_reduce`` instead.

A step is ``begin()`` -> forward -> backward -> ``finish()``.  With a single process the collectives
are skipped and only the direct dW write remains.  The whole step, collectives included, can be
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
                 average: bool = True, direct: bool = True, overlap: bool = True, collective: str = "auto",
                 defer: bool = True):
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
        self.collective = "nccl"
        self._p2p = None
        self.flat = None
        # "auto": peer memory up to 1 GiB of gradients (every rank takes the same decision from the same sizes);
        # beyond that NCCL's bandwidth-optimal algorithms are the better tool and the mapping is left alone
        if (collective == "p2p" or (collective == "auto" and self.numel * 4 <= (1 << 30))) and dev.type == "cuda" \
                and dt == torch.float32 and self._world() > 1 and dist.get_backend(group) == "nccl":
            try:
                self._setup_p2p(dev)
                self.collective = "p2p"
            except Exception as e:  # no peer mapping on this system: the library collective still works
                if collective == "p2p":
                    raise
                import warnings
                warnings.warn(f"FlatGradients: symmetric memory unavailable ({type(e).__name__}: {str(e)[:120]}); "
                              "using dist.all_reduce")
                self._p2p = None
                self.flat = None
        if self.flat is None:
            self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        if collective == "none":
            self.collective = "none"
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
        self.defer_dw = defer
        self._deferred: list = []      # (numel, order, parameter, launch) of dW launches handed to defer()

    # -- plumbing -------------------------------------------------------------------------
    def _setup_p2p(self, dev) -> None:
        """Flat buffer + flag block in symmetric memory, peer pointer tables for csrc/p2p.cu."""
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        lib = _lib.load()
        group = self.group if self.group is not None else dist.group.WORLD
        flat = symm_mem.empty(self.numel, dtype=torch.float32, device=dev)
        flat.zero_()
        words = (lib.hol_p2p_flags_bytes() + 3) // 4
        flags = symm_mem.empty(words, dtype=torch.int32, device=dev)
        flags.zero_()
        torch.cuda.synchronize(dev)
        h_buf = symm_mem.rendezvous(flat, group)
        h_flg = symm_mem.rendezvous(flags, group)
        dist.barrier(group)  # every rank's flag block is zeroed before anyone signals into it
        self.flat = flat
        self._p2p = dict(lib=lib, flags=flags, handles=(h_buf, h_flg), bufs_dev=int(h_buf.buffer_ptrs_dev),
                         flags_dev=int(h_flg.buffer_ptrs_dev), rank=int(h_buf.rank), world=int(h_buf.world_size))

    def check(self) -> None:
        """Raise if a peer-memory exchange gave up waiting for a rank (csrc/p2p.cu); synchronises."""
        if self._p2p is not None and int(self._p2p["flags"][33].item()) != 0:
            raise RuntimeError("FlatGradients: a rank did not arrive at a peer-memory all-reduce (timeout)")

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
        if world == 1 or self.collective == "none":  # "none": diagnostic only (bench.py --collective none)
            return
        if self._p2p is not None:
            import ctypes

            from . import _lib
            q = self._p2p
            with torch.cuda.device(self.flat.device):
                _lib.check(q["lib"].hol_p2p_allreduce_f32(
                    ctypes.c_void_p(q["bufs_dev"]), ctypes.c_void_p(q["flags_dev"]), q["rank"], q["world"], lo, hi - lo,
                    (1.0 / world) if self.average else 1.0,
                    ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)), "hol_p2p_allreduce_f32")
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
        if self._deferred:
            raise RuntimeError("FlatGradients.begin(): the previous backward left deferred dW launches behind; "
                               "every backward must be followed by finish()")
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
        if id(p) in self._done or self._is_deferred(p):
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

    def wants_deferred(self, p: torch.nn.Parameter) -> bool:
        """True if the layer owning `p` should compute only dX now and hand its dW launch to ``defer()``: the
        exchange is overlapped (side stream, more than one process) and a parameter with a larger gradient
        has not produced it yet -- that one's exchange is what the deferred dW kernels will cover."""
        if not (self.defer_dw and self.direct):
            return False
        if self.defer_dw != "always" and (self._side is None or self._world() == 1):
            return False
        n = p.numel()
        return any(q.numel() > n and id(q) not in self._done and not self._is_deferred(q) and q is not p
                   for q in self.params)

    def _is_deferred(self, p: torch.nn.Parameter) -> bool:
        return any(q is p for _n, _k, q, _f in self._deferred)

    def defer(self, p: torch.nn.Parameter, launch) -> None:
        """`launch()` queues the dW kernels of `p` (writing its slice); run by ``finish()``, then ``ready(p)``."""
        self._deferred.append((p.numel(), len(self._deferred), p, launch))

    def _run_deferred(self) -> None:
        todo, self._deferred = sorted(self._deferred, key=lambda t: (-t[0], t[1])), []
        for _n, _k, p, launch in todo:
            launch()
            self.ready(p)

    def finish(self) -> None:
        """End of the backward: run the deferred dW launches, reduce what autograd accumulated and join the
        side stream."""
        self._run_deferred()
        self._alias()
        rest = [i for i in self._range if i not in self._done]
        if self._world() > 1 and rest:
            # Nothing is left to overlap with, but the exchanges of a step stay on ONE stream: the peer-memory
            # kernels share a flag block and a device-side epoch, so two of them must never run concurrently
            # (every rank has to walk the same sequence of barriers).
            if self._side is not None:
                ev = torch.cuda.Event()
                ev.record()
                self._side_used = True
                with torch.cuda.stream(self._side):
                    self._side.wait_event(ev)
                    for lo, hi in self._ranges(rest):
                        self._reduce(lo, hi)
            else:
                for lo, hi in self._ranges(rest):
                    self._reduce(lo, hi)
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
