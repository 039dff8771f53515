"""Data-parallel host logic on CPU: world_size-2 gloo processes exercise the flat-gradient
allreduce and the batch sharding used by bench.py / the N>1 path (the kernels themselves need a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import high_order_layers_torch_b200 as hol
from high_order_layers_torch_b200.parallel import FlatGradients, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(1234)  # identical replicas
        net = hol.HighOrderMLP(layer_type="continuous", n=4, in_width=20, out_width=10, hidden_layers=1,
                               hidden_width=16, in_segments=4, out_segments=4, hidden_segments=4,
                               normalization=hol.MaxAbsNormalization)
        flat = FlatGradients(net.parameters())
        # first/last layers are registered twice in HighOrderMLP: they must be counted once
        assert flat.numel == sum(p.numel() for p in {id(p): p for p in net.parameters()}.values())
        assert len(flat.params) == 3
        flat.zero_()
        # emulate the dW kernels writing rank-dependent gradients (accumulated in place into the views)
        for k, p in enumerate(flat.params):
            assert p.grad.data_ptr() >= flat.flat.data_ptr()
            p.grad.add_(torch.full_like(p, float(rank + 1) * (k + 1)))
        flat.allreduce(average=True)
        for k, p in enumerate(flat.params):
            want = (k + 1) * sum(r + 1 for r in range(world)) / world
            assert torch.allclose(p.grad, torch.full_like(p, want)), (rank, k)
        # a second step reuses the same buffer
        flat.zero_()
        assert not flat.flat.any()
        flat.params[0].grad.add_(1.0)
        flat.allreduce(average=False)
        assert torch.allclose(flat.params[0].grad, torch.full_like(flat.params[0], float(world)))
        # the layers' gradient-sink protocol (ops._pw_bwd): target() hands out the parameter's slice, the dW
        # kernel overwrites it, ready() starts the slice's all-reduce at once, finish() reduces the rest
        flat.average = True
        flat.begin()
        p0, p1, p2 = flat.params
        out = flat.target(p0)
        assert out is not None and out.data_ptr() == p0.grad.data_ptr()
        out.copy_(torch.full_like(p0, float(rank + 1)))
        flat.ready(p0)
        p1.grad.add_(10.0 * (rank + 1))          # autograd route
        p2.grad.add_(100.0 * (rank + 1))
        flat.finish()
        mean = sum(r + 1 for r in range(world)) / world
        assert torch.allclose(p0.grad, torch.full_like(p0, mean))
        assert torch.allclose(p1.grad, torch.full_like(p1, 10.0 * mean))
        assert torch.allclose(p2.grad, torch.full_like(p2, 100.0 * mean))
        flat.begin()                              # the sink-written slice is not re-zeroed (it is overwritten) ...
        assert torch.allclose(p0.grad, torch.full_like(p0, mean)) and not p1.grad.any() and not p2.grad.any()
        try:                                      # ... and a second gradient for it within one step is refused
            flat.ready(p0)
            flat.target(p0)
            raise AssertionError("expected RuntimeError")
        except RuntimeError:
            pass
        # set_to_none breaks the aliasing; begin()/finish() re-home a freshly allocated gradient
        p1.grad = None
        flat.begin()
        assert p1.grad.data_ptr() == flat.flat[flat._range[id(p1)][0]:].data_ptr()
        lo, hi = shard_range(65536 + 3, rank, world)
        t = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(t)
        assert t.item() == 65536 + 3
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))


def test_shard_range_partitions_batch():
    for total in (0, 1, 7, 8192, 65536, 65539):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                covered += list(range(lo, hi)) if total < 100 else [hi - lo]
            if total < 100:
                assert covered == list(range(total))
            else:
                assert sum(covered) == total and max(covered) - min(covered) <= 1


def test_largest_message_first_deferral_order(monkeypatch):
    """parallel.FlatGradients: a layer that is followed (in backward order) by a parameter with a larger
    gradient defers its dW launch; finish() runs the deferred launches largest first and starts each slice's
    exchange right behind it.  Host logic only: the side stream / world size are stubbed."""
    import torch

    from high_order_layers_torch_b200.parallel import FlatGradients

    w1 = torch.nn.Parameter(torch.zeros(128, 784, 13))   # config 3: 5.2 MB
    w2 = torch.nn.Parameter(torch.zeros(128, 128, 13))
    w3 = torch.nn.Parameter(torch.zeros(10, 128, 13))
    flat = FlatGradients([w1, w2, w3])
    log = []
    monkeypatch.setattr(flat, "_world", lambda: 8)
    monkeypatch.setattr(flat, "_reduce", lambda lo, hi: log.append(("reduce", hi - lo)))
    flat._side = None
    flat.begin()
    assert not flat.wants_deferred(w3)            # no overlap stream: nothing to cover, nothing deferred
    flat._side = object()
    monkeypatch.setattr(flat, "ready", lambda p: (log.append(("ready", p.numel())), flat._done.add(id(p))))
    # backward order: layer 3, layer 2, layer 1
    assert flat.wants_deferred(w3) and flat.target(w3) is not None
    flat.defer(w3, lambda: log.append(("dw", w3.numel())))
    try:
        flat.target(w3)
        raise AssertionError("a second gradient for a deferred parameter must be refused")
    except RuntimeError:
        pass
    assert flat.wants_deferred(w2)
    flat.defer(w2, lambda: log.append(("dw", w2.numel())))
    assert not flat.wants_deferred(w1)            # the largest one: computed and exchanged at once
    flat.ready(w1)
    flat._side = None                              # (finish() would join the stub stream)
    flat.finish()
    assert log == [("ready", w1.numel()), ("dw", w2.numel()), ("ready", w2.numel()),
                   ("dw", w3.numel()), ("ready", w3.numel())], log
    flat.begin()                                   # nothing left behind
    # a single layer (config 2) or a single process never defers
    solo = FlatGradients([torch.nn.Parameter(torch.zeros(4, 4, 5))])
    solo._side = object()
    monkeypatch.setattr(solo, "_world", lambda: 8)
    assert not solo.wants_deferred(solo.params[0])
    monkeypatch.setattr(flat, "_world", lambda: 1)
    flat._side = object()
    assert not flat.wants_deferred(w3)
