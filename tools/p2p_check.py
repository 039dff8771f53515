#!/usr/bin/env python
"""Multi-GPU self-check of the peer-memory all-reduce (csrc/p2p.cu) against NCCL; run under torchrun:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_check.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from high_order_layers_torch_b200.parallel import FlatGradients

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in ((128, 784, 13), (7,), (128, 128, 13), (10, 128, 13), (3, 5))]
flat = FlatGradients(params, collective="p2p")
assert flat.collective == "p2p", flat.collective
ok = True
for it in range(3):
    gen = torch.Generator(device=dev).manual_seed(100 * it + rank)
    flat.begin()
    local = torch.randn(flat.numel, device=dev, generator=gen)
    flat.flat.copy_(local)
    ref = local.clone()
    dist.all_reduce(ref)                      # NCCL sum
    ref /= world
    # per-"layer" exchange on the side stream for the first and third parameter, the rest in finish()
    flat._done.clear()
    for p in (params[2], params[0]):
        flat.ready(p)
    flat.finish()
    torch.cuda.synchronize()
    err = (flat.flat - ref).abs().max().item() / ref.abs().max().item()
    # every rank holds bit-identical results
    gathered = [torch.empty_like(flat.flat) for _ in range(world)]
    dist.all_gather(gathered, flat.flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = ok and err < 1e-6 and same
    if rank == 0:
        print(f"iter {it}: rel err vs NCCL {err:.2e}, identical on all ranks: {same}")
flat.check()
# graph capture + replay, timing against NCCL at the three gradient sizes of the config-3 MLP
for n in (10 * 128 * 13, 128 * 128 * 13, 128 * 784 * 13):
    buf_nccl = torch.randn(n, device=dev)
    for mode, ovr in (("p2p (graph replay)", 0), ("nccl", None)):
        if ovr is not None:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    flat._reduce(0, n)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            dist.barrier()
            with torch.cuda.graph(g):
                flat._reduce(0, n)
            fn = g.replay
        else:
            fn = lambda: dist.all_reduce(buf_nccl)
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{n * 4 / 1e6:6.2f} MB on {world} GPUs, {mode}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per all-reduce")
        flat.check()
        if ovr is not None:
            del g
flat.check()
if rank == 0:
    print("P2P_CHECK", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
