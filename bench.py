#!/usr/bin/env python
"""Benchmark of the piecewise-polynomial hot path (fwd+bwd samples/s) -- see DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config all|c1|c2|c3|c4|c5] [--impl ours|reference]

Default (`--config all`, N=1 and every N of the scaling run): the headline line is the workload
BASELINE.json's metric is quoted on -- "piecewise-poly MLP 1/2/4/8 x B200" = config 3, the
invariant-MNIST-shaped HighOrderMLP 784->128->128->10 (continuous, n=4, segments=4, max_abs), batch
8192 per GPU, whole step (forward, CE loss, backward, per-layer NCCL all-reduce of dW overlapped
with the rest of the backward) captured in one CUDA graph -- and the same JSON line carries the other
named configs as sub-blocks under "configs": c2 (single 1024->1024 n=5 S=8 layer, batch 65536), c4
(LeNet-style conv net, batch 1024) and one c5 sweep point (4096->4096 n=4 S=64, batch 32768), each
with its own ms_per_step / value / roofline / plan, measured the same way in the same process.
`--config cX` runs a single workload as the line itself.

A step = zero grads, forward, backward with a supplied upstream gradient or loss (dX and dW), and
for N>1 the NCCL all-reduce of dW (weak scaling: per-GPU batch fixed).  One JSON line on rank 0.

`--impl reference` times the reference's own CPU implementation of the same path on the host
cores (the unmodified reference from baseline/_ref when it is present, otherwise the numpy
oracle port) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: SIMT FMA issue peak at the max SM clock


# ------------------------------------------------------------------------------------------
# workload definitions
# ------------------------------------------------------------------------------------------
CONFIGS = {
    "c2": dict(kind="layer", layer="discontinuous", n=5, S=8, I=1024, O=1024, B=65536,
               name="c2: PiecewiseDiscontinuousPolynomial 1024->1024 n=5 segments=8, batch 65536/GPU, fwd+bwd(dX,dW)"),
    "c5": dict(kind="layer", layer="continuous", n=4, S=8, I=4096, O=4096, B=32768,
               name="c5: PiecewisePolynomial 4096->4096 sweep point (n, segments from --n/--segments), batch 32768/GPU"),
    "c3": dict(kind="mlp", layer="continuous", n=4, S=4, widths=(784, 128, 128, 10), B=8192,
               name="c3: HighOrderMLP 784->128->128->10 continuous n=4 segments=4 max_abs, batch 8192/GPU, CE loss"),
    "c1": dict(kind="mlp", layer="continuous", n=3, S=2, widths=(1, 10, 10, 1), B=1024,
               name="c1: HighOrderMLP 1->10->10->1 continuous n=3 segments=2 max_abs, batch 1024, MSE loss"),
    "c4": dict(kind="conv", n=3, S=4, B=1024,
               name="c4: LeNet-style continuous2d net (3->6 k5, pool, 6->16 k5, pool, fc 400->100), n=3 segments=4, "
                    "32x32x3, batch 1024/GPU, CE loss"),
}


def nb_of(n, S, disc):
    return n * S if disc else (n - 1) * S + 1


def layer_bytes_flops(B, I, O, n, S, disc, first_layer=False):
    """Algorithmic bytes / flops of one layer-step."""
    nb = nb_of(n, S, disc)
    by = 4 * (3 * B * I + 2 * B * O + 3 * I * O * nb) - (4 * B * I if first_layer else 0)
    fl = 6 * B * I * O * n
    return by, fl


def kernel_bytes_flops(kernel, B, I, O, n, S, disc):
    nb = nb_of(n, S, disc)
    if kernel == "forward":
        return 4 * (B * I + I * O * nb + B * O), 2 * B * I * O * n
    if kernel == "dx":
        return 4 * (B * O + 2 * B * I + I * O * nb), 2 * B * I * O * n
    if kernel == "dw":
        return 4 * (B * O + B * I + I * O * nb), 2 * B * I * O * n
    raise KeyError(kernel)


# ------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def count_launches(layers, first_needs_dx=True, conv=False, fused_norm=False):
    """Kernels of csrc/ launched per step for a stack of layers [(B, I, O, n, S, disc), ...];
    `fused_norm`: every layer but the last runs with its normalisation folded in (HighOrderMLP)."""
    from high_order_layers_torch_b200 import ops
    total = 0
    for k, (B, I, O, n, S, disc) in enumerate(layers):
        want_dx = first_needs_dx or k > 0
        fn = fused_norm and k + 1 < len(layers)
        total += ops.launch_count(0, B, I, O, n, S, disc, fused_norm=fn)
        total += ops.launch_count(1, B, I, O, n, S, disc, want_dx=want_dx, want_dw=True, prepared=True, fused_norm=fn)
        if conv:
            total += 1  # the conv prep is two kernels (per-pixel prep, unfold)
        if conv and want_dx:
            total += 1  # col2im
    return total


def tc_gemm_flops(kind, B, I, O, n, S, disc):
    """Dense tensor-core flops EXECUTED by the split-fp16 GEMM of one pass (3 MMAs per k-step),
    with the tile padding of csrc/tc_gemm.cu (tc_plan)."""
    up = lambda a, b: (a + b - 1) // b * b
    nb = nb_of(n, S, disc)
    kin = up(nb, 8) if kind == "dx" else nb  # Phi / W / dW pack the inputs back to back; dX pads them to 8 columns
    K = I * kin
    bn = 256 if O > 128 else (128 if O > 64 else 64)
    Bp = up(B, 512) if B > 256 else up(B, 32)
    if kind == "forward":
        M, N, Kc = up(Bp, 128), up(O, bn), up(K, 64)
    elif kind == "dx":
        bnx = (256 // (2 * kin)) * (2 * kin)
        M, N, Kc = up(Bp, 128), up(K, bnx), up(O, 64)
    else:
        M, N, Kc = up(K, 128), up(O, bn), up(Bp, 64)
    return 3 * 2 * M * N * Kc


def defer_mode(args):
    """FlatGradients(defer=...): largest-message-first by default, off / forced for the diagnostics."""
    return False if args.no_defer else ("always" if args.force_defer else True)


def make_inputs(torch, shape, gen, device):
    """x = 2*rand-1 with 1 % of the entries drawn from uniform(-1.25, 1.25)."""
    x = torch.rand(shape, generator=gen, device=device) * 2 - 1
    wide = torch.rand(shape, generator=gen, device=device) * 2.5 - 1.25
    pick = torch.rand(shape, generator=gen, device=device) < 0.01
    return torch.where(pick, wide, x)


def build_workload(torch, cfg, device, rank, args):
    import high_order_layers_torch_b200 as hol
    from high_order_layers_torch_b200.parallel import FlatGradients

    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    wl = {"cfg": cfg}
    if cfg["kind"] == "layer":
        n, S, I, O, B = cfg["n"], cfg["S"], cfg["I"], cfg["O"], cfg["B"]
        disc = cfg["layer"] == "discontinuous"
        torch.manual_seed(1234)  # identical replicas on every rank
        layer = hol.high_order_fc_layers(cfg["layer"], n=n, in_features=I, out_features=O, segments=S,
                                         device=str(device))
        with torch.no_grad():
            layer.w.uniform_(-1, 1)
        x = make_inputs(torch, (B, I), gen, device).requires_grad_(True)
        gy = torch.randn(B, O, generator=gen, device=device)
        flat = FlatGradients(layer.parameters(), collective=args.collective, defer=defer_mode(args))

        def step():
            flat.begin()
            x.grad = None
            y = layer(x)
            y.backward(gy)
            flat.finish()

        def e2e_compute(dev):
            xd = dev["x"].detach().requires_grad_(True)
            flat.begin()
            y = layer(xd)
            loss = (y * dev["gy"]).sum()     # d loss / d y = gy: the same backward as the device-resident step
            loss.backward()
            flat.finish()
            return loss

        by, fl = layer_bytes_flops(B, I, O, n, S, disc)
        wl.update(model=layer, flat=flat, step=step, e2e_compute=e2e_compute, samples=B,
                  host=dict(x=x.detach().cpu().pin_memory(), gy=gy.cpu().pin_memory()),
                  h2d=4 * B * I + 4 * B * O, d2h=4, bytes=by, flops=fl,
                  launches=count_launches([(B, I, O, n, S, disc)]),
                  layers=[(B, I, O, n, S, disc)], x=x, gy=gy)
        return wl

    if cfg["kind"] == "mlp":
        n, S, B = cfg["n"], cfg["S"], cfg["B"]
        wd = cfg["widths"]
        torch.manual_seed(1234)
        net = hol.HighOrderMLP(layer_type=cfg["layer"], n=n, in_width=wd[0], out_width=wd[-1],
                               hidden_layers=len(wd) - 3, hidden_width=wd[1], in_segments=S, out_segments=S,
                               hidden_segments=S, normalization=hol.MaxAbsNormalization, device=str(device))
        with torch.no_grad():
            for p in net.parameters():
                p.uniform_(-1, 1)
        x = make_inputs(torch, (B, wd[0]), gen, device)
        classify = wd[-1] > 1
        target = (torch.randint(0, wd[-1], (B,), generator=gen, device=device) if classify
                  else torch.randn(B, wd[-1], generator=gen, device=device))
        lossf = torch.nn.functional.cross_entropy if classify else torch.nn.functional.mse_loss
        flat = FlatGradients(net.parameters(), collective=args.collective, defer=defer_mode(args))

        def step():
            flat.begin()
            loss = lossf(net(x), target)
            loss.backward()
            flat.finish()
            return loss

        def e2e_compute(dev):
            flat.begin()
            loss = lossf(net(dev["x"]), dev["t"])
            loss.backward()
            flat.finish()
            return loss

        dims = list(zip(wd[:-1], wd[1:]))
        by = fl = 0
        for k, (i_, o_) in enumerate(dims):
            b_, f_ = layer_bytes_flops(B, i_, o_, n, S, False, first_layer=(k == 0))
            by += b_; fl += f_
        nl = len(dims)
        wl.update(model=net, flat=flat, step=step, e2e_compute=e2e_compute, samples=B,
                  host=dict(x=x.cpu().pin_memory(), t=target.cpu().pin_memory()),
                  h2d=x.numel() * 4 + target.numel() * target.element_size(), d2h=4, bytes=by, flops=fl,
                  # the normalisation between consecutive layers rides in the layer calls (hol_pw_forward_norm_f32)
                  launches=count_launches([(B, i_, o_, n, S, False) for i_, o_ in dims], first_needs_dx=False,
                                          fused_norm=True),
                  layers=[(B, i_, o_, n, S, False) for i_, o_ in dims], x=x, first_needs_dx=False)
        return wl

    if cfg["kind"] == "conv":
        n, S, B = cfg["n"], cfg["S"], cfg["B"]
        nn = torch.nn
        torch.manual_seed(1234)

        class Net(nn.Module):  # topology of the reference's examples/cifar10.py:54-102 with 100 classes
            def __init__(self):
                super().__init__()
                self.conv1 = hol.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=3,
                                                               out_channels=6, kernel_size=5, device=str(device))
                self.conv2 = hol.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=6,
                                                               out_channels=16, kernel_size=5, device=str(device))
                self.pool = nn.MaxPool2d(2, 2)
                self.fc = nn.Linear(16 * 5 * 5, 100)

            def forward(self, x):
                x = self.pool(self.conv1(x))
                x = self.pool(self.conv2(x))
                return self.fc(x.flatten(1))

        net = Net().to(device)
        with torch.no_grad():
            net.conv1.conv.weight.uniform_(-0.2, 0.2)
            net.conv2.conv.weight.uniform_(-0.2, 0.2)
        x = make_inputs(torch, (B, 3, 32, 32), gen, device)
        target = torch.randint(0, 100, (B,), generator=gen, device=device)
        flat = FlatGradients(net.parameters(), collective=args.collective, defer=defer_mode(args))

        def step():
            flat.begin()
            loss = torch.nn.functional.cross_entropy(net(x), target)
            loss.backward()
            flat.finish()
            return loss

        def e2e_compute(dev):
            flat.begin()
            loss = torch.nn.functional.cross_entropy(net(dev["x"]), dev["t"])
            loss.backward()
            flat.finish()
            return loss

        rows1, rows2 = B * 28 * 28, B * 10 * 10
        b1, f1 = layer_bytes_flops(rows1, 75, 6, n, S, False, first_layer=True)
        b2, f2 = layer_bytes_flops(rows2, 150, 16, n, S, False)
        wl.update(model=net, flat=flat, step=step, e2e_compute=e2e_compute, samples=B,
                  host=dict(x=x.cpu().pin_memory(), t=target.cpu().pin_memory()),
                  h2d=x.numel() * 4 + target.numel() * 8, d2h=4, bytes=b1 + b2, flops=f1 + f2,
                  launches=count_launches([(rows1, 75, 6, n, S, False), (rows2, 150, 16, n, S, False)],
                                          first_needs_dx=False, conv=True),
                  layers=[(rows1, 75, 6, n, S, False), (rows2, 150, 16, n, S, False)], x=x, first_needs_dx=False)
        return wl
    raise KeyError(cfg["kind"])


SUB_STEPS = {"c2": 5, "c4": 20, "c5": 2, "c3": 20, "c1": 20}  # timed steps of a workload run as a sub-block


def time_workload(torch, dist, key, cfg, device, rank, world, steps, warmup, args, with_e2e=True):
    """Build one workload, time `steps` steps of it (CUDA events, barrier + synchronize on both sides,
    max over ranks), optionally its end-to-end leg, and its per-kernel roofline block."""
    local_rank = device.index
    wl = build_workload(torch, cfg, device, rank, args)
    wl["key"] = key if not (args.n or args.segments or args.batch) or key != args.config else None
    step = wl["step"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    # Small-step workloads (MLP / conv nets: ~1 ms of GPU work in ~50 launches) are launch-latency
    # bound in eager mode: capture the whole step (forward, backward, the overlapped all-reduces) in
    # one CUDA graph and replay it.  The single-layer configs run for tens of ms and stay eager.
    graphed, graph = False, None
    if args.graph == "on" or (args.graph == "auto" and cfg["kind"] in ("mlp", "conv")):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            step = graph.replay
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            graphed = True
        except Exception as e:  # capture is an optimisation, never a requirement
            if rank == 0:
                print(f"bench.py: CUDA graph capture failed ({type(e).__name__}: {str(e)[:120]}); running eagerly",
                      file=sys.stderr)
            torch.cuda.synchronize()
            graph = None
            step = wl["step"]
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public module API: pinned host inputs -> H2D -> fwd+bwd -> loss D2H
    # Every step copies ITS inputs from pinned host memory and reads ITS loss back.  Three device input
    # buffers and three pinned loss words: the copies of steps k+1 and k+2 run on a second stream while step
    # k computes (the usual input prefetch; with only two buffers the copy of step k+1 has to fit exactly
    # inside step k, and under the step's own HBM traffic it does not: 0.47 ms alone, ~0.64 ms next to the
    # 0.59 ms config-3 step), and the host reads the loss of step k-1 while step k runs (the usual
    # one-step-lagged logging read) on a third stream, so neither the host nor a copy idles the GPU between
    # steps.  The first copy and the last read-back are exposed.
    e2e_steps, e2e_ms = 0, 0.0
    if with_e2e:
        NBUF = 3
        # launch-bound workloads (~1 ms steps): enough steps that the one exposed input copy at the start is not
        # a tenth of the measurement
        e2e_steps = max(steps, 50) if graphed else max(1, min(steps, 20))
        host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(NBUF)]
        main_stream = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        d2h_stream = torch.cuda.Stream()   # the loss read-back: neither behind the next input copy nor in front of the next step
        dev_bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in wl["host"].items()}
                    for _ in range(NBUF)]

        # launch-bound workloads replay the captured step here too: one graph per input buffer
        e2e_graphs = None
        if graphed:
            try:
                e2e_graphs = []
                for slot in range(NBUF):
                    for k, v in wl["host"].items():
                        dev_bufs[slot][k].copy_(v)
                    torch.cuda.synchronize()
                    g2 = torch.cuda.CUDAGraph()
                    # the graphs never run concurrently: one memory pool, so the workspaces of a step sit at the
                    # same addresses whichever input buffer it reads (what a training loop's allocator gives)
                    with torch.cuda.graph(g2, pool=(e2e_graphs[0][0].pool() if e2e_graphs else None)):
                        loss2 = wl["e2e_compute"](dev_bufs[slot])
                    e2e_graphs.append((g2, loss2))
                torch.cuda.synchronize()
            except Exception as e:
                if rank == 0:
                    print(f"bench.py: e2e graph capture failed ({type(e).__name__}: {str(e)[:120]}); eager e2e",
                          file=sys.stderr)
                torch.cuda.synchronize()
                e2e_graphs = None

        def run_e2e(nsteps):
            ready = [torch.cuda.Event() for _ in range(NBUF)]
            consumed = [None] * NBUF
            landed = [None] * NBUF
            losses = []

            def prefetch(slot):
                with torch.cuda.stream(copy_stream):
                    if consumed[slot] is not None:
                        copy_stream.wait_event(consumed[slot])
                    for k, v in wl["host"].items():
                        dev_bufs[slot][k].copy_(v, non_blocking=True)
                    ready[slot].record(copy_stream)

            for k in range(min(NBUF - 1, nsteps)):
                prefetch(k)
            for k in range(nsteps):
                slot = k % NBUF
                if k + NBUF - 1 < nsteps:
                    prefetch((k + NBUF - 1) % NBUF)   # the buffer of step k-1, whose compute is already queued
                main_stream.wait_event(ready[slot])
                if landed[slot] is not None:   # the loss word of this slot's graph has been read (NBUF steps ago)
                    main_stream.wait_event(landed[slot])
                if e2e_graphs is not None:
                    e2e_graphs[slot][0].replay()
                    loss = e2e_graphs[slot][1]
                else:
                    loss = wl["e2e_compute"](dev_bufs[slot])
                ev = torch.cuda.Event()
                ev.record(main_stream)
                consumed[slot] = ev
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(ev)
                    src = loss.detach()
                    src.record_stream(d2h_stream)
                    host_loss[slot].copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(d2h_stream)
                    landed[slot] = ev
                if k >= 1:  # the previous step's loss, read while this step runs
                    prev = (k - 1) % NBUF
                    landed[prev].synchronize()
                    losses.append(float(host_loss[prev]))
            last = (nsteps - 1) % NBUF
            landed[last].synchronize()
            losses.append(float(host_loss[last]))
            assert len(losses) == nsteps and all(v == v for v in losses), "e2e: a step's loss was not read back"
            return losses[-1]

        run_e2e(2)
        barrier()
        t0 = time.perf_counter()
        run_e2e(e2e_steps)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        # the input copy alone (idle GPU): when it is as long as the step, the end-to-end rate is the PCIe rate
        hc0, hc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            hc0.record(copy_stream)
            for _ in range(4):
                for k, v in wl["host"].items():
                    dev_bufs[0][k].copy_(v, non_blocking=True)
            hc1.record(copy_stream)
        copy_stream.synchronize()
        h2d_alone_ms = hc0.elapsed_time(hc1) / 4
        if e2e_graphs is not None:
            for g2, _ in e2e_graphs:
                g2.reset()
        del dev_bufs, e2e_graphs

    tm = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = (float(v) for v in tm.tolist())
    ms_per_step = ms_total / steps
    res = {
        "workload": cfg["name"], "per_gpu_batch": cfg["B"], "global_batch": cfg["B"] * world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "value": wl["samples"] * world / (ms_per_step / 1e3),
        "unit": "samples/s", "cuda_graph": graphed, "gpu_launches_per_step": wl["launches"],
        "l2": "inputs larger than L2 (no flush needed)" if wl["bytes"] > 3 * 126e6
              else "working set fits L2; no flush between steps",
        "clocks": clocks, "collective": wl["flat"].collective if world > 1 else "none",
    }
    wl["flat"].check()
    if with_e2e:
        res["e2e"] = {"value": wl["samples"] * world / (e2e_ms / e2e_steps / 1e3), "unit": "samples/s",
                      "h2d_bytes_per_step": wl["h2d"], "d2h_bytes_per_step": wl["d2h"], "steps": e2e_steps,
                      "ms_per_step": e2e_ms / e2e_steps, "h2d_ms_alone": h2d_alone_ms,
                      "h2d_gbs_alone": wl["h2d"] / (h2d_alone_ms * 1e6),
                      "note": "module API; pinned-host inputs copied H2D and the loss read back D2H every step; "
                              "the copies of steps k+1 / k+2 overlap the compute of step k (three input buffers) and the host reads the "
                              "loss of step k-1 while step k runs (one-step-lagged read, every step's loss is read)"
                              + ("; the step is replayed from a CUDA graph captured per input buffer" if graphed else "")}
    peaks = load_peaks()
    res["step_algorithmic"] = {
        "bytes": wl["bytes"], "flops": wl["flops"],
        "hbm_frac_of_measured": wl["bytes"] / (ms_per_step / 1e3) / (peaks["hbm_gbs"] * 1e9),
        "fp32_simt_frac_of_74.4TF": wl["flops"] / (ms_per_step / 1e3) / (FP32_PEAK_TFLOPS * 1e12),
        "binding_roof": "compute (tensor pipe on the tcgen05 path, fp32 FMA issue / shared memory on the gather "
                        "path), not HBM"}
    if rank == 0:
        try:
            res["roofline"] = measure_roofline(torch, wl, ms_per_step)
        except torch.OutOfMemoryError as e:  # the largest sweep points leave no room for a second workspace
            res["roofline"] = {"error": f"out of memory in the diagnostic pass: {str(e)[:80]}"}
    # release everything of this workload before the next one is built
    if graph is not None:
        graph.reset()
    del graph, step, wl
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    barrier()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this build has no CPU path; use --impl reference for the CPU baseline)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a hung collective must abort the run, not burn the box: the NCCL watchdog kills the process
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=240))
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    if args.plan_overrides:
        from high_order_layers_torch_b200 import ops as _ops
        _ops.set_plan_overrides(args.plan_overrides)
        if rank == 0:
            print(f"bench.py: planner overrides {args.plan_overrides:#x} (diagnostic run)", file=sys.stderr)
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    head_key = "c3" if args.config == "all" else args.config
    cfg = dict(CONFIGS[head_key])
    if args.config != "all":
        if args.n:
            cfg["n"] = args.n
        if args.segments:
            cfg["S"] = args.segments
        if args.batch:
            cfg["B"] = args.batch
    head = time_workload(torch, dist, head_key, cfg, device, rank, world, steps, warmup, args)

    subs = {}
    if args.config == "all":
        for key, over in (("c2", {}), ("c4", {}), ("c5", {"n": 4, "S": 64})):
            scfg = dict(CONFIGS[key])
            scfg.update(over)
            if over:
                scfg["name"] = scfg["name"].replace("sweep point (n, segments from --n/--segments)",
                                                    f"sweep point n={over['n']} segments={over['S']}")
            try:
                subs[key if not over else f"c5_n{over['n']}_s{over['S']}"] = time_workload(
                    torch, dist, key, scfg, device, rank, world, SUB_STEPS[key], 3, args, with_e2e=(key != "c5"))
            except Exception as e:  # a sub-block never costs the headline line
                torch.cuda.synchronize()
                subs[key] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    cpu = None
    ref_cuda = None
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(cfg, budget_s=args.cpu_budget, key=head_key)
        if world == 1 and not args.no_ref_cuda:
            ref_cuda = ref_torch_cuda(cfg, head_key)
            if "c2" in subs and "error" not in subs["c2"]:
                subs["c2"]["ref_torch_cuda"] = ref_torch_cuda(dict(CONFIGS["c2"]), "c2")
    if world > 1:
        dist.barrier()

    if rank == 0:
        out = {
            "metric": "fwd+bwd samples/sec", "value": head["value"], "unit": "samples/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": head["workload"], "per_gpu_batch": head["per_gpu_batch"],
                       "global_batch": head["global_batch"], "parallelism": f"dp{world}", "l2": head["l2"],
                       "allreduce": ("per-layer fp32 mean all-reduce of dW on a side stream, issued right after the layer's "
                                     "dW kernels and overlapped with the rest of the backward; " +
                                     ("own peer-memory kernels over NVLink (csrc/p2p.cu)" if head["collective"] == "p2p"
                                      else "NCCL" if head["collective"] == "nccl"
                                      else "SKIPPED (--collective none: diagnostic run, not a valid data-parallel number)"))
                       if world > 1 else "none", "cuda_graph": head["cuda_graph"]},
            "e2e": head.get("e2e"),
            "gpu_launches": head["gpu_launches_per_step"] * steps,
            "gpu_launches_per_step": head["gpu_launches_per_step"],
            "clocks": head["clocks"],
            "roofline": head.get("roofline"),
            "cpu_baseline": cpu,
            "ref_torch_cuda": ref_cuda,
            "step_algorithmic": head["step_algorithmic"],
            "configs": subs,
        }
        print(json.dumps(out))
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # Round 1: with a CUDA graph that captured the all-reduce still alive, destroy_process_group() never
        # returned.  The graphs are reset above (time_workload); should the teardown hang all the same, a
        # watchdog ends the process -- every rank has passed the barrier and rank 0 has printed its line.
        def _bail():
            time.sleep(20.0)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture of the same command; keyed by (config, dominant pass).  None where no
# capture exists.
NCU_TRAFFIC = {
    ("c2", "forward"): 20.051674e9 + 0.257603e9,
    ("c2", "dw"): 20.371956e9 + 0.167338e9,
    ("c2", "dx"): 4.371416e9 + 0.270849e9,
    ("c3", "forward"): 340.866e6 + 12.938e6,   # layer 1 (784 -> 128), the layer the roofline block profiles
    ("c3", "dw"): 347.057e6 + 30.321e6,
}
NCU_TRAFFIC_SOURCE = "profiles/r2_c2_ncu_full_summary_final.md / r2_c3_ncu_full_summary_final.md (ncu --set full, same command)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        p["source"] = "measured (MEASURED_PEAKS.json)"
        return p
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def measure_roofline(torch, wl, ms_per_step):
    """Per-kernel CUDA-event times of the widest layer of the workload (hol_pw_profile_f32 records
    events between the kernels and around every GEMM launch on the launching stream).

    `frac` is the contract's roofline fraction of the dominant kernel: its ALGORITHMIC flops
    (tensor-core GEMM: 2*B*I*O*n of the pass) or bytes (gather kernels) per launch divided by
    its measured duration and by the measured peak.  The tensor pipe's utilisation by the flops the
    split-fp16 dense form actually executes (3*Kin/n times the algorithmic ones) is reported beside
    it as `tensor_pipe_util`.  Only passes the real step runs are considered (the first layer of a
    network has no dX)."""
    from high_order_layers_torch_b200 import ops

    peaks = load_peaks()
    layers = wl["layers"]
    k_big = max(range(len(layers)), key=lambda k: layers[k][0] * layers[k][1] * layers[k][2] * layers[k][3])
    B, I, O, n, S, disc = layers[k_big]
    runs_dx = wl.get("first_needs_dx", True) or k_big > 0
    dev = wl["x"].device
    if wl["cfg"]["kind"] == "layer":
        # reuse the workload's own tensors (the 4096x4096 sweep points do not fit twice) and drop its gradients
        x, gy, w = wl["x"].detach(), wl["gy"], wl["model"].w.detach()
        for prm in wl["flat"].params:
            prm.grad = None
        wl["flat"].flat = None
        wl["x"].grad = None
        torch.cuda.empty_cache()
    else:
        gen = torch.Generator(device=dev).manual_seed(99)
        x = make_inputs(torch, (B, I), gen, dev)
        w = torch.rand(O, I, nb_of(n, S, disc), generator=gen, device=dev) * 2 - 1
        gy = torch.randn(B, O, generator=gen, device=dev)
    acc = {}
    reps = 3 if B * I * O * n < 2e12 else 1
    ops.profile_layer(x, w, gy, n, S, 2.0, disc, None)  # warm
    for _ in range(reps):
        for k, v in ops.profile_layer(x, w, gy, n, S, 2.0, disc, None).items():
            acc[k] = acc.get(k, 0.0) + v / reps
    layer = {"B": B, "I": I, "O": O, "n": n, "segments": S, "discontinuous": disc,
             "passes_in_the_step": ["forward", "dw"] + (["dx"] if runs_dx else [])}
    plan = ops.plan_flags(B, I, O, n, S, disc)
    tc_of = {"forward": plan["tc_forward"], "dx": plan["tc_dx"], "dw": plan["tc_dw"]}
    gemm_key = {"forward": "gemm_fwd", "dx": "gemm_dx", "dw": "gemm_dw"}
    kname = {"forward": "pw_fwd_kernel", "dx": "pw_dx_kernel", "dw": "pw_dw_dense_kernel" if plan["dense_dw"] else "pw_dw_kernel"}
    passes = {}
    for ps in layer["passes_in_the_step"]:
        by, fl = kernel_bytes_flops(ps, B, I, O, n, S, disc)
        ms = acc[gemm_key[ps]] if tc_of[ps] and acc.get(gemm_key[ps], 0.0) > 0.0 else acc[ps]
        t = max(ms, 1e-6) / 1e3
        traffic = NCU_TRAFFIC.get((wl.get("key"), ps))
        passes[ps] = {"kernel": f"pw_tc_gemm_kernel ({ps})" if tc_of[ps] else kname[ps], "launch_ms": ms,
                      "stage_ms": acc[ps], "algorithmic_bytes": by, "algorithmic_flops": fl,
                      "useful_tflops": fl / t / 1e12, "algorithmic_gbs": by / t / 1e9,
                      "hbm_frac": by / t / 1e9 / peaks["hbm_gbs"],
                      "fp32_simt_frac_of_74.4TF": fl / t / 1e12 / FP32_PEAK_TFLOPS,
                      "traffic": traffic, "traffic_ratio": (traffic / by) if traffic else None}
        if tc_of[ps]:
            executed = tc_gemm_flops(ps, B, I, O, n, S, disc)
            tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
            passes[ps]["tensor_pipe_util"] = {"executed_flops": executed, "executed_tflops": executed / t / 1e12,
                                              "peak": tpeak, "frac": executed / t / 1e12 / tpeak}
    dom = max(passes, key=lambda k: passes[k]["launch_ms"])
    d = passes[dom]
    hbm_note = ("this contraction is compute bound (arithmetic intensity >> ridge), so the HBM fraction of its "
                "algorithmic bytes is small by construction")
    if tc_of[dom]:
        tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        out = {"bound": "tensor", "kernel": f"pw_tc_gemm_kernel ({dom} pass: split-fp16 hi*hi + hi*lo + lo*hi)",
               "achieved": d["useful_tflops"], "peak": tpeak, "unit": "TFLOP/s", "frac": d["useful_tflops"] / tpeak,
               "definition": "algorithmic flops of the pass (2*B*I*O*n) / CUDA-event duration of the GEMM launch / "
                             "measured cuBLAS bf16 rate; the executed-flops utilisation is tensor_pipe_util",
               "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step); burst "
                              f"{peaks['bf16_tflops']}",
               "tensor_pipe_util": d["tensor_pipe_util"]}
    else:
        out = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["algorithmic_gbs"], "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "frac": d["hbm_frac"], "peak_source": peaks["source"],
               "fp32": {"achieved_tflops": d["useful_tflops"], "peak_tflops": FP32_PEAK_TFLOPS,
                        "frac": d["fp32_simt_frac_of_74.4TF"]},
               "note": hbm_note + "; the gather kernels are bound by one shared-memory word per FMA: see fp32.frac"}
    out.update({"traffic": d["traffic"], "traffic_ratio": d["traffic_ratio"],
                "traffic_source": NCU_TRAFFIC_SOURCE if d["traffic"] else None,
                "dominant_pass": dom, "launch_ms": d["launch_ms"], "passes": passes, "kernel_ms": acc,
                "layer": layer, "plan": plan})
    return out


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------
def _reference_modules():
    """The unmodified reference from baseline/_ref (pip --target install, git-ignored), or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "high_order_layers_torch")):
        return None
    import types
    import torch
    if "pytorch_lightning" not in sys.modules:
        stub = types.ModuleType("pytorch_lightning")
        stub.LightningModule = type("LightningModule", (torch.nn.Module,), {})
        sys.modules["pytorch_lightning"] = stub
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import high_order_layers_torch.layers as L  # noqa
        import high_order_layers_torch.networks as N  # noqa
        return types.SimpleNamespace(layers=L, networks=N)
    except Exception as e:  # pragma: no cover
        print(f"bench.py: reference import failed ({e}); using the oracle port", file=sys.stderr)
        return None


def cpu_step_factory(cfg, chunk, device="cpu"):
    """One bounded step (fwd+bwd on `chunk` samples) of the workload through the UNMODIFIED reference
    modules (or the numpy oracle port when baseline/_ref is absent); returns (fn, kind, samples).
    device="cuda" runs the reference's own torch-CUDA path: construction and calls happen under
    `with torch.device("cuda")` (the reference builds its node tables on the default device)."""
    import numpy as np
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    R = _reference_modules()
    rng = np.random.default_rng(0)
    if device != "cpu":
        if R is None:
            raise RuntimeError("the torch-CUDA reference arm needs baseline/_ref")
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        inner, _, _ = _on_device(cfg, chunk, device, R, rng)
        return inner, "reference", chunk
    if cfg["kind"] == "layer":
        n, S, I, O = cfg["n"], cfg["S"], cfg["I"], cfg["O"]
        disc = cfg["layer"] == "discontinuous"
        x = rng.uniform(-1, 1, size=(chunk, I)).astype(np.float32)
        gy = rng.standard_normal(size=(chunk, O)).astype(np.float32)
        if R is not None:
            layer = R.layers.high_order_fc_layers(cfg["layer"], n=n, in_features=I, out_features=O, segments=S)
            xt = torch.from_numpy(x).requires_grad_(True)
            gt = torch.from_numpy(gy)

            def fn():
                layer.w.grad = None
                xt.grad = None
                layer(xt).backward(gt)
            return fn, "reference", chunk
        from oracle import pw_oracle as orc
        w = rng.uniform(-1, 1, size=(O, I, nb_of(n, S, disc))).astype(np.float32)

        def fn():
            orc.pw_forward(x, w, n, S, 2.0, disc, dtype=np.float32)
            orc.pw_backward(gy, x, w, n, S, 2.0, disc, dtype=np.float32)
        return fn, "port", chunk
    if cfg["kind"] == "mlp":
        n, S, wd = cfg["n"], cfg["S"], cfg["widths"]
        x = rng.uniform(-1, 1, size=(chunk, wd[0])).astype(np.float32)
        if R is not None:
            net = R.networks.HighOrderMLP(layer_type=cfg["layer"], n=n, in_width=wd[0], out_width=wd[-1],
                                          hidden_layers=len(wd) - 3, hidden_width=wd[1], in_segments=S, out_segments=S,
                                          hidden_segments=S, normalization=R.layers.MaxAbsNormalization)
            xt = torch.from_numpy(x)
            classify = wd[-1] > 1
            tgt = torch.randint(0, wd[-1], (chunk,)) if classify else torch.randn(chunk, wd[-1])
            lossf = torch.nn.functional.cross_entropy if classify else torch.nn.functional.mse_loss

            def fn():
                net.zero_grad(set_to_none=True)
                lossf(net(xt), tgt).backward()
            return fn, "reference", chunk
        from oracle import pw_oracle as orc
        ws = [rng.uniform(-1, 1, size=(o, i, nb_of(n, S, False))).astype(np.float32) for i, o in zip(wd[:-1], wd[1:])]
        gy = rng.standard_normal(size=(chunk, wd[-1])).astype(np.float32)

        def fn():
            orc.mlp_forward_backward(x, ws, gy, n, S, dtype=np.float32)
        return fn, "port", chunk
    if cfg["kind"] == "conv":
        n, S = cfg["n"], cfg["S"]
        x = rng.uniform(-1, 1, size=(chunk, 3, 32, 32)).astype(np.float32)
        if R is not None:
            nn = torch.nn
            c1 = R.layers.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=3, out_channels=6,
                                                        kernel_size=5)
            c2 = R.layers.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=6,
                                                        out_channels=16, kernel_size=5)
            pool, fc = nn.MaxPool2d(2, 2), nn.Linear(400, 100)
            xt = torch.from_numpy(x)
            tgt = torch.randint(0, 100, (chunk,))
            params = list(c1.parameters()) + list(c2.parameters()) + list(fc.parameters())

            def fn():
                for p in params:
                    p.grad = None
                out = fc(pool(c2(pool(c1(xt)))).flatten(1))
                torch.nn.functional.cross_entropy(out, tgt).backward()
            return fn, "reference", chunk
        from oracle import pw_oracle as orc
        w1 = rng.uniform(-1, 1, size=(6, 3 * nb_of(n, S, False), 5, 5)).astype(np.float32)

        def fn():
            orc.pw_conv2d_forward(x, w1, n, S, dtype=np.float32)
        return fn, "port", chunk
    raise KeyError(cfg["kind"])


def _on_device(cfg, chunk, device, R, rng):
    """The reference modules of the workload on `device`; returns (step fn, module, tensors)."""
    import numpy as np
    import torch

    dev = torch.device(device)
    with dev:
        if cfg["kind"] == "layer":
            n, S, I, O = cfg["n"], cfg["S"], cfg["I"], cfg["O"]
            layer = R.layers.high_order_fc_layers(cfg["layer"], n=n, in_features=I, out_features=O, segments=S,
                                                  device=device).to(dev)
            with torch.no_grad():
                layer.w.uniform_(-1, 1)
            xt = torch.from_numpy(rng.uniform(-1, 1, size=(chunk, I)).astype(np.float32)).to(dev).requires_grad_(True)
            gt = torch.from_numpy(rng.standard_normal(size=(chunk, O)).astype(np.float32)).to(dev)

            def fn():
                with dev:
                    layer.w.grad = None
                    xt.grad = None
                    y = layer(xt)
                    y.backward(gt)
                    return y
            return fn, layer, dict(x=xt, gy=gt)
        if cfg["kind"] == "mlp":
            n, S, wd = cfg["n"], cfg["S"], cfg["widths"]
            net = R.networks.HighOrderMLP(layer_type=cfg["layer"], n=n, in_width=wd[0], out_width=wd[-1],
                                          hidden_layers=len(wd) - 3, hidden_width=wd[1], in_segments=S, out_segments=S,
                                          hidden_segments=S, normalization=R.layers.MaxAbsNormalization,
                                          device=device).to(dev)
            xt = torch.from_numpy(rng.uniform(-1, 1, size=(chunk, wd[0])).astype(np.float32)).to(dev)
            classify = wd[-1] > 1
            tgt = torch.randint(0, wd[-1], (chunk,), device=dev) if classify else torch.randn(chunk, wd[-1], device=dev)
            lossf = torch.nn.functional.cross_entropy if classify else torch.nn.functional.mse_loss

            def fn():
                with dev:
                    net.zero_grad(set_to_none=True)
                    loss = lossf(net(xt), tgt)
                    loss.backward()
                    return loss
            return fn, net, dict(x=xt, t=tgt)
        if cfg["kind"] == "conv":
            n, S = cfg["n"], cfg["S"]
            nn = torch.nn
            c1 = R.layers.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=3, out_channels=6,
                                                        kernel_size=5, device=device)
            c2 = R.layers.high_order_convolution_layers("continuous2d", n=n, segments=S, in_channels=6,
                                                        out_channels=16, kernel_size=5, device=device)
            pool, fc = nn.MaxPool2d(2, 2), nn.Linear(400, 100)
            mods = nn.ModuleList([c1, c2, fc]).to(dev)
            xt = torch.from_numpy(rng.uniform(-1, 1, size=(chunk, 3, 32, 32)).astype(np.float32)).to(dev)
            tgt = torch.randint(0, 100, (chunk,), device=dev)

            def fn():
                with dev:
                    mods.zero_grad(set_to_none=True)
                    loss = torch.nn.functional.cross_entropy(fc(pool(c2(pool(c1(xt)))).flatten(1)), tgt)
                    loss.backward()
                    return loss
            return fn, mods, dict(x=xt, t=tgt)
    raise KeyError(cfg["kind"])


def ref_torch_cuda(cfg, key, budget_s=8.0):
    """The reference's own torch-CUDA path on this GPU (north_star: reported next to ours): the
    unmodified modules under `with torch.device("cuda")`, fp32, TF32 off, micro-batched so that its
    [B, I, O, n] gathered tensors fit, CUDA-event timed.  For single-layer workloads the same
    micro-batch is also pushed through our layer with the reference's weights (parity in the run)."""
    import numpy as np
    import torch

    R = _reference_modules()
    if R is None:
        return {"unavailable": "baseline/_ref not installed"}
    chunk = GPU_REF_CHUNK[key]
    rng = np.random.default_rng(0)
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        fn, mod, tens = _on_device(cfg, chunk, "cuda", R, rng)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t0 = 0, time.perf_counter()
        ev0.record()
        while True:
            fn()
            reps += 1
            if reps % 4 == 0:
                torch.cuda.synchronize()
            if time.perf_counter() - t0 > budget_s or reps >= 200:
                break
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        out = {"value": chunk * reps / (ms / 1e3), "unit": "samples/s", "micro_batch": chunk, "steps": reps,
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
               "sample": f"{reps} micro-batches of {chunk} samples (fwd+bwd, fp32, TF32 off) through the unmodified "
                         f"reference modules under torch.device('cuda'), CUDA events, {ms / 1e3:.1f} s"}
        if cfg["kind"] == "layer":
            import high_order_layers_torch_b200 as hol
            ours = hol.high_order_fc_layers(cfg["layer"], n=cfg["n"], in_features=cfg["I"], out_features=cfg["O"],
                                            segments=cfg["S"], device="cuda")
            with torch.no_grad():
                ours.w.copy_(mod.w)
            y_ref = fn()
            x2 = tens["x"].detach().clone().requires_grad_(True)
            y = ours(x2)
            y.backward(tens["gy"])
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            out["parity_rel_err"] = {"y": rel(y, y_ref), "dx": rel(x2.grad, tens["x"].grad),
                                     "dw": rel(ours.w.grad, mod.w.grad)}
        return out
    except Exception as e:  # a reported baseline, never a reason to lose the bench line
        torch.cuda.synchronize()
        return {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    finally:
        torch.cuda.empty_cache()


CPU_CHUNK = {"c2": 64, "c5": 8, "c3": 256, "c1": 1024, "c4": 128}
GPU_REF_CHUNK = {"c2": 256, "c5": 16, "c3": 2048, "c1": 1024, "c4": 1024}


def cpu_baseline(cfg, budget_s=15.0, key=None):
    """Time the CPU implementation on a bounded sample of the workload (about `budget_s` seconds)."""
    key = key or [k for k, v in CONFIGS.items() if v["name"] == cfg["name"]][0]
    chunk = CPU_CHUNK[key]
    fn, kind, samples = cpu_step_factory(cfg, chunk)
    fn()  # warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        fn()
        reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 50:
            break
    import torch
    return {"value": samples * reps / el, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{reps} micro-batches of {samples} samples of the same layer shape (fwd+bwd, fp32, "
                      f"{'unmodified reference modules on torch CPU' if kind == 'reference' else 'numpy oracle port'}), "
                      f"{el:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    key = "c3" if args.config == "all" else args.config   # the headline workload of the default run
    cfg = dict(CONFIGS[key])
    if args.n:
        cfg["n"] = args.n
    if args.segments:
        cfg["S"] = args.segments
    chunk = CPU_CHUNK[key]
    fn, kind, samples = cpu_step_factory(cfg, chunk)
    import torch
    warmup = max(1, min(args.warmup, 3))
    steps = max(1, args.steps)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        fn()
        done += 1
        if time.perf_counter() - t0 > 150:  # keep the whole run within a few minutes
            break
    el = time.perf_counter() - t0
    value = samples * done / el
    sample = (f"each step = one micro-batch of {samples} samples of the workload's layer shapes, fwd+bwd fp32 on the "
              f"host CPU ({'unmodified reference via baseline/_ref' if kind == 'reference' else 'numpy oracle port'})")
    print(json.dumps({
        "impl": "reference", "metric": "fwd+bwd samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": done, "warmup": warmup, "ms_per_step": el / done * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "per_gpu_batch": cfg["B"], "global_batch": cfg["B"] * world,
                   "parallelism": "cpu", "note": "CPU reference arm; rank 0 only"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="all", choices=["all"] + sorted(CONFIGS),
                    help="all (default): config 3 as the headline + c2 / c4 / one c5 point as sub-blocks")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override n (c5 sweep)")
    ap.add_argument("--segments", type=int, default=0, help="override segments (c5 sweep)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's torch-CUDA leg")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl", "none"],
                    help="dW all-reduce: our peer-memory kernels over NVLink (p2p; auto = p2p when available) or NCCL")
    ap.add_argument("--no-defer", action="store_true",
                    help="data parallel: every layer's dW in backward order (no largest-message-first deferral)")
    ap.add_argument("--force-defer", action="store_true", help="defer dW even with one process (diagnostic)")
    ap.add_argument("--plan-overrides", type=int, default=0,
                    help="diagnostic HOL_PLAN_* mask for the library's planner (include/hol_b200.h), e.g. 32 = no auxiliary stream")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="capture the step in a CUDA graph (auto: MLP / conv workloads only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
