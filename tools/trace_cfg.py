"""Kernel-level time table of one bench step for any config (diagnostic): python tools/trace_cfg.py c3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--segments", type=int, default=0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--rows", type=int, default=30)
a = ap.parse_args()
cfg = dict(bench.CONFIGS[a.config])
if a.n:
    cfg["n"] = a.n
if a.segments:
    cfg["S"] = a.segments
if a.batch:
    cfg["B"] = a.batch
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.build_workload(torch, cfg, dev, 0, a)
for _ in range(3):
    wl["step"]()
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    wl["step"]()
t1.record()
torch.cuda.synchronize()
print(f"{a.config}: {t0.elapsed_time(t1) / 5:.3f} ms/step (events, 5 steps)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl["step"]()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=90))
