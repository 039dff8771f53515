"""Accuracy of the split-fp16 tcgen05 path against the fp64 oracle as a function of the number of
k-chunks accumulated in TMEM between drains (hol_set_plan_overrides, HOL_PLAN_FLUSH_SHIFT): the tensor core accumulates with
round-toward-zero, so the bias grows with the chain length; worst case = same-sign terms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from high_order_layers_torch_b200 import ops
from oracle import pw_oracle as orc

n, S, I, O, B, disc = 5, 8, 1024, 256, 512, True
rng = np.random.default_rng(0)
nodes = ops.lobatto_nodes(n).numpy()
for label, wgen, ggen in (("random signs", lambda s: rng.uniform(-1, 1, size=s), lambda s: rng.standard_normal(size=s)),
                          ("same sign", lambda s: rng.uniform(0.5, 1, size=s), lambda s: rng.uniform(0.5, 1, size=s))):
    x = rng.uniform(-1, 1, size=(B, I)).astype(np.float32)
    w = wgen((O, I, n * S)).astype(np.float32)
    gy = ggen((B, O)).astype(np.float32)
    y64 = orc.pw_forward(x, w, n, S, 2.0, disc, None, nodes=nodes)
    dx64, dw64 = orc.pw_backward(gy, x, w, n, S, 2.0, disc, None, nodes=nodes)
    for fc in sys.argv[1:] or ["2", "4", "8"]:
        ops.set_plan_overrides(int(fc) << 8)
        xt = torch.from_numpy(x).cuda().requires_grad_(True)
        wt = torch.from_numpy(w).cuda().requires_grad_(True)
        y = ops.piecewise_polynomial(xt, wt, n, S, 2.0, disc, None)
        y.backward(torch.from_numpy(gy).cuda())
        rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        print(f"{label:13s} flush={fc}: rel err y {rel(y.detach().cpu().numpy(), y64):.2e}  "
              f"dx {rel(xt.grad.cpu().numpy(), dx64):.2e}  dw {rel(wt.grad.cpu().numpy(), dw64):.2e}")
