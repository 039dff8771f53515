#!/bin/bash
# GPU job C: smoke first (a hang in a new kernel must not burn the box), then tests, bench, ncu
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; rc=$?; echo "smoke rc=$rc"; tail -5 gpurun_out/r2_smoke.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python - > gpurun_out/r2_cluster_probe.txt 2>&1 <<'PY'
import torch, numpy as np, time
from high_order_layers_torch_b200 import ops, _lib
from oracle import pw_oracle as orc
# a shape large enough for two-CTA clusters: 1024 rows x 256 -> 512, n=3 S=4
torch.manual_seed(0)
B,I,O,n,S=4096,256,512,3,4
x=(torch.rand(B,I,device='cuda')*2.4-1.2).requires_grad_(True); w=(torch.rand(O,I,9,device='cuda')*2-1).requires_grad_(True); gy=torch.randn(B,O,device='cuda')
for mask,name in ((0,'cluster'),(_lib.HOL_PLAN_NO_CLUSTER,'single')):
    ops.set_plan_overrides(mask)
    x.grad=None; w.grad=None
    y=ops.piecewise_polynomial(x,w,n,S,2.0,False,None); y.backward(gy); torch.cuda.synchronize()
    nodes=ops.lobatto_nodes(n).numpy()
    y64=orc.pw_forward(x.detach()[:128].cpu().numpy(), w.detach().cpu().numpy(), n,S,2.0,False,None,nodes=nodes)
    rel=lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
    print(name, 'y err', rel(y.detach()[:128].cpu().numpy(), y64), 'dx sum', float(x.grad.abs().sum()), 'dw sum', float(w.grad.abs().sum()))
print('plan', ops.plan_flags(B,I,O,n,S,False))
PY
echo "probe rc=$?"; cat gpurun_out/r2_cluster_probe.txt | tail -5
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2_pytest4.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest4.txt
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_all_c.json 2> gpurun_out/r2_bench_all_c.err; echo "bench all rc=$?"; tail -c 600 gpurun_out/r2_bench_all_c.err
