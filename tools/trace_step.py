"""Which torch op launches which kernel in one c2 step (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import high_order_layers_torch_b200 as hol
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layer = hol.PiecewiseDiscontinuousPolynomial(n=5, in_features=1024, out_features=1024, segments=8, device="cuda")
x = (torch.rand(B, 1024, device="cuda") * 2 - 1).requires_grad_(True)
gy = torch.randn(B, 1024, device="cuda")
for _ in range(2):
    layer(x).backward(gy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    layer.w.grad = None; x.grad = None
    layer(x).backward(gy)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
