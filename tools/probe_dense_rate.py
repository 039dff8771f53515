"""Sizing probe (not part of the product): what a dense fp16 tensor-core GEMM of the segment-expanded
shape costs on this B200, to decide whether the tcgen05 path can beat the gather kernels."""
import torch, time
dev = "cuda"
def t_mm(M, K, N, reps=5):
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    b = torch.randn(K, N, device=dev, dtype=torch.float16)
    torch.mm(a, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): torch.mm(a, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"M={M} K={K} N={N}: {ms:.3f} ms  {2*M*K*N/ms/1e9:.0f} TFLOP/s")
t_mm(65536, 40960, 1024)      # c2 forward, one pass (x3 for split fp16)
t_mm(65536, 1024, 40960)      # c2 dX shape
t_mm(40960, 65536, 1024)      # c2 dW shape
t_mm(8192, 784 * 16, 128)     # c3 layer 1 forward
t_mm(8192, 4096, 4096)
