"""CUDA-event timings of the streaming glue kernels at their largest S-512 shapes."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.randn(16, 512, 512, 128, device="cuda", dtype=torch.bfloat16)
g = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
st = capi.groupnorm_stats(x, 32)
y = torch.empty_like(x)
ms = timeit(lambda: capi.groupnorm_apply(x, st, g, b, 32, 1e-6, True, out=y))
print(f"gn_apply 16x512x512x128: {ms:.3f} ms {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
ms = timeit(lambda: capi.groupnorm_stats(x, 32))
print(f"gn_stats 16x512x512x128: {ms:.3f} ms {x.numel() * 2 / ms / 1e6:.0f} GB/s")
s = torch.randn(16, 1, 4096, 4096, device="cuda", dtype=torch.float32)
p = torch.empty(s.shape, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: capi.softmax_rows(s, 4096, 1 / math.sqrt(512), out=p))
print(f"softmax_rows 16x4096x4096: {ms:.3f} ms {s.numel() * 6 / ms / 1e6:.0f} GB/s")
