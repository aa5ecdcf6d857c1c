"""A/B timing of the fused attention kernels: v3 (128 keys per block) vs v2 (64 keys per block)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi

def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

for name, (B, H, M, N, D) in (("FLUX joint", (4, 24, 4608, 4608, 128)), ("UNet 64x64", (16, 5, 4096, 4096, 64)),
                               ("UNet 32x32", (16, 10, 1024, 1024, 64)), ("UNet 16x16", (16, 20, 256, 256, 64)),
                               ("UNet cross 64x64", (16, 5, 4096, 77, 64))):
    q = torch.randn(B, M, H, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, N, H, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, N, H, D, device="cuda", dtype=torch.bfloat16)
    o = torch.empty_like(q)
    fl = 4.0 * B * H * M * N * D
    res = []
    for variant, nm in ((2, "v2"), (3, "v3"), (4, "v4 (d=128 only)"), (5, "v5")):
        capi.lib().omgsr_flash_attn_set_variant(variant)
        ms = timeit(lambda: capi.flash_attn(q, k, v, D ** -0.5, out=o))
        res.append(f"{nm} {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s")
    capi.lib().omgsr_flash_attn_set_variant(0)
    print(f"{name:18s} B={B} H={H} M={M} N={N} d={D}: " + " | ".join(res), flush=True)
