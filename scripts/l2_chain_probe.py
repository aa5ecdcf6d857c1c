"""does chunking the batch so that the GroupNorm-apply output stays in L2 for the following conv pay off?"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
B, H, C = 16, 512, 128
if len(sys.argv) > 1: H, C = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(B, H, H, C, device="cuda", dtype=torch.bfloat16)
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
w = pack_conv_weight(torch.randn(C, C, 3, 3) / math.sqrt(9 * C)).cuda()
st = capi.groupnorm_stats(x, 32)
z = torch.empty_like(x); out = torch.empty_like(x)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def whole():
    capi.groupnorm_apply(x, st, g, b, 32, 1e-6, True, out=z)
    capi.conv2d(z, w, b, out=out)
print(f"H={H} C={C} whole batch: {timeit(whole):.3f} ms")
for chunk in (1, 2, 4, 8):
    ztmp = torch.empty(chunk, H, H, C, device="cuda", dtype=torch.bfloat16)
    def chunked():
        for i in range(0, B, chunk):
            capi.groupnorm_apply(x[i:i + chunk], st[i:i + chunk], g, b, 32, 1e-6, True, out=ztmp)
            capi.conv2d(ztmp, w, b, out=out[i:i + chunk])
    print(f"H={H} C={C} chunk={chunk} images ({ztmp.numel() * 2 / 2**20:.0f} MiB temp): {timeit(chunked):.3f} ms")
