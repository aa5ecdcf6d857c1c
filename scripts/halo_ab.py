"""A/B timing of the 128-channel halo convolution: row-shared schedule (N = 256 instructions) vs the round-1 kernel.
Run twice: OMGSR_HALO_ROW_SHARED=0|1 python scripts/halo_ab.py; OMGSR_HALO_ROLLING=0|1 selects the rolling-row kernel."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight, pack_upsample_weight

def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

B, H, C = 16, 512, 128
x = torch.randn(B, H, H, C, device="cuda", dtype=torch.bfloat16)
w = pack_conv_weight(torch.randn(C, C, 3, 3) / 34).cuda(); b = torch.zeros(C, device="cuda")
out = torch.empty_like(x); res = torch.randn_like(x)
st = torch.empty(B, 32, 2, device="cuda")
coef = capi.groupnorm_coef(capi.groupnorm_stats(x), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), B, H * H, 32, 1e-6)
fl = 2.0 * C * C * 9 * B * H * H
for name, fn in (("plain", lambda: capi.conv2d(x, w, b, out=out)),
                 ("+gn_stats", lambda: capi.conv2d(x, w, b, out=out, gn_stats=st)),
                 ("+res+gn_stats", lambda: capi.conv2d(x, w, b, res=res, out=out, gn_stats=st)),
                 ("gn_in", lambda: capi.conv2d_gn(x, coef, w, b, out=out)),
                 ("gn_in+gn_stats", lambda: capi.conv2d_gn(x, coef, w, b, out=out, gn_stats=st)),
                 ("gn_in+res+gn_stats", lambda: capi.conv2d_gn(x, coef, w, b, res=res, out=out, gn_stats=st))):
    ms = timeit(fn)
    print(f"3x3 128->128 @512^2 x16 {name:20s} {capi.last_conv_variant():22s} {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
x2 = torch.randn(B, 256, 256, C, device="cuda", dtype=torch.bfloat16)
wu = pack_upsample_weight(torch.randn(C, C, 3, 3) / 34).cuda()
o2 = torch.empty(B, 512, 512, C, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: capi.conv2d(x2, wu, b, upsample2x=True, out=o2, gn_stats=st))
print(f"upsample2x+3x3 128->128 256^2->512^2 x16 {capi.last_conv_variant():22s} {ms:.3f} ms  {2.0*C*C*16*B*256*256/ms/1e9:.0f} TFLOP/s (4 x 2x2 convs)")
