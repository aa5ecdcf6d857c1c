import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, math, torch
from omgsr_b200 import capi
from omgsr_b200.pack import pack_conv_weight
mode = sys.argv[1] if len(sys.argv) > 1 else "generic"
B, H, W, Cin, Cout = 16, 512, 512, 128, 128
if len(sys.argv) > 2:
    B, H, W, Cin, Cout = [int(v) for v in sys.argv[2:7]]
x = torch.randn(B, H, W, Cin, device="cuda", dtype=torch.bfloat16)
w = pack_conv_weight(torch.randn(Cout, Cin, 3, 3) / math.sqrt(9 * Cin)).cuda()
b = torch.zeros(Cout, device="cuda")
out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
fl = capi.CONV_NO_HALO if mode == "generic" else 0
if mode == "halo_gn_res":  # the heaviest variant of the step: fused input GroupNorm + residual + output statistics
    coef = capi.groupnorm_coef(capi.groupnorm_stats(x, 32), torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda"),
                               B, H * W, 32, 1e-6)
    res = torch.randn_like(out)
    st = torch.empty(B, 32, 2, device="cuda")
    for _ in range(3):
        capi.conv2d_gn(x, coef, w, b, res=res, out=out, gn_stats=st)
    torch.cuda.synchronize()
    sys.exit(0)
if mode == "halo_gn":  # GroupNorm + SiLU fused into the operand path
    coef = capi.groupnorm_coef(capi.groupnorm_stats(x, 32), torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda"),
                               B, H * W, 32, 1e-6)
    for _ in range(3):
        capi.conv2d_gn(x, coef, w, b, out=out)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    capi.conv2d(x, w, b, out=out, flags=fl)
torch.cuda.synchronize()
