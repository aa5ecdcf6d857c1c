"""a few launches of one GEMM shape through the C ABI (ncu target): gemm_probe.py M N K [res] [act]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi
M, N, K = [int(v) for v in sys.argv[1:4]]
res = len(sys.argv) > 4 and sys.argv[4] == "1"
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) / math.sqrt(K)
bias = None if (len(sys.argv) > 6 and sys.argv[6] == "nobias") else torch.zeros(N, device="cuda")
No = N // 2 if act == capi.ACT_GEGLU else N
out = torch.empty(M, No, device="cuda", dtype=torch.bfloat16)
r = torch.randn(M, No, device="cuda", dtype=torch.bfloat16) if res else None
for _ in range(5):
    capi.gemm(a, w, bias, res=r, out=out, act=act)
torch.cuda.synchronize()
