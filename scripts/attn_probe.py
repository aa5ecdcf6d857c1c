"""a few flash-attention launches (ncu target): attn_probe.py B H M N D"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omgsr_b200 import capi
B, H, M, N, D = [int(v) for v in sys.argv[1:6]]
q = torch.randn(B, M, H, D, device="cuda", dtype=torch.bfloat16)
k = torch.randn(B, N, H, D, device="cuda", dtype=torch.bfloat16)
v = torch.randn(B, N, H, D, device="cuda", dtype=torch.bfloat16)
o = torch.empty_like(q)
for _ in range(4):
    capi.flash_attn(q, k, v, 1 / math.sqrt(D), out=o)
torch.cuda.synchronize()
