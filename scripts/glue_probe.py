"""one GroupNorm apply (+stats) at the largest VAE shape and one flash-attention call at the FLUX shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import math, torch
from omgsr_b200 import capi
x = torch.randn(16, 512, 512, 128, device="cuda", dtype=torch.bfloat16)
g = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
for _ in range(2):
    capi.groupnorm(x, g, b, 32, 1e-6, True)
q = torch.randn(4, 4608, 24, 128, device="cuda", dtype=torch.bfloat16)
k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(2):
    capi.flash_attn(q, k, v, 1 / math.sqrt(128))
torch.cuda.synchronize()
