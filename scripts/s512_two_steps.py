"""exactly two OMGSR-S-512 batch-16 steps through the module (ncu launch-list target: the second one is summarised)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infer.omgsr_s_infer_model import OMGSR_S_Infer
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
x = (torch.rand(16, 3, 512, 512, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda().bfloat16()
ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(2)).cuda().bfloat16()
with torch.no_grad():
    for _ in range(2):
        z = m._encode(x)
        out = m._forward_no_tile(z, ctx)
torch.cuda.synchronize()
print("ok", float(out.float().std()))
