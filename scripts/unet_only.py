"""UNet stage of OMGSR-S-512 alone (batch 16): target for an ncu launch list."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infer.omgsr_s_infer_model import OMGSR_S_Infer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
z = torch.randn(B, 64, 64, 8, device="cuda").bfloat16()[..., :4] if False else None
x = (torch.rand(B, 3, 512, 512) * 2 - 1).cuda().bfloat16()
ctx = torch.randn(1, 77, 1024).cuda().bfloat16()
with torch.no_grad():
    z = m._encode(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(2):
        eh = m.unet.eps_hat_nhwc(z, ctx)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
