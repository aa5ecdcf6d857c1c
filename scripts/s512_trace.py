"""per-layer table of the OMGSR-S-512 step: run with OMGSR_TRACE=1; every GEMM / conv launch prints its shape, the
kernel variant and its CUDA-event time to stderr. Stage markers are written between the stages of the last pass."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infer.omgsr_s_infer_model import OMGSR_S_Infer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
x = (torch.rand(B, 3, 512, 512) * 2 - 1).cuda().bfloat16()
ctx = torch.randn(1, 77, 1024).cuda().bfloat16()
with torch.no_grad():
    for it in range(3):
        sys.stderr.write(f"MARK pass {it} enc\n"); sys.stderr.flush()
        z = m._encode(x); torch.cuda.synchronize()
        sys.stderr.write(f"MARK pass {it} unet\n"); sys.stderr.flush()
        eh = m.unet.eps_hat_nhwc(z, ctx); torch.cuda.synchronize()
        sys.stderr.write(f"MARK pass {it} dec\n"); sys.stderr.flush()
        img = m._decode_from_eps(z, eh); torch.cuda.synchronize()
