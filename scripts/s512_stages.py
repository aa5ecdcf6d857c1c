"""stage split of the OMGSR-S-512 step (batch 16): VAE encode / UNet / VAE decode, CUDA events."""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infer.omgsr_s_infer_model import OMGSR_S_Infer
from omgsr_b200 import capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with contextlib.redirect_stdout(io.StringIO()):
    m = OMGSR_S_Infer("random://sd21?seed=0", "random://lora?seed=1", 273, "cuda", torch.bfloat16)
x = (torch.rand(B, 3, 512, 512) * 2 - 1).cuda().bfloat16()
ctx = torch.randn(1, 77, 1024).cuda().bfloat16()
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {"enc": [], "unet": [], "dec": []}
with torch.no_grad():
    for it in range(6):
        e = [ev() for _ in range(4)]
        e[0].record(); z = m._encode(x)
        e[1].record(); eh = m.unet.eps_hat_nhwc(z, ctx)
        e[2].record(); img = m._decode_from_eps(z, eh)
        e[3].record(); torch.cuda.synchronize()
        if it >= 2:
            acc["enc"].append(e[0].elapsed_time(e[1])); acc["unet"].append(e[1].elapsed_time(e[2])); acc["dec"].append(e[2].elapsed_time(e[3]))
ms = {k: sum(v) / len(v) for k, v in acc.items()}
fl = {"enc": 1.1167, "unet": 0.8043, "dec": 2.5145}
print(json.dumps({"batch": B, "ms": ms, "tflops_algorithmic": {k: fl[k] * B / ms[k] * 1e3 for k in ms},
                  "total_ms": sum(ms.values()), "images_per_s": B / sum(ms.values()) * 1e3}))
